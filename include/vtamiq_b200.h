/* vtamiq_b200 — C-ABI of the B200 (sm_100a) VTAMIQ inference hot path.
 *
 * The reference (ch-andrei/VTAMIQ) is pure Python/PyTorch and has no FFI of its own; the functions
 * below are what a `ctypes` binding inside the reference's modules would call in place of the ATen
 * ops on the path.  Each entry cites the reference code it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types; every buffer is a caller-allocated DEVICE
 *     pointer unless the parameter is documented as host memory;
 *   - every call takes an explicit stream (a cudaStream_t passed as void*), is asynchronous with
 *     respect to the host, allocates nothing, and is capturable in a CUDA graph;
 *   - return value: 0 = ok, negative = error (VTQ_ERR_*); the message is kept per handle and read
 *     with vtq_last_error_string();
 *   - "16-bit" operands are IEEE fp16 (VTQ_F16) or bfloat16 (VTQ_BF16), chosen per call; all
 *     accumulation, the residual stream, LayerNorm, softmax statistics, DiffNet and the head are fp32;
 *   - row-major everywhere; a "token row" is one 768-wide hidden vector.  The ref and dist streams
 *     of a batch of B pairs are stacked as 2B sequences: sequence index = img * B + b (img 0 = ref).
 */
#ifndef VTAMIQ_B200_H_
#define VTAMIQ_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTQ_ABI_VERSION 5

enum vtq_status {
  VTQ_OK = 0,
  VTQ_ERR_INVALID = -1, /* bad argument / unsupported shape */
  VTQ_ERR_CUDA = -2,    /* CUDA runtime or driver error (sticky errors included) */
  VTQ_ERR_NO_DEVICE = -3 /* no sm_100 device: there is NO CPU fallback */
};

enum vtq_dtype { VTQ_F16 = 0, VTQ_BF16 = 1 };

/* GEMM epilogues (vtq_gemm) */
enum vtq_epilogue {
  VTQ_EPI_BIAS_H = 0,      /* out16 = acc + bias                       (QKV projection)            */
  VTQ_EPI_BIAS_GELU_H = 1, /* out16 = gelu_erf(acc + bias)             (fc1)                        */
  VTQ_EPI_BIAS_F32 = 2,    /* out32 = acc + bias                       (patch embedding)            */
  VTQ_EPI_BIAS_RESID_F32 = 3 /* out32 += gamma * (acc + bias), gamma optional (attn.out, fc2)       */
};

typedef struct vtq_ctx vtq_ctx;

/* ---- handle -------------------------------------------------------------------------------- */
int vtq_abi_version(void);
/* Fails with VTQ_ERR_NO_DEVICE unless `device` is a compute-capability 10.x GPU. */
int vtq_create(vtq_ctx** out, int device);
int vtq_destroy(vtq_ctx* ctx);
const char* vtq_last_error_string(const vtq_ctx* ctx); /* ctx may be NULL: error of the last failed vtq_create */
/* number of kernels launched through this handle since creation (monotonic) */
unsigned long long vtq_launch_count(const vtq_ctx* ctx);
/* Coordinate-range status of the gather kernels (vtq_patch_gather*): 1 if any launch since the last reset saw a
 * patch origin outside [0, H-16] x [0, W-16] (or NaN) — torch raises IndexError for those in the reference's gather
 * (data/patch_sampling.py:531-545); here the origin is clamped into the image and this flag is raised instead.  The
 * flag lives in pinned, mapped host memory: reading it needs no synchronisation, but it only reflects kernels that
 * have finished.  reset != 0 clears it. */
int vtq_coord_status(vtq_ctx* ctx, int reset);
/* TMA descriptor cache of this handle (descriptors are encoded once per distinct (pointer, shape, box) and re-used) */
int vtq_tensor_map_stats(const vtq_ctx* ctx, unsigned long long* hits, unsigned long long* misses);
/* Traversal direction of the launches that follow on this handle (sticky): 0 = rows / tiles / work items in increasing
 * order, 1 = decreasing.  Results are identical either way.  A chain of kernels that alternates the direction starts
 * every kernel on the data its predecessor wrote LAST, which is what is still in the 126 MB L2 (the activations of one
 * encoder block are 0.65 GB at cfg2): vtq_layernorm, vtq_gemm / vtq_gemm_ln (CTA-pair kernel) and vtq_attention_fwd
 * honour it; the other entry points ignore it. */
int vtq_set_reverse(vtq_ctx* ctx, int reverse);

/* bytes of scratch vtq_diffnet_head needs for a batch of B pairs */
int64_t vtq_workspace_bytes(const vtq_ctx* ctx, int B, int hidden);

/* ---- K1: patch extraction --------------------------------------------------------------------
 * replaces data/patch_sampling.py:529-545 (gather closure), :559-568 (uv), :572-574 (scale ids).
 *
 * images   [n_img][3][H][W] fp32, one pyramid level.  Image i uses coordinate set (i % n_set).
 * samples  [n_set][2][n] float64 — row 0 = y (top), row 1 = x (left) of each patch, un-truncated, exactly
 *          as stratified_grid_sampling returns them (patch_sampling.py:224).
 * Patch p of this level lands in slot (patch_offset + p) of N_total.  Any output pointer may be NULL.
 *   patches_f32 [n_img][N_total][3][16][16]  bit-exact copy of the reference gather
 *   patches_16  [n_img][N_total][768]        same values rounded to `dtype` (A operand of the embed GEMM)
 *   pos         [n_img][N_total][2]  fp32 uv, = clamp((s + 8) / (dim - 8), 0, 1 - 1e-6) evaluated in fp64
 *   scales      [n_img][N_total]     fp32 scale id (the reference casts its int ids to fp32, train.py:254)
 */
int vtq_patch_gather(vtq_ctx* ctx, const float* images, int n_img, int H, int W, const double* samples, int n_set,
                     int n, int patch_offset, int N_total, float* patches_f32, void* patches_16, int dtype,
                     float* pos, float* scales, int scale_id, void* stream);

/* K1 with the image transform fused in (SURVEY §8f "next" #2).  images [n_img][H][W][3] uint8 (as decoded); each
 * pixel goes through the reference's transform_img arithmetic — x/255, then (x-0.5)/0.5, fp32, in that order
 * (data/utils.py:76,:94; data/patch_datasets.py:51-52) — inside the gather, so outputs are bit-identical to gathering
 * from the transformed fp32 tensor.  Same slot / uv / scale-id conventions as vtq_patch_gather (level 0 of a pyramid
 * is gathered straight from the uint8 image; coarser levels come from vtq_avgpool2x2_u8 / vtq_avgpool2x2). */
int vtq_patch_gather_u8(vtq_ctx* ctx, const uint8_t* images, int n_img, int H, int W, const double* samples, int n_set,
                        int n, int patch_offset, int N_total, float* patches_f32, void* patches_16, int dtype,
                        float* pos, float* scales, int scale_id, void* stream);

/* Coordinate sampler on the device (SURVEY 8f "next" #3): the law of the reference's default sampler,
 * PatchSampler(grid_type=GRID_TYPE_PERTURBED_SIMPLE) -> stratified_grid_sampling (data/patch_sampling.py:236-237,
 * :308-327, :362-376), for `batch` images in one launch: n DISTINCT points of the height x width grid
 * (width = ceil(sqrt(n / (h/w))), height = ceil(width * h/w)) chosen uniformly, jittered by U(-2a, 2a) cells, clipped,
 * scaled to [0, h-ho] x [0, w-wo].  key2: two 64-bit words ON THE DEVICE (the caller draws them with its own
 * generator; Philox4x32-10 is keyed by them).  out [batch][2][n] float64, row 0 = y — the layout vtq_patch_gather takes.
 * Parity with numpy's RNG stream is statistical by necessity (tests/golden/sampler_draws.npz). */
int vtq_sample_grid(vtq_ctx* ctx, const void* key2, int batch, int h, int w, int ho, int wo, int n,
                    double perturbed_amount, double* out, void* stream);

/* The same transform for whole images: uint8 [n_img][H][W][3] -> fp32 [n_img][3][H][W]. */
int vtq_normalize_u8(vtq_ctx* ctx, const uint8_t* src, float* dst, int n_img, int H, int W, void* stream);

/* Pyramid level 1 straight from the decoded image: transform (as above) then 2x2 mean, i.e.
 * AvgPool2d(2)(transform(img)) of data/patch_sampling.py:552,:600 without materialising the fp32 level-0 image.
 * src uint8 [n_img][H][W][3] -> dst fp32 [n_img][3][H/2][W/2].  Needs W % 8 == 0. */
int vtq_avgpool2x2_u8(vtq_ctx* ctx, const uint8_t* src, float* dst, int n_img, int H, int W, void* stream);

/* 2x2 mean pyramid level, floor mode, summation order ((a00+a01)+a10)+a11 then *0.25
 * replaces nn.AvgPool2d(2) at data/patch_sampling.py:552,:600.  src [planes][H][W] -> dst [planes][H/2][W/2] */
int vtq_avgpool2x2(vtq_ctx* ctx, const float* src, float* dst, int planes, int H, int W, void* stream);

/* fp32 patches (the reference's model input, [rows][768]) -> 16-bit GEMM operand */
int vtq_cast_rows(vtq_ctx* ctx, const float* src, void* dst16, int64_t n_elems, int dtype, void* stream);

/* ---- K2: embedding assembly -------------------------------------------------------------------
 * replaces modules/VisionTransformer/transformer.py:396-400 (scale index), :417-426 (uv index),
 * :507-524 (tokens), :536-558 (sums + concat).
 *
 * proj      [n_seq*N][768] fp32   patch projection incl. conv bias (output of vtq_gemm, VTQ_EPI_BIAS_F32)
 * pos       [n_seq*N][2]   fp32 uv;     scales [n_seq*N] fp32 or NULL
 * pos_table [grid*grid+1][768]; scale_table [num_scales+1][768] or NULL
 * cls_token [768] or NULL, extra_tokens [n_extra][768] or NULL;  T = (cls?1:0) + n_extra
 * x         [n_seq][T+N][768] fp32 residual stream (written)
 * pos_idx / scale_idx  optional int32 [n_seq*N] dumps of the gathered row indices (parity tests)
 */
int vtq_embed_assemble(vtq_ctx* ctx, const float* proj, const float* pos, const float* scales,
                       const float* pos_table, int grid, const float* scale_table, int num_scales,
                       const float* cls_token, const float* extra_tokens, int n_extra, int n_seq, int N,
                       int hidden, float* x, int32_t* pos_idx, int32_t* scale_idx, void* stream);

/* ---- K4: LayerNorm, fp32 rows -> 16-bit rows ----------------------------------------------------
 * replaces nn.LayerNorm(768, eps=1e-6) at transformer.py:253-254,:276,:281 (encoder_norm: see K7).
 * Row r is read at x + r * x_stride (elements; 0 = hidden) and written densely at out16 + r * hidden. */
int vtq_layernorm(vtq_ctx* ctx, const float* x, int64_t x_stride, const float* weight, const float* bias, float eps,
                  int64_t rows, int hidden, void* out16, int dtype, void* stream);

/* ---- K3/K5: dense projection on tcgen05 ----------------------------------------------------------
 * out = epilogue(A[M][K] * W[N][K]^T + bias[N]).  A, W 16-bit (lda/ldw = K unless lda given), fp32 accumulate
 * in TMEM.  replaces F.conv2d patch projection (transformer.py:532), Linear query/key/value (:154-156, fused
 * as one N=3*hidden GEMM), Linear out (:169) + residual (:279), fc1+GELU (:213), fc2 (:214) + residual (:284).
 * K % 64 == 0, N % 64 == 0; any M >= 1.  ldo in elements of the output type. */
int vtq_gemm(vtq_ctx* ctx, const void* A, int64_t lda, const void* W, const float* bias, int M, int N, int K,
             int dtype, int epilogue, void* out, int64_t ldo, const float* gamma, void* stream);

/* ---- K3/K5 + K4 folded: the encoder's LayerNorms carried by the GEMMs either side of them -------------
 * replaces the same lines as vtq_gemm plus the LayerNorm between them (transformer.py:276,:281) without a
 * separate pass over the fp32 residual stream.  CTA-pair kernel only: M >= 256.  Exactly one side per call:
 *
 * produce (ln_out != NULL; epilogue must be VTQ_EPI_BIAS_RESID_F32):
 *     out32 += gamma * (acc + bias) as a read-add-write of the caller's rows; additionally
 *     raw16_out [M][N]                    the new rows rounded to 16 bit, NOT normalised
 *     ln_out    [vtq_gemm_ln_slots(N)][M][2]  per row partial (sum, sum of squares) over a fixed column subset
 *                                         per slot (deterministic: no atomics); the slots add up to the full row.
 * consume (ln_in != NULL; epilogue VTQ_EPI_BIAS_H or VTQ_EPI_BIAS_GELU_H):
 *     A = raw16 rows, ln_in = ln_in_slots slots as written above (or by vtq_rowstats_cast: 1 slot), and the caller
 *     pre-folds the LayerNorm's affine into the operands:  W'[n][k] = W[n][k] * ln_w[k]   (then rounded to 16 bit),
 *     bias'[n] = bias[n] + sum_k ln_b[k] W[n][k],  ln_colsum[n] = sum_k W'[n][k]  (of the rounded W').  The kernel
 *     computes  out16 = epi( rstd_r * (A W'^T - mean_r * ln_colsum) + bias' ),  mean/rstd over K columns with ln_eps,
 *     which equals epi( LN(x) W^T + bias ). */
int vtq_gemm_ln(vtq_ctx* ctx, const void* A, int64_t lda, const void* W, const float* bias, int M, int N, int K,
                int dtype, int epilogue, void* out, int64_t ldo, const float* gamma, const float* ln_in,
                int ln_in_slots, const float* ln_colsum, float ln_eps, void* raw16_out, float* ln_out,
                void* stream);
int vtq_gemm_ln_slots(int N);

/* fp32 rows -> raw 16-bit copy + (sum, sum of squares) per row in slot 0 of ln_out [1][rows][2]: the entry of a
 * folded-LayerNorm chain (the rows embed_assemble wrote).  hidden 768 or 1024. */
int vtq_rowstats_cast(vtq_ctx* ctx, const float* x, int64_t rows, int hidden, void* raw16_out, float* ln_out,
                      int dtype, void* stream);

/* ---- K6: fused multi-head self-attention -----------------------------------------------------------
 * replaces transformer.py:158-166: softmax(Q K^T / sqrt(64)) V per (sequence, head), never materialising S x S.
 * qkv [n_seq*S][3*heads*64] 16-bit rows = [q | k | v];  out [n_seq*S][heads*64] 16-bit (head-concatenated).
 * q_rows: 0 = all S query rows; otherwise only query rows [0, q_rows) of every sequence are produced (rounded up
 * to the 256-row work granule) — the last encoder block needs the prefix tokens only (transformer.py:634). */
int vtq_attention_fwd(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads, int dtype, int q_rows,
                      void* stream);

/* Diagnostics: same as vtq_attention_fwd, and CTA 0 writes clock64() stamps of its pipeline events into
 * trace[3][512] (device memory, zero it first): row 0 = MMA issue thread (after every Q K^T / P V commit), rows 1,2 =
 * softmax warpgroups A,B (7 events per key tile: wait S, S ready, S in registers, max done, prev P V retired,
 * SFU turn acquired, P published; +1 per work item).  Used by scripts/attn_trace.py to attribute stalls; not part of the hot path. */
int vtq_attention_fwd_trace(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads, int dtype,
                            long long* trace, void* stream);

/* ---- K7: final LayerNorm of the quality token + difference ----------------------------------------
 * replaces encoder_norm on the used token (transformer.py:376,:634) and modules/vtamiq/vtamiq.py:104-111.
 * x_ref, x_dist [B][S][hidden] fp32 (two blocks of the stacked residual stream; the pairwise mode of train.py:286-287
 * scores two distorted blocks against ONE encoded reference block);
 * diff[b] = gamma * (LN(x_ref[b][token]) - LN(x_dist[b][token])); gamma may be NULL. */
int vtq_cls_diff(vtq_ctx* ctx, const float* x_ref, const float* x_dist, int B, int S, int hidden, int token,
                 const float* ln_weight, const float* ln_bias, float eps, const float* gamma, float* diff,
                 void* stream);

/* ---- K8: DiffNet + quality head ------------------------------------------------------------------------
 * replaces modules/RCAN/channel_attention.py:13-86 as instantiated by modules/vtamiq/vtamiq.py:12-23, and the
 * q_predictor of vtamiq.py:71-77,:114-117.  All fp32.
 * params: HOST array of device pointers, in this order:
 *   for g in [0,num_rgs): for r in [0,num_rcabs): prelu_a, W1, b1, Wdown, bdown, Wup, bup ; then Wg, bg
 *   then Wf, bf (final conv; both NULL if num_rgs == 0 i.e. calibrate=False)
 *   then Wh, bh, prelu_h, Wq, bq (head: hidden -> hidden/4 -> 1)
 * diff [B][hidden] (read), q [B] (written), workspace >= vtq_workspace_bytes(B, hidden). */
int vtq_diffnet_head(vtq_ctx* ctx, const float* diff, const void* const* params, int n_params, int num_rgs,
                     int num_rcabs, int hidden, int ca_hidden, int head_hidden, int B, float* q, void* workspace,
                     void* stream);

/* ---- K8 with a backward pass: frozen-encoder fine-tuning (SURVEY §8f "next" #1, first slice) --------------------
 * replaces, for the parameters behind the encoder, what torch.autograd does in train.py:317-322 when the encoder is
 * frozen (modules/vtamiq/vtamiq.py:81-92, modules/VisionTransformer/backbone.py:62-106).  All fp32.
 *
 * vtq_tail_train_fwd: q = head(DiffNet(gamma * d0)) like vtq_diffnet_head, but every activation the backward pass
 *   needs is kept in `saved` (vtq_tail_saved_floats(...) floats).
 *     d0         [B][hidden]  LN(cls_ref) - LN(cls_dist), i.e. vtq_cls_diff with gamma = NULL
 *     gamma      [hidden] or NULL (diff_scale=False)
 *     drop_scale [num_rgs][B] or NULL: DropPath of each ResidualGroup branch (channel_attention.py:26-29) as a per-pair
 *                factor mask/keep_prob; NULL = evaluation mode (identity)
 *     workspace  >= 32 KB (barrier counters of the fused decoder)
 * vtq_tail_bwd: given dq [B] = dLoss/dq, writes dLoss/dparam for every entry of `params` into the matching entry of
 *   `grads` (HOST array of device pointers, same order and shapes as params; an entry may be NULL = not wanted),
 *   dgamma [hidden] (or NULL) and d_d0 [B][hidden] (or NULL; the gradient that would flow on into the encoder).
 *   Gradients are written, not accumulated.  workspace >= vtq_tail_bwd_workspace_bytes(B, hidden).
 *   Reductions run in a fixed order (no atomics): results are bit-reproducible. */
int64_t vtq_tail_saved_floats(int B, int num_rgs, int num_rcabs, int hidden, int ca_hidden, int head_hidden);
int64_t vtq_tail_bwd_workspace_bytes(int B, int hidden);
int vtq_tail_train_fwd(vtq_ctx* ctx, const float* d0, const float* gamma, const void* const* params, int n_params,
                       int num_rgs, int num_rcabs, int hidden, int ca_hidden, int head_hidden, int B,
                       const float* drop_scale, float* saved, float* q, void* workspace, void* stream);
int vtq_tail_bwd(vtq_ctx* ctx, const float* dq, const float* d0, const float* gamma, const void* const* params,
                 void* const* grads, int n_params, int num_rgs, int num_rcabs, int hidden, int ca_hidden,
                 int head_hidden, int B, const float* drop_scale, const float* saved, float* dgamma, float* d_d0,
                 void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VTAMIQ_B200_H_ */
