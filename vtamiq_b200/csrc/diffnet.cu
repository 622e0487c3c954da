// K8 — DiffNet (RCAN residual groups with channel attention on a length-1 signal) + quality head, fp32.
//
// On a (B, C, 1) signal every 1x1 Conv1d is a matrix-vector product per pair and AdaptiveAvgPool1d(1) is the
// identity, so the whole decoder is a chain of small dense layers with fused pre-activation / gate / skip:
//   RCAB : y = W1 prelu_a(x) + b1 ;  h = relu(Wd y + bd) ;  x' = x + y * sigmoid(Wu h + bu)
//   RG   : g' = g + (Wg RCAB^n(g) + bg)
//   tail : z = Wf RG^m(d) + bf ;  q = wq . prelu(Wh z + bh) + bq
// The chain is strictly sequential (each layer needs the complete previous activation), so the whole decoder runs
// as ONE persistent cooperative kernel with a device-scope barrier between layers (diffnet_fused_kernel).
// Reference: modules/RCAN/channel_attention.py:13-86, modules/vtamiq/vtamiq.py:12-23,:71-77,:114-117.
#include <cstdlib>

#include "common.cuh"
#include "host.h"

namespace vtq {

enum : int { PRE_NONE = 0, PRE_PRELU = 1 };
enum : int { DEPI_NONE = 0, DEPI_RELU = 1, DEPI_ADD = 2, DEPI_GATE = 3, DEPI_PRELU = 4 };

constexpr int DENSE_PAIRS = 32;    // pairs per tile (one per lane in the epilogue)
constexpr int DENSE_THREADS = 256;  // 8 warps
constexpr int MAX_LAYERS = 64;

struct DenseLayer {
  const float* W;          // [out_dim][in_dim]
  const float* bias;       // [out_dim]
  const float* in;         // [B][in_dim]
  float* out;              // [B][out_dim]
  const float* res;        // [B][out_dim] or null
  const float* gate;       // [B][out_dim] or null
  const float* pre_param;  // PReLU slope applied to the input, or null
  const float* epi_param;  // PReLU slope applied to the output, or null
  float* aux;              // training: GATE -> the sigmoid gate, PRELU -> the pre-activation; null at inference
  const float* row_scale;  // training: ADD -> per-pair scale of the branch (DropPath mask / keep_prob); or null
  int in_dim, out_dim, pre, epi;
};
struct LayerList {
  DenseLayer l[MAX_LAYERS];
  int n;
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Whole decoder + head in ONE persistent launch.  grid = (G, T): the G CTAs of a column split every layer's output
// channels and meet at a device-scope barrier between layers (each layer needs the complete previous activation);
// the T columns work on different 32-pair tiles.  Per layer and CTA: stage pre(in[tile]) in smem (L1-bypassing
// cp.async: other CTAs wrote it); warp w takes a pair of channels; lane i owns input columns {128 j + 4 i .. +3}
// (conflict-free 128-bit LDS of weights and activations), accumulating even/odd columns in packed fp32x2 FMAs;
// a 31-shuffle transposing reduction leaves pair p's sum on lane p; fused bias / gate / skip epilogue.
// The CTA's weight rows are copied to smem with cp.async issued BEFORE it waits at the barrier (weights do not
// depend on activations), so the DRAM latency of the weight stream overlaps the barrier.
constexpr int DENSE_KV = 8;       // float4 per lane per weight row: in_dim <= 1024
constexpr int DENSE_WROWS = 16;   // weight rows staged in smem per round (8 warps x 2 channels)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// stage rows [c, c+n) of W (n <= DENSE_WROWS) into ws[n][in_dim] with 16-byte async copies
__device__ __forceinline__ void stage_weights(float* ws, const float* __restrict__ W, int c, int n, int in_dim) {
  const int nvec = in_dim >> 2;
  const float4* src = reinterpret_cast<const float4*>(W + static_cast<size_t>(c) * in_dim);
  for (int idx = threadIdx.x; idx < n * nvec; idx += DENSE_THREADS) cp_async16(reinterpret_cast<float4*>(ws) + idx, src + idx);
}

__global__ void __launch_bounds__(DENSE_THREADS, 1)
    diffnet_fused_kernel(const __grid_constant__ LayerList L, int B, int max_dim, unsigned* __restrict__ counters) {
  extern __shared__ __align__(16) float smem_f[];
  float* xs = smem_f;                              // [DENSE_PAIRS][in_dim]   activations of this tile
  float* ws = smem_f + DENSE_PAIRS * max_dim;      // [DENSE_WROWS][in_dim]   weight rows of this round
  const int G = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (B + DENSE_PAIRS - 1) / DENSE_PAIRS;
  unsigned* counter = counters + blockIdx.y * 32;  // one 128 B line per column
  unsigned target = 0;

  for (int tile = blockIdx.y; tile < n_tiles; tile += gridDim.y) {
    const int b0 = tile * DENSE_PAIRS;
    const int nb = min(DENSE_PAIRS, B - b0);
    for (int li = 0; li < L.n; ++li) {
      const DenseLayer& ly = L.l[li];
      const int in_dim = ly.in_dim, out_dim = ly.out_dim;
      const int nvec = in_dim >> 2;
      const int chunk = (out_dim + G - 1) / G;
      const int c0 = blockIdx.x * chunk;
      const int nch = max(0, min(chunk, out_dim - c0));

      // round 0 of this CTA's weight rows: in flight while we wait at the barrier (weights do not depend on it)
      if (nch > 0) stage_weights(ws, ly.W, c0, min(nch, DENSE_WROWS), in_dim);
      cp_async_commit();
      // wait until every CTA of this column has published the previous layer
      if (threadIdx.x == 0 && target > 0) {
        unsigned spins = 0;
        while (ld_acquire(counter) < target) {
          if (++spins > (1u << 26)) __trap();
        }
      }
      __syncthreads();

      if (nch > 0) {
        // activations: async 16-byte copies straight to smem (L2 -> smem, L1 bypassed: other CTAs wrote them)
        for (int idx = threadIdx.x; idx < nb * nvec; idx += DENSE_THREADS)
          cp_async16(reinterpret_cast<float4*>(xs) + idx,
                     reinterpret_cast<const float4*>(ly.in + static_cast<size_t>(b0) * in_dim) + idx);
        cp_async_commit();
        cp_async_wait_all();
        if (ly.pre == PRE_PRELU) {  // each thread fixes up exactly the elements it copied
          const float a_pre = __ldg(ly.pre_param);
          for (int idx = threadIdx.x; idx < nb * nvec; idx += DENSE_THREADS) {
            float4 t = reinterpret_cast<float4*>(xs)[idx];
            t.x = t.x > 0.f ? t.x : a_pre * t.x;
            t.y = t.y > 0.f ? t.y : a_pre * t.y;
            t.z = t.z > 0.f ? t.z : a_pre * t.z;
            t.w = t.w > 0.f ? t.w : a_pre * t.w;
            reinterpret_cast<float4*>(xs)[idx] = t;
          }
        }
        for (int idx = nb * nvec + threadIdx.x; idx < DENSE_PAIRS * nvec; idx += DENSE_THREADS)
          reinterpret_cast<float4*>(xs)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);  // ragged last tile
        __syncthreads();

        for (int r0 = 0; r0 < nch; r0 += DENSE_WROWS) {
          const int nr = min(DENSE_WROWS, nch - r0);
          if (r0 > 0) {  // later rounds (a CTA owns more than 16 channels only when many tiles share the SMs)
            __syncthreads();
            stage_weights(ws, ly.W, c0 + r0, nr, in_dim);
            cp_async_commit();
            cp_async_wait_all();
            __syncthreads();
          }
          if (warp * 2 < nr) {
            const int ch_a = c0 + r0 + warp * 2;
            const bool has_b = (warp * 2 + 1) < nr;
            const float* wra = ws + (warp * 2) * in_dim;
            const float* wrb = has_b ? wra + in_dim : wra;
            f32x2 acc_a[DENSE_PAIRS], acc_b[DENSE_PAIRS];  // (even-column sum, odd-column sum)
#pragma unroll
            for (int p = 0; p < DENSE_PAIRS; ++p) acc_a[p] = acc_b[p] = 0ull;
#pragma unroll 1
            for (int k = lane * 4; k < in_dim; k += 128) {
              const float4 wa = *reinterpret_cast<const float4*>(wra + k);
              const float4 wb = *reinterpret_cast<const float4*>(wrb + k);
              const f32x2 a01 = f2_pack(wa.x, wa.y), a23 = f2_pack(wa.z, wa.w);
              const f32x2 b01 = f2_pack(wb.x, wb.y), b23 = f2_pack(wb.z, wb.w);
#pragma unroll
              for (int p = 0; p < DENSE_PAIRS; ++p) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + p * in_dim + k);
                const f32x2 x01 = f2_pack(xv.x, xv.y), x23 = f2_pack(xv.z, xv.w);
                acc_a[p] = f2_fma(a01, x01, acc_a[p]);
                acc_a[p] = f2_fma(a23, x23, acc_a[p]);
                acc_b[p] = f2_fma(b01, x01, acc_b[p]);
                acc_b[p] = f2_fma(b23, x23, acc_b[p]);
              }
            }
            float ra[DENSE_PAIRS], rb[DENSE_PAIRS];
#pragma unroll
            for (int p = 0; p < DENSE_PAIRS; ++p) {
              float lo, hi;
              f2_unpack(acc_a[p], lo, hi);
              ra[p] = lo + hi;
              f2_unpack(acc_b[p], lo, hi);
              rb[p] = lo + hi;
            }
            // transposing butterfly: after the 5 steps lane p holds the sum over lanes of r[p]
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool upper = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < off; ++i) {
                const float sa = upper ? ra[i] : ra[i + off];
                const float ka = upper ? ra[i + off] : ra[i];
                ra[i] = ka + __shfl_xor_sync(0xffffffffu, sa, off);
                const float sb = upper ? rb[i] : rb[i + off];
                const float kb = upper ? rb[i + off] : rb[i];
                rb[i] = kb + __shfl_xor_sync(0xffffffffu, sb, off);
              }
            }
            if (lane < nb) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (h == 1 && !has_b) break;
                const int ch = ch_a + h;
                const size_t oi = static_cast<size_t>(b0 + lane) * out_dim + ch;
                float v = (h ? rb[0] : ra[0]) + __ldg(ly.bias + ch);
                if (ly.epi == DEPI_RELU) v = fmaxf(v, 0.f);
                else if (ly.epi == DEPI_ADD) {
                  if (ly.row_scale != nullptr) v *= __ldg(ly.row_scale + b0 + lane);
                  v = __ldcg(ly.res + oi) + v;
                } else if (ly.epi == DEPI_GATE) {
                  const float sg = 1.0f / (1.0f + expf(-v));
                  if (ly.aux != nullptr) ly.aux[oi] = sg;
                  v = __ldcg(ly.res + oi) + __ldcg(ly.gate + oi) * sg;
                } else if (ly.epi == DEPI_PRELU) {
                  if (ly.aux != nullptr) ly.aux[oi] = v;
                  const float a = __ldg(ly.epi_param);
                  v = v > 0.f ? v : a * v;
                }
                ly.out[oi] = v;
              }
            }
          }
        }
      } else {
        cp_async_wait_all();
      }
      // publish this layer: every thread's stores, then one release increment per CTA
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
      }
      target += G;
    }
  }
}

}  // namespace vtq

namespace vtq {

// Offsets (in floats) of the activations the training forward keeps for the backward pass; see tail_train.cu.
TailSaved tail_saved_layout(int B, int num_rgs, int num_rcabs, int hidden, int ca_hidden, int head_hidden) {
  TailSaved t;
  size_t o = 0;
  const size_t bh = static_cast<size_t>(B) * hidden;
  t.d = o; o += bh;
  t.rcab0 = o;
  t.rcab_stride = 3 * bh + static_cast<size_t>(B) * ca_hidden;   // y, sg, xo (hidden wide) + hc (ca_hidden wide)
  t.group_stride = static_cast<size_t>(num_rcabs) * t.rcab_stride + bh;  // + gout
  o += static_cast<size_t>(num_rgs) * t.group_stride;
  t.z = o; o += (num_rgs > 0 ? bh : 0);
  t.u = o; o += static_cast<size_t>(B) * head_hidden;
  t.hh = o; o += static_cast<size_t>(B) * head_hidden;
  t.total = o;
  return t;
}

int check_tail_args(vtq_ctx* ctx, const void* const* params, int n_params, int num_rgs, int num_rcabs,
                           int hidden, int ca_hidden, int head_hidden, int B) {
  VTQ_CHECK_ARG(ctx, B >= 1 && hidden % 4 == 0 && hidden <= 128 * DENSE_KV && head_hidden % 4 == 0 && head_hidden >= 4,
                "shape (hidden must be a multiple of 4, <= 1024)");
  VTQ_CHECK_ARG(ctx, num_rgs == 0 || (ca_hidden % 4 == 0 && ca_hidden >= 4),
                "channel-attention width must be a multiple of 4");
  VTQ_CHECK_ARG(ctx, ca_hidden <= hidden && head_hidden <= hidden, "squeeze widths");
  VTQ_CHECK_ARG(ctx, num_rgs >= 0 && (num_rgs == 0 || num_rcabs >= 1), "each residual group needs >= 1 RCAB");
  const int smem = (DENSE_PAIRS + DENSE_WROWS) * hidden * static_cast<int>(sizeof(float));
  VTQ_CHECK_ARG(ctx, smem <= ctx->smem_optin, "hidden too large for the staging tile");
  const int expect = num_rgs * (num_rcabs * 7 + 2) + 2 + 5;
  VTQ_CHECK_ARG(ctx, n_params == expect, "parameter list length");
  const int n_layers = num_rgs * (num_rcabs * 3 + 1) + (num_rgs > 0 ? 1 : 0) + 2;
  VTQ_CHECK_ARG(ctx, n_layers <= MAX_LAYERS, "too many DiffNet layers for one launch");
  for (int i = 0; i < n_params; ++i) {
    const bool final_conv = (i == num_rgs * (num_rcabs * 7 + 2) || i == num_rgs * (num_rcabs * 7 + 2) + 1);
    VTQ_CHECK_ARG(ctx, params[i] != nullptr || (final_conv && num_rgs == 0), "null parameter");
  }
  return VTQ_OK;
}

// Launch the fused decoder on a prepared layer list (cooperative grid: T columns of G CTAs, all co-resident).
static int launch_layer_list(vtq_ctx* ctx, const LayerList& L, int B, int hidden, unsigned* counters, cudaStream_t st) {
  const int smem = (DENSE_PAIRS + DENSE_WROWS) * hidden * static_cast<int>(sizeof(float));
  const int n_tiles = (B + DENSE_PAIRS - 1) / DENSE_PAIRS;
  int T = n_tiles < 8 ? n_tiles : 8;
  int G = ctx->num_sms / T;
  static const int g_cap = [] {   // CTAs per column.  Measured at B = 32: 24 -> 0.56 ms, 48 -> 0.39, 64 -> 0.36, 96 -> 0.302,
    const char* e = std::getenv("VTQ_DIFFNET_G");   // 128 -> 0.294, 148 -> 0.300: per-CTA weight streaming, not the barrier
    return e ? std::atoi(e) : 128;                  // arrivals, sets the layer time
  }();
  if (G > g_cap) G = g_cap;
  if (G < 1) G = 1;
  if (int rc = ensure_dyn_smem(ctx, diffnet_fused_kernel, ctx->smem_optin < 200 * 1024 ? ctx->smem_optin : 200 * 1024,
                               "diffnet: cudaFuncSetAttribute")) return rc;
  cudaError_t e = cudaMemsetAsync(counters, 0, 32768, st);
  if (e != cudaSuccess) return check_cuda(ctx, e, "diffnet: cudaMemsetAsync");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G, T, 1);
  cfg.blockDim = dim3(DENSE_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, diffnet_fused_kernel, L, B, hidden, counters);
  if (e != cudaSuccess) return check_cuda(ctx, e, "diffnet: cooperative launch");
  VTQ_CHECK_LAUNCH(ctx, "diffnet fused launch");
  return VTQ_OK;
}

static void add_layer(LayerList& L, const float* in, int in_dim, const float* W, const float* bias, int out_dim,
                      int pre, const float* pre_param, int epi, float* out, const float* res, const float* gate,
                      const float* epi_param, float* aux = nullptr, const float* row_scale = nullptr) {
  DenseLayer& d = L.l[L.n++];
  d.W = W; d.bias = bias; d.in = in; d.out = out; d.res = res; d.gate = gate;
  d.pre_param = pre_param; d.epi_param = epi_param; d.aux = aux; d.row_scale = row_scale;
  d.in_dim = in_dim; d.out_dim = out_dim; d.pre = pre; d.epi = epi;
}

// Training forward of the tail: the same fused decoder, but every activation the backward pass needs lands in its own
// slot of `saved` (tail_saved_layout) instead of four rotating buffers; the sigmoid gates and the head's
// pre-activation are kept (aux), and DropPath scales the residual-group branches per pair (drop_scale).
int launch_tail_train_forward(vtq_ctx* ctx, const float* d_scaled_in_saved, const void* const* params, int n_params,
                              int num_rgs, int num_rcabs, int hidden, int ca_hidden, int head_hidden, int B,
                              const float* drop_scale, float* saved, float* q, unsigned* counters, cudaStream_t st) {
  (void)n_params;
  const TailSaved t = tail_saved_layout(B, num_rgs, num_rcabs, hidden, ca_hidden, head_hidden);
  const size_t bh = static_cast<size_t>(B) * hidden;
  auto P = [&](int i) { return static_cast<const float*>(params[i]); };
  LayerList L;
  L.n = 0;
  int pi = 0;
  const float* g_in = d_scaled_in_saved;
  for (int g = 0; g < num_rgs; ++g) {
    float* gbase = saved + t.rcab0 + static_cast<size_t>(g) * t.group_stride;
    const float* x_in = g_in;
    for (int r = 0; r < num_rcabs; ++r) {
      float* y = gbase + static_cast<size_t>(r) * t.rcab_stride;
      float* sg = y + bh;
      float* xo = sg + bh;
      float* hc = xo + bh;
      const float *a = P(pi), *W1 = P(pi + 1), *b1 = P(pi + 2), *Wd = P(pi + 3), *bd = P(pi + 4), *Wu = P(pi + 5),
                  *bu = P(pi + 6);
      pi += 7;
      add_layer(L, x_in, hidden, W1, b1, hidden, PRE_PRELU, a, DEPI_NONE, y, nullptr, nullptr, nullptr);
      add_layer(L, y, hidden, Wd, bd, ca_hidden, PRE_NONE, nullptr, DEPI_RELU, hc, nullptr, nullptr, nullptr);
      add_layer(L, hc, ca_hidden, Wu, bu, hidden, PRE_NONE, nullptr, DEPI_GATE, xo, x_in, y, nullptr, sg);
      x_in = xo;
    }
    float* gout = gbase + static_cast<size_t>(num_rcabs) * t.rcab_stride;
    add_layer(L, x_in, hidden, P(pi), P(pi + 1), hidden, PRE_NONE, nullptr, DEPI_ADD, gout, g_in, nullptr, nullptr,
              nullptr, drop_scale ? drop_scale + static_cast<size_t>(g) * B : nullptr);
    pi += 2;
    g_in = gout;
  }
  const float* z = g_in;
  if (num_rgs > 0) {
    add_layer(L, g_in, hidden, P(pi), P(pi + 1), hidden, PRE_NONE, nullptr, DEPI_NONE, saved + t.z, nullptr, nullptr,
              nullptr);
    z = saved + t.z;
  }
  pi += 2;
  add_layer(L, z, hidden, P(pi), P(pi + 1), head_hidden, PRE_NONE, nullptr, DEPI_PRELU, saved + t.hh, nullptr, nullptr,
            P(pi + 2), saved + t.u);
  add_layer(L, saved + t.hh, head_hidden, P(pi + 3), P(pi + 4), 1, PRE_NONE, nullptr, DEPI_NONE, q, nullptr, nullptr,
            nullptr);
  return launch_layer_list(ctx, L, B, hidden, counters, st);
}

}  // namespace vtq

using namespace vtq;

extern "C" int64_t vtq_workspace_bytes(const vtq_ctx* ctx, int B, int hidden) {
  (void)ctx;
  if (B < 0 || hidden < 0) return 0;
  // 32 KB of barrier counters, then x, y, g (hidden wide) + one hidden-wide scratch for squeeze / head activations
  return 32768 + static_cast<int64_t>(4) * B * hidden * static_cast<int64_t>(sizeof(float));
}

extern "C" int vtq_diffnet_head(vtq_ctx* ctx, const float* diff, const void* const* params, int n_params,
                                int num_rgs, int num_rcabs, int hidden, int ca_hidden, int head_hidden, int B,
                                float* q, void* workspace, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, diff && params && q && workspace, "null pointer");
  if (int rc = check_tail_args(ctx, params, n_params, num_rgs, num_rcabs, hidden, ca_hidden, head_hidden, B)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (diffnet_cluster_eligible(ctx, hidden, ca_hidden, head_hidden)) {  // the fast path: one launch of 16-CTA clusters
    const int rc = launch_diffnet_cluster(ctx, diff, params, num_rgs, num_rcabs, hidden, ca_hidden, head_hidden, B, q, st);
    if (rc <= 0) return rc;  // 1: clusters of that size cannot be scheduled here -> cooperative kernel below
  }
  const size_t plane = static_cast<size_t>(B) * hidden;
  unsigned* counters = static_cast<unsigned*>(workspace);        // first 32 KB: barrier counters
  float* xbuf = reinterpret_cast<float*>(static_cast<char*>(workspace) + 32768);
  float* ybuf = xbuf + plane;
  float* gbuf = ybuf + plane;
  float* hbuf = gbuf + plane;
  auto P = [&](int i) { return static_cast<const float*>(params[i]); };

  LayerList L;
  L.n = 0;
  int pi = 0;
  const float* g_in = diff;  // group input (skip source)
  for (int g = 0; g < num_rgs; ++g) {
    const float* x_in = g_in;
    for (int r = 0; r < num_rcabs; ++r) {
      const float *a = P(pi), *W1 = P(pi + 1), *b1 = P(pi + 2), *Wd = P(pi + 3), *bd = P(pi + 4), *Wu = P(pi + 5),
                  *bu = P(pi + 6);
      pi += 7;
      add_layer(L, x_in, hidden, W1, b1, hidden, PRE_PRELU, a, DEPI_NONE, ybuf, nullptr, nullptr, nullptr);
      add_layer(L, ybuf, hidden, Wd, bd, ca_hidden, PRE_NONE, nullptr, DEPI_RELU, hbuf, nullptr, nullptr, nullptr);
      add_layer(L, hbuf, ca_hidden, Wu, bu, hidden, PRE_NONE, nullptr, DEPI_GATE, xbuf, x_in, ybuf, nullptr);
      x_in = xbuf;
    }
    // NOTE: the group conv reads x_in (= xbuf) and writes gbuf while its skip source g_in may BE gbuf (groups >= 1):
    // each output element reads res[oi] and writes out[oi] at the same index in the same thread, so that is safe.
    add_layer(L, x_in, hidden, P(pi), P(pi + 1), hidden, PRE_NONE, nullptr, DEPI_ADD, gbuf, g_in, nullptr, nullptr);
    pi += 2;
    g_in = gbuf;
  }
  const float* z = g_in;
  if (num_rgs > 0) {
    add_layer(L, g_in, hidden, P(pi), P(pi + 1), hidden, PRE_NONE, nullptr, DEPI_NONE, ybuf, nullptr, nullptr, nullptr);
    z = ybuf;
  }
  pi += 2;
  add_layer(L, z, hidden, P(pi), P(pi + 1), head_hidden, PRE_NONE, nullptr, DEPI_PRELU, hbuf, nullptr, nullptr,
            P(pi + 2));
  add_layer(L, hbuf, head_hidden, P(pi + 3), P(pi + 4), 1, PRE_NONE, nullptr, DEPI_NONE, q, nullptr, nullptr, nullptr);
  return launch_layer_list(ctx, L, B, hidden, counters, st);
}
