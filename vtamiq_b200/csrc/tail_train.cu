// Training support for the tail of the path (SURVEY §8f "next" #1, first slice): DiffNet + quality head with a
// backward pass, so that the reference's frozen-encoder fine-tuning (backbone.py:62-106 set_freeze_state;
// train.py:317-322 loss.backward()) runs on this library: the encoder stays forward-only on the inference kernels,
// the parameters behind it (diff_scale.gamma, quality_decoder.*, q_predictor.*) receive gradients.
//
// Per pair (vectors of `hidden` channels unless noted; channel_attention.py:13-86, vtamiq.py:104-117):
//   d   = gamma (.) d0                                   d0 = LN(cls_ref) - LN(cls_dist)  (vtq_cls_diff, gamma = NULL)
//   RCAB:  a = prelu_alpha(x) ; y = W1 a + b1 ; h = relu(Wd y + bd) ; s = sigmoid(Wu h + bu) ; x' = x + y (.) s
//   RG:    g' = g + m (.) (Wg RCAB^n(g) + bg)            m = DropPath mask / keep_prob per pair (1 when not training)
//   tail:  z = Wf RG^k(d) + bf ; u = Wh z + bh ; hh = prelu(u) ; q = wq . hh + bq
// Forward = the fused cooperative decoder of diffnet.cu writing every activation into its own slot (TailSaved).
// Backward = the chain rule walked layer by layer with three fp32 building blocks: a tiled SIMT GEMM
// (dX = dY W and dW = dY^T X), a handful of element-wise kernels, and fixed-order column / block reductions
// (bias, PReLU-slope and gamma gradients) — no atomics, so gradients are bit-reproducible run to run.
// Everything is fp32: this is 30 MFLOP per pair next to a 190 GFLOP encoder, launch-latency bound.
#include "common.cuh"
#include "host.h"

namespace vtq {

// ------------------------------------------------------------------------------------------------
// C[M][N] = A * B with A(m,k) = TRANS_A ? A[k*lda + m] : A[m*lda + k],  B(k,n) = B[k*ldb + n]  (fp32, guarded edges)
//   b_prelu != null: B elements pass through PReLU with that slope on load (the RCAB's pre-activation)
//   mode EP_STORE: C = acc ; EP_ACCUM: C += acc ; EP_RELU_MASK: C = mask[m][n] > 0 ? acc : 0
// ------------------------------------------------------------------------------------------------
enum : int { EP_STORE = 0, EP_ACCUM = 1, EP_RELU_MASK = 2 };
constexpr int SG_BM = 32, SG_BN = 64, SG_BK = 16, SG_THREADS = 128;

template <bool TRANS_A>
__global__ void __launch_bounds__(SG_THREADS) sgemm_kernel(const float* __restrict__ A, int lda,
                                                           const float* __restrict__ Bm, int ldb,
                                                           float* __restrict__ Cm, int ldc, int M, int N, int K,
                                                           const float* __restrict__ b_prelu, int mode,
                                                           const float* __restrict__ mask) {
  __shared__ float As[SG_BK][SG_BM + 1];
  __shared__ float Bs[SG_BK][SG_BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int tm = (tid >> 4) * 4, tn = (tid & 15) * 4;
  const float slope = b_prelu != nullptr ? __ldg(b_prelu) : 1.0f;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    // A tile: 32 x 16 = 512 elements, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * SG_THREADS;
      int m, k;
      if (TRANS_A) { m = idx & (SG_BM - 1); k = idx >> 5; }   // contiguous along m
      else { k = idx & (SG_BK - 1); m = idx >> 4; }           // contiguous along k
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < K) v = TRANS_A ? A[static_cast<size_t>(gk) * lda + gm] : A[static_cast<size_t>(gm) * lda + gk];
      As[k][m] = v;
    }
    // B tile: 16 x 64 = 1024 elements, 8 per thread, contiguous along n
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int idx = tid + e * SG_THREADS;
      const int n = idx & (SG_BN - 1), k = idx >> 6;
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < K) {
        v = Bm[static_cast<size_t>(gk) * ldb + gn];
        if (b_prelu != nullptr) v = v > 0.f ? v : slope * v;
      }
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][tm + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tn + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + tm + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tn + j;
      if (gn >= N) continue;
      const size_t o = static_cast<size_t>(gm) * ldc + gn;
      float v = acc[i][j];
      if (mode == EP_ACCUM) v += Cm[o];
      else if (mode == EP_RELU_MASK) v = mask[o] > 0.f ? v : 0.f;
      Cm[o] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// element-wise pieces (grid-stride)
// ------------------------------------------------------------------------------------------------
__global__ void ew_scale_cols_kernel(float* __restrict__ out, const float* __restrict__ in,
                                     const float* __restrict__ gamma, size_t n, int H) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = gamma != nullptr ? in[i] * __ldg(gamma + (i % H)) : in[i];
}
__global__ void ew_scale_rows_kernel(float* __restrict__ out, const float* __restrict__ in,
                                     const float* __restrict__ row_scale, size_t n, int H) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = row_scale != nullptr ? in[i] * __ldg(row_scale + (i / H)) : in[i];
}
__global__ void ew_add_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
                              size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = a[i] + b[i];
}
// x' = x + y (.) s with s = sigmoid(pre):  dy = dx' (.) s ;  dpre = dx' (.) y (.) s (1 - s)
__global__ void ew_gate_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ sg,
                                   const float* __restrict__ y, float* __restrict__ dy, float* __restrict__ dpre,
                                   size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float g = dx[i], s = sg[i];
    dy[i] = g * s;
    dpre[i] = g * y[i] * s * (1.0f - s);
  }
}
// dhh[b][j] = dq[b] * wq[j]
__global__ void ew_outer_kernel(float* __restrict__ out, const float* __restrict__ dq, const float* __restrict__ wq,
                                size_t n, int J) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = dq[i / J] * __ldg(wq + (i % J));
}

// PReLU backward (torch semantics: x > 0 ? g : slope * g ; slope gradient += x > 0 ? 0 : x * g).
//   out = (base ? base : 0) + dact (.) prelu'(x);  partials[block] = sum over the block's elements of the slope term
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 512;
__global__ void __launch_bounds__(RED_THREADS) ew_prelu_bwd_kernel(const float* __restrict__ dact,
                                                                   const float* __restrict__ x,
                                                                   const float* __restrict__ slope_p,
                                                                   const float* __restrict__ base,
                                                                   float* __restrict__ out,
                                                                   float* __restrict__ partials, size_t n) {
  __shared__ float red[RED_THREADS];
  const float slope = __ldg(slope_p);
  float part = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float g = dact[i], xv = x[i];
    const float gi = xv > 0.f ? g : slope * g;
    out[i] = base != nullptr ? base[i] + gi : gi;
    part += xv > 0.f ? 0.f : xv * g;
  }
  red[threadIdx.x] = part;
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = red[0];
}
// out[0] = sum of `count` partials, fixed order (one block)
__global__ void __launch_bounds__(RED_THREADS) reduce_partials_kernel(float* __restrict__ out,
                                                                      const float* __restrict__ partials, int count) {
  __shared__ float red[RED_THREADS];
  float part = 0.f;
  for (int i = threadIdx.x; i < count; i += RED_THREADS) part += partials[i];
  red[threadIdx.x] = part;
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}
// out[c] = sum_b in[b][c] (* in2[b][c]);  one thread per column, rows in order (deterministic)
__global__ void colsum_kernel(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ in2,
                              int B, int N) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const size_t i = static_cast<size_t>(b) * N + c;
    s += in2 != nullptr ? in[i] * in2[i] : in[i];
  }
  out[c] = s;
}

// ------------------------------------------------------------------------------------------------
// host-side launch helpers
// ------------------------------------------------------------------------------------------------
struct TailLauncher {
  vtq_ctx* ctx;
  cudaStream_t st;
  int rc = VTQ_OK;
  float* partials;  // RED_MAX_BLOCKS floats

  static unsigned ew_blocks(size_t n) {
    const size_t b = (n + 255) / 256;
    return static_cast<unsigned>(b < 1 ? 1 : (b > 2048 ? 2048 : b));
  }
  void done(const char* what) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && rc == VTQ_OK) rc = check_cuda(ctx, e, what);
  }
  // dX[M=B][N=in] = dY[B][out] * W[out][in]
  void bwd_input(const float* dY, const float* W, float* dX, int B, int out_dim, int in_dim, int mode = EP_STORE,
                 const float* mask = nullptr) {
    dim3 grid((in_dim + SG_BN - 1) / SG_BN, (B + SG_BM - 1) / SG_BM);
    sgemm_kernel<false><<<grid, SG_THREADS, 0, st>>>(dY, out_dim, W, in_dim, dX, in_dim, B, in_dim, out_dim, nullptr,
                                                     mode, mask);
    done("tail bwd_input");
  }
  // dW[out][in] = dY[B][out]^T * X[B][in] ; db[out] = sum_b dY
  void bwd_weight(const float* dY, const float* X, float* dW, float* db, int B, int out_dim, int in_dim,
                  const float* x_prelu = nullptr) {
    if (dW != nullptr) {
      dim3 grid((in_dim + SG_BN - 1) / SG_BN, (out_dim + SG_BM - 1) / SG_BM);
      sgemm_kernel<true><<<grid, SG_THREADS, 0, st>>>(dY, out_dim, X, in_dim, dW, in_dim, out_dim, in_dim, B, x_prelu,
                                                      EP_STORE, nullptr);
      done("tail bwd_weight");
    }
    if (db != nullptr) colsum(db, dY, nullptr, B, out_dim);
  }
  void colsum(float* out, const float* in, const float* in2, int B, int N) {
    colsum_kernel<<<(N + 127) / 128, 128, 0, st>>>(out, in, in2, B, N);
    done("tail colsum");
  }
  void prelu_bwd(const float* dact, const float* x, const float* slope, const float* base, float* out, float* dslope,
                 size_t n) {
    unsigned blocks = ew_blocks(n);
    if (blocks > RED_MAX_BLOCKS) blocks = RED_MAX_BLOCKS;
    ew_prelu_bwd_kernel<<<blocks, RED_THREADS, 0, st>>>(dact, x, slope, base, out, partials, n);
    done("tail prelu_bwd");
    if (dslope != nullptr) {
      reduce_partials_kernel<<<1, RED_THREADS, 0, st>>>(dslope, partials, static_cast<int>(blocks));
      done("tail reduce_partials");
    }
  }
};

}  // namespace vtq

using namespace vtq;

extern "C" int64_t vtq_tail_saved_floats(int B, int num_rgs, int num_rcabs, int hidden, int ca_hidden,
                                         int head_hidden) {
  if (B < 0 || num_rgs < 0 || num_rcabs < 0 || hidden < 0 || ca_hidden < 0 || head_hidden < 0) return 0;
  return static_cast<int64_t>(tail_saved_layout(B, num_rgs, num_rcabs, hidden, ca_hidden, head_hidden).total);
}

extern "C" int64_t vtq_tail_bwd_workspace_bytes(int B, int hidden) {
  if (B < 0 || hidden < 0) return 0;
  // dg, dx, dy, dps, da, dt (hidden wide) + dh, du, dhh (<= hidden wide) + block partials
  return (static_cast<int64_t>(9) * B * hidden + RED_MAX_BLOCKS) * static_cast<int64_t>(sizeof(float));
}

extern "C" int vtq_tail_train_fwd(vtq_ctx* ctx, const float* d0, const float* gamma, const void* const* params,
                                  int n_params, int num_rgs, int num_rcabs, int hidden, int ca_hidden,
                                  int head_hidden, int B, const float* drop_scale, float* saved, float* q,
                                  void* workspace, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, d0 && params && saved && q && workspace, "null pointer");
  if (int rc = check_tail_args(ctx, params, n_params, num_rgs, num_rcabs, hidden, ca_hidden, head_hidden, B)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const TailSaved t = tail_saved_layout(B, num_rgs, num_rcabs, hidden, ca_hidden, head_hidden);
  const size_t bh = static_cast<size_t>(B) * hidden;
  ew_scale_cols_kernel<<<TailLauncher::ew_blocks(bh), 256, 0, st>>>(saved + t.d, d0, gamma, bh, hidden);
  VTQ_CHECK_LAUNCH(ctx, "tail scale launch");
  return launch_tail_train_forward(ctx, saved + t.d, params, n_params, num_rgs, num_rcabs, hidden, ca_hidden,
                                   head_hidden, B, drop_scale, saved, q, static_cast<unsigned*>(workspace), st);
}

extern "C" int vtq_tail_bwd(vtq_ctx* ctx, const float* dq, const float* d0, const float* gamma,
                            const void* const* params, void* const* grads, int n_params, int num_rgs, int num_rcabs,
                            int hidden, int ca_hidden, int head_hidden, int B, const float* drop_scale,
                            const float* saved, float* dgamma, float* d_d0, void* workspace, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, dq && d0 && params && grads && saved && workspace, "null pointer");
  if (int rc = check_tail_args(ctx, params, n_params, num_rgs, num_rcabs, hidden, ca_hidden, head_hidden, B)) return rc;
  const TailSaved t = tail_saved_layout(B, num_rgs, num_rcabs, hidden, ca_hidden, head_hidden);
  const int H = hidden, CA = ca_hidden, HH = head_hidden;
  const size_t bh = static_cast<size_t>(B) * H;
  float* w = static_cast<float*>(workspace);
  float *dg = w, *dx = w + bh, *dy = w + 2 * bh, *dps = w + 3 * bh, *da = w + 4 * bh, *dt = w + 5 * bh,
        *dh = w + 6 * bh, *du = w + 7 * bh, *dhh = w + 8 * bh;
  TailLauncher L{ctx, static_cast<cudaStream_t>(stream), VTQ_OK, w + 9 * bh};
  cudaStream_t st = L.st;
  auto P = [&](int i) { return static_cast<const float*>(params[i]); };
  auto G = [&](int i) { return static_cast<float*>(grads[i]); };  // may be null: that gradient is not wanted
  const int tail0 = num_rgs * (num_rcabs * 7 + 2);  // index of Wf
  const int head0 = tail0 + 2;                      // Wh, bh, prelu_h, Wq, bq

  // ---- head: q = wq . prelu(u) + bq,  u = Wh z + bh
  const float* z = num_rgs > 0 ? saved + t.z
                               : saved + t.d;  // calibrate=False: the head reads the scaled difference directly
  const size_t bj = static_cast<size_t>(B) * HH;
  ew_outer_kernel<<<TailLauncher::ew_blocks(bj), 256, 0, st>>>(dhh, dq, P(head0 + 3), bj, HH);
  L.done("tail outer");
  if (G(head0 + 3)) L.bwd_weight(dq, saved + t.hh, G(head0 + 3), nullptr, B, 1, HH);   // dwq [1][HH]
  if (G(head0 + 4)) L.colsum(G(head0 + 4), dq, nullptr, B, 1);                         // dbq
  L.prelu_bwd(dhh, saved + t.u, P(head0 + 2), nullptr, du, G(head0 + 2), bj);
  L.bwd_weight(du, z, G(head0), G(head0 + 1), B, HH, H);
  L.bwd_input(du, P(head0), dg, B, HH, H);             // dg = dL/dz
  // ---- final conv z = Wf g + bf
  if (num_rgs > 0) {
    const float* g_last = saved + t.rcab0 + static_cast<size_t>(num_rgs - 1) * t.group_stride +
                          static_cast<size_t>(num_rcabs) * t.rcab_stride;
    L.bwd_weight(dg, g_last, G(tail0), G(tail0 + 1), B, H, H);
    L.bwd_input(dg, P(tail0), dx, B, H, H);
    cudaMemcpyAsync(dg, dx, bh * sizeof(float), cudaMemcpyDeviceToDevice, st);   // dg = dL/dg_last
  }
  // ---- residual groups, last to first
  for (int g = num_rgs - 1; g >= 0; --g) {
    const float* gbase = saved + t.rcab0 + static_cast<size_t>(g) * t.group_stride;
    const float* g_in = g == 0 ? saved + t.d
                               : saved + t.rcab0 + static_cast<size_t>(g - 1) * t.group_stride +
                                     static_cast<size_t>(num_rcabs) * t.rcab_stride;
    const int pg = g * (num_rcabs * 7 + 2);
    // g' = g + m (.) (Wg x_R + bg)
    ew_scale_rows_kernel<<<TailLauncher::ew_blocks(bh), 256, 0, st>>>(
        dt, dg, drop_scale ? drop_scale + static_cast<size_t>(g) * B : nullptr, bh, H);
    L.done("tail rowscale");
    const float* x_R = gbase + static_cast<size_t>(num_rcabs - 1) * t.rcab_stride + 2 * bh;  // xo of the last RCAB
    L.bwd_weight(dt, x_R, G(pg + num_rcabs * 7), G(pg + num_rcabs * 7 + 1), B, H, H);
    L.bwd_input(dt, P(pg + num_rcabs * 7), dx, B, H, H);   // dx = dL/dx_R
    for (int r = num_rcabs - 1; r >= 0; --r) {
      const float* y = gbase + static_cast<size_t>(r) * t.rcab_stride;
      const float* sg = y + bh;
      const float* hc = y + 3 * bh;
      const float* x_in = r == 0 ? g_in : gbase + static_cast<size_t>(r - 1) * t.rcab_stride + 2 * bh;
      const int pr = pg + r * 7;  // prelu_a, W1, b1, Wd, bd, Wu, bu
      ew_gate_bwd_kernel<<<TailLauncher::ew_blocks(bh), 256, 0, st>>>(dx, sg, y, dy, dps, bh);
      L.done("tail gate_bwd");
      L.bwd_weight(dps, hc, G(pr + 5), G(pr + 6), B, H, CA);              // Wu [H][CA]
      L.bwd_input(dps, P(pr + 5), dh, B, H, CA, EP_RELU_MASK, hc);         // d(pre-relu h) = (dps Wu) (.) [h > 0]
      L.bwd_weight(dh, y, G(pr + 3), G(pr + 4), B, CA, H);                // Wd [CA][H]
      L.bwd_input(dh, P(pr + 3), dy, B, CA, H, EP_ACCUM);                  // dy += dh Wd
      L.bwd_weight(dy, x_in, G(pr + 1), G(pr + 2), B, H, H, P(pr));        // W1: X = prelu(x_in)
      L.bwd_input(dy, P(pr + 1), da, B, H, H);                             // da = dy W1
      L.prelu_bwd(da, x_in, P(pr), dx, dx, G(pr), bh);                     // dx = dx + da (.) prelu'(x_in) ; d alpha
    }
    ew_add_kernel<<<TailLauncher::ew_blocks(bh), 256, 0, st>>>(dg, dg, dx, bh);   // skip + branch
    L.done("tail add");
  }
  // ---- d = gamma (.) d0
  if (dgamma != nullptr) L.colsum(dgamma, dg, d0, B, H);
  if (d_d0 != nullptr) {
    ew_scale_cols_kernel<<<TailLauncher::ew_blocks(bh), 256, 0, st>>>(d_d0, dg, gamma, bh, H);
    L.done("tail d_d0");
  }
  return L.rc;
}
