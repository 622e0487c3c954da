// K8 — DiffNet (RCAN residual groups with channel attention on a length-1 signal) + quality head, fp32.
//
// On a (B, C, 1) signal every 1x1 Conv1d is a matrix-vector product per pair and AdaptiveAvgPool1d(1) is the
// identity, so the whole decoder is a chain of small dense layers with fused pre-activation / gate / skip:
//   RCAB : y = W1 prelu_a(x) + b1 ;  h = relu(Wd y + bd) ;  x' = x + y * sigmoid(Wu h + bu)
//   RG   : g' = g + (Wg RCAB^n(g) + bg)
//   tail : z = Wf RG^m(d) + bf ;  q = wq . prelu(Wh z + bh) + bq
// One launch per dense layer (the chain is strictly sequential); vtq_diffnet_head issues the whole chain from
// C in one call so the host sees a single operator, and the chain is CUDA-graph capturable.
// Reference: modules/RCAN/channel_attention.py:13-86, modules/vtamiq/vtamiq.py:12-23,:71-77,:114-117.
#include "common.cuh"
#include "host.h"

namespace vtq {

enum : int { PRE_NONE = 0, PRE_PRELU = 1 };
enum : int { DEPI_NONE = 0, DEPI_RELU = 1, DEPI_ADD = 2, DEPI_GATE = 3, DEPI_PRELU = 4 };

constexpr int DENSE_PAIRS = 32;   // pairs per block (one per lane in the epilogue)
constexpr int DENSE_WARPS = 8;    // output channels per block (one per warp)

// out[b][o] = epi( sum_i W[o][i] * pre(in[b][i]) + bias[o] )
// block: stage pre(in[b0:b0+32][:]) in smem; warp w owns channel o = blockIdx.x*8 + w; lanes split the
// reduction (stride-32, conflict-free LDS, coalesced weight reads), 32 running sums (one per pair) per lane.
template <int PRE, int EPI>
__global__ void __launch_bounds__(DENSE_WARPS * 32) dense_kernel(
    const float* __restrict__ in, int in_dim, const float* __restrict__ W, const float* __restrict__ bias,
    int out_dim, int B, const float* __restrict__ pre_param, float* __restrict__ out,
    const float* __restrict__ res, const float* __restrict__ gate, const float* __restrict__ epi_param) {
  extern __shared__ float xs[];  // [DENSE_PAIRS][in_dim]
  const int b0 = blockIdx.y * DENSE_PAIRS;
  const int nb = min(DENSE_PAIRS, B - b0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  float a_pre = 0.f;
  if constexpr (PRE == PRE_PRELU) a_pre = __ldg(pre_param);
  const int nvec = in_dim >> 2;
  for (int idx = threadIdx.x; idx < DENSE_PAIRS * nvec; idx += blockDim.x) {
    const int p = idx / nvec, v = idx % nvec;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < nb) {
      t = __ldg(reinterpret_cast<const float4*>(in + static_cast<size_t>(b0 + p) * in_dim) + v);
      if constexpr (PRE == PRE_PRELU) {
        t.x = t.x > 0.f ? t.x : a_pre * t.x;
        t.y = t.y > 0.f ? t.y : a_pre * t.y;
        t.z = t.z > 0.f ? t.z : a_pre * t.z;
        t.w = t.w > 0.f ? t.w : a_pre * t.w;
      }
    }
    reinterpret_cast<float4*>(xs)[idx] = t;
  }
  __syncthreads();

  const int o = blockIdx.x * DENSE_WARPS + warp;
  if (o >= out_dim) return;
  float acc[DENSE_PAIRS];
#pragma unroll
  for (int p = 0; p < DENSE_PAIRS; ++p) acc[p] = 0.f;
  const float* wrow = W + static_cast<size_t>(o) * in_dim;
  for (int k = lane; k < in_dim; k += 32) {
    const float w = __ldg(wrow + k);
#pragma unroll
    for (int p = 0; p < DENSE_PAIRS; ++p) acc[p] = fmaf(w, xs[p * in_dim + k], acc[p]);
  }
  float mine = 0.f;
#pragma unroll
  for (int p = 0; p < DENSE_PAIRS; ++p) {
    float v = acc[p];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == p) mine = v;
  }
  if (lane < nb) {
    const size_t oi = static_cast<size_t>(b0 + lane) * out_dim + o;
    float v = mine + __ldg(bias + o);
    if constexpr (EPI == DEPI_RELU) v = fmaxf(v, 0.f);
    if constexpr (EPI == DEPI_ADD) v = res[oi] + v;
    if constexpr (EPI == DEPI_GATE) v = res[oi] + gate[oi] * (1.0f / (1.0f + expf(-v)));
    if constexpr (EPI == DEPI_PRELU) {
      const float a = __ldg(epi_param);
      v = v > 0.f ? v : a * v;
    }
    out[oi] = v;
  }
}

template <int PRE, int EPI>
static int dense(vtq_ctx* ctx, const float* in, int in_dim, const float* W, const float* bias, int out_dim, int B,
                 const float* pre_param, float* out, const float* res, const float* gate, const float* epi_param,
                 cudaStream_t st) {
  auto kern = dense_kernel<PRE, EPI>;
  const int smem = DENSE_PAIRS * in_dim * static_cast<int>(sizeof(float));
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return check_cuda(ctx, e, "diffnet: cudaFuncSetAttribute");
    configured = smem;
  }
  dim3 grid((out_dim + DENSE_WARPS - 1) / DENSE_WARPS, (B + DENSE_PAIRS - 1) / DENSE_PAIRS);
  kern<<<grid, DENSE_WARPS * 32, smem, st>>>(in, in_dim, W, bias, out_dim, B, pre_param, out, res, gate, epi_param);
  VTQ_CHECK_LAUNCH(ctx, "diffnet dense launch");
  return VTQ_OK;
}

}  // namespace vtq

using namespace vtq;

extern "C" int64_t vtq_workspace_bytes(const vtq_ctx* ctx, int B, int hidden) {
  (void)ctx;
  if (B < 0 || hidden < 0) return 0;
  // x, y, g (hidden wide) + one hidden-wide scratch for the squeeze / head activations
  return static_cast<int64_t>(4) * B * hidden * static_cast<int64_t>(sizeof(float));
}

extern "C" int vtq_diffnet_head(vtq_ctx* ctx, const float* diff, const void* const* params, int n_params,
                                int num_rgs, int num_rcabs, int hidden, int ca_hidden, int head_hidden, int B,
                                float* q, void* workspace, void* stream) {
  if (!ctx) return VTQ_ERR_INVALID;
  VTQ_CHECK_ARG(ctx, diff && params && q && workspace, "null pointer");
  VTQ_CHECK_ARG(ctx, B >= 1 && hidden % 4 == 0 && head_hidden % 4 == 0 && head_hidden >= 4, "shape");
  VTQ_CHECK_ARG(ctx, num_rgs == 0 || (ca_hidden % 4 == 0 && ca_hidden >= 4),
                "channel-attention width must be a multiple of 4");
  VTQ_CHECK_ARG(ctx, ca_hidden <= hidden && head_hidden <= hidden, "squeeze widths");
  VTQ_CHECK_ARG(ctx, num_rgs >= 0 && (num_rgs == 0 || num_rcabs >= 1), "each residual group needs >= 1 RCAB");
  VTQ_CHECK_ARG(ctx, DENSE_PAIRS * hidden * 4 <= ctx->smem_optin, "hidden too large for the staging tile");
  const int expect = num_rgs * (num_rcabs * 7 + 2) + 2 + 5;
  VTQ_CHECK_ARG(ctx, n_params == expect, "parameter list length");
  for (int i = 0; i < n_params; ++i) {
    const bool final_conv = (i == num_rgs * (num_rcabs * 7 + 2) || i == num_rgs * (num_rcabs * 7 + 2) + 1);
    VTQ_CHECK_ARG(ctx, params[i] != nullptr || (final_conv && num_rgs == 0), "null parameter");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t plane = static_cast<size_t>(B) * hidden;
  float* xbuf = static_cast<float*>(workspace);
  float* ybuf = xbuf + plane;
  float* gbuf = ybuf + plane;
  float* hbuf = gbuf + plane;
  auto P = [&](int i) { return static_cast<const float*>(params[i]); };

  int pi = 0;
  int rc;
  const float* g_in = diff;  // group input (skip source)
  for (int g = 0; g < num_rgs; ++g) {
    const float* x_in = g_in;
    for (int r = 0; r < num_rcabs; ++r) {
      const float *a = P(pi), *W1 = P(pi + 1), *b1 = P(pi + 2), *Wd = P(pi + 3), *bd = P(pi + 4), *Wu = P(pi + 5),
                  *bu = P(pi + 6);
      pi += 7;
      if ((rc = dense<PRE_PRELU, DEPI_NONE>(ctx, x_in, hidden, W1, b1, hidden, B, a, ybuf, nullptr, nullptr,
                                            nullptr, st)))
        return rc;
      if ((rc = dense<PRE_NONE, DEPI_RELU>(ctx, ybuf, hidden, Wd, bd, ca_hidden, B, nullptr, hbuf, nullptr, nullptr,
                                           nullptr, st)))
        return rc;
      if ((rc = dense<PRE_NONE, DEPI_GATE>(ctx, hbuf, ca_hidden, Wu, bu, hidden, B, nullptr, xbuf, x_in, ybuf,
                                           nullptr, st)))
        return rc;
      x_in = xbuf;
    }
    const float *Wg = P(pi), *bg = P(pi + 1);
    pi += 2;
    if ((rc = dense<PRE_NONE, DEPI_ADD>(ctx, x_in, hidden, Wg, bg, hidden, B, nullptr, gbuf, g_in, nullptr, nullptr,
                                        st)))
      return rc;
    g_in = gbuf;
  }
  const float* z = g_in;
  {
    const float *Wf = P(pi), *bf = P(pi + 1);
    pi += 2;
    if (num_rgs > 0) {
      if ((rc = dense<PRE_NONE, DEPI_NONE>(ctx, g_in, hidden, Wf, bf, hidden, B, nullptr, ybuf, nullptr, nullptr,
                                           nullptr, st)))
        return rc;
      z = ybuf;
    }
  }
  const float *Wh = P(pi), *bh = P(pi + 1), *ah = P(pi + 2), *Wq = P(pi + 3), *bq = P(pi + 4);
  if ((rc = dense<PRE_NONE, DEPI_PRELU>(ctx, z, hidden, Wh, bh, head_hidden, B, nullptr, hbuf, nullptr, nullptr, ah,
                                        st)))
    return rc;
  if ((rc = dense<PRE_NONE, DEPI_NONE>(ctx, hbuf, head_hidden, Wq, bq, 1, B, nullptr, q, nullptr, nullptr, nullptr,
                                       st)))
    return rc;
  return VTQ_OK;
}
