// HBM-bound row kernels of the path: fp32->16-bit cast, K2 embedding assembly, K4 LayerNorm, K7 quality-token
// LayerNorm + difference.  (K1 — patch gather, pyramid, uint8 transform — lives in gather.cu.)
// All are one-pass, coalesced, 128-bit on the wide side; none has data reuse worth staging in smem.
#include <cstdlib>

#include "common.cuh"
#include "host.h"

namespace vtq {

// ------------------------------------------------------------------------------------------------
// fp32 -> 16-bit, 8 elements per thread (2 x 128-bit loads, 1 x 128-bit store)
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void cast_rows_kernel(const float* __restrict__ src, void* __restrict__ dst, size_t n8) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n8) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
  const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
  uint4 o = make_uint4(pack2<DT>(a.x, a.y), pack2<DT>(a.z, a.w), pack2<DT>(b.x, b.y), pack2<DT>(b.z, b.w));
  reinterpret_cast<uint4*>(dst)[i] = o;
}

// ------------------------------------------------------------------------------------------------
// K2: embedding assembly.  One warp per output token row (hidden/128 float4 per lane).
//   token rows : cls + pos_table[0]  |  extra_tokens[k]
//   patch rows : (proj + pos_table[idx]) + scale_table[sidx]      (same association as the reference)
// Index arithmetic is the reference's fp32 arithmetic: floor(u*g)*g + floor(v*g) + 1; clamp(s,0,ns-1)+1.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_assemble_kernel(
    const float* __restrict__ proj, const float* __restrict__ pos, const float* __restrict__ scales,
    const float* __restrict__ pos_table, int grid_w, const float* __restrict__ scale_table, int num_scales,
    const float* __restrict__ cls_token, const float* __restrict__ extra_tokens, int T, int n_seq, int N, int hidden,
    float* __restrict__ x, int32_t* __restrict__ pos_idx, int32_t* __restrict__ scale_idx) {
  const int S = T + N;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= static_cast<size_t>(n_seq) * S) return;
  const int lane = threadIdx.x & 31;
  const int seq = static_cast<int>(row / S);
  const int tok = static_cast<int>(row % S);
  float4* dst = reinterpret_cast<float4*>(x + row * hidden);
  const int nvec = hidden >> 2;
  if (tok < T) {
    const bool is_cls = (cls_token != nullptr) && tok == 0;
    const float* base = is_cls ? cls_token : extra_tokens + static_cast<size_t>(tok - (cls_token ? 1 : 0)) * hidden;
    for (int v = lane; v < nvec; v += 32) {
      float4 a = __ldg(reinterpret_cast<const float4*>(base) + v);
      if (is_cls && pos_table != nullptr) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(pos_table) + v);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      dst[v] = a;
    }
    return;
  }
  const size_t pr = static_cast<size_t>(seq) * N + (tok - T);
  const float* prow = proj + pr * hidden;
  const float* ptab = nullptr;
  const float* stab = nullptr;
  if (pos_table != nullptr) {
    const float g = static_cast<float>(grid_w);
    const float fu = floorf(__fmul_rn(pos[pr * 2 + 0], g));
    const float fv = floorf(__fmul_rn(pos[pr * 2 + 1], g));
    const float fidx = __fadd_rn(__fadd_rn(__fmul_rn(fu, g), fv), 1.0f);
    // uv outside [0, 1) (or NaN) would index outside the table — the reference raises IndexError there.  The row
    // index is clamped to the table (memory-safe); the dump keeps the UNclamped value so a caller can detect it.
    const long long raw_idx = static_cast<long long>(fidx);
    const long long last = static_cast<long long>(grid_w) * grid_w;
    const long long idx = raw_idx < 0 ? 0 : (raw_idx > last ? last : raw_idx);
    if (lane == 0 && pos_idx != nullptr) pos_idx[pr] = static_cast<int32_t>(raw_idx);
    ptab = pos_table + static_cast<size_t>(idx) * hidden;
  }
  if (scale_table != nullptr) {
    float s = scales[pr];
    s = fminf(fmaxf(s, 0.0f), static_cast<float>(num_scales - 1)) + 1.0f;
    const long long sidx = static_cast<long long>(s);
    if (lane == 0 && scale_idx != nullptr) scale_idx[pr] = static_cast<int32_t>(sidx);
    stab = scale_table + static_cast<size_t>(sidx) * hidden;
  }
  for (int v = lane; v < nvec; v += 32) {
    float4 a = __ldg(reinterpret_cast<const float4*>(prow) + v);
    if (ptab != nullptr) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ptab) + v);
      a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
    }
    if (stab != nullptr) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(stab) + v);
      a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
    }
    dst[v] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// K4: LayerNorm over `hidden` (<= 1024, multiple of 128): one warp per row, the row stays in registers,
// two-pass mean / variance with warp shuffles, 16-bit output (the A operand of the next GEMM).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ld_f4_hint(const float4* p, uint64_t hint) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(hint));
  return v;
}

// One warp per row, rows in registers, two-pass mean / variance (the reference's arithmetic order per row is kept:
// results are bit-identical whatever the grid).  A warp walks rows with a grid-wide stride and has the NEXT row's
// loads in flight while it reduces and writes the current one: the one-row-per-warp version ran at 0.78 of the copy
// bandwidth (every warp sat out a full DRAM latency between its load and its first add, `long_scoreboard` on the first
// FADD in ncu), and a LayerNorm launch is 9 % of the cfg2 step.
template <int DT, int VPL /* float4 per lane */>
__global__ void __launch_bounds__(256, 3) layernorm_kernel(const float* __restrict__ x, size_t x_stride,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           float eps, size_t rows, void* __restrict__ out16,
                                                           uint64_t hint_x, int reverse) {
  constexpr int HIDDEN = VPL * 128;
  pdl_launch_dependents();
  pdl_wait();
  const size_t stride = static_cast<size_t>(gridDim.x) * (blockDim.x >> 5);
  size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  // reverse: walk the rows from the END.  The GEMM in front of a LayerNorm wrote x in increasing row order, so the
  // last rows are the ones still in L2; and the rows this kernel writes last (the first ones) are what the GEMM
  // behind it reads first.
  const size_t flip = reverse ? rows - 1 : 0;
#define LN_ROW(r) (reverse ? flip - (r) : (r))
  const int lane = threadIdx.x & 31;
  float4 v[VPL], nx[VPL];
  {
    const float4* src = reinterpret_cast<const float4*>(x + LN_ROW(row) * x_stride);
#pragma unroll
    for (int k = 0; k < VPL; ++k) v[k] = ld_f4_hint(src + lane + 32 * k, hint_x);  // keep x resident in L2
  }
  while (true) {
    const size_t next = row + stride;
    const bool has_next = next < rows;
    if (has_next) {
      const float4* src = reinterpret_cast<const float4*>(x + LN_ROW(next) * x_stride);
#pragma unroll
      for (int k = 0; k < VPL; ++k) nx[k] = ld_f4_hint(src + lane + 32 * k, hint_x);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    const float mean = warp_sum(s) * (1.0f / HIDDEN);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const float dx = v[k].x - mean, dy = v[k].y - mean, dz = v[k].z - mean, dw = v[k].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / HIDDEN) + eps);
    uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out16) + LN_ROW(row) * HIDDEN);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * k);
      const float4 be = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * k);
      const float y0 = (v[k].x - mean) * rstd * g.x + be.x;
      const float y1 = (v[k].y - mean) * rstd * g.y + be.y;
      const float y2 = (v[k].z - mean) * rstd * g.z + be.z;
      const float y3 = (v[k].w - mean) * rstd * g.w + be.w;
      __stcg(dst + lane + 32 * k, make_uint2(pack2<DT>(y0, y1), pack2<DT>(y2, y3)));  // L2 only: next GEMM's TMA reads it
    }
    if (!has_next) break;
#pragma unroll
    for (int k = 0; k < VPL; ++k) v[k] = nx[k];
    row = next;
  }
#undef LN_ROW
}

// ------------------------------------------------------------------------------------------------
// Entry of the folded-LayerNorm chain (vtq_gemm_ln): fp32 rows -> RAW 16-bit copy + per-row (sum, sum of squares)
// in statistics slot 0.  One warp per row.
// ------------------------------------------------------------------------------------------------
template <int DT, int VPL>
__global__ void __launch_bounds__(256) rowstats_cast_kernel(const float* __restrict__ x, size_t rows,
                                                            void* __restrict__ raw16, float2* __restrict__ stats) {
  constexpr int HIDDEN = VPL * 128;
  const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* src = reinterpret_cast<const float4*>(x + row * HIDDEN);
  uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(raw16) + row * HIDDEN);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const float4 v = src[lane + 32 * k];
    s1 += (v.x + v.y) + (v.z + v.w);
    s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
    __stcg(dst + lane + 32 * k, make_uint2(pack2<DT>(v.x, v.y), pack2<DT>(v.z, v.w)));
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) stats[row] = make_float2(s1, s2);
}

// ------------------------------------------------------------------------------------------------
// K7: diff[b] = gamma * (LN(x[b][token]) - LN(x[B+b][token])), fp32.  One warp per pair.
// Only the quality-token rows need the encoder_norm; 1/sqrt (not rsqrt.approx) keeps this fp32-faithful.
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(128) cls_diff_kernel(const float* __restrict__ x_ref,
                                                       const float* __restrict__ x_dist, int B, int S, int token,
                                                       const float* __restrict__ w, const float* __restrict__ b,
                                                       float eps, const float* __restrict__ gamma,
                                                       float* __restrict__ diff) {
  constexpr int HIDDEN = VPL * 128;
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= B) return;
  const int lane = threadIdx.x & 31;
  float4 y[2][VPL];
#pragma unroll
  for (int img = 0; img < 2; ++img) {
    const size_t row = static_cast<size_t>(pair) * S + token;
    const float4* src = reinterpret_cast<const float4*>((img == 0 ? x_ref : x_dist) + row * HIDDEN);
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      v[k] = src[lane + 32 * k];
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mean = warp_sum(s) / HIDDEN;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const float dx = v[k].x - mean, dy = v[k].y - mean, dz = v[k].z - mean, dw = v[k].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / HIDDEN + eps);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * k);
      const float4 be = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * k);
      y[img][k].x = (v[k].x - mean) * rstd * g.x + be.x;
      y[img][k].y = (v[k].y - mean) * rstd * g.y + be.y;
      y[img][k].z = (v[k].z - mean) * rstd * g.z + be.z;
      y[img][k].w = (v[k].w - mean) * rstd * g.w + be.w;
    }
  }
  float4* dst = reinterpret_cast<float4*>(diff + static_cast<size_t>(pair) * HIDDEN);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    float4 d = make_float4(y[0][k].x - y[1][k].x, y[0][k].y - y[1][k].y, y[0][k].z - y[1][k].z,
                           y[0][k].w - y[1][k].w);
    if (gamma != nullptr) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * k);
      d.x *= g.x; d.y *= g.y; d.z *= g.z; d.w *= g.w;
    }
    dst[lane + 32 * k] = d;
  }
}

}  // namespace vtq

// ================================================================================================
// C-ABI wrappers
// ================================================================================================
using namespace vtq;

extern "C" int vtq_cast_rows(vtq_ctx* ctx, const float* src, void* dst16, int64_t n_elems, int dtype,
                             void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, src && dst16, "null pointer");
  VTQ_CHECK_ARG(ctx, n_elems >= 0 && n_elems % 8 == 0, "element count must be a multiple of 8");
  VTQ_CHECK_ARG(ctx, (reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst16)) % 16 == 0,
                "16-byte alignment");
  VTQ_CHECK_ARG(ctx, dtype == VTQ_F16 || dtype == VTQ_BF16, "dtype");
  if (n_elems == 0) return VTQ_OK;
  const size_t n8 = static_cast<size_t>(n_elems) / 8;
  const unsigned blocks = static_cast<unsigned>((n8 + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == VTQ_F16) cast_rows_kernel<DT_F16><<<blocks, 256, 0, st>>>(src, dst16, n8);
  else cast_rows_kernel<DT_BF16><<<blocks, 256, 0, st>>>(src, dst16, n8);
  VTQ_CHECK_LAUNCH(ctx, "cast_rows launch");
  return VTQ_OK;
}

extern "C" int vtq_embed_assemble(vtq_ctx* ctx, const float* proj, const float* pos, const float* scales,
                                  const float* pos_table, int grid, const float* scale_table, int num_scales,
                                  const float* cls_token, const float* extra_tokens, int n_extra, int n_seq, int N,
                                  int hidden, float* x, int32_t* pos_idx, int32_t* scale_idx, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, proj && x, "null pointer");
  VTQ_CHECK_ARG(ctx, hidden % 4 == 0 && hidden >= 4, "hidden must be a multiple of 4");
  VTQ_CHECK_ARG(ctx, n_seq >= 1 && N >= 1 && n_extra >= 0, "shape");
  VTQ_CHECK_ARG(ctx, pos_table == nullptr || (pos != nullptr && grid >= 1), "pos table needs uv and a grid width");
  VTQ_CHECK_ARG(ctx, scale_table == nullptr || (scales != nullptr && num_scales >= 1),
                "Model uses scale embedding but scales is passed as None.");
  VTQ_CHECK_ARG(ctx, n_extra == 0 || extra_tokens != nullptr, "extra tokens pointer");
  const int T = (cls_token ? 1 : 0) + n_extra;
  const size_t rows = static_cast<size_t>(n_seq) * (T + N);
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  embed_assemble_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      proj, pos, scales, pos_table, grid, scale_table, num_scales, cls_token, extra_tokens, T, n_seq, N, hidden, x,
      pos_idx, scale_idx);
  VTQ_CHECK_LAUNCH(ctx, "embed_assemble launch");
  return VTQ_OK;
}

extern "C" int vtq_layernorm(vtq_ctx* ctx, const float* x, int64_t x_stride, const float* weight, const float* bias,
                             float eps, int64_t rows, int hidden, void* out16, int dtype, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, x && weight && bias && out16, "null pointer");
  VTQ_CHECK_ARG(ctx, hidden == 768 || hidden == 1024, "hidden must be 768 or 1024");
  VTQ_CHECK_ARG(ctx, rows >= 0, "rows");
  VTQ_CHECK_ARG(ctx, dtype == VTQ_F16 || dtype == VTQ_BF16, "dtype");
  if (x_stride <= 0) x_stride = hidden;
  VTQ_CHECK_ARG(ctx, x_stride >= hidden && x_stride % 4 == 0, "x_stride must be >= hidden and a multiple of 4");
  if (rows == 0) return VTQ_OK;
  unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  static const int ln_blocks_per_sm = [] {   // resident blocks per SM that walk the rows (0 = one row per warp, A/B)
    const char* e = std::getenv("VTQ_LN_BLOCKS_PER_SM");
    return e ? std::atoi(e) : 3;
  }();
  if (ln_blocks_per_sm > 0 && blocks > static_cast<unsigned>(ctx->num_sms * ln_blocks_per_sm))
    blocks = static_cast<unsigned>(ctx->num_sms * ln_blocks_per_sm);
  const int reverse = (ctx->reverse_next && x_stride == hidden) ? 1 : 0;   // dense row blocks only (not strided token rows)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t r = static_cast<size_t>(rows);
  const size_t xs = static_cast<size_t>(x_stride);
  const uint64_t hint_x = l2_hints_enabled() ? L2_EVICT_LAST : L2_EVICT_NORMAL;
  cudaError_t le;
  if (hidden == 768) {
    if (dtype == VTQ_F16) le = launch_pdl(layernorm_kernel<DT_F16, 6>, dim3(blocks), dim3(256), 0, st, x, xs, weight, bias, eps, r, out16, hint_x, reverse);
    else le = launch_pdl(layernorm_kernel<DT_BF16, 6>, dim3(blocks), dim3(256), 0, st, x, xs, weight, bias, eps, r, out16, hint_x, reverse);
  } else {
    if (dtype == VTQ_F16) le = launch_pdl(layernorm_kernel<DT_F16, 8>, dim3(blocks), dim3(256), 0, st, x, xs, weight, bias, eps, r, out16, hint_x, reverse);
    else le = launch_pdl(layernorm_kernel<DT_BF16, 8>, dim3(blocks), dim3(256), 0, st, x, xs, weight, bias, eps, r, out16, hint_x, reverse);
  }
  if (le != cudaSuccess) return check_cuda(ctx, le, "layernorm launch");
  VTQ_CHECK_LAUNCH(ctx, "layernorm launch");
  return VTQ_OK;
}

extern "C" int vtq_rowstats_cast(vtq_ctx* ctx, const float* x, int64_t rows, int hidden, void* raw16_out,
                                 float* ln_out, int dtype, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, x && raw16_out && ln_out, "null pointer");
  VTQ_CHECK_ARG(ctx, hidden == 768 || hidden == 1024, "hidden must be 768 or 1024");
  VTQ_CHECK_ARG(ctx, rows >= 1, "rows");
  VTQ_CHECK_ARG(ctx, dtype == VTQ_F16 || dtype == VTQ_BF16, "dtype");
  VTQ_CHECK_ARG(ctx, (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(raw16_out) |
                      reinterpret_cast<uintptr_t>(ln_out)) % 16 == 0, "pointers must be 16-byte aligned");
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t r = static_cast<size_t>(rows);
  float2* stats = reinterpret_cast<float2*>(ln_out);
  if (hidden == 768) {
    if (dtype == VTQ_F16) rowstats_cast_kernel<DT_F16, 6><<<blocks, 256, 0, st>>>(x, r, raw16_out, stats);
    else rowstats_cast_kernel<DT_BF16, 6><<<blocks, 256, 0, st>>>(x, r, raw16_out, stats);
  } else {
    if (dtype == VTQ_F16) rowstats_cast_kernel<DT_F16, 8><<<blocks, 256, 0, st>>>(x, r, raw16_out, stats);
    else rowstats_cast_kernel<DT_BF16, 8><<<blocks, 256, 0, st>>>(x, r, raw16_out, stats);
  }
  VTQ_CHECK_LAUNCH(ctx, "rowstats_cast launch");
  return VTQ_OK;
}

extern "C" int vtq_cls_diff(vtq_ctx* ctx, const float* x_ref, const float* x_dist, int B, int S, int hidden, int token,
                            const float* ln_weight, const float* ln_bias, float eps, const float* gamma,
                            float* diff, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, x_ref && x_dist && ln_weight && ln_bias && diff, "null pointer");
  VTQ_CHECK_ARG(ctx, hidden == 768 || hidden == 1024, "hidden must be 768 or 1024");
  VTQ_CHECK_ARG(ctx, B >= 1 && S >= 1 && token >= 0 && token < S, "shape");
  const unsigned blocks = static_cast<unsigned>((B + 3) / 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (hidden == 768) cls_diff_kernel<6><<<blocks, 128, 0, st>>>(x_ref, x_dist, B, S, token, ln_weight, ln_bias, eps, gamma, diff);
  else cls_diff_kernel<8><<<blocks, 128, 0, st>>>(x_ref, x_dist, B, S, token, ln_weight, ln_bias, eps, gamma, diff);
  VTQ_CHECK_LAUNCH(ctx, "cls_diff launch");
  return VTQ_OK;
}
