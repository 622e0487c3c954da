// K1 — patch extraction (HBM-bound): patch gather (+uv, scale ids) from fp32 CHW or uint8 HWC images, the reference's
// image transform (x/255, (x-.5)/.5) fused in, and the chained 2x2-mean pyramid.
// Reference semantics: data/patch_sampling.py:529-545 (gather closure), :559-568 (uv), :572-574 (scale ids),
// :552,:600 (nn.AvgPool2d(2) pyramid); data/utils.py:76,:94 + data/patch_datasets.py:51-52 (transform).
//
// Traffic model.  A 16-pixel patch row is 64 B (fp32) / 48 B (uint8 HWC) at an arbitrary 4 B / 1 B alignment, so it
// straddles 2-3 of DRAM's 32-byte sectors: sector granularity alone makes the read traffic ~1.44x (fp32) / ~1.7x (u8)
// the algorithmic bytes.  The vector kernels below request exactly those sectors, once: one thread owns one patch row,
// pulls the covering 32-byte-aligned span with 256-bit loads (the third only when the row really reaches into it),
// rotates it into place in registers and writes its 16 outputs with one (16-bit) or two (fp32) 256-bit stores —
// consecutive threads write consecutive 32-byte chunks of the patch matrix.  The scalar kernels (one 4-byte / 1-byte
// load per pixel) remain as the generic path for buffers that are not 32-byte aligned / padded.
//
// Coordinates are trusted by the reference (an out-of-range index raises IndexError in torch).  Here they are clamped
// into the image (memory-safe) and a flag word in pinned host memory is set; the host mirror raises IndexError.
#include "common.cuh"
#include "host.h"

namespace vtq {

constexpr int PATCH = 16;
constexpr int PATCH_ELEMS = 3 * PATCH * PATCH;  // 768

// ------------------------------------------------------------------------------------------------
// shared pieces
// ------------------------------------------------------------------------------------------------
struct PatchOrigin {
  int y0, x0;
  double sy, sx;
};

// torch advanced indexing with float64 indices truncates toward zero (patch_sampling.py:531-545 casts the sampled
// coordinates with .astype(int)); valid origins are [0, H-16] x [0, W-16].
__device__ __forceinline__ PatchOrigin load_origin(const double* __restrict__ samples, int set, int n, int p, int H,
                                                   int W, int* __restrict__ oob_flag) {
  PatchOrigin o;
  o.sy = __ldg(samples + (static_cast<size_t>(set) * 2 + 0) * n + p);
  o.sx = __ldg(samples + (static_cast<size_t>(set) * 2 + 1) * n + p);
  const bool ok = o.sy >= 0.0 && o.sx >= 0.0 && o.sy < static_cast<double>(H - PATCH + 1) &&
                  o.sx < static_cast<double>(W - PATCH + 1);  // false for NaN
  if (!ok && oob_flag != nullptr) *reinterpret_cast<volatile int*>(oob_flag) = 1;
  const double cy = ok ? o.sy : fmin(fmax(o.sy == o.sy ? o.sy : 0.0, 0.0), static_cast<double>(H - PATCH));
  const double cx = ok ? o.sx : fmin(fmax(o.sx == o.sx ? o.sx : 0.0, 0.0), static_cast<double>(W - PATCH));
  o.y0 = static_cast<int>(cy);
  o.x0 = static_cast<int>(cx);
  return o;
}

// uv = clamp((sample + P/2) / (dim - P/2), 0, 1 - 1e-6) in float64, rounded once to fp32 (patch_sampling.py:559-568;
// the divisor is an fp32 value in the reference)
__device__ __forceinline__ void store_uv(float* __restrict__ pos, float* __restrict__ scales, size_t slot,
                                         const PatchOrigin& o, int H, int W, float scale_id) {
  if (pos != nullptr) {
    const double hi = 1.0 - 1e-6;
    const double u = (o.sy + 8.0) / static_cast<double>(static_cast<float>(H - 8));
    const double v = (o.sx + 8.0) / static_cast<double>(static_cast<float>(W - 8));
    pos[slot * 2 + 0] = __double2float_rn(fmin(fmax(u, 0.0), hi));
    pos[slot * 2 + 1] = __double2float_rn(fmin(fmax(v, 0.0), hi));
  }
  if (scales != nullptr) scales[slot] = scale_id;
}

__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// The reference transform of one decoded byte: t = u / 255 (to_tensor), then (t - 0.5) / 0.5 (normalize), fp32, each
// step rounded.  IEEE-exact forms: the scalar kernels use the division intrinsics; the vector kernels use
//   q0 = u*r, e = fma(-q0, 255, u), q = fma(e, r, q0)       (r = RN(1/255); == RN(u/255) for every u in 0..255)
//   z  = fma(q, 2, -1)                                       (== RN(RN(q - 0.5) / 0.5): scaling by 2 is exact)
// — proven for all 256 inputs by tests/test_host_logic.py::test_u8_transform_is_exact, and on the GPU by the bit-exact
// gather tests.
__device__ __forceinline__ float normalize_u8(uint8_t u) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(u), 255.0f), 0.5f), 0.5f);
}
__device__ __forceinline__ f32x2 normalize_u8x2(float u0, float u1) {
  const float r = 0.003921568859368563f;  // RN(1/255)
  const f32x2 u = f2_pack(u0, u1);
  const f32x2 q0 = f2_mul(u, f2_pack(r, r));
  const f32x2 e = f2_fma(q0, f2_pack(-255.0f, -255.0f), u);
  const f32x2 q = f2_fma(e, f2_pack(r, r), q0);
  return f2_fma(q, f2_pack(2.0f, 2.0f), f2_pack(-1.0f, -1.0f));
}
// byte k of `w` as an exact float: (0x4B000000 | byte) is 2^23 + byte
template <int K>
__device__ __forceinline__ float byte_to_float(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | K)) - 8388608.0f;
}

// ------------------------------------------------------------------------------------------------
// generic (scalar) gather kernels: block = patch, 192 threads x 4 pixels.  Any alignment.
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(192) patch_gather_kernel(const float* __restrict__ images, int H, int W,
                                                           const double* __restrict__ samples, int n_set, int n,
                                                           int patch_offset, int N_total,
                                                           float* __restrict__ patches_f32,
                                                           void* __restrict__ patches_16, float* __restrict__ pos,
                                                           float* __restrict__ scales, float scale_id,
                                                           int* __restrict__ oob_flag) {
  const int p = blockIdx.x;
  const int img = blockIdx.y;
  const PatchOrigin og = load_origin(samples, img % n_set, n, p, H, W, oob_flag);
  const size_t slot = static_cast<size_t>(img) * N_total + patch_offset + p;
  const int t = threadIdx.x;       // 0..191: (c, i, j4)
  const int c = t >> 6;            // channel
  const int i = (t >> 2) & 15;     // row inside the patch
  const int j4 = (t & 3) * 4;      // first of 4 columns
  const float* src = images + ((static_cast<size_t>(img) * 3 + c) * H + (og.y0 + i)) * W + og.x0 + j4;
  const float v0 = __ldg(src + 0), v1 = __ldg(src + 1), v2 = __ldg(src + 2), v3 = __ldg(src + 3);
  const size_t o = slot * PATCH_ELEMS + static_cast<size_t>(t) * 4;
  if (patches_f32 != nullptr) *reinterpret_cast<float4*>(patches_f32 + o) = make_float4(v0, v1, v2, v3);
  if (patches_16 != nullptr)
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(patches_16) + o) = make_uint2(pack2<DT>(v0, v1), pack2<DT>(v2, v3));
  if (t == 0) store_uv(pos, scales, slot, og, H, W, scale_id);
}

template <int DT>
__global__ void __launch_bounds__(192) patch_gather_u8_kernel(const uint8_t* __restrict__ images, int H, int W,
                                                              const double* __restrict__ samples, int n_set, int n,
                                                              int patch_offset, int N_total,
                                                              float* __restrict__ patches_f32,
                                                              void* __restrict__ patches_16, float* __restrict__ pos,
                                                              float* __restrict__ scales, float scale_id,
                                                              int* __restrict__ oob_flag) {
  const int p = blockIdx.x;
  const int img = blockIdx.y;
  const PatchOrigin og = load_origin(samples, img % n_set, n, p, H, W, oob_flag);
  const size_t slot = static_cast<size_t>(img) * N_total + patch_offset + p;
  const int t = threadIdx.x;    // (c, i, j4)
  const int c = t >> 6;
  const int i = (t >> 2) & 15;
  const int j4 = (t & 3) * 4;
  const uint8_t* src = images + ((static_cast<size_t>(img) * H + (og.y0 + i)) * W + og.x0 + j4) * 3 + c;  // HWC
  const float v0 = normalize_u8(__ldg(src)), v1 = normalize_u8(__ldg(src + 3)), v2 = normalize_u8(__ldg(src + 6)),
              v3 = normalize_u8(__ldg(src + 9));
  const size_t o = slot * PATCH_ELEMS + static_cast<size_t>(t) * 4;
  if (patches_f32 != nullptr) *reinterpret_cast<float4*>(patches_f32 + o) = make_float4(v0, v1, v2, v3);
  if (patches_16 != nullptr)
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(patches_16) + o) = make_uint2(pack2<DT>(v0, v1), pack2<DT>(v2, v3));
  if (t == 0) store_uv(pos, scales, slot, og, H, W, scale_id);
}

// ------------------------------------------------------------------------------------------------
// vector gather, fp32 CHW source.  thread = (patch, channel, row): 48 threads per patch, 4 patches per block.
// Preconditions (checked by the launcher): image base and output bases 32-byte aligned, total image floats % 8 == 0.
// ------------------------------------------------------------------------------------------------
constexpr int GV_PATCHES = 4;  // patches per 192-thread block

template <int DT>
__global__ void __launch_bounds__(192) patch_gather_vec_kernel(const float* __restrict__ images, int H, int W,
                                                               const double* __restrict__ samples, int n_set, int n,
                                                               int patch_offset, int N_total,
                                                               float* __restrict__ patches_f32,
                                                               void* __restrict__ patches_16, float* __restrict__ pos,
                                                               float* __restrict__ scales, float scale_id,
                                                               int* __restrict__ oob_flag) {
  const int r = threadIdx.x % 48;  // (c, i) = row r of the 48 x 16 patch matrix
  const int p = blockIdx.x * GV_PATCHES + threadIdx.x / 48;
  if (p >= n) return;
  const int img = blockIdx.y;
  const PatchOrigin og = load_origin(samples, img % n_set, n, p, H, W, oob_flag);
  const size_t slot = static_cast<size_t>(img) * N_total + patch_offset + p;
  const int c = r >> 4, i = r & 15;
  const float* src = images + ((static_cast<size_t>(img) * 3 + c) * H + (og.y0 + i)) * W + og.x0;
  const uintptr_t a = reinterpret_cast<uintptr_t>(src);
  const int s = static_cast<int>((a >> 2) & 7);  // floats between the 32-byte boundary and the first pixel
  const uint8_t* base = reinterpret_cast<const uint8_t*>(a & ~static_cast<uintptr_t>(31));
  uint32_t f[24];
  ldg256(base, *reinterpret_cast<uint32_t(*)[8]>(&f[0]));
  ldg256(base + 32, *reinterpret_cast<uint32_t(*)[8]>(&f[8]));
  if (s != 0) ldg256(base + 64, *reinterpret_cast<uint32_t(*)[8]>(&f[16]));
  else {
#pragma unroll
    for (int k = 16; k < 24; ++k) f[k] = 0u;
  }
  uint32_t o[16];
#define VTQ_ROT(SV)                                \
  case SV:                                         \
    _Pragma("unroll") for (int k = 0; k < 16; ++k) o[k] = f[k + SV]; \
    break;
  switch (s) {
    VTQ_ROT(0) VTQ_ROT(1) VTQ_ROT(2) VTQ_ROT(3) VTQ_ROT(4) VTQ_ROT(5) VTQ_ROT(6)
    default:
#pragma unroll
      for (int k = 0; k < 16; ++k) o[k] = f[k + 7];
      break;
  }
#undef VTQ_ROT
  const size_t e0 = slot * PATCH_ELEMS + static_cast<size_t>(r) * 16;
  if (patches_f32 != nullptr) {
    stg256(patches_f32 + e0, *reinterpret_cast<uint32_t(*)[8]>(&o[0]));
    stg256(patches_f32 + e0 + 8, *reinterpret_cast<uint32_t(*)[8]>(&o[8]));
  }
  if (patches_16 != nullptr) {
    uint32_t h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = pack2<DT>(__uint_as_float(o[2 * k]), __uint_as_float(o[2 * k + 1]));
    stg256(reinterpret_cast<uint16_t*>(patches_16) + e0, h);
  }
  if (r == 0) store_uv(pos, scales, slot, og, H, W, scale_id);
}

// ------------------------------------------------------------------------------------------------
// vector gather, uint8 HWC source with the transform fused.  thread = (patch, row): the row's 48 interleaved bytes
// arrive with two (three) 256-bit loads, are rotated into place with funnel shifts, de-interleaved, transformed and
// written as three 32-byte (16-bit) / 64-byte (fp32) channel rows.  12 patches per 192-thread block.
// Preconditions: image base and outputs 32-byte aligned, total image bytes % 32 == 0.
// ------------------------------------------------------------------------------------------------
constexpr int GU_PATCHES = 12;

template <int C>
__device__ __forceinline__ void u8_row_channel(const uint32_t (&w)[12], float (&v)[16]) {
  // pixel j, channel C sits at byte 3j + C of the 48-byte row
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const int b0 = 3 * j + C, b1 = 3 * (j + 1) + C;
    float u0, u1;
    switch (b0 & 3) {
      case 0: u0 = byte_to_float<0>(w[b0 >> 2]); break;
      case 1: u0 = byte_to_float<1>(w[b0 >> 2]); break;
      case 2: u0 = byte_to_float<2>(w[b0 >> 2]); break;
      default: u0 = byte_to_float<3>(w[b0 >> 2]); break;
    }
    switch (b1 & 3) {
      case 0: u1 = byte_to_float<0>(w[b1 >> 2]); break;
      case 1: u1 = byte_to_float<1>(w[b1 >> 2]); break;
      case 2: u1 = byte_to_float<2>(w[b1 >> 2]); break;
      default: u1 = byte_to_float<3>(w[b1 >> 2]); break;
    }
    f2_unpack(normalize_u8x2(u0, u1), v[j], v[j + 1]);
  }
}

template <int DT>
__global__ void __launch_bounds__(192) patch_gather_u8_vec_kernel(const uint8_t* __restrict__ images, int H, int W,
                                                                  const double* __restrict__ samples, int n_set,
                                                                  int n, int patch_offset, int N_total,
                                                                  float* __restrict__ patches_f32,
                                                                  void* __restrict__ patches_16,
                                                                  float* __restrict__ pos, float* __restrict__ scales,
                                                                  float scale_id, int* __restrict__ oob_flag) {
  const int i = threadIdx.x & 15;
  const int p = blockIdx.x * GU_PATCHES + (threadIdx.x >> 4);
  if (p >= n) return;
  const int img = blockIdx.y;
  const PatchOrigin og = load_origin(samples, img % n_set, n, p, H, W, oob_flag);
  const size_t slot = static_cast<size_t>(img) * N_total + patch_offset + p;
  const uint8_t* src = images + ((static_cast<size_t>(img) * H + (og.y0 + i)) * W + og.x0) * 3;
  const uintptr_t a = reinterpret_cast<uintptr_t>(src);
  const int off = static_cast<int>(a & 31);
  const uint8_t* base = reinterpret_cast<const uint8_t*>(a & ~static_cast<uintptr_t>(31));
  uint32_t f[24];
  ldg256(base, *reinterpret_cast<uint32_t(*)[8]>(&f[0]));
  ldg256(base + 32, *reinterpret_cast<uint32_t(*)[8]>(&f[8]));
  if (off > 16) ldg256(base + 64, *reinterpret_cast<uint32_t(*)[8]>(&f[16]));
  else {
#pragma unroll
    for (int k = 16; k < 24; ++k) f[k] = 0u;
  }
  const int ws = off >> 2;
  const uint32_t bs = static_cast<uint32_t>(off & 3) * 8;
  uint32_t w[12];
#define VTQ_ROT(SV)                                                                                          \
  case SV:                                                                                                   \
    _Pragma("unroll") for (int k = 0; k < 12; ++k) w[k] = __funnelshift_r(f[k + SV], f[k + SV + 1], bs);    \
    break;
  switch (ws) {
    VTQ_ROT(0) VTQ_ROT(1) VTQ_ROT(2) VTQ_ROT(3) VTQ_ROT(4) VTQ_ROT(5) VTQ_ROT(6)
    default:
#pragma unroll
      for (int k = 0; k < 12; ++k) w[k] = __funnelshift_r(f[k + 7], f[k + 8], bs);
      break;
  }
#undef VTQ_ROT
  const size_t e0 = slot * PATCH_ELEMS + static_cast<size_t>(i) * 16;
  float v[16];
#define VTQ_EMIT(C)                                                                                      \
  u8_row_channel<C>(w, v);                                                                               \
  if (patches_f32 != nullptr) {                                                                          \
    uint32_t lo[8], hi[8];                                                                               \
    _Pragma("unroll") for (int k = 0; k < 8; ++k) { lo[k] = __float_as_uint(v[k]); hi[k] = __float_as_uint(v[8 + k]); } \
    stg256(patches_f32 + e0 + C * 256, lo);                                                              \
    stg256(patches_f32 + e0 + C * 256 + 8, hi);                                                          \
  }                                                                                                      \
  if (patches_16 != nullptr) {                                                                           \
    uint32_t h[8];                                                                                       \
    _Pragma("unroll") for (int k = 0; k < 8; ++k) h[k] = pack2<DT>(v[2 * k], v[2 * k + 1]);              \
    stg256(reinterpret_cast<uint16_t*>(patches_16) + e0 + C * 256, h);                                   \
  }
  VTQ_EMIT(0)
  VTQ_EMIT(1)
  VTQ_EMIT(2)
#undef VTQ_EMIT
  if (i == 0) store_uv(pos, scales, slot, og, H, W, scale_id);
}

// ------------------------------------------------------------------------------------------------
// uint8 HWC -> normalised fp32 CHW.  Vector form: 4 pixels (12 bytes) per thread, one float4 per channel plane.
// ------------------------------------------------------------------------------------------------
__global__ void normalize_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int H, int W,
                                    size_t total) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;  // over [img][c][y][x]
  if (idx >= total) return;
  const int x = static_cast<int>(idx % W);
  size_t rem = idx / W;
  const int y = static_cast<int>(rem % H);
  rem /= H;
  const int c = static_cast<int>(rem % 3);
  const size_t img = rem / 3;
  dst[idx] = normalize_u8(__ldg(src + ((img * H + y) * W + x) * 3 + c));
}

__global__ void __launch_bounds__(256) normalize_u8_vec_kernel(const uint8_t* __restrict__ src,
                                                               float* __restrict__ dst, size_t plane /* H*W */,
                                                               size_t quads /* n_img * H*W / 4 */) {
  const size_t q = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (q >= quads) return;
  const size_t pix = q * 4;  // first of 4 pixels; plane % 4 == 0 so they share an image
  const size_t img = pix / plane, in_plane = pix % plane;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src + pix * 3);
  uint32_t w[12];
  w[0] = __ldg(s);
  w[1] = __ldg(s + 1);
  w[2] = __ldg(s + 2);
  float4 o[3];
  // bytes: p0 = (0,1,2) p1 = (3,4,5) p2 = (6,7,8) p3 = (9,10,11)
  {
    float a, b, c2, d;
    f2_unpack(normalize_u8x2(byte_to_float<0>(w[0]), byte_to_float<3>(w[0])), a, b);
    f2_unpack(normalize_u8x2(byte_to_float<2>(w[1]), byte_to_float<1>(w[2])), c2, d);
    o[0] = make_float4(a, b, c2, d);
    f2_unpack(normalize_u8x2(byte_to_float<1>(w[0]), byte_to_float<0>(w[1])), a, b);
    f2_unpack(normalize_u8x2(byte_to_float<3>(w[1]), byte_to_float<2>(w[2])), c2, d);
    o[1] = make_float4(a, b, c2, d);
    f2_unpack(normalize_u8x2(byte_to_float<2>(w[0]), byte_to_float<1>(w[1])), a, b);
    f2_unpack(normalize_u8x2(byte_to_float<0>(w[2]), byte_to_float<3>(w[2])), c2, d);
    o[2] = make_float4(a, b, c2, d);
  }
  float* d0 = dst + img * 3 * plane + in_plane;
  *reinterpret_cast<float4*>(d0) = o[0];
  *reinterpret_cast<float4*>(d0 + plane) = o[1];
  *reinterpret_cast<float4*>(d0 + 2 * plane) = o[2];
}

// ------------------------------------------------------------------------------------------------
// 2x2 mean, floor mode; the summation tree and the exact /4 follow ATen's avg_pool2d (kh outer, kw inner):
// ((a00 + a01) + a10) + a11, then * 0.25.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pool4(float a00, float a01, float a10, float a11) {
  return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a00, a01), a10), a11), 0.25f);
}

__global__ void avgpool2x2_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int Ho,
                                  int Wo, size_t total) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int xo = static_cast<int>(idx % Wo);
  const size_t rem = idx / Wo;
  const int yo = static_cast<int>(rem % Ho);
  const size_t plane = rem / Ho;
  const float* s = src + (plane * H + 2 * yo) * static_cast<size_t>(W) + 2 * xo;
  dst[idx] = pool4(__ldg(s), __ldg(s + 1), __ldg(s + W), __ldg(s + W + 1));
}

// 4 outputs per thread: two 256-bit row loads, one 128-bit store.  W % 8 == 0, 32-byte aligned planes.
__global__ void __launch_bounds__(256) avgpool2x2_vec_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                             int H, int W, int Ho, int Wo, size_t total4) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= total4) return;
  const int wq = Wo >> 2;
  const int xq = static_cast<int>(idx % wq);
  const size_t rem = idx / wq;
  const int yo = static_cast<int>(rem % Ho);
  const size_t plane = rem / Ho;
  const float* s = src + (plane * H + 2 * yo) * static_cast<size_t>(W) + 8 * xq;
  uint32_t a[8], b[8];
  ldg256(s, a);
  ldg256(s + W, b);
  float4 o;
  o.x = pool4(__uint_as_float(a[0]), __uint_as_float(a[1]), __uint_as_float(b[0]), __uint_as_float(b[1]));
  o.y = pool4(__uint_as_float(a[2]), __uint_as_float(a[3]), __uint_as_float(b[2]), __uint_as_float(b[3]));
  o.z = pool4(__uint_as_float(a[4]), __uint_as_float(a[5]), __uint_as_float(b[4]), __uint_as_float(b[5]));
  o.w = pool4(__uint_as_float(a[6]), __uint_as_float(a[7]), __uint_as_float(b[6]), __uint_as_float(b[7]));
  *reinterpret_cast<float4*>(dst + (plane * Ho + yo) * static_cast<size_t>(Wo) + 4 * xq) = o;
}

// Level 1 of the pyramid straight from the decoded uint8 HWC image: transform each of the four pixels, then the
// 2x2 mean — the fp32 level-0 image is never materialised.  4 outputs per channel per thread (8 input pixels =
// 24 bytes per row).  W % 8 == 0, base 8-byte aligned.
__global__ void __launch_bounds__(256) avgpool2x2_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                            int H, int W, int Ho, int Wo, size_t total4) {
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;  // over [img][yo][xo/4]
  if (idx >= total4) return;
  const int wq = Wo >> 2;
  const int xq = static_cast<int>(idx % wq);
  const size_t rem = idx / wq;
  const int yo = static_cast<int>(rem % Ho);
  const size_t img = rem / Ho;
  const uint8_t* s0 = src + ((img * H + 2 * yo) * static_cast<size_t>(W) + 8 * xq) * 3;
  const uint8_t* s1 = s0 + static_cast<size_t>(W) * 3;
  uint32_t r0[6], r1[6];
#pragma unroll
  for (int k = 0; k < 6; k += 2) {
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(s0) + (k >> 1));
    const uint2 b = __ldg(reinterpret_cast<const uint2*>(s1) + (k >> 1));
    r0[k] = a.x; r0[k + 1] = a.y;
    r1[k] = b.x; r1[k + 1] = b.y;
  }
  float t0[24], t1[24];  // transformed bytes of both rows, interleaved order (pixel j, channel c at 3j + c)
#pragma unroll
  for (int b = 0; b < 24; b += 2) {
    float u0, u1, v0, v1;
    switch (b & 3) {
      case 0: u0 = byte_to_float<0>(r0[b >> 2]); u1 = byte_to_float<1>(r0[b >> 2]);
              v0 = byte_to_float<0>(r1[b >> 2]); v1 = byte_to_float<1>(r1[b >> 2]); break;
      default: u0 = byte_to_float<2>(r0[b >> 2]); u1 = byte_to_float<3>(r0[b >> 2]);
               v0 = byte_to_float<2>(r1[b >> 2]); v1 = byte_to_float<3>(r1[b >> 2]); break;
    }
    f2_unpack(normalize_u8x2(u0, u1), t0[b], t0[b + 1]);
    f2_unpack(normalize_u8x2(v0, v1), t1[b], t1[b + 1]);
  }
  const size_t plane = static_cast<size_t>(Ho) * Wo;
  float* d = dst + img * 3 * plane + static_cast<size_t>(yo) * Wo + 4 * xq;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float4 o;
    o.x = pool4(t0[0 + c], t0[3 + c], t1[0 + c], t1[3 + c]);
    o.y = pool4(t0[6 + c], t0[9 + c], t1[6 + c], t1[9 + c]);
    o.z = pool4(t0[12 + c], t0[15 + c], t1[12 + c], t1[15 + c]);
    o.w = pool4(t0[18 + c], t0[21 + c], t1[18 + c], t1[21 + c]);
    *reinterpret_cast<float4*>(d + c * plane) = o;
  }
}

static bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

static bool force_scalar_gather() {
  static const bool on = [] {
    const char* e = std::getenv("VTQ_GATHER_SCALAR");
    return e != nullptr && e[0] == '1';   // A/B switch: the generic one-load-per-pixel kernels
  }();
  return on;
}

}  // namespace vtq

// ================================================================================================
// C-ABI wrappers
// ================================================================================================
using namespace vtq;


// ------------------------------------------------------------------------------------------------
// Coordinate sampler (SURVEY 8f "next" #3): the reference's default law, PatchSampler(GRID_TYPE_PERTURBED_SIMPLE) ->
// stratified_grid_sampling (data/patch_sampling.py:236-237, :308-327, :362-376), for a whole batch on the device.
// One block per image: a uniform random permutation of the `height x width` grid cells (64-bit keys = 32 random bits |
// cell index, bitonic sort in shared memory), the first n cells are the image's n DISTINCT grid points; each is
// jittered by U(-2a, 2a) cells, moved to the cell centre, clipped to [0, 1] and scaled to [0, h-ho] x [0, w-wo].
// Randomness: Philox4x32-10 keyed by two 64-bit words the caller drew with its own generator (read on the device).
// ------------------------------------------------------------------------------------------------
namespace vtq {
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {  // uniform in [0, 1) with 53 random bits
  const unsigned long long v = ((static_cast<unsigned long long>(hi) << 32) | lo) >> 11;
  return static_cast<double>(v) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(256) sample_grid_kernel(const unsigned long long* __restrict__ key, int h, int w, int ho,
                                                          int wo, int n, int height, int width, int padded,
                                                          double amount, double* __restrict__ out) {
  extern __shared__ unsigned long long sg_keys[];  // [padded]
  const int img = blockIdx.x;
  const int cells = height * width;
  const uint32_t k0 = static_cast<uint32_t>(key[0]), k1 = static_cast<uint32_t>(key[0] >> 32);
  const uint32_t s0 = static_cast<uint32_t>(key[1]), s1 = static_cast<uint32_t>(key[1] >> 32);
  for (int c = threadIdx.x; c < padded; c += blockDim.x) {
    unsigned long long v = ~0ull;
    if (c < cells) {
      uint32_t r[4];
      philox4x32_10(static_cast<uint32_t>(c), static_cast<uint32_t>(img), s0, s1 ^ 0x5A17u, k0, k1, r);
      v = (static_cast<unsigned long long>(r[0]) << 32) | static_cast<uint32_t>(c);
    }
    sg_keys[c] = v;
  }
  __syncthreads();
  for (int k = 2; k <= padded; k <<= 1) {       // bitonic sort, ascending
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < padded; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = sg_keys[i], b = sg_keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { sg_keys[i] = b; sg_keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  double* oy = out + static_cast<size_t>(img) * 2 * n;
  double* ox = oy + n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = static_cast<int>(sg_keys[i] & 0xffffffffull);
    const double gy = static_cast<double>(c % height);  // the reference flattens its (2, width, height) grid this way
    const double gx = static_cast<double>(c / height);
    uint32_t r[4];
    philox4x32_10(static_cast<uint32_t>(i), static_cast<uint32_t>(img), s0, s1 ^ 0xC0FFEEu, k0, k1, r);
    const double jy = (2.0 * u01_53(r[0], r[1]) - 1.0) * (2.0 * amount);
    const double jx = (2.0 * u01_53(r[2], r[3]) - 1.0) * (2.0 * amount);
    const double py = fmin(fmax((gy + jy) / height + 0.5 / height, 0.0), 1.0);
    const double px = fmin(fmax((gx + jx) / width + 0.5 / width, 0.0), 1.0);
    oy[i] = py * (h - ho);
    ox[i] = px * (w - wo);
  }
}
}  // namespace vtq

extern "C" int vtq_patch_gather(vtq_ctx* ctx, const float* images, int n_img, int H, int W, const double* samples,
                                int n_set, int n, int patch_offset, int N_total, float* patches_f32,
                                void* patches_16, int dtype, float* pos, float* scales, int scale_id,
                                void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, images && samples, "null pointer");
  VTQ_CHECK_ARG(ctx, H >= PATCH && W >= PATCH, "image smaller than one patch");
  VTQ_CHECK_ARG(ctx, n_img >= 1 && n_set >= 1 && n_img % n_set == 0, "n_img must be a multiple of n_set");
  VTQ_CHECK_ARG(ctx, n >= 0 && patch_offset >= 0 && patch_offset + n <= N_total, "patch range");
  VTQ_CHECK_ARG(ctx, n_img <= 65535, "n_img <= 65535");
  VTQ_CHECK_ARG(ctx, dtype == VTQ_F16 || dtype == VTQ_BF16, "dtype");
  if (n == 0) return VTQ_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sid = static_cast<float>(scale_id);
  const size_t total = static_cast<size_t>(n_img) * 3 * H * W;
  const bool vec = !force_scalar_gather() && aligned32(images) && total % 8 == 0 &&
                   (patches_f32 == nullptr || aligned32(patches_f32)) &&
                   (patches_16 == nullptr || aligned32(patches_16));
  if (vec) {
    dim3 grid((n + GV_PATCHES - 1) / GV_PATCHES, n_img);
    if (dtype == VTQ_F16)
      patch_gather_vec_kernel<DT_F16><<<grid, 192, 0, st>>>(images, H, W, samples, n_set, n, patch_offset, N_total,
                                                            patches_f32, patches_16, pos, scales, sid, ctx->oob_flag_dev);
    else
      patch_gather_vec_kernel<DT_BF16><<<grid, 192, 0, st>>>(images, H, W, samples, n_set, n, patch_offset, N_total,
                                                             patches_f32, patches_16, pos, scales, sid, ctx->oob_flag_dev);
  } else {
    dim3 grid(n, n_img);
    if (dtype == VTQ_F16)
      patch_gather_kernel<DT_F16><<<grid, 192, 0, st>>>(images, H, W, samples, n_set, n, patch_offset, N_total,
                                                        patches_f32, patches_16, pos, scales, sid, ctx->oob_flag_dev);
    else
      patch_gather_kernel<DT_BF16><<<grid, 192, 0, st>>>(images, H, W, samples, n_set, n, patch_offset, N_total,
                                                         patches_f32, patches_16, pos, scales, sid, ctx->oob_flag_dev);
  }
  VTQ_CHECK_LAUNCH(ctx, "patch_gather launch");
  return VTQ_OK;
}

extern "C" int vtq_patch_gather_u8(vtq_ctx* ctx, const uint8_t* images, int n_img, int H, int W,
                                   const double* samples, int n_set, int n, int patch_offset, int N_total,
                                   float* patches_f32, void* patches_16, int dtype, float* pos, float* scales,
                                   int scale_id, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, images && samples, "null pointer");
  VTQ_CHECK_ARG(ctx, H >= PATCH && W >= PATCH, "image smaller than one patch");
  VTQ_CHECK_ARG(ctx, n_img >= 1 && n_set >= 1 && n_img % n_set == 0, "n_img must be a multiple of n_set");
  VTQ_CHECK_ARG(ctx, n >= 0 && patch_offset >= 0 && patch_offset + n <= N_total, "patch range");
  VTQ_CHECK_ARG(ctx, n_img <= 65535, "n_img <= 65535");
  VTQ_CHECK_ARG(ctx, dtype == VTQ_F16 || dtype == VTQ_BF16, "dtype");
  if (n == 0) return VTQ_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float sid = static_cast<float>(scale_id);
  const size_t total = static_cast<size_t>(n_img) * 3 * H * W;
  const bool vec = !force_scalar_gather() && aligned32(images) && total % 32 == 0 &&
                   (patches_f32 == nullptr || aligned32(patches_f32)) &&
                   (patches_16 == nullptr || aligned32(patches_16));
  if (vec) {
    dim3 grid((n + GU_PATCHES - 1) / GU_PATCHES, n_img);
    if (dtype == VTQ_F16)
      patch_gather_u8_vec_kernel<DT_F16><<<grid, 192, 0, st>>>(images, H, W, samples, n_set, n, patch_offset, N_total,
                                                               patches_f32, patches_16, pos, scales, sid, ctx->oob_flag_dev);
    else
      patch_gather_u8_vec_kernel<DT_BF16><<<grid, 192, 0, st>>>(images, H, W, samples, n_set, n, patch_offset, N_total,
                                                                patches_f32, patches_16, pos, scales, sid, ctx->oob_flag_dev);
  } else {
    dim3 grid(n, n_img);
    if (dtype == VTQ_F16)
      patch_gather_u8_kernel<DT_F16><<<grid, 192, 0, st>>>(images, H, W, samples, n_set, n, patch_offset, N_total,
                                                           patches_f32, patches_16, pos, scales, sid, ctx->oob_flag_dev);
    else
      patch_gather_u8_kernel<DT_BF16><<<grid, 192, 0, st>>>(images, H, W, samples, n_set, n, patch_offset, N_total,
                                                            patches_f32, patches_16, pos, scales, sid, ctx->oob_flag_dev);
  }
  VTQ_CHECK_LAUNCH(ctx, "patch_gather_u8 launch");
  return VTQ_OK;
}

extern "C" int vtq_normalize_u8(vtq_ctx* ctx, const uint8_t* src, float* dst, int n_img, int H, int W, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, src && dst, "null pointer");
  VTQ_CHECK_ARG(ctx, n_img >= 1 && H >= 1 && W >= 1, "shape");
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(n_img) * 3 * plane;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (plane % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const size_t quads = static_cast<size_t>(n_img) * plane / 4;
    normalize_u8_vec_kernel<<<static_cast<unsigned>((quads + 255) / 256), 256, 0, st>>>(src, dst, plane, quads);
  } else {
    normalize_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(src, dst, H, W, total);
  }
  VTQ_CHECK_LAUNCH(ctx, "normalize_u8 launch");
  return VTQ_OK;
}

extern "C" int vtq_avgpool2x2(vtq_ctx* ctx, const float* src, float* dst, int planes, int H, int W, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, src && dst, "null pointer");
  VTQ_CHECK_ARG(ctx, planes >= 1 && H >= 2 && W >= 2, "shape");
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = static_cast<size_t>(planes) * Ho * Wo;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (W % 8 == 0 && aligned32(src) && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const size_t total4 = total / 4;
    avgpool2x2_vec_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, st>>>(src, dst, H, W, Ho, Wo, total4);
  } else {
    avgpool2x2_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(src, dst, H, W, Ho, Wo, total);
  }
  VTQ_CHECK_LAUNCH(ctx, "avgpool2x2 launch");
  return VTQ_OK;
}

extern "C" int vtq_avgpool2x2_u8(vtq_ctx* ctx, const uint8_t* src, float* dst, int n_img, int H, int W,
                                 void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, src && dst, "null pointer");
  VTQ_CHECK_ARG(ctx, n_img >= 1 && H >= 2 && W >= 8, "shape");
  VTQ_CHECK_ARG(ctx, W % 8 == 0, "vtq_avgpool2x2_u8 needs W % 8 == 0 (use vtq_normalize_u8 + vtq_avgpool2x2 otherwise)");
  VTQ_CHECK_ARG(ctx, (reinterpret_cast<uintptr_t>(src) & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                "alignment (src 8 B, dst 16 B)");
  const int Ho = H / 2, Wo = W / 2;
  const size_t total4 = static_cast<size_t>(n_img) * Ho * (Wo / 4);
  avgpool2x2_u8_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, dst, H, W, Ho, Wo, total4);
  VTQ_CHECK_LAUNCH(ctx, "avgpool2x2_u8 launch");
  return VTQ_OK;
}

extern "C" int vtq_sample_grid(vtq_ctx* ctx, const void* key2, int batch, int h, int w, int ho, int wo, int n,
                               double perturbed_amount, double* out, void* stream) {
  VTQ_ENTER(ctx);
  VTQ_CHECK_ARG(ctx, key2 && out, "null pointer");
  VTQ_CHECK_ARG(ctx, batch >= 1 && n >= 1 && h > ho && w > wo && ho >= 1 && wo >= 1, "shape");
  const double aspect = static_cast<double>(h) / static_cast<double>(w);
  int width = static_cast<int>(ceil(sqrt(static_cast<double>(n) / aspect)));
  if (width < 1) width = 1;
  const int height = static_cast<int>(ceil(width * aspect));
  const long long cells = static_cast<long long>(height) * width;
  VTQ_CHECK_ARG(ctx, cells >= n, "grid smaller than the sample count");
  int padded = 1;
  while (padded < cells) padded <<= 1;
  const int smem = padded * 8;
  VTQ_CHECK_ARG(ctx, smem <= ctx->smem_optin, "too many samples per image for the shared-memory sort (n <= ~25000)");
  if (int rc = vtq::ensure_dyn_smem(ctx, vtq::sample_grid_kernel, ctx->smem_optin, "sample_grid: cudaFuncSetAttribute")) return rc;
  vtq::sample_grid_kernel<<<batch, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const unsigned long long*>(key2), h, w, ho, wo, n, height, width, padded, perturbed_amount, out);
  VTQ_CHECK_LAUNCH(ctx, "sample_grid launch");
  return VTQ_OK;
}
