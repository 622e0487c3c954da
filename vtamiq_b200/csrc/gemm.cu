// K3/K5 — dense projection  out = epilogue(A[M][K] · W[N][K]^T + bias)  on the 5th-gen tensor cores.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   : A tile 128x64 and W tile BNx64 (16-bit, 128B swizzle) into a smem ring
//   warp 1      MMA issuer     : tcgen05.mma.cta_group::1.kind::f16, M=128 N=BN K=16, fp32 accumulators in
//                                TMEM, two accumulator buffers so the epilogue of tile i overlaps tile i+1
//   warps 2..5  epilogue       : tcgen05.ld (one TMEM lane = one output row per thread), bias / erf-GELU /
//                                LayerScale in registers, 128B-swizzled staging in smem, then per-warp TMA
//                                store — or TMA reduce-add for the fp32 residual stream, so the residual
//                                read-modify-write never travels through the SM.
// Replaces the ATen addmm/conv calls listed in include/vtamiq_b200.h (vtq_gemm).
#include "common.cuh"
#include "host.h"

namespace vtq {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 x 16-bit = one 128-byte swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_STAGING_BYTES = 4 * 2 * 4096;  // 4 epilogue warps x 2 buffers x (32 rows x 128 B)

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;  // double-buffered accumulator
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_STAGING_BYTES + 256 /*barriers*/ + 1024 /*align*/;
};

enum : int { EPI_H = 0, EPI_F32 = 1 };

// erf-GELU on two values at once, fp32 throughout (packed FFMA2).  erf(z) = z * P(u), u = z^2 * (2/3.5^2) - 1,
// P = degree-12 Chebyshev fit of erf(z)/z on |z| <= 3.5 re-expanded in u (well conditioned on [-1,1]);
// |z| is clamped to 3.5 where erf is within 7.4e-7 of +-1.  Max |erf error| 6.5e-7, max |gelu error| 1.7e-6
// (tests/test_host_logic.py::test_gelu_polynomial_accuracy restates and checks the same coefficients).
// The epilogue is issue-bound next to a K=768 mainloop: this is ~11 issue slots per element instead of ~25 for
// erff().
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  constexpr float kC[13] = {4.038730562e-01f, -2.007001489e-01f, 1.467439830e-01f, -1.146334782e-01f,
                            8.848482370e-02f, -6.463798881e-02f, 4.461858794e-02f, -3.037289716e-02f,
                            1.790029742e-02f, -6.848857272e-03f, 3.642286174e-03f, -4.138993565e-03f,
                            1.783549204e-03f};
  const float z0 = fminf(fmaxf(x0 * 0.70710678118654752440f, -3.5f), 3.5f);
  const float z1 = fminf(fmaxf(x1 * 0.70710678118654752440f, -3.5f), 3.5f);
  const f32x2 z = f2_pack(z0, z1);
  const f32x2 u = f2_fma(f2_mul(z, z), f2_pack(0.16326530612244897f, 0.16326530612244897f), f2_pack(-1.0f, -1.0f));
  f32x2 acc = f2_pack(kC[12], kC[12]);
#pragma unroll
  for (int i = 11; i >= 0; --i) acc = f2_fma(acc, u, f2_pack(kC[i], kC[i]));
  const f32x2 e = f2_mul(z, acc);                       // erf(x / sqrt2)
  const f32x2 h = f2_mul(f2_pack(x0, x1), f2_pack(0.5f, 0.5f));
  f2_unpack(f2_fma(h, e, h), x0, x1);                   // 0.5 x (1 + erf)
}

template <int DT, int BN, int EPI, bool GELU>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const float* __restrict__ bias,
                const float* __restrict__ gamma, int M, int N, int K, int accumulate) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms are 1024 B: align the ring and the staging buffers.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + GEMM_STAGING_BYTES);
  uint64_t* full_bar = bars;                          // [STAGES]
  uint64_t* empty_bar = bars + Cfg::STAGES;           // [STAGES]
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;        // [2]
  uint64_t* acc_empty = bars + 2 * Cfg::STAGES + 2;   // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = K / GEMM_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / num_n) * GEMM_BM;
        const int n0 = (t % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = ring + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m0);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n0);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(DT, GEMM_BM, BN, 0, 0);
      uint32_t stage = 0, phase = 0;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const uint32_t acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t da = umma_smem_desc(sa, 16, 1024);
          const uint64_t db = umma_smem_desc(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle row: +2 in the (addr >> 4) field
            umma_f16_ss(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);  // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ------------------------------- epilogue -----------------------------------
    const int ew = warp - 2;             // staging slot
    const int lane_grp = warp & 3;       // TMEM lanes this warp may touch: [32*lane_grp, +32)
    uint8_t* my_staging = staging + ew * 8192;
    const uint32_t swz = static_cast<uint32_t>(lane & 7);
    uint32_t n_store = 0;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int m0 = (t / num_n) * GEMM_BM;
      const int n0 = (t % num_n) * BN;
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + acc * BN;
      const int row0 = m0 + lane_grp * 32;  // first output row of this warp's 32-row slab

      if constexpr (EPI == EPI_F32) {
        // 32 fp32 columns (128 B per row) per staged box
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int ncol = n0 + c * 32;
          if (ncol >= N) break;
          uint32_t r[32];
          tmem_ld32(t_row + c * 32, r);
          tmem_wait_ld();
          if (c == BN / 32 - 1 || ncol + 32 >= N) {  // last read of this accumulator
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
          }
          uint8_t* buf = my_staging + (n_store & 1) * 4096;
          if (lane == 0) tma_wait_group_read<1>();
          __syncwarp();
          const uint32_t row_addr = smem_u32(buf) + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 bv = __ldg(reinterpret_cast<const float4*>(bias + ncol) + j);
            float v0 = __uint_as_float(r[4 * j + 0]) + bv.x;
            float v1 = __uint_as_float(r[4 * j + 1]) + bv.y;
            float v2 = __uint_as_float(r[4 * j + 2]) + bv.z;
            float v3 = __uint_as_float(r[4 * j + 3]) + bv.w;
            if (gamma != nullptr) {
              float4 gv = __ldg(reinterpret_cast<const float4*>(gamma + ncol) + j);
              v0 *= gv.x; v1 *= gv.y; v2 *= gv.z; v3 *= gv.w;
            }
            st_shared_v4(row_addr + ((static_cast<uint32_t>(j) ^ swz) << 4), __float_as_uint(v0),
                         __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (accumulate) tma_reduce_add_2d(&tmO, buf, ncol, row0);
            else tma_store_2d(&tmO, buf, ncol, row0);
            tma_commit_group();
          }
          ++n_store;
        }
      } else {
        // 64 16-bit columns (128 B per row) per staged box = two TMEM loads
#pragma unroll 1
        for (int c = 0; c < BN / 64; ++c) {
          const int ncol = n0 + c * 64;
          if (ncol >= N) break;
          uint8_t* buf = my_staging + (n_store & 1) * 4096;
          if (lane == 0) tma_wait_group_read<1>();
          __syncwarp();
          const uint32_t row_addr = smem_u32(buf) + lane * 128;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[32];
            tmem_ld32(t_row + c * 64 + hh * 32, r);
            tmem_wait_ld();
            if (hh == 1 && (c == BN / 64 - 1 || ncol + 64 >= N)) {
              tc_fence_before();
              mbar_arrive(&acc_empty[acc]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + ncol + hh * 32) + 2 * j);
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + ncol + hh * 32) + 2 * j + 1);
              float v[8];
              v[0] = __uint_as_float(r[8 * j + 0]) + b0.x;
              v[1] = __uint_as_float(r[8 * j + 1]) + b0.y;
              v[2] = __uint_as_float(r[8 * j + 2]) + b0.z;
              v[3] = __uint_as_float(r[8 * j + 3]) + b0.w;
              v[4] = __uint_as_float(r[8 * j + 4]) + b1.x;
              v[5] = __uint_as_float(r[8 * j + 5]) + b1.y;
              v[6] = __uint_as_float(r[8 * j + 6]) + b1.z;
              v[7] = __uint_as_float(r[8 * j + 7]) + b1.w;
              if constexpr (GELU) {
#pragma unroll
                for (int e = 0; e < 8; e += 2) gelu_erf2(v[e], v[e + 1]);
              }
              const uint32_t chunk = static_cast<uint32_t>(hh * 4 + j);
              st_shared_v4(row_addr + ((chunk ^ swz) << 4), pack2<DT>(v[0], v[1]), pack2<DT>(v[2], v[3]),
                           pack2<DT>(v[4], v[5]), pack2<DT>(v[6], v[7]));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmO, buf, ncol, row0);
            tma_commit_group();
          }
          ++n_store;
        }
      }
    }
    if (lane == 0) tma_wait_group<0>();  // all bulk stores retired before the CTA's smem goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int DT, int BN, int EPI, bool GELU>
static int launch_one(vtq_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                      const float* bias, const float* gamma, int M, int N, int K, int accumulate,
                      cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<DT, BN, EPI, GELU>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(ctx, e, "gemm: cudaFuncSetAttribute");
    configured = true;
  }
  const int num_tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + BN - 1) / BN);
  const int grid = num_tiles < ctx->num_sms ? num_tiles : ctx->num_sms;
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, tmO, bias, gamma, M, N, K, accumulate);
  VTQ_CHECK_LAUNCH(ctx, "gemm launch");
  return VTQ_OK;
}

int launch_gemm(vtq_ctx* ctx, const void* A, int64_t lda, const void* W, const float* bias, int M, int N, int K,
                int dtype, int epilogue, void* out, int64_t ldo, const float* gamma, cudaStream_t st) {
  VTQ_CHECK_ARG(ctx, A && W && bias && out, "null pointer");
  VTQ_CHECK_ARG(ctx, M >= 1 && N >= 64 && K >= 64, "empty problem");
  VTQ_CHECK_ARG(ctx, K % GEMM_BK == 0, "K must be a multiple of 64");
  VTQ_CHECK_ARG(ctx, N % 64 == 0, "N must be a multiple of 64");
  VTQ_CHECK_ARG(ctx, dtype == VTQ_F16 || dtype == VTQ_BF16, "dtype must be VTQ_F16 or VTQ_BF16");
  VTQ_CHECK_ARG(ctx, epilogue >= VTQ_EPI_BIAS_H && epilogue <= VTQ_EPI_BIAS_RESID_F32, "unknown epilogue");
  if (lda <= 0) lda = K;
  VTQ_CHECK_ARG(ctx, lda >= K && lda % 8 == 0, "lda must be >= K and a multiple of 8 elements");
  const bool out32 = epilogue == VTQ_EPI_BIAS_F32 || epilogue == VTQ_EPI_BIAS_RESID_F32;
  if (ldo <= 0) ldo = N;
  VTQ_CHECK_ARG(ctx, ldo >= N && (ldo * (out32 ? 4 : 2)) % 16 == 0, "ldo must be >= N and 16-byte aligned");
  VTQ_CHECK_ARG(ctx, (reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) |
                      reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias)) % 16 == 0,
                "pointers must be 16-byte aligned");
  VTQ_CHECK_ARG(ctx, gamma == nullptr || reinterpret_cast<uintptr_t>(gamma) % 16 == 0, "gamma alignment");

  // Wide N (QKV, fc1) uses 128x256 tiles; N=768 GEMMs use 128x128 tiles to cut wave quantisation.
  const int BN = (N % 256 == 0 && N >= 1536) ? 256 : 128;
  VTQ_CHECK_ARG(ctx, N % BN == 0 || N % 64 == 0, "N tiling");

  CUtensorMap tmA, tmB, tmO;
  const CUtensorMapDataType dt16 = tm_dtype16(dtype);
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {GEMM_BK, GEMM_BM};
    int rc = make_tensor_map(ctx, &tmA, dt16, 2, A, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {GEMM_BK, static_cast<uint32_t>(BN)};
    int rc = make_tensor_map(ctx, &tmB, dt16, 2, W, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldo) * (out32 ? 4 : 2)};
    uint32_t box[2] = {out32 ? 32u : 64u, 32u};
    int rc = make_tensor_map(ctx, &tmO, out32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt16, 2, out, dims, strides, box);
    if (rc) return rc;
  }
  const int acc = epilogue == VTQ_EPI_BIAS_RESID_F32 ? 1 : 0;
  if (epilogue != VTQ_EPI_BIAS_RESID_F32) gamma = nullptr;

#define VTQ_GEMM_DISPATCH(DTV, BNV)                                                                             \
  switch (epilogue) {                                                                                           \
    case VTQ_EPI_BIAS_H: return launch_one<DTV, BNV, EPI_H, false>(ctx, tmA, tmB, tmO, bias, gamma, M, N, K, 0, st); \
    case VTQ_EPI_BIAS_GELU_H: return launch_one<DTV, BNV, EPI_H, true>(ctx, tmA, tmB, tmO, bias, gamma, M, N, K, 0, st); \
    default: return launch_one<DTV, BNV, EPI_F32, false>(ctx, tmA, tmB, tmO, bias, gamma, M, N, K, acc, st);    \
  }
  if (dtype == VTQ_F16) {
    if (BN == 256) { VTQ_GEMM_DISPATCH(DT_F16, 256) } else { VTQ_GEMM_DISPATCH(DT_F16, 128) }
  } else {
    if (BN == 256) { VTQ_GEMM_DISPATCH(DT_BF16, 256) } else { VTQ_GEMM_DISPATCH(DT_BF16, 128) }
  }
#undef VTQ_GEMM_DISPATCH
}

}  // namespace vtq
