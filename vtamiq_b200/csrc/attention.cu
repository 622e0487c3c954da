// K6 — fused multi-head self-attention, softmax(Q K^T / 8) V, head_dim 64, no mask.
//
// One CTA per (PAIR of 128-row query tiles, head, sequence); ref and dist sequences of the whole batch go through
// one launch (sequence index = img * B + b).  The S x S score matrix lives only in TMEM / registers:
//   warp 0      TMA producer : both Q tiles once, then K/V tiles (128 keys x 64) through a 3-deep smem ring that
//                              the two query tiles share
//   warp 1      MMA issuer   : S_t = Q_t K^T (tcgen05.mma M128 N128 K16 x4, both operands K-major) and
//                              O_t += P_t V (M128 N64 K16 x8, A = P from smem, B = V as an MN-major operand — V is
//                              consumed exactly as the QKV GEMM wrote it, no transpose pass), for t = A, B
//   warps 4..7  softmax A    : one query row per thread.  The whole 128-key score row is pulled from TMEM into
//   warps 8..11 softmax B      registers ONCE and the S buffer is released immediately, so Q_t K^T of the next key
//                              tile runs underneath this tile's exponentials; running max / sum in fp32, lazy
//                              rescaling (O is only corrected when the max grows by > 2^8), P rounded to 16 bits
//                              into 128B-swizzled smem, final O / l through per-warp TMA stores.
// Registers are re-partitioned with setmaxnreg: the producer warpgroup drops to 56, the softmax warpgroups grow to
// 224 (56*128 + 224*256 = 168*384, the launch allocation) so a whole score row (128 fp32) fits.  TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384).
// Replaces modules/VisionTransformer/transformer.py:158-166 (matmul, /sqrt(d), softmax, matmul, permute copy).
#include <cstdlib>

#include "common.cuh"
#include "host.h"

namespace vtq {

constexpr int ATT_BQ = 128;   // query rows per tile (two tiles per CTA)
constexpr int ATT_BKV = 128;  // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 3 * 128;  // warpgroup 0: TMA + MMA warps (+2 idle), warpgroups 1, 2: softmax A, B
constexpr int ATT_TILE_BYTES = 128 * ATT_D * 2;  // 16 KB: a 128-row x 64 x 16-bit tile
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_P_BYTES = ATT_BQ * ATT_BKV * 2;  // 32 KB per query tile
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES * (2 + 2 * ATT_KV_STAGES) + 2 * ATT_P_BYTES + 256;
static_assert(ATT_SMEM_BYTES <= 227 * 1024, "smem budget");
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_TMEM_S = 0;    // + t * 128
constexpr uint32_t ATT_TMEM_O = 256;  // + t * 64

template <int DT>
__global__ void __launch_bounds__(ATT_THREADS, 1)
    attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, int S,
                     int heads) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzle atoms need a 1024 B aligned base
  uint8_t* sQ = smem;                                  // [2 tiles]
  uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;               // [stages]
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;   // [stages]
  uint8_t* sP = sV + ATT_KV_STAGES * ATT_TILE_BYTES;   // [2 tiles]; reused as output staging at the end
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_P_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // [3]
  uint64_t* kv_empty = bars + 4;      // [3]
  uint64_t* s_full = bars + 7;        // [2] S_t(j) complete                      (MMA commit)
  uint64_t* s_free = bars + 9;        // [2] S_t(j) copied to registers           (128 arrivals)
  uint64_t* p_full = bars + 11;       // [2] P_t(j) in smem, O_t rescaled         (128 arrivals)
  uint64_t* pv_done = bars + 13;      // [2] O_t += P_t(j) V(j) complete          (MMA commit)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * ATT_BQ);
  const int head = blockIdx.y;
  const int seq = blockIdx.z;
  const int hidden = heads * ATT_D;
  const int nkv = (S + ATT_BKV - 1) / ATT_BKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 128);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      tma_load_3d(sQ, &tmQKV, q_full, head * ATT_D, q0, seq);
      tma_load_3d(sQ + ATT_TILE_BYTES, &tmQKV, q_full, head * ATT_D, q0 + ATT_BQ, seq);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % ATT_KV_STAGES;
        const uint32_t ph = (j / ATT_KV_STAGES) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
        tma_load_3d(sK + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], hidden + head * ATT_D, j * ATT_BKV, seq);
        tma_load_3d(sV + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], 2 * hidden + head * ATT_D, j * ATT_BKV, seq);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_f16(DT, ATT_BQ, ATT_BKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_f16(DT, ATT_BQ, ATT_D, 0, 1);  // B (=V) is MN-major

      auto issue_qk = [&](int t, int j) {
        const uint64_t dQ = umma_smem_desc(smem_u32(sQ + t * ATT_TILE_BYTES), 16, 1024);
        const uint64_t dK = umma_smem_desc(smem_u32(sK + (j % ATT_KV_STAGES) * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_f16_ss(tmem_base + ATT_TMEM_S + t * 128, dQ + uint64_t(k * 2), dK + uint64_t(k * 2), idesc_qk,
                      k ? 1u : 0u);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {
        // O_t (+)= P_t(j) V(j): 8 k-steps of 16 keys.  P: two 64-key K-major blocks of 16 KB.  V: rows = keys,
        // 128 B apart, 8-key groups 1024 B apart -> one k-step advances the start address by 2048 B.
        const uint32_t aP = smem_u32(sP + t * ATT_P_BYTES);
        const uint32_t aV = smem_u32(sV + (j % ATT_KV_STAGES) * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
          const uint64_t dP = umma_smem_desc(aP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
          const uint64_t dV = umma_smem_desc(aV + kk * 2048, 1024, 1024);
          umma_f16_ss(tmem_base + ATT_TMEM_O + t * 64, dP, dV, idesc_pv, (j | kk) ? 1u : 0u);
        }
        umma_commit(&pv_done[t]);
      };

      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      for (int j = 0; j < nkv; ++j) {
        const bool more = j + 1 < nkv;
        if (more) mbar_wait(&kv_full[(j + 1) % ATT_KV_STAGES], ((j + 1) / ATT_KV_STAGES) & 1);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (more) {  // next score tile as soon as the softmax group has pulled S_t(j) into registers
            mbar_wait(&s_free[t], j & 1);
            tc_fence_after();
            issue_qk(t, j + 1);
          }
          mbar_wait(&p_full[t], j & 1);
          tc_fence_after();
          issue_pv(t, j);
        }
        umma_commit(&kv_empty[j % ATT_KV_STAGES]);  // K(j), V(j) no longer needed once everything above retires
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------- softmax / correction / output --------------
    const int t = (warp - 4) >> 2;         // query tile of this warpgroup
    const int lane_grp = warp & 3;         // TMEM lane quarter of this warp
    const int row = lane_grp * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
    const uint32_t tS = t_lane + ATT_TMEM_S + t * 128;
    const uint32_t tO = t_lane + ATT_TMEM_O + t * 64;
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const uint32_t p_row = smem_u32(sP + t * ATT_P_BYTES) + row * 128;
    const float c = 0.125f * 1.44269504088896340736f;  // (1/sqrt(64)) * log2(e)
    const f32x2 c2 = f2_pack(c, c);

    float m = -INFINITY;  // reference max (raw score domain) that P and O are currently scaled by
    float l = 0.f;        // running sum of exp
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t r[128];
      tmem_ld32(tS + 0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
      tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
      tmem_ld32(tS + 64, *reinterpret_cast<uint32_t(*)[32]>(&r[64]));
      tmem_ld32(tS + 96, *reinterpret_cast<uint32_t(*)[32]>(&r[96]));
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&s_free[t]);  // S_t may be overwritten by Q_t K(j+1)^T from here on

      const int kv_valid = S - j * ATT_BKV;  // keys of this tile that exist (>= 1)
      if (kv_valid < ATT_BKV) {              // CTA-uniform: only the last key tile of a ragged sequence
#pragma unroll
        for (int e = 0; e < 128; ++e)
          if (e >= kv_valid) r[e] = 0xff800000u;  // -inf -> exp2 gives exactly 0
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int e = 0; e < 128; e += 8) {
        mx0 = fmax3(mx0, __uint_as_float(r[e + 0]), __uint_as_float(r[e + 1]));
        mx1 = fmax3(mx1, __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
        mx2 = fmax3(mx2, __uint_as_float(r[e + 4]), __uint_as_float(r[e + 5]));
        mx3 = fmax3(mx3, __uint_as_float(r[e + 6]), __uint_as_float(r[e + 7]));
      }
      float m_new = fmaxf(fmaxf(m, fmax3(mx0, mx1, mx2)), mx3);
      // Lazy rescale: keep the old reference max while the true max grew by < 2^8 in the exp2 domain — P then
      // stays <= 256 (exact in fp16/bf16 range, fp32 sums) and O needs no correction.  First tile: m = -inf.
      if ((m_new - m) * c <= 8.0f) m_new = m;

      if (j > 0) {
        // P_t(j-1) V(j-1) must have retired before P_t is overwritten or O_t is touched
        mbar_wait(&pv_done[t], (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, m_new != m)) {
          const float alpha = ex2_approx((m - m_new) * c);
          l *= alpha;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t o[32];
            tmem_ld32(tO + hh * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st32(tO + hh * 32, o);
          }
          tmem_wait_st();
        }
      }
      m = m_new;
      const float nmc = -m_new * c;
      const f32x2 nmc2 = f2_pack(nmc, nmc);

      // p = exp2(s*c - m*c): packed FFMA2 + MUFU.EX2, packed partial sums, 16-bit P into swizzled smem
      f32x2 sum_a = 0ull, sum_b = 0ull;
#pragma unroll
      for (int cc = 0; cc < ATT_BKV / 8; ++cc) {  // 8 keys = one 16-byte chunk of this row
        uint32_t pk[4];
#pragma unroll
        for (int q = 0; q < 4; q += 2) {
          const int e = cc * 8 + q * 2;
          float t0, t1, t2, t3;
          f2_unpack(f2_fma(f2_pack(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), c2, nmc2), t0, t1);
          f2_unpack(f2_fma(f2_pack(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])), c2, nmc2), t2, t3);
          const float p0 = ex2_approx(t0), p1 = ex2_approx(t1), p2 = ex2_approx(t2), p3 = ex2_approx(t3);
          sum_a = f2_add(sum_a, f2_pack(p0, p1));
          sum_b = f2_add(sum_b, f2_pack(p2, p3));
          pk[q] = pack2<DT>(p0, p1);
          pk[q + 1] = pack2<DT>(p2, p3);
        }
        const uint32_t blk = p_row + (cc >> 3) * 16384;  // 64-key K-major block
        st_shared_v4(blk + ((static_cast<uint32_t>(cc & 7) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
      }
      {
        float s0, s1;
        f2_unpack(f2_add(sum_a, sum_b), s0, s1);
        l += s0 + s1;
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }

    // output: O / l -> 16 bit -> staging (this tile's P buffer is free once the last P V retired) -> TMA store
    mbar_wait(&pv_done[t], (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    uint8_t* stage_out = sP + t * ATT_P_BYTES + lane_grp * 4096;
    const uint32_t o_row = smem_u32(stage_out) + lane * 128;
    const uint32_t oswz = static_cast<uint32_t>(lane & 7);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t o[32];
      tmem_ld32(tO + hh * 32, o);
      tmem_wait_ld();
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * jj + e]) * inv_l;
        const uint32_t chunk = static_cast<uint32_t>(hh * 4 + jj);
        st_shared_v4(o_row + ((chunk ^ oswz) << 4), pack2<DT>(v[0], v[1]), pack2<DT>(v[2], v[3]),
                     pack2<DT>(v[4], v[5]), pack2<DT>(v[6], v[7]));
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    const int out_row0 = q0 + t * ATT_BQ + lane_grp * 32;
    if (lane == 0 && out_row0 < S) {
      tma_store_3d(&tmO, stage_out, head * ATT_D, out_row0, seq);  // rows >= S are clipped by the tensor map
      tma_commit_group();
      tma_wait_group<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
}

int launch_attention(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads, int dtype,
                     cudaStream_t st) {
  VTQ_CHECK_ARG(ctx, qkv && out, "null pointer");
  VTQ_CHECK_ARG(ctx, n_seq >= 1 && S >= 1 && heads >= 1, "empty problem");
  VTQ_CHECK_ARG(ctx, n_seq <= 65535 && heads <= 65535, "grid limits: n_seq, heads <= 65535");
  VTQ_CHECK_ARG(ctx, dtype == VTQ_F16 || dtype == VTQ_BF16, "dtype must be VTQ_F16 or VTQ_BF16");
  VTQ_CHECK_ARG(ctx, (reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
                "pointers must be 16-byte aligned");
  const uint64_t hidden = static_cast<uint64_t>(heads) * ATT_D;
  const CUtensorMapDataType dt16 = tm_dtype16(dtype);
  CUtensorMap tmQKV, tmO;
  {
    uint64_t dims[3] = {3 * hidden, static_cast<uint64_t>(S), static_cast<uint64_t>(n_seq)};
    uint64_t strides[2] = {3 * hidden * 2, static_cast<uint64_t>(S) * 3 * hidden * 2};
    uint32_t box[3] = {ATT_D, 128, 1};
    int rc = make_tensor_map(ctx, &tmQKV, dt16, 3, qkv, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {hidden, static_cast<uint64_t>(S), static_cast<uint64_t>(n_seq)};
    uint64_t strides[2] = {hidden * 2, static_cast<uint64_t>(S) * hidden * 2};
    uint32_t box[3] = {ATT_D, 32, 1};
    int rc = make_tensor_map(ctx, &tmO, dt16, 3, out, dims, strides, box);
    if (rc) return rc;
  }
  dim3 grid((S + 2 * ATT_BQ - 1) / (2 * ATT_BQ), heads, n_seq);
  static bool configured[2] = {false, false};
  if (dtype == VTQ_F16) {
    if (!configured[0]) {
      cudaError_t e = cudaFuncSetAttribute(attention_kernel<DT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           ATT_SMEM_BYTES);
      if (e != cudaSuccess) return check_cuda(ctx, e, "attention: cudaFuncSetAttribute");
      configured[0] = true;
    }
    attention_kernel<DT_F16><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(tmQKV, tmO, S, heads);
  } else {
    if (!configured[1]) {
      cudaError_t e = cudaFuncSetAttribute(attention_kernel<DT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           ATT_SMEM_BYTES);
      if (e != cudaSuccess) return check_cuda(ctx, e, "attention: cudaFuncSetAttribute");
      configured[1] = true;
    }
    attention_kernel<DT_BF16><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(tmQKV, tmO, S, heads);
  }
  VTQ_CHECK_LAUNCH(ctx, "attention launch");
  return VTQ_OK;
}

}  // namespace vtq
