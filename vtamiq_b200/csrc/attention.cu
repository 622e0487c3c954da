// K6 — fused multi-head self-attention, softmax(Q K^T / 8) V, head_dim 64, no mask.
//
// One CTA per (query tile of 128 rows, head, sequence); ref and dist sequences of the whole batch go through
// one launch (sequence index = img * B + b).  The S x S score matrix lives only in TMEM:
//   warp 0      TMA producer : Q tile once, then K/V tiles (128 keys x 64) through a 2-deep smem ring
//   warp 1      MMA issuer   : S = Q K^T  (tcgen05.mma M128 N128 K16 x4, both operands K-major),
//                              O += P V   (M128 N64 K16 x8, A = P from smem, B = V as an MN-major operand —
//                              V is consumed exactly as the QKV GEMM wrote it, no transpose pass)
//   warps 2..5  softmax      : one query row per thread; S read from TMEM, running max / sum in fp32,
//                              P rounded to 16 bits into 128B-swizzled smem, O rescaled in TMEM when the max
//                              moves, final O / l written through per-warp TMA stores.
// TMEM: S = columns [0,128), O = [128,192) (256 allocated -> two CTAs co-reside on an SM and overlap each
// other's softmax and MMA phases).
// Replaces modules/VisionTransformer/transformer.py:158-166 (matmul, /sqrt(d), softmax, matmul, permute copy).
#include "common.cuh"
#include "host.h"

namespace vtq {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 192;
constexpr int ATT_TILE_BYTES = 128 * ATT_D * 2;  // 16 KB: a 128-row x 64 x 16-bit tile
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_P_BYTES = ATT_BQ * ATT_BKV * 2;  // 32 KB
// 112 KB of tiles + 128 B of barriers: two CTAs (+1 KB system reserve each) fit the SM's 228 KB.  The dynamic
// smem base is 1024-aligned by declaration (checked at kernel entry), so no alignment slack is budgeted.
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES) + ATT_P_BYTES + 128;
static_assert(2 * (ATT_SMEM_BYTES + 1024) <= 228 * 1024, "two attention CTAs must co-reside on one SM");
constexpr uint32_t ATT_TMEM_COLS = 256;
constexpr uint32_t ATT_TMEM_S = 0;
constexpr uint32_t ATT_TMEM_O = 128;

template <int DT>
__global__ void __launch_bounds__(ATT_THREADS, 2)
    attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, int S,
                     int heads) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzle atoms need a 1024 B aligned base
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_TILE_BYTES;                   // [stages]
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;   // [stages]
  uint8_t* sP = sV + ATT_KV_STAGES * ATT_TILE_BYTES;   // P tile; reused as output staging at the end
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + ATT_P_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // [2]
  uint64_t* kv_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;      // 1: S(j) complete (and, by in-order commit, P V(j-1) complete)
  uint64_t* p_full = bars + 6;      // 1: P(j) in smem, S(j) consumed, O rescaled  (128 arrivals)
  uint64_t* o_full = bars + 7;      // 1: last P V complete
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BQ;
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
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_3d(sQ, &tmQKV, q_full, head * ATT_D, q0, seq);
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
      const uint32_t tS = tmem_base + ATT_TMEM_S;
      const uint32_t tO = tmem_base + ATT_TMEM_O;
      const uint64_t dQ = umma_smem_desc(smem_u32(sQ), 16, 1024);
      const uint32_t aP = smem_u32(sP);

      auto issue_pv = [&](int j) {
        // O (+)= P(j) V(j): 8 k-steps of 16 keys.  P: two 64-key K-major blocks of 16 KB.  V: rows = keys,
        // 128 B apart, 8-key groups 1024 B apart -> one k-step advances the start address by 2048 B.
        const uint32_t aV = smem_u32(sV + (j % ATT_KV_STAGES) * ATT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
          const uint64_t dP = umma_smem_desc(aP + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
          const uint64_t dV = umma_smem_desc(aV + kk * 2048, 1024, 1024);
          umma_f16_ss(tO, dP, dV, idesc_pv, (j | kk) ? 1u : 0u);
        }
        umma_commit(&kv_empty[j % ATT_KV_STAGES]);
      };

      mbar_wait(q_full, 0);
      for (int j = 0; j < nkv; ++j) {
        const int st = j % ATT_KV_STAGES;
        mbar_wait(&kv_full[st], (j / ATT_KV_STAGES) & 1);
        if (j > 0) {
          mbar_wait(p_full, (j - 1) & 1);
          tc_fence_after();
          issue_pv(j - 1);
        }
        tc_fence_after();
        const uint64_t dK = umma_smem_desc(smem_u32(sK + st * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) {
          umma_f16_ss(tS, dQ + uint64_t(k * 2), dK + uint64_t(k * 2), idesc_qk, k ? 1u : 0u);
        }
        umma_commit(s_full);
      }
      mbar_wait(p_full, (nkv - 1) & 1);
      tc_fence_after();
      issue_pv(nkv - 1);
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    // ------------------------------- softmax / correction / output --------------
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const uint32_t p_row = smem_u32(sP) + row * 128;
    const float c = 0.125f * 1.44269504088896340736f;  // (1/sqrt(64)) * log2(e)

    float m = -INFINITY;  // reference max of the raw (unscaled) scores that P and O are currently scaled by
    float l = 0.f;        // running sum of exp
    const f32x2 c2 = f2_pack(c, c);
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv_valid = S - j * ATT_BKV;  // keys of this tile that exist (>= 1)

      // pass 1: row maximum (3-input max; only the chunk straddling the sequence end pays for masking)
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < ATT_BKV / 32; ++cc) {
        const int rem = kv_valid - cc * 32;
        if (rem <= 0) break;
        uint32_t r[32];
        tmem_ld32(t_lane + ATT_TMEM_S + cc * 32, r);
        tmem_wait_ld();
        if (rem < 32) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (e >= rem) r[e] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          mx0 = fmax3(mx0, __uint_as_float(r[e]), __uint_as_float(r[e + 1]));
          mx1 = fmax3(mx1, __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
        }
      }
      float m_new = fmaxf(m, fmaxf(mx0, mx1));
      // Lazy rescale: keep the old reference max while the true max grew by < 2^8 in the exp2 domain — P then
      // stays <= 256 (exact in fp16/bf16 range, fp32 sums) and O needs no correction.  First tile: m = -inf.
      if ((m_new - m) * c <= 8.0f) m_new = m;
      if (j > 0 && __any_sync(0xffffffffu, m_new != m)) {
        // O correction (P V(j-1) has completed: s_full(j) was committed after it)
        const float alpha = ex2_approx((m - m_new) * c);
        l *= alpha;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[32];
          tmem_ld32(t_lane + ATT_TMEM_O + hh * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
          tmem_st32(t_lane + ATT_TMEM_O + hh * 32, r);
        }
        tmem_wait_st();
      }
      m = m_new;
      const float nmc = -m_new * c;
      const f32x2 nmc2 = f2_pack(nmc, nmc);

      // pass 2: p = exp2(s*c - m*c) (packed FFMA2 + MUFU.EX2), packed row sums, 16-bit P into swizzled smem
      f32x2 sum2 = 0ull;
#pragma unroll 1
      for (int cc = 0; cc < ATT_BKV / 32; ++cc) {
        const uint32_t blk = p_row + (cc >> 1) * 16384;
        const int rem = kv_valid - cc * 32;
        if (rem <= 0) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t chunk = static_cast<uint32_t>((cc & 1) * 4 + jj);
            st_shared_v4(blk + ((chunk ^ swz) << 4), 0u, 0u, 0u, 0u);
          }
          continue;
        }
        uint32_t r[32];
        tmem_ld32(t_lane + ATT_TMEM_S + cc * 32, r);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float t0, t1;
          f2_unpack(f2_fma(f2_pack(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), c2, nmc2), t0, t1);
          float p0 = ex2_approx(t0), p1 = ex2_approx(t1);
          if (rem < 32) {  // warp-uniform; only the straddling chunk
            if (e >= rem) p0 = 0.f;
            if (e + 1 >= rem) p1 = 0.f;
          }
          sum2 = f2_add(sum2, f2_pack(p0, p1));
          pk[e >> 1] = pack2<DT>(p0, p1);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t chunk = static_cast<uint32_t>((cc & 1) * 4 + jj);
          st_shared_v4(blk + ((chunk ^ swz) << 4), pk[4 * jj + 0], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
        }
      }
      {
        float s0, s1;
        f2_unpack(sum2, s0, s1);
        l += s0 + s1;
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // output: O / l -> 16 bit -> staging (the P buffer is free once o_full fires) -> TMA store
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    uint8_t* stage_out = sP + (warp - 2) * 4096;
    const uint32_t o_row = smem_u32(stage_out) + lane * 128;
    const uint32_t oswz = static_cast<uint32_t>(lane & 7);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t r[32];
      tmem_ld32(t_lane + ATT_TMEM_O + hh * 32, r);
      tmem_wait_ld();
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * jj + e]) * inv_l;
        const uint32_t chunk = static_cast<uint32_t>(hh * 4 + jj);
        st_shared_v4(o_row + ((chunk ^ oswz) << 4), pack2<DT>(v[0], v[1]), pack2<DT>(v[2], v[3]),
                     pack2<DT>(v[4], v[5]), pack2<DT>(v[6], v[7]));
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    const int out_row0 = q0 + lane_grp * 32;
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
  dim3 grid((S + ATT_BQ - 1) / ATT_BQ, heads, n_seq);
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
