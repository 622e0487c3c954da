// K6 (round-1 kernel, kept for A/B runs behind VTQ_ATTN_V3=1) — fused multi-head self-attention, softmax(Q K^T / 8) V, head_dim 64, no mask.
//
// Persistent: one CTA per SM loops over work items = (sequence, head, PAIR of 128-row query tiles); ref and dist
// sequences of the whole batch go through one launch (sequence index = img * B + b).  All pipelines (Q double
// buffer, K/V ring, S/P/O hand-offs) run across work-item boundaries, so loads and Q K^T of the next item overlap
// the tail of the current one.  The S x S score matrix lives only in TMEM / registers:
//   warp 0      TMA producer : both Q tiles once, then K/V tiles (128 keys x 64) through a 3-deep smem ring that
//                              the two query tiles share
//   warp 1, 2   MMA issuers  : one thread per query tile (t = A, B): the MMAs of a tile form a dependent chain
//                              (each accumulates into the previous one's tile, ~125 cycles per instruction at these
//                              sizes), so a single in-order issuer serialised both tiles' chains; S_t = Q_t K^T is
//                              issued one key tile AHEAD, also across work-item boundaries.
//                              S_t = Q_t K^T (tcgen05.mma M128 N128 K16 x4, both operands K-major) and
//                              O_t += P_t V (M128 N64 K16 x8, A = P_t read straight from TENSOR MEMORY, B = V as an
//                              MN-major smem operand — V is consumed exactly as the QKV GEMM wrote it, no transpose
//                              pass)
//   warps 4..7  softmax A    : one query row per thread.  The whole 128-key score row is pulled from TMEM into
//   warps 8..11 softmax B      registers ONCE and the S buffer is released immediately, so Q_t K^T of the next key
//                              tile runs underneath this tile's exponentials; running max / sum in fp32, lazy
//                              rescaling (O is only corrected when the max grows by > 2^8), P rounded to 16 bits
//                              and written back to TMEM with tcgen05.st (row = lane, two keys per column: the
//                              K-major A layout of tcgen05.mma), so P never touches shared memory; final O / l
//                              through per-warp TMA stores.
// Registers are re-partitioned with setmaxnreg: the producer warpgroup drops to 56, the softmax warpgroups grow to
// 224 (56*128 + 224*256 = 168*384, the launch allocation) so a whole score row (128 fp32) fits.
// TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512).
// Replaces modules/VisionTransformer/transformer.py:158-166 (matmul, /sqrt(d), softmax, matmul, permute copy).
#include <cstdlib>

#include "common.cuh"
#include "host.h"

namespace vtq {
namespace v3 {

constexpr int ATT_BQ = 128;   // query rows per tile (two tiles per CTA)
constexpr int ATT_BKV = 128;  // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 3 * 128;  // warpgroup 0: TMA + two MMA warps (+1 idle), warpgroups 1, 2: softmax A, B
constexpr int ATT_TILE_BYTES = 128 * ATT_D * 2;  // 16 KB: a 128-row x 64 x 16-bit tile
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_OUT_BYTES = ATT_BQ * ATT_D * 2;  // 16 KB output staging per query tile (4 warps x 32 rows x 128 B)
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES * (4 + 2 * ATT_KV_STAGES) + 2 * ATT_OUT_BYTES + 256;
static_assert(ATT_SMEM_BYTES <= 227 * 1024, "smem budget");
constexpr int ATT_XU_RELEASE_CHUNK = 11;  // of 16 eight-key chunks per row
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_TMEM_S = 0;    // + t * 128
constexpr uint32_t ATT_TMEM_O = 256;  // + t * 64
constexpr uint32_t ATT_TMEM_P = 384;  // + t * 64: P_t as the K-major A operand of P V (two 16-bit keys per column)

// Diagnostics (vtq_attention_fwd_trace): CTA 0 records clock64() at pipeline events; slot layout
// trace[role * 512 + event_index], role 0 = MMA thread of tile A, 1 = softmax A (warp 4 lane 0), 2 = softmax B.
#define ATT_TRACE(role, idx)                                                                     \
  do {                                                                                           \
    if (trace != nullptr && blockIdx.x == 0 && (idx) < 512) trace[(role) * 512 + (idx)] = clock64(); \
  } while (0)

template <int DT>
__global__ void __launch_bounds__(ATT_THREADS, 1)
    attention_v3_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, int S,
                     int heads, int n_seq, int q_rows, uint64_t hint_qkv, long long* __restrict__ trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzle atoms need a 1024 B aligned base
  uint8_t* sQ = smem;                                  // [2 buffers][2 tiles]
  uint8_t* sK = sQ + 4 * ATT_TILE_BYTES;               // [stages]
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;   // [stages]
  uint8_t* sO = sV + ATT_KV_STAGES * ATT_TILE_BYTES;   // [2 tiles] output staging of each work item
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * ATT_OUT_BYTES);
  uint64_t* q_full = bars;            // [2] Q pair of a work item landed            (TMA tx)
  uint64_t* q_empty = bars + 2;       // [2] all Q K^T of that work item retired     (2 MMA commits)
  uint64_t* kv_full = bars + 4;       // [3]
  uint64_t* kv_empty = bars + 7;      // [3]                                         (2 MMA commits)
  uint64_t* s_full = bars + 10;       // [2] S_t(n) complete                         (MMA commit)
  uint64_t* s_free = bars + 12;       // [2] S_t(n) copied to registers              (128 arrivals)
  uint64_t* p_full = bars + 14;       // [2] P_t(n) in tensor memory, O_t rescaled   (128 arrivals)
  uint64_t* pv_done = bars + 16;      // [2] O_t += P_t(n) V complete                (MMA commit)
  uint64_t* o_free = bars + 18;       // [2] O_t of a finished work item read out    (128 arrivals)
  uint64_t* xu_turn = bars + 20;      // [2] exponential phases of the two groups alternate (4 warp arrivals)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int hidden = heads * ATT_D;
  const int nkv = (S + ATT_BKV - 1) / ATT_BKV;
  const int nqp = (q_rows + 2 * ATT_BQ - 1) / (2 * ATT_BQ);  // query-tile pairs per (sequence, head)
  const int n_items = nqp * heads * n_seq;               // work item = (seq, head, query pair), pair fastest

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], 2);
    }
    for (int s = 0; s < ATT_KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 128);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_done[t], 1);
      mbar_init(&o_free[t], 128);
      mbar_init(&xu_turn[t], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_launch_dependents();
  pdl_wait();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ------------------------------- TMA producer -------------------------------
      if (lane == 0) {
        uint32_t it = 0, kvc = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
          const int qp = w % nqp;
          const int head = (w / nqp) % heads;
          const int seq = w / (nqp * heads);
          const int q0 = qp * (2 * ATT_BQ);
          const uint32_t qb = it & 1;
          mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[qb], 2 * ATT_TILE_BYTES);
          tma_load_3d_hint(sQ + (2 * qb) * ATT_TILE_BYTES, &tmQKV, &q_full[qb], head * ATT_D, q0, seq, hint_qkv);
          tma_load_3d_hint(sQ + (2 * qb + 1) * ATT_TILE_BYTES, &tmQKV, &q_full[qb], head * ATT_D, q0 + ATT_BQ, seq,
                           hint_qkv);
          for (int j = 0; j < nkv; ++j, ++kvc) {
            const uint32_t st = kvc % ATT_KV_STAGES;
            mbar_wait(&kv_empty[st], ((kvc / ATT_KV_STAGES) & 1) ^ 1);
            mbar_expect_tx(&kv_full[st], 2 * ATT_TILE_BYTES);
            tma_load_3d_hint(sK + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], hidden + head * ATT_D, j * ATT_BKV, seq,
                             hint_qkv);
            tma_load_3d_hint(sV + st * ATT_TILE_BYTES, &tmQKV, &kv_full[st], 2 * hidden + head * ATT_D, j * ATT_BKV,
                             seq, hint_qkv);
          }
        }
      }
      __syncwarp();
    } else if (warp <= 2) {
      // ------------------------------- MMA issuers: warp 1 -> query tile A, warp 2 -> query tile B ------------
      if (lane == 0) {
        const int mt = warp - 1;
        constexpr uint32_t idesc_qk = umma_idesc_f16(DT, ATT_BQ, ATT_BKV, 0, 0);
        constexpr uint32_t idesc_pv = umma_idesc_f16(DT, ATT_BQ, ATT_D, 0, 1);  // B (=V) is MN-major
        uint32_t n_s[2] = {0, 0};  // score tiles issued per query tile (global over work items)
        uint32_t n_p[2] = {0, 0};  // P V products issued per query tile
        int tr = 0;

        auto issue_qk = [&](int t, uint32_t qb, uint32_t kv_idx) {
          if (n_s[t] > 0) {  // the softmax group must have pulled the previous S_t into registers
            mbar_wait(&s_free[t], (n_s[t] - 1) & 1);
            tc_fence_after();
          }
          const uint64_t dQ = umma_smem_desc(smem_u32(sQ + (2 * qb + t) * ATT_TILE_BYTES), 16, 1024);
          const uint64_t dK = umma_smem_desc(smem_u32(sK + (kv_idx % ATT_KV_STAGES) * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_f16_ss(tmem_base + ATT_TMEM_S + t * 128, dQ + uint64_t(k * 2), dK + uint64_t(k * 2), idesc_qk,
                        k ? 1u : 0u);
          umma_commit(&s_full[t]);
          ++n_s[t];
          if (mt == 0) ATT_TRACE(0, tr++);
        };
        auto issue_pv = [&](int t, uint32_t kv_idx, bool first) {
          // O_t (+)= P_t V: 8 k-steps of 16 keys.  P: tensor memory, 8 columns per k-step.  V: rows = keys, 128 B
          // apart, 8-key groups 1024 B apart -> one k-step advances the start address by 2048 B.
          mbar_wait(&p_full[t], n_p[t] & 1);
          tc_fence_after();
          const uint32_t aV = smem_u32(sV + (kv_idx % ATT_KV_STAGES) * ATT_TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
            const uint64_t dV = umma_smem_desc(aV + kk * 2048, 1024, 1024);
            umma_f16_ts(tmem_base + ATT_TMEM_O + t * 64, tmem_base + ATT_TMEM_P + t * 64 + kk * 8, dV, idesc_pv,
                        (!first || kk) ? 1u : 0u);
          }
          umma_commit(&pv_done[t]);
          ++n_p[t];
          if (mt == 0) ATT_TRACE(0, tr++);
        };

        // One flat walk over this CTA's key tiles, ACROSS work items: g = global key-tile index (= K/V ring
        // counter), Q K^T always runs one tile ahead of P V — also over a work-item boundary, so the first score
        // tile of the next item is computed underneath the last exponentials / the output of the current one.
        const uint32_t my_items = static_cast<uint32_t>((n_items - static_cast<int>(blockIdx.x) + gridDim.x - 1) / gridDim.x);
        const uint32_t total = my_items * static_cast<uint32_t>(nkv);
        uint32_t qk_it = 0, qk_j = 0;  // (work item, key tile) of the next Q K^T pair to issue
        auto issue_qk_pair = [&](uint32_t g) {
          const uint32_t qb = qk_it & 1;
          if (qk_j == 0) mbar_wait(&q_full[qb], (qk_it >> 1) & 1);
          mbar_wait(&kv_full[g % ATT_KV_STAGES], (g / ATT_KV_STAGES) & 1);
          tc_fence_after();
          issue_qk(mt, qb, g);
          if (++qk_j == static_cast<uint32_t>(nkv)) {
            umma_commit(&q_empty[qb]);  // every Q K^T of this work item has been issued
            qk_j = 0;
            ++qk_it;
          }
        };
        issue_qk_pair(0);
        uint32_t it = 0, j = 0;  // (work item, key tile) of the P V being issued
        for (uint32_t g = 0; g < total; ++g) {
          const bool more = g + 1 < total;
          if (more) {
            const uint32_t qb = qk_it & 1;
            if (qk_j == 0) mbar_wait(&q_full[qb], (qk_it >> 1) & 1);
            mbar_wait(&kv_full[(g + 1) % ATT_KV_STAGES], ((g + 1) / ATT_KV_STAGES) & 1);
            tc_fence_after();
            issue_qk(mt, qk_it & 1, g + 1);  // next score tile runs underneath this tile's exponentials
          }
          if (j == 0 && it > 0) {            // the first P V of a work item overwrites O_t: previous O read out?
            mbar_wait(&o_free[mt], (it - 1) & 1);
            tc_fence_after();
          }
          issue_pv(mt, g, j == 0);
          if (more && ++qk_j == static_cast<uint32_t>(nkv)) {
            umma_commit(&q_empty[qk_it & 1]);  // every Q K^T of that work item (this tile) has been issued
            qk_j = 0;
            ++qk_it;
          }
          umma_commit(&kv_empty[g % ATT_KV_STAGES]);  // K(g), V(g) free once both issuers' MMAs on them retire
          if (++j == static_cast<uint32_t>(nkv)) {
            j = 0;
            ++it;
          }
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------- softmax / correction / output --------------
    const int t = (warp - 4) >> 2;         // query tile of this warpgroup
    const int lane_grp = warp & 3;         // TMEM lane quarter of this warp
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
    const uint32_t tS = t_lane + ATT_TMEM_S + t * 128;
    const uint32_t tO = t_lane + ATT_TMEM_O + t * 64;
    const uint32_t tP = t_lane + ATT_TMEM_P + t * 64;
    uint8_t* stage_out = sO + t * ATT_OUT_BYTES + lane_grp * 4096;  // this warp's 32 output rows
    const uint32_t o_row = smem_u32(stage_out) + lane * 128;
    const uint32_t oswz = static_cast<uint32_t>(lane & 7);
    const float c = 0.125f * 1.44269504088896340736f;  // (1/sqrt(64)) * log2(e)
    const f32x2 c2 = f2_pack(c, c);

    uint32_t n = 0;  // score tiles consumed by this warpgroup (global over work items)
    int tr = 0;
    const bool tracer = (lane == 0) && (lane_grp == 0);
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int qp = w % nqp;
      const int head = (w / nqp) % heads;
      const int seq = w / (nqp * heads);
      const int q0 = qp * (2 * ATT_BQ);
      float m = -INFINITY;  // reference max (raw score domain) that P and O are currently scaled by
      float l = 0.f;        // running sum of exp
      bool s_ok = false;    // early probe of the next S tile (issued before the exponentials of the current one)
      for (int j = 0; j < nkv; ++j, ++n) {
        if (tracer) ATT_TRACE(1 + t, tr++);  // 0: start waiting for S
        if (!s_ok) mbar_wait(&s_full[t], n & 1);
        tc_fence_after();
        if (tracer) ATT_TRACE(1 + t, tr++);  // 1: S ready
        uint32_t r[128];
        tmem_ld32(tS + 0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_ld32(tS + 64, *reinterpret_cast<uint32_t(*)[32]>(&r[64]));
        tmem_ld32(tS + 96, *reinterpret_cast<uint32_t(*)[32]>(&r[96]));
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&s_free[t]);  // S_t may be overwritten by the next Q_t K^T from here on
        if (tracer) ATT_TRACE(1 + t, tr++);  // 2: S in registers
        // probe the two hand-offs needed before the exponentials now; their round trips hide under the row max
        const uint32_t turn_parity = (t == 0) ? ((n & 1) ^ 1) : (n & 1);
        const bool pv_ok = (j == 0) || mbar_test(&pv_done[t], (n - 1) & 1);
        const bool turn_ok = mbar_test(&xu_turn[t], turn_parity);

        const int kv_valid = S - j * ATT_BKV;  // keys of this tile that exist (>= 1)
        if (kv_valid < ATT_BKV) {              // CTA-uniform: only the last key tile of a ragged sequence
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            if (kv_valid < (ch + 1) * 32) {    // uniform: chunks entirely inside the sequence are skipped
#pragma unroll
              for (int e = ch * 32; e < ch * 32 + 32; ++e)
                if (e >= kv_valid) r[e] = 0xff800000u;  // -inf -> exp2 gives exactly 0
            }
          }
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

        if (tracer) ATT_TRACE(1 + t, tr++);  // 3: row max done
        if (j > 0) {
          // P_t V of the previous tile must have retired before P_t is overwritten or O_t is touched
          if (!pv_ok) mbar_wait(&pv_done[t], (n - 1) & 1);
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
        if (tracer) ATT_TRACE(1 + t, tr++);  // 4: previous P V retired (+ rescale)
        m = m_new;
        const float nmc = -m_new * c;
        const f32x2 nmc2 = f2_pack(nmc, nmc);

        // The two softmax groups take turns on the SFU (16 ex2/clk/SM is the binding pipe): while one group
        // runs its 128 exponentials per thread, the other does its latency-bound part (S load, max, hand-offs).
        if (!turn_ok) mbar_wait(&xu_turn[t], turn_parity);
        if (tracer) ATT_TRACE(1 + t, tr++);  // 5: SFU turn acquired
        s_ok = (j + 1 < nkv) && mbar_test(&s_full[t], (n + 1) & 1);  // consumed at the top of the next tile
        // p = exp2(s*c - m*c): packed FFMA2 + MUFU.EX2, packed partial sums, 16-bit P into tensor memory
        f32x2 sum_a = 0ull, sum_b = 0ull;
        uint32_t pq[8];  // 16 keys of P, packed; stored to TMEM as one 8-column piece (= one k-step of P V)
#pragma unroll
        for (int cc = 0; cc < ATT_BKV / 8; ++cc) {  // 8 keys per step
#pragma unroll
          for (int q = 0; q < 4; q += 2) {
            const int e = cc * 8 + q * 2;
            float t0, t1, t2, t3;
            f2_unpack(f2_fma(f2_pack(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), c2, nmc2), t0, t1);
            f2_unpack(f2_fma(f2_pack(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])), c2, nmc2), t2, t3);
            const float p0 = ex2_approx(t0), p1 = ex2_approx(t1), p2 = ex2_approx(t2), p3 = ex2_approx(t3);
            sum_a = f2_add(sum_a, f2_pack(p0, p1));
            sum_b = f2_add(sum_b, f2_pack(p2, p3));
            pq[(cc & 1) * 4 + q] = pack2<DT>(p0, p1);
            pq[(cc & 1) * 4 + q + 1] = pack2<DT>(p2, p3);
          }
          if (cc & 1) tmem_st8(tP + (cc >> 1) * 8, pq);
          if (cc == ATT_XU_RELEASE_CHUNK) {
            // hand the SFU to the other group about one barrier wake-up latency before this group's last
            // exponentials issue
            __syncwarp();
            if (lane == 0) mbar_arrive(&xu_turn[t ^ 1]);
          }
        }
        {
          float s0, s1;
          f2_unpack(f2_add(sum_a, sum_b), s0, s1);
          l += s0 + s1;
        }
        tmem_wait_st();  // P_t is in tensor memory
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        if (tracer) ATT_TRACE(1 + t, tr++);  // 6: P published
      }

      // output: O / l -> 16 bit -> this warp's staging rows -> TMA store
      if (lane == 0) tma_wait_group_read<0>();  // the previous work item's store has finished reading them
      __syncwarp();
      mbar_wait(&pv_done[t], (n - 1) & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld32(tO, o0);
      tmem_ld32(tO + 32, o1);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&o_free[t]);  // the next work item's first P V may overwrite O_t
      const float inv_l = __frcp_rn(l);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float v[8], u[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          v[e] = __uint_as_float(o0[8 * jj + e]) * inv_l;
          u[e] = __uint_as_float(o1[8 * jj + e]) * inv_l;
        }
        st_shared_v4(o_row + ((static_cast<uint32_t>(jj) ^ oswz) << 4), pack2<DT>(v[0], v[1]), pack2<DT>(v[2], v[3]),
                     pack2<DT>(v[4], v[5]), pack2<DT>(v[6], v[7]));
        st_shared_v4(o_row + ((static_cast<uint32_t>(4 + jj) ^ oswz) << 4), pack2<DT>(u[0], u[1]),
                     pack2<DT>(u[2], u[3]), pack2<DT>(u[4], u[5]), pack2<DT>(u[6], u[7]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      const int out_row0 = q0 + t * ATT_BQ + lane_grp * 32;
      if (lane == 0 && out_row0 < S) {
        tma_store_3d(&tmO, stage_out, head * ATT_D, out_row0, seq);  // rows >= S are clipped by the tensor map
        tma_commit_group();
      }
      if (tracer) ATT_TRACE(1 + t, tr++);  // 7: work item output issued
    }
    if (lane == 0) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
}

}  // namespace v3

using namespace v3;

int launch_attention_v3(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads, int dtype,
                     int q_rows, cudaStream_t st, long long* trace) {
  VTQ_CHECK_ARG(ctx, qkv && out, "null pointer");
  VTQ_CHECK_ARG(ctx, n_seq >= 1 && S >= 1 && heads >= 1, "empty problem");
  VTQ_CHECK_ARG(ctx, q_rows >= 0 && q_rows <= S, "q_rows must be in [0, S]");
  if (q_rows == 0) q_rows = S;
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
  const long long n_items = static_cast<long long>((q_rows + 2 * ATT_BQ - 1) / (2 * ATT_BQ)) * heads * n_seq;
  VTQ_CHECK_ARG(ctx, n_items < (1ll << 30), "too many work items");
  dim3 grid(static_cast<unsigned>(n_items < ctx->num_sms ? n_items : ctx->num_sms));
  // q|k|v rows are dead after this kernel: let them leave L2 first (keeps the residual stream resident)
  const uint64_t hint_qkv = l2_hints_enabled() ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
  if (dtype == VTQ_F16) {
    if (int rc = ensure_dyn_smem(ctx, attention_v3_kernel<DT_F16>, ATT_SMEM_BYTES, "attention: cudaFuncSetAttribute")) return rc;
    cudaError_t le = launch_pdl(attention_v3_kernel<DT_F16>, grid, dim3(ATT_THREADS), ATT_SMEM_BYTES, st, tmQKV, tmO, S,
                                heads, n_seq, q_rows, hint_qkv, trace);
    if (le != cudaSuccess) return check_cuda(ctx, le, "attention launch");
  } else {
    if (int rc = ensure_dyn_smem(ctx, attention_v3_kernel<DT_BF16>, ATT_SMEM_BYTES, "attention: cudaFuncSetAttribute")) return rc;
    cudaError_t le = launch_pdl(attention_v3_kernel<DT_BF16>, grid, dim3(ATT_THREADS), ATT_SMEM_BYTES, st, tmQKV, tmO, S,
                                heads, n_seq, q_rows, hint_qkv, trace);
    if (le != cudaSuccess) return check_cuda(ctx, le, "attention launch");
  }
  VTQ_CHECK_LAUNCH(ctx, "attention launch");
  return VTQ_OK;
}

}  // namespace vtq
