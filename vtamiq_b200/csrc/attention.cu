// K6 — fused multi-head self-attention, softmax(Q K^T / 8) V, head_dim 64, no mask.
//
// Persistent: one CTA per SM loops over work items = (sequence, head, PAIR of 128-row query tiles); ref and dist
// sequences of the whole batch go through one launch (sequence index = img * B + b).  All pipelines (Q double
// buffer, K/V ring, S/P/O hand-offs) run across work-item boundaries.  The S x S score matrix lives only in TMEM /
// registers.  768 threads:
//   warp 0        TMA producer : both Q tiles once, then K/V tiles (128 keys x 64) through a 3-deep smem ring that
//                                the two query tiles share
//   warp 1, 2     MMA issuers  : one thread per query tile (t = A, B).  S_t = Q_t K^T (tcgen05.mma M128 N128 K16 x4,
//                                both operands K-major) is issued one key tile AHEAD, also across work-item
//                                boundaries; O_t += P_t V (M128 N64 K16 x8, A = P_t read straight from TENSOR
//                                MEMORY, B = V as an MN-major smem operand — consumed exactly as the QKV GEMM wrote
//                                it, no transpose pass).
//   warps 4..19   softmax      : FOUR warps per SM sub-partition.  Warpgroup (t, h) owns key half h (64 keys) of
//                                every score tile of query tile t, one query row per thread: 64 scores TMEM ->
//                                registers, row max (the two halves of a row exchange theirs through shared memory
//                                and a 256-thread named barrier, then take identical lazy-rescale decisions), exp2
//                                (packed FFMA2 + MUFU.EX2), 16-bit P packed in place and written back to tensor
//                                memory with one tcgen05.st (the K-major A operand of P V), which is only then
//                                made to wait for the previous P V.  A single warp per sub-partition cannot keep the
//                                SFU pipe (16 ex2/clk/SM, the binding unit at head_dim 64) busy through its own
//                                TMEM / barrier latencies; four free-running warps get much closer.  Lazy
//                                rescaling: O is only corrected when the max grew by > 2^8 (done by half 0).
//   warps 20..23  epilogue     : O_t / (l_0 + l_1) of a finished work item: TMEM -> registers -> 16 bit -> per-warp
//                                TMA store, while the softmax warps are already in the next work item.
// Registers are re-partitioned with setmaxnreg: 40 (producer) / 96 (softmax) / 56 (helper) = the 768 x 80 the CTA
// is launched with (setmaxnreg only moves registers inside the pool the launch allocated).
// TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512).
// Replaces modules/VisionTransformer/transformer.py:158-166 (matmul, /sqrt(d), softmax, matmul, permute copy).
#include <cstdlib>

#include "common.cuh"
#include "host.h"

namespace vtq {

int launch_attention_v3(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads, int dtype, int q_rows,
                        cudaStream_t st, long long* trace);
int launch_attention_v5(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads, int dtype, int q_rows,
                        cudaStream_t st, long long* trace);

constexpr int ATT_BQ = 128;   // query rows per tile (two tiles per CTA)
constexpr int ATT_BKV = 128;  // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 6 * 128;  // warpgroup 0: TMA + two MMA warps (+1 idle), 1..4: softmax (t, h), 5: helper
constexpr int ATT_TILE_BYTES = 128 * ATT_D * 2;  // 16 KB: a 128-row x 64 x 16-bit tile
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_OUT_BYTES = ATT_BQ * ATT_D * 2;  // 16 KB output staging per query tile (4 warps x 32 rows x 128 B)
constexpr int ATT_MAX_BYTES = 2 * 2 * 2 * ATT_BQ * 4;  // half-row maxima: [tile parity][query tile][key half][row]
constexpr int ATT_L_BYTES = 2 * 2 * 2 * ATT_BQ * 4;    // row sums:   [item parity][query tile][key half][row]
constexpr int ATT_SMEM_BYTES =
    ATT_TILE_BYTES * (4 + 2 * ATT_KV_STAGES) + 2 * ATT_OUT_BYTES + ATT_MAX_BYTES + ATT_L_BYTES + 256;
static_assert(ATT_SMEM_BYTES <= 227 * 1024, "smem budget");
// The two query tiles take turns on the SFU: the exponentials of one tile (two warps per sub-partition) saturate the
// pipe by themselves, so tile B's exponentials start when tile A's are mostly done and vice versa — each tile's TMEM
// load / row max / hand-offs then run underneath the other tile's exponentials instead of both tiles idling the
// pipe together.  A tile hands the turn over after this many of its 8 chunks (one barrier wake-up ahead of its end).
#ifndef ATT_TURN_RELEASE
#define ATT_TURN_RELEASE 5   // -1: free-running (A/B switch)
#endif
// Part of the exponentials runs on the FMA pipe instead of the SFU (16 ex2/clk/SM is the binding unit): Cody-Waite
// range reduction (round-to-nearest through the 1.5 * 2^23 magic constant) + a minimax polynomial for 2^f on
// [-0.5, 0.5] in packed fp32x2, exponent re-inserted with one integer shift-add.  Degree 3: max relative error
// 7.5e-5 (a sixth of the fp16 rounding step of P), degree 4: 2.7e-6.  ATT_POLY_MASK selects the key PAIRS that take
// this path: bit (4 * (chunk & 3) + pair) over four consecutive 8-key chunks; 0 = all on the SFU.
#ifndef ATT_POLY_MASK
#define ATT_POLY_MASK 0x0000
#endif
#ifndef ATT_POLY_DEG
#define ATT_POLY_DEG 3
#endif
#ifndef ATT_RELAXED_NS
#define ATT_RELAXED_NS 0   // > 0: TMA producer and epilogue warps sleep this long between barrier probes (A/B)
#endif
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr uint32_t ATT_TMEM_S = 0;    // + t * 128
constexpr uint32_t ATT_TMEM_O = 256;  // + t * 64
constexpr uint32_t ATT_TMEM_P = 384;  // + t * 64: P_t as the K-major A operand of P V (two 16-bit keys per column)

// Diagnostics (vtq_attention_fwd_trace): CTA 0 records clock64() at pipeline events; slot layout
// trace[role * 512 + event_index], role 0 = MMA thread of tile A, 1 = softmax (A, half 0).
#define ATT_TRACE(role, idx)                                                                     \
  do {                                                                                           \
    if constexpr (TRACE) {                                                                       \
      if (trace != nullptr && blockIdx.x == 0 && (idx) < 512) trace[(role) * 512 + (idx)] = clock64(); \
    }                                                                                            \
  } while (0)

__device__ __forceinline__ void att_wait_slack(uint64_t* bar, uint32_t parity) {
#if ATT_RELAXED_NS > 0
  mbar_wait_relaxed(bar, parity, ATT_RELAXED_NS);
#else
  mbar_wait(bar, parity);
#endif
}

// max over 32 scores (columns c0 .. c0+31 of the tile), keys >= kv_valid masked out
__device__ __forceinline__ void att_fold_max(const uint32_t (&v)[32], int c0, int kv_valid, float& a0, float& a1) {
  if (kv_valid >= c0 + 32) {
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      a0 = fmax3(a0, __uint_as_float(v[e + 0]), __uint_as_float(v[e + 1]));
      a1 = fmax3(a1, __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (c0 + e < kv_valid) a0 = fmaxf(a0, __uint_as_float(v[e]));
  }
}

// 2^x for two arguments on the FMA pipe (x <= ~8; arguments below -125 — masked keys are -inf — flush to 2^-125)
__device__ __forceinline__ void att_exp2_poly2(f32x2 x2, float& p0, float& p1) {
  float x0, x1;
  f2_unpack(x2, x0, x1);
  x2 = f2_pack(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const f32x2 magic = f2_pack(12582912.f, 12582912.f), nmagic = f2_pack(-12582912.f, -12582912.f);
  const f32x2 fa = f2_add(x2, magic);                                 // integer part in the low mantissa bits
  const f32x2 xf = f2_fma(f2_add(fa, nmagic), f2_pack(-1.f, -1.f), x2);  // x - round(x) in [-0.5, 0.5]
#if ATT_POLY_DEG == 4
  f32x2 p = f2_fma(f2_pack(0.009570097550749779f, 0.009570097550749779f), xf, f2_pack(0.05591785907745361f, 0.05591785907745361f));
  p = f2_fma(p, xf, f2_pack(0.240247443318367f, 0.240247443318367f));
  p = f2_fma(p, xf, f2_pack(0.6931217908859253f, 0.6931217908859253f));
  p = f2_fma(p, xf, f2_pack(0.9999992847442627f, 0.9999992847442627f));
#else
  f32x2 p = f2_fma(f2_pack(0.05517163500189781f, 0.05517163500189781f), xf, f2_pack(0.2426111251115799f, 0.2426111251115799f));
  p = f2_fma(p, xf, f2_pack(0.6932609677314758f, 0.6932609677314758f));
  p = f2_fma(p, xf, f2_pack(0.9999280571937561f, 0.9999280571937561f));
#endif
  float q0, q1, f0, f1;
  f2_unpack(p, q0, q1);
  f2_unpack(fa, f0, f1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(f0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(f1) << 23));
}

template <int DT, bool TRACE>
__global__ void __launch_bounds__(ATT_THREADS, 1)
    attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, int S,
                     int heads, int n_seq, int q_rows, uint64_t hint_qkv, int rev, long long* __restrict__ trace) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128B-swizzle atoms need a 1024 B aligned base
  uint8_t* sQ = smem;                                  // [2 buffers][2 tiles]
  uint8_t* sK = sQ + 4 * ATT_TILE_BYTES;               // [stages]
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;   // [stages]
  uint8_t* sO = sV + ATT_KV_STAGES * ATT_TILE_BYTES;   // [2 tiles] output staging of each work item
  float* sMax = reinterpret_cast<float*>(sO + 2 * ATT_OUT_BYTES);  // [tile parity][t][h][row]
  float* sL = sMax + ATT_MAX_BYTES / 4;                            // [item parity][t][h][row]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 2 * ATT_OUT_BYTES + ATT_MAX_BYTES + ATT_L_BYTES);
  uint64_t* q_full = bars;            // [2] Q pair of a work item landed            (TMA tx)
  uint64_t* q_empty = bars + 2;       // [2] all Q K^T of that work item retired     (2 MMA commits)
  uint64_t* kv_full = bars + 4;       // [3]
  uint64_t* kv_empty = bars + 7;      // [3]                                         (2 MMA commits)
  uint64_t* s_full = bars + 10;       // [2] S_t(n) complete                         (MMA commit)
  uint64_t* s_free = bars + 12;       // [2] S_t(n) copied to registers              (8 warp arrivals: both key halves)
  uint64_t* p_full = bars + 14;       // [2] P_t(n) in tensor memory, O_t rescaled   (8 warp arrivals)
  uint64_t* pv_done = bars + 16;      // [2] O_t += P_t(n) V complete                (MMA commit)
  uint64_t* o_free = bars + 18;       // [2] O_t of a finished work item read out    (4 warp arrivals, epilogue)
  uint64_t* xu_turn = bars + 20;      // [2] the other query tile's exponentials are mostly done (8 warp arrivals)
  uint64_t* l_full = bars + 22;       // [2] both halves' row sums of a work item in sL (8 warp arrivals)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 24);

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
      mbar_init(&s_free[t], 8);
      mbar_init(&p_full[t], 8);
      mbar_init(&pv_done[t], 1);
      mbar_init(&o_free[t], 4);
      mbar_init(&l_full[t], 8);
      mbar_init(&xu_turn[t], 8);
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      // ------------------------------- TMA producer -------------------------------
      if (lane == 0) {
        uint32_t it = 0, kvc = 0;
        for (int w0 = blockIdx.x; w0 < n_items; w0 += gridDim.x, ++it) {
          const int w = rev ? n_items - 1 - w0 : w0;   // work-item walk direction (vtq_set_reverse)
          const int qp = w % nqp;
          const int head = (w / nqp) % heads;
          const int seq = w / (nqp * heads);
          const int q0 = qp * (2 * ATT_BQ);
          const uint32_t qb = it & 1;
          att_wait_slack(&q_empty[qb], ((it >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[qb], 2 * ATT_TILE_BYTES);
          tma_load_3d_hint(sQ + (2 * qb) * ATT_TILE_BYTES, &tmQKV, &q_full[qb], head * ATT_D, q0, seq, hint_qkv);
          tma_load_3d_hint(sQ + (2 * qb + 1) * ATT_TILE_BYTES, &tmQKV, &q_full[qb], head * ATT_D, q0 + ATT_BQ, seq,
                           hint_qkv);
          for (int j = 0; j < nkv; ++j, ++kvc) {
            const uint32_t st = kvc % ATT_KV_STAGES;
            att_wait_slack(&kv_empty[st], ((kvc / ATT_KV_STAGES) & 1) ^ 1);
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
        uint32_t n_s = 0;  // score tiles issued for this query tile (global over work items)
        uint32_t n_p = 0;  // P V products issued
        int tr = 0;

        auto issue_qk = [&](uint32_t qb, uint32_t kv_idx) {
          if (n_s > 0) {  // both key halves of the previous S_t must be in registers
            mbar_wait(&s_free[mt], (n_s - 1) & 1);
            tc_fence_after();
          }
          const uint64_t dQ = umma_smem_desc(smem_u32(sQ + (2 * qb + mt) * ATT_TILE_BYTES), 16, 1024);
          const uint64_t dK = umma_smem_desc(smem_u32(sK + (kv_idx % ATT_KV_STAGES) * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_f16_ss(tmem_base + ATT_TMEM_S + mt * 128, dQ + uint64_t(k * 2), dK + uint64_t(k * 2), idesc_qk,
                        k ? 1u : 0u);
          umma_commit(&s_full[mt]);
          ++n_s;
          if (mt == 0) ATT_TRACE(0, tr++);
        };
        auto issue_pv = [&](uint32_t kv_idx, bool first) {
          // O_t (+)= P_t V: 8 k-steps of 16 keys.  P: tensor memory, 8 columns per k-step.  V: rows = keys, 128 B
          // apart, 8-key groups 1024 B apart -> one k-step advances the start address by 2048 B.
          mbar_wait(&p_full[mt], n_p & 1);
          tc_fence_after();
          const uint32_t aV = smem_u32(sV + (kv_idx % ATT_KV_STAGES) * ATT_TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < ATT_BKV / 16; ++kk) {
            const uint64_t dV = umma_smem_desc(aV + kk * 2048, 1024, 1024);
            umma_f16_ts(tmem_base + ATT_TMEM_O + mt * 64, tmem_base + ATT_TMEM_P + mt * 64 + kk * 8, dV, idesc_pv,
                        (!first || kk) ? 1u : 0u);
          }
          umma_commit(&pv_done[mt]);
          ++n_p;
          if (mt == 0) ATT_TRACE(0, tr++);
        };

        // One flat walk over this CTA's key tiles, ACROSS work items: g = global key-tile index (= K/V ring
        // counter), Q K^T always runs one tile ahead of P V — also over a work-item boundary, so the first score
        // tile of the next item (and its row maxima) are ready before the current item's last exponentials end.
        const uint32_t my_items = static_cast<uint32_t>((n_items - static_cast<int>(blockIdx.x) + gridDim.x - 1) / gridDim.x);
        const uint32_t total = my_items * static_cast<uint32_t>(nkv);
        uint32_t qk_it = 0, qk_j = 0;  // (work item, key tile) of the next Q K^T to issue
        {
          mbar_wait(&q_full[0], 0);
          mbar_wait(&kv_full[0], 0);
          tc_fence_after();
          issue_qk(0, 0);
          if (++qk_j == static_cast<uint32_t>(nkv)) {
            umma_commit(&q_empty[0]);
            qk_j = 0;
            ++qk_it;
          }
        }
        uint32_t it = 0, j = 0;  // (work item, key tile) of the P V being issued
        for (uint32_t g = 0; g < total; ++g) {
          const bool more = g + 1 < total;
          if (more) {
            const uint32_t qb = qk_it & 1;
            if (qk_j == 0) mbar_wait(&q_full[qb], (qk_it >> 1) & 1);
            mbar_wait(&kv_full[(g + 1) % ATT_KV_STAGES], ((g + 1) / ATT_KV_STAGES) & 1);
            tc_fence_after();
            issue_qk(qb, g + 1);  // next score tile runs underneath this tile's exponentials
          }
          if (j == 0 && it > 0) {            // the first P V of a work item overwrites O_t: previous O read out?
            mbar_wait(&o_free[mt], (it - 1) & 1);
            tc_fence_after();
          }
          issue_pv(g, j == 0);
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
  } else if (warp < 20) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    // ------------------------------- softmax: warpgroup (t, h) ------------------
    const int t = (warp - 4) >> 3;         // query tile
    const int h = ((warp - 4) >> 2) & 1;   // key half of every score tile
    const int lane_grp = warp & 3;         // TMEM lane quarter of this warp
    const int row = lane_grp * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
    const uint32_t tS = t_lane + ATT_TMEM_S + t * 128 + h * 64;
    const uint32_t tO = t_lane + ATT_TMEM_O + t * 64;
    const uint32_t tP = t_lane + ATT_TMEM_P + t * 64 + h * 32;
    const float c = 0.125f * 1.44269504088896340736f;  // (1/sqrt(64)) * log2(e)
    const f32x2 c2 = f2_pack(c, c);

    uint32_t n = 0;   // score tiles consumed (global over work items)
    uint32_t li = 0;  // work items finished by this CTA
    int tr = 0;
    const bool tracer = TRACE && (lane == 0) && (warp == 4);
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++li) {
      float m = -INFINITY;  // reference max (raw score domain) that P and O are currently scaled by
      float l = 0.f;        // running sum of exp over this thread's key halves
      for (int j = 0; j < nkv; ++j, ++n) {
        if (tracer) ATT_TRACE(1, tr++);  // 0: tile start
        mbar_wait(&s_full[t], n & 1);
        tc_fence_after();
        uint32_t r[64];
        tmem_ld32(tS + 0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);  // (8 warps) S_t may be overwritten by the next Q_t K^T
        if (tracer) ATT_TRACE(1, tr++);  // 1: S in registers

        const int kv_valid = S - j * ATT_BKV - h * 64;  // keys of this half tile that exist (may be <= 0)
        if (kv_valid < 64) {                            // warp-uniform: only the last key tile of a ragged sequence
#pragma unroll
          for (int e = 0; e < 64; ++e)
            if (e >= kv_valid) r[e] = 0xff800000u;  // -inf -> exp2 gives exactly 0
        }
        // row max: this thread's 64 keys, then the other key half's through shared memory (slot = tile parity:
        // the partner reads slot n before it can pass the barrier of tile n+1, and slot n is rewritten after that)
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 64; e += 8) {
          mx0 = fmax3(mx0, __uint_as_float(r[e + 0]), __uint_as_float(r[e + 1]));
          mx1 = fmax3(mx1, __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
          mx2 = fmax3(mx2, __uint_as_float(r[e + 4]), __uint_as_float(r[e + 5]));
          mx3 = fmax3(mx3, __uint_as_float(r[e + 6]), __uint_as_float(r[e + 7]));
        }
        float tile_max = fmaxf(fmax3(mx0, mx1, mx2), mx3);
        sMax[((n & 1) * 4 + t * 2 + h) * ATT_BQ + row] = tile_max;
        named_bar_sync(1 + t, 256);
        tile_max = fmaxf(tile_max, sMax[((n & 1) * 4 + t * 2 + (h ^ 1)) * ATT_BQ + row]);
        bool pv_ok = (n == 0);
        float m_new = fmaxf(m, tile_max);
        // Lazy rescale: keep the old reference max while the true max grew by < 2^8 in the exp2 domain — P then
        // stays <= 256 (exact in fp16/bf16 range, fp32 sums) and O needs no correction.  First tile: m = -inf.
        // Both key halves of a row see the same (m, tile_max) and therefore take the same decision.
        if ((m_new - m) * c <= 8.0f) m_new = m;
        if (j > 0 && __any_sync(0xffffffffu, m_new != m)) {  // rare
          const float alpha = ex2_approx((m - m_new) * c);   // exactly 1 for rows whose reference did not move
          l *= alpha;
          if (h == 0) {
            // P_t V of the previous tile must have retired before O_t is touched
            if (!pv_ok) mbar_wait(&pv_done[t], (n - 1) & 1);
            pv_ok = true;
            tc_fence_after();
#pragma unroll 1
            for (int hh = 0; hh < 4; ++hh) {
              uint32_t o[16];
              tmem_ld16(tO + hh * 16, o);
              tmem_wait_ld();
#pragma unroll
              for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
              tmem_st16(tO + hh * 16, o);
            }
            tmem_wait_st();
          }
        }
        m = m_new;
        const float nmc = -m_new * c;
        const f32x2 nmc2 = f2_pack(nmc, nmc);
#if ATT_TURN_RELEASE >= 0
        mbar_wait(&xu_turn[t], (t == 0) ? ((n & 1) ^ 1) : (n & 1));
#endif
        if (tracer) ATT_TRACE(1, tr++);  // 2: exponentials start

        // p = exp2(s*c - m*c): packed FFMA2 + MUFU.EX2, packed partial sums; the 16-bit P (two keys per word) goes
        // back into the registers of the scores it came from: chunk cc (8 keys) -> r[4cc .. 4cc+3]
        f32x2 sum_a = 0ull, sum_b = 0ull;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
#pragma unroll
          for (int q = 0; q < 4; q += 2) {
            const int e = cc * 8 + q * 2;
            const f32x2 a01 = f2_fma(f2_pack(__uint_as_float(r[e]), __uint_as_float(r[e + 1])), c2, nmc2);
            const f32x2 a23 = f2_fma(f2_pack(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3])), c2, nmc2);
            float p0, p1, p2, p3;
            if ((ATT_POLY_MASK >> (4 * (cc & 3) + q)) & 1) {
              att_exp2_poly2(a01, p0, p1);
            } else {
              float t0, t1;
              f2_unpack(a01, t0, t1);
              p0 = ex2_approx(t0), p1 = ex2_approx(t1);
            }
            if ((ATT_POLY_MASK >> (4 * (cc & 3) + q + 1)) & 1) {
              att_exp2_poly2(a23, p2, p3);
            } else {
              float t2, t3;
              f2_unpack(a23, t2, t3);
              p2 = ex2_approx(t2), p3 = ex2_approx(t3);
            }
            sum_a = f2_add(sum_a, f2_pack(p0, p1));
            sum_b = f2_add(sum_b, f2_pack(p2, p3));
            r[cc * 4 + q] = pack2<DT>(p0, p1);
            r[cc * 4 + q + 1] = pack2<DT>(p2, p3);
          }
#if ATT_TURN_RELEASE >= 0
          if (cc == ATT_TURN_RELEASE) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&xu_turn[t ^ 1]);
          }
#endif
        }
        {
          float s0, s1;
          f2_unpack(f2_add(sum_a, sum_b), s0, s1);
          l += s0 + s1;
        }
        // P_t is single-buffered: it waited in registers for the previous P_t V to retire
        if (!pv_ok) mbar_wait(&pv_done[t], (n - 1) & 1);
        tc_fence_after();
        tmem_st32(tP, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        if (tracer) ATT_TRACE(1, tr++);  // 3: P published
      }
      // hand the partial denominators to the epilogue warpgroup and move on to the next work item
      sL[((li & 1) * 4 + t * 2 + h) * ATT_BQ + row] = l;
      __syncwarp();
      if (lane == 0) mbar_arrive(&l_full[t]);
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // ------------------------------- epilogue: O_t / (l_0 + l_1) -> 16 bit -> TMA store ----
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
    const uint32_t oswz = static_cast<uint32_t>(lane & 7);
    uint32_t li = 0;
    for (int w0 = blockIdx.x; w0 < n_items; w0 += gridDim.x, ++li) {
      const int w = rev ? n_items - 1 - w0 : w0;
      const int qp = w % nqp;
      const int head = (w / nqp) % heads;
      const int seq = w / (nqp * heads);
      const int q0 = qp * (2 * ATT_BQ);
      const uint32_t n_last = (li + 1) * static_cast<uint32_t>(nkv) - 1;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        uint8_t* stage_out = sO + t * ATT_OUT_BYTES + lane_grp * 4096;  // this warp's 32 output rows
        const uint32_t o_row = smem_u32(stage_out) + lane * 128;
        // l_full first: it implies every earlier P_t V of this tile has retired, so the parity wait on pv_done
        // below can only be satisfied by the work item's LAST product
        att_wait_slack(&l_full[t], li & 1);
        const float* lp = sL + ((li & 1) * 4 + t * 2) * ATT_BQ + row;
        const float inv_l = __frcp_rn(lp[0] + lp[ATT_BQ]);
        att_wait_slack(&pv_done[t], n_last & 1);
        tc_fence_after();
        if (lane == 0) tma_wait_group_read<1>();  // the store that last used this staging block has read it
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t o[32];
          tmem_ld32(t_lane + ATT_TMEM_O + t * 64 + hh * 32, o);
          tmem_wait_ld();
          if (hh == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[t]);  // the next work item's first P V may overwrite O_t
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * jj + e]) * inv_l;
            st_shared_v4(o_row + ((static_cast<uint32_t>(hh * 4 + jj) ^ oswz) << 4), pack2<DT>(v[0], v[1]),
                         pack2<DT>(v[2], v[3]), pack2<DT>(v[4], v[5]), pack2<DT>(v[6], v[7]));
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        const int out_row0 = q0 + t * ATT_BQ + lane_grp * 32;
        if (lane == 0) {
          if (out_row0 < S) tma_store_3d(&tmO, stage_out, head * ATT_D, out_row0, seq);  // rows >= S are clipped
          tma_commit_group();
        }
      }
    }
    if (lane == 0) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
}

static int attention_variant() {  // A/B switches: VTQ_ATTN_V3=1 (round-1 kernel), VTQ_ATTN_V5=1 (intermediate)
  static const int v = [] {
    const char* e3 = std::getenv("VTQ_ATTN_V3");
    const char* e5 = std::getenv("VTQ_ATTN_V5");
    return (e3 && e3[0] == '1') ? 3 : ((e5 && e5[0] == '1') ? 5 : 0);
  }();
  return v;
}

int launch_attention(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads, int dtype,
                     int q_rows, cudaStream_t st, long long* trace) {
  if (attention_variant() == 3) return launch_attention_v3(ctx, qkv, out, n_seq, S, heads, dtype, q_rows, st, trace);
  if (attention_variant() == 5) return launch_attention_v5(ctx, qkv, out, n_seq, S, heads, dtype, q_rows, st, trace);
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
  auto go = [&](auto kern) -> int {
    if (int rc = ensure_dyn_smem(ctx, kern, ATT_SMEM_BYTES, "attention: cudaFuncSetAttribute")) return rc;
    cudaError_t le = launch_pdl(kern, grid, dim3(ATT_THREADS), ATT_SMEM_BYTES, st, tmQKV, tmO, S, heads, n_seq, q_rows,
                                hint_qkv, ctx->reverse_next, trace);
    return le != cudaSuccess ? check_cuda(ctx, le, "attention launch") : VTQ_OK;
  };
  int rc;   // the diagnostic stamps are compiled out of the production kernel (-3.4 % kernel time)
  if (dtype == VTQ_F16) rc = trace ? go(attention_kernel<DT_F16, true>) : go(attention_kernel<DT_F16, false>);
  else rc = trace ? go(attention_kernel<DT_BF16, true>) : go(attention_kernel<DT_BF16, false>);
  if (rc) return rc;
  VTQ_CHECK_LAUNCH(ctx, "attention launch");
  return VTQ_OK;
}

}  // namespace vtq
