// K8 (inference) — DiffNet + quality head as ONE launch of 16-CTA thread-block clusters, fp32.
//
// The decoder is a strictly sequential chain of 55 small dense layers on a [pairs][768] signal (diffnet.cu has the
// algebra).  The round-1 kernel spread every layer over ~96 CTAs and paid, per layer, a device-scope barrier through
// L2 (~2 us) plus a 98 KB re-read of the whole activation tile into every CTA: 6.8 us per layer, 0.37 ms at B = 32.
// Here one CLUSTER of 16 CTAs owns a tile of 8 pairs for the whole chain and nothing between two layers touches
// global memory:
//   * the tile's activation vector lives in every CTA's shared memory (two buffers, layer parity).  A layer's output
//     channels are split over the 16 CTAs (48 of 768); each CTA sends its [8 pairs][48] slice to all 16 CTAs with
//     cp.async.bulk shared::cta -> shared::cluster copies that complete_tx on the RECEIVER's mbarrier.  A CTA starts
//     layer l+1 when its own barrier has seen all 16 slices — there is no cluster-wide barrier on the chain (a CTA
//     can only be sending layer l+1's input after every CTA has finished computing layer l-1, whose input buffer it
//     overwrites);
//   * weights do not depend on activations: each CTA streams its 48 rows in two K halves through a 2-stage ring of
//     bulk copies (one cp.async.bulk per row, mbarrier complete_tx) that runs ahead across layer boundaries;
//   * the skip / gate operands of the RCAB and group epilogues (x, y, g) are element-wise: every thread keeps the two
//     elements it owns in registers for the whole chain;
//   * math: warp w owns 6 channels x 8 pairs (one pass over the activations serves 6 channels: the kernel is
//     shared-memory-bandwidth bound otherwise), lane i owns input columns {128 j + 4 i ..}, packed fp32x2 FMAs, and
//     two transposing shuffle butterflies leave (channel, pair) sums on single lanes.
// Different clusters take different 8-pair tiles (B = 32: four clusters; large B: as many clusters as fit, looping).
// Reference: modules/RCAN/channel_attention.py:13-86, modules/vtamiq/vtamiq.py:12-23,:71-77,:114-117.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "host.h"

namespace vtq {

constexpr int DC_CL = 16;        // CTAs per cluster (non-portable size, allowed on sm_100)
constexpr int DC_PAIRS = 8;      // pairs per cluster tile
constexpr int DC_THREADS = 256;  // 8 warps
constexpr int DC_WCH = 6;        // channels per warp
constexpr int DC_MAX_CHUNK = 8 * DC_WCH;  // a CTA owns at most 48 channels of a layer (hidden <= 768)
constexpr int DC_SLICE_PAD = 16;          // floats between two CTAs' slices of the activation vector (bank spread)
constexpr int DC_NARROW_MAX = 128;         // widest vector that may need the row re-layout (chunk % 4 != 0)
constexpr int DC_MAX_LAYERS = 64;

enum : int { CE_NONE = 0, CE_SAVE_Y = 1, CE_RELU = 2, CE_GATE = 3, CE_ADD = 4, CE_PRELU = 5, CE_OUT = 6 };

struct CLayer {
  const float* W;          // [out_dim][in_dim]
  const float* bias;       // [out_dim]
  const float* pre_param;  // PReLU slope applied to the input vector, or null
  const float* epi_param;  // PReLU slope applied to the output (CE_PRELU), or null
  int in_dim, out_dim, epi;
};
struct CLayerList {
  CLayer l[DC_MAX_LAYERS];
  int n;
};

__device__ __forceinline__ uint32_t dc_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t dc_mapa(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
// local shared memory -> another CTA's shared memory; the bytes are credited to the receiver's mbarrier
__device__ __forceinline__ void dc_bulk_to_cluster(uint32_t dst_cluster_addr, uint32_t src_addr, uint32_t bytes,
                                                   uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster_addr),
               "r"(src_addr), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}
// global -> this CTA's shared memory, one bulk copy (TMA engine); bytes are credited to a local mbarrier
__device__ __forceinline__ void dc_bulk_from_global(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dc_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ int dc_chunk(int dim) { return (dim + DC_CL - 1) / DC_CL; }

// Transposing butterfly over 32 values: afterwards lane i holds the sum over all lanes of v[i] (in v[0]).
__device__ __forceinline__ void dc_butterfly32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float s = upper ? v[i] : v[i + off];
      const float kp = upper ? v[i + off] : v[i];
      v[i] = kp + __shfl_xor_sync(0xffffffffu, s, off);
    }
  }
}
// The same for 16 values: lanes i and i + 16 both end up with the sum over all lanes of v[i].
__device__ __forceinline__ void dc_butterfly16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float s = upper ? v[i] : v[i + off];
      const float kp = upper ? v[i + off] : v[i];
      v[i] = kp + __shfl_xor_sync(0xffffffffu, s, off);
    }
  }
}

__global__ void __launch_bounds__(DC_THREADS, 1)
    diffnet_cluster_kernel(const __grid_constant__ CLayerList L, const float* __restrict__ diff, float* __restrict__ q,
                           int B, int hidden, int dbg) {
  extern __shared__ __align__(128) float dc_smem[];
  const int chunk_h = dc_chunk(hidden);
  const int xs_floats = DC_CL * (DC_PAIRS * chunk_h + DC_SLICE_PAD);  // one activation buffer (sliced layout)
  const int ws_floats = DC_MAX_CHUNK * ((hidden / 2 + 127) / 128 * 128);  // one ring stage: 48 rows x half of K
  float* xs = dc_smem;                       // [2][16 slices][8 pairs][chunk] (+ pad between slices)
  float* ws = xs + 2 * xs_floats;            // [2][rows][K half]
  float* stage = ws + 2 * ws_floats;         // [2][8 pairs][chunk]  this CTA's output slice (source of the copies)
  float* rowbuf = stage + 2 * DC_PAIRS * DC_MAX_CHUNK;  // [8 pairs][<= 128]: row layout of a narrow input vector
  uint64_t* full = reinterpret_cast<uint64_t*>(rowbuf + DC_PAIRS * DC_NARROW_MAX);  // [2] all 16 slices of a buffer landed
  const int rank = static_cast<int>(dc_cluster_rank());
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (B + DC_PAIRS - 1) / DC_PAIRS;
  const int my_pair = lane & 7;   // output elements this lane owns after the butterflies: (channel lane/8, pair) and,
  const int my_ca = lane >> 3;    // on lanes < 16, (channel 4 + lane/8, pair)

  uint64_t* wfull = full + 2;  // [2] a weight ring stage has landed
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&wfull[0], 1);
    mbar_init(&wfull[1], 1);
    fence_barrier_init();
  }
  dc_cluster_sync();  // barriers are initialised before any CTA of the cluster can send to them

  // ---- weight ring: flat sequence of (layer, K half) this CTA needs; half h covers k-iterations [h*J0, ...) ----
  auto slice_of = [&](int out_dim, int& c0, int& nch) {
    const int chunk = dc_chunk(out_dim);
    c0 = rank * chunk;
    nch = max(0, min(chunk, out_dim - c0));
  };
  auto khalf = [&](int in_dim, int h, int& k0, int& klen) {
    const int J = (in_dim + 127) >> 7, J0 = (J + 1) >> 1;
    k0 = h ? J0 * 128 : 0;
    const int k1 = h ? in_dim : min(in_dim, J0 * 128);
    klen = max(0, k1 - k0);
  };
  auto next_round = [&](int& l, int& h) -> bool {
    while (l < L.n) {
      int c0, nch, k0, klen;
      slice_of(L.l[l].out_dim, c0, nch);
      if (nch > 0) {
        while (h < 2) {
          khalf(L.l[l].in_dim, h, k0, klen);
          if (klen > 0) return true;
          ++h;
        }
      }
      ++l;
      h = 0;
    }
    return false;
  };
  // one bulk copy per weight row (a 16-byte cp.async stream tops out near 30 GB/s per SM; the copy engine does not)
  auto prefetch = [&](int l, int h, int stg) {
    if (warp != 0 || (dbg & 8)) return;
    int c0, nch, k0, klen;
    slice_of(L.l[l].out_dim, c0, nch);
    khalf(L.l[l].in_dim, h, k0, klen);
    const int in_dim = L.l[l].in_dim;
    const float* src = L.l[l].W + static_cast<size_t>(c0) * in_dim + k0;
    float* dst = ws + stg * ws_floats;
    if (lane == 0) mbar_expect_tx(&wfull[stg], static_cast<uint32_t>(nch * klen * 4));
    __syncwarp();
    for (int row = lane; row < nch; row += 32)
      dc_bulk_from_global(dst + row * klen, src + static_cast<size_t>(row) * in_dim, static_cast<uint32_t>(klen * 4),
                          &wfull[stg]);
  };

  uint32_t seq = 0;               // weight rounds consumed by this CTA (ring stage = seq & 1, phase = seq >> 1)
  uint32_t full_phase0 = 0, full_phase1 = 0;  // phases of full[0] / full[1] this CTA has consumed
  for (int tile = blockIdx.y; tile < n_tiles; tile += gridDim.y) {
    const int b0 = tile * DC_PAIRS;
    const int nb = min(DC_PAIRS, B - b0);
    // the tile's input vector into buffer 0, sliced layout [slice s][pair][chunk_h]
    {
      const int cv = chunk_h >> 2;
      for (int idx = threadIdx.x; idx < DC_CL * DC_PAIRS * cv; idx += DC_THREADS) {
        const int s = idx / (DC_PAIRS * cv), e = idx - s * (DC_PAIRS * cv);
        const int p = e / cv, c4 = e - p * cv;
        const int col = s * chunk_h + c4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < nb && col < hidden) v = __ldg(reinterpret_cast<const float4*>(diff + static_cast<size_t>(b0 + p) * hidden + col));
        *reinterpret_cast<float4*>(xs + s * (DC_PAIRS * chunk_h + DC_SLICE_PAD) + p * chunk_h + c4 * 4) = v;
      }
    }
    // element-wise operands this thread owns for the whole chain (hidden-wide layers share one channel mapping)
    float Xa = 0.f, Xb = 0.f, Ya = 0.f, Yb = 0.f, Ga = 0.f, Gb = 0.f;
    {
      int c0, nch;
      slice_of(hidden, c0, nch);
      const int cha = warp * DC_WCH + my_ca, chb = warp * DC_WCH + 4 + my_ca;
      if (cha < nch && my_pair < nb) Xa = Ga = __ldg(diff + static_cast<size_t>(b0 + my_pair) * hidden + c0 + cha);
      if (lane < 16 && chb < nch && my_pair < nb) Xb = Gb = __ldg(diff + static_cast<size_t>(b0 + my_pair) * hidden + c0 + chb);
    }
    int pl = 0, ph = 0;  // (layer, K half) of the weight round in flight
    if (next_round(pl, ph)) prefetch(pl, ph, seq & 1);
    __syncthreads();

    for (int li = 0; li < L.n; ++li) {
      const CLayer& ly = L.l[li];
      const int in_dim = ly.in_dim, out_dim = ly.out_dim;
      const int chunk_out = dc_chunk(out_dim);
      int c0, nch;
      slice_of(out_dim, c0, nch);
      const bool sends = (ly.epi != CE_OUT);
      // arm the barrier of the buffer this layer's output goes to (all 16 slices, padded to the chunk)
      if (sends && threadIdx.x == 0 && !(dbg & 2))
        mbar_expect_tx(&full[(li + 1) & 1], static_cast<uint32_t>(DC_CL * DC_PAIRS * chunk_out * 4));
      // this layer's input: buffer li & 1 (layer 0: loaded above; later layers: wait for the 16 slices)
      if (li > 0 && !(dbg & 2)) {
        if (li & 1) {
          mbar_wait(&full[1], full_phase1 & 1);
          ++full_phase1;
        } else {
          mbar_wait(&full[0], full_phase0 & 1);
          ++full_phase0;
        }
      }
      int chunk_in = dc_chunk(in_dim);
      int slice_stride = DC_PAIRS * chunk_in + DC_SLICE_PAD;
      float* cur = xs + (li & 1) * xs_floats;
      if (nch > 0) {
        if (chunk_in & 3) {  // narrow vector (the squeeze output): re-lay it out as plain rows for 128-bit reads
          for (int idx = threadIdx.x; idx < DC_PAIRS * in_dim; idx += DC_THREADS) {
            const int p = idx / in_dim, k = idx - p * in_dim;
            rowbuf[idx] = cur[(k / chunk_in) * slice_stride + p * chunk_in + (k % chunk_in)];
          }
          cur = rowbuf;
          chunk_in = in_dim;
          slice_stride = 0;
          __syncthreads();
        }
        if (ly.pre_param != nullptr) {  // PReLU on this CTA's copy of the input vector (pad floats included: harmless)
          const float a = __ldg(ly.pre_param);
          const int n4 = (slice_stride ? DC_CL * slice_stride : DC_PAIRS * in_dim) >> 2;
          for (int idx = threadIdx.x; idx < n4; idx += DC_THREADS) {
            float4 t = reinterpret_cast<float4*>(cur)[idx];
            t.x = t.x > 0.f ? t.x : a * t.x;
            t.y = t.y > 0.f ? t.y : a * t.y;
            t.z = t.z > 0.f ? t.z : a * t.z;
            t.w = t.w > 0.f ? t.w : a * t.w;
            reinterpret_cast<float4*>(cur)[idx] = t;
          }
        }
        // ---- 6 channels x 8 pairs per warp, K in two halves as the weight ring delivers them ----
        f32x2 acc[DC_WCH * DC_PAIRS];
#pragma unroll
        for (int i = 0; i < DC_WCH * DC_PAIRS; ++i) acc[i] = 0ull;
        const int ch_w = warp * DC_WCH;       // first channel (inside the CTA's slice) of this warp
        const bool active = ch_w < nch;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          int k0, klen;
          khalf(in_dim, h, k0, klen);
          if (klen <= 0) continue;
          if (!(dbg & 8)) mbar_wait(&wfull[seq & 1], (seq >> 1) & 1);  // this half's rows have landed
          __syncthreads();                             // every warp is done with the other ring stage
          const float* wst = ws + (seq & 1) * ws_floats;
          {
            int nl = li, nh = h + 1;
            if (next_round(nl, nh)) prefetch(nl, nh, (seq + 1) & 1);
          }
          ++seq;
          if (active && !(dbg & 1)) {
#pragma unroll 1
            for (int k = k0 + lane * 4; k < k0 + klen; k += 128) {
              float4 w[DC_WCH];
#pragma unroll
              for (int c = 0; c < DC_WCH; ++c) {
                const int row = min(ch_w + c, nch - 1);  // rows past the slice repeat the last one (results unused)
                w[c] = *reinterpret_cast<const float4*>(wst + row * klen + (k - k0));
              }
              const int sl = k / chunk_in;
              const float* xk = cur + sl * slice_stride + (k - sl * chunk_in);
#pragma unroll
              for (int p = 0; p < DC_PAIRS; ++p) {
                const float4 xv = *reinterpret_cast<const float4*>(xk + p * chunk_in);
                const f32x2 x01 = f2_pack(xv.x, xv.y), x23 = f2_pack(xv.z, xv.w);
#pragma unroll
                for (int c = 0; c < DC_WCH; ++c) {
                  acc[c * DC_PAIRS + p] = f2_fma(f2_pack(w[c].x, w[c].y), x01, acc[c * DC_PAIRS + p]);
                  acc[c * DC_PAIRS + p] = f2_fma(f2_pack(w[c].z, w[c].w), x23, acc[c * DC_PAIRS + p]);
                }
              }
            }
          }
        }
        float va = 0.f, vb = 0.f;
        if (active && !(dbg & 4)) {
          float r32[32], r16[16];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float lo, hi;
            f2_unpack(acc[i], lo, hi);
            r32[i] = lo + hi;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float lo, hi;
            f2_unpack(acc[32 + i], lo, hi);
            r16[i] = lo + hi;
          }
          dc_butterfly32(r32, lane);  // lane i: channel ch_w + i / 8, pair i % 8
          dc_butterfly16(r16, lane);  // lane i (and i + 16): channel ch_w + 4 + (i % 16) / 8, pair i % 8
          va = r32[0];
          vb = r16[0];
        }
        float* st = stage + (li & 1) * DC_PAIRS * DC_MAX_CHUNK;
        // fused epilogue on the (up to) two elements of this lane
        auto finish = [&](float v, int chl, float& Xr, float& Yr, float& Gr) {
          v += __ldg(ly.bias + c0 + chl);
          switch (ly.epi) {
            case CE_SAVE_Y: Yr = v; break;
            case CE_RELU: v = fmaxf(v, 0.f); break;
            case CE_GATE: {
              const float sg = 1.0f / (1.0f + expf(-v));
              v = Xr + Yr * sg;
              Xr = v;
              break;
            }
            case CE_ADD:
              v = Gr + v;
              Gr = v;
              Xr = v;
              break;
            case CE_PRELU: {
              const float a = __ldg(ly.epi_param);
              v = v > 0.f ? v : a * v;
              break;
            }
            default: break;
          }
          if (ly.epi == CE_OUT) {
            if (my_pair < nb) q[b0 + my_pair] = v;  // out_dim == 1: cluster rank 0, warp 0, lanes 0..7
          } else {
            st[my_pair * chunk_out + chl] = v;
          }
        };
        if (active) {
          const int cha = ch_w + my_ca, chb = ch_w + 4 + my_ca;
          if (cha < nch) finish(va, cha, Xa, Ya, Ga);
          if (lane < 16 && chb < nch) finish(vb, chb, Xb, Yb, Gb);
        }
      }
      if (sends) {
        // the slice is complete in `st`: hand it to the async proxy and send it to all 16 CTAs (this one included)
        fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0 && lane < DC_CL && !(dbg & 2)) {
          const uint32_t bytes = static_cast<uint32_t>(DC_PAIRS * chunk_out * 4);
          const uint32_t src = smem_u32(stage + (li & 1) * DC_PAIRS * DC_MAX_CHUNK);
          float* nxt = xs + ((li + 1) & 1) * xs_floats;
          const uint32_t dst = smem_u32(nxt + rank * (DC_PAIRS * chunk_out + DC_SLICE_PAD));
          dc_bulk_to_cluster(dc_mapa(dst, lane), src, bytes, dc_mapa(smem_u32(&full[(li + 1) & 1]), lane));
        }
      }
    }
    dc_cluster_sync();  // the next tile reuses both buffers: every CTA must be through this tile's last layer
  }
  dc_cluster_sync();    // no CTA may exit while a peer can still be sending into its shared memory
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static void cadd(CLayerList& L, const float* W, const float* bias, int in_dim, int out_dim, const float* pre_param,
                 int epi, const float* epi_param = nullptr) {
  CLayer& d = L.l[L.n++];
  d.W = W; d.bias = bias; d.pre_param = pre_param; d.epi_param = epi_param;
  d.in_dim = in_dim; d.out_dim = out_dim; d.epi = epi;
}

static int dc_smem_bytes(int hidden) {
  const int chunk_h = (hidden + DC_CL - 1) / DC_CL;
  const int xs_floats = DC_CL * (DC_PAIRS * chunk_h + DC_SLICE_PAD);
  const int ws_floats = DC_MAX_CHUNK * ((hidden / 2 + 127) / 128 * 128);
  return (2 * xs_floats + 2 * ws_floats + 2 * DC_PAIRS * DC_MAX_CHUNK + DC_PAIRS * DC_NARROW_MAX) *
             static_cast<int>(sizeof(float)) + 64;  // + 4 mbarriers
}

// true if this problem can run on the cluster kernel (otherwise the caller uses the cooperative kernel of diffnet.cu)
bool diffnet_cluster_eligible(const vtq_ctx* ctx, int hidden, int ca_hidden, int head_hidden) {
  static const bool off = [] {
    const char* e = std::getenv("VTQ_DIFFNET_COOP");
    return e && e[0] == '1';
  }();
  if (std::getenv("VTQ_DEBUG"))
    fprintf(stderr, "diffnet cluster eligibility: off %d hidden %d ca %d head %d smem %d / optin %d\n", int(off), hidden,
            ca_hidden, head_hidden, dc_smem_bytes(hidden), ctx->smem_optin);
  // Measured slower than the cooperative kernel on B200 (0.33 vs 0.30 ms at B = 32, 1.58 vs 0.73 ms at B = 256;
  // profiles/r02_diffnet_notes.md): opt-in only.
  static const bool on = [] {
    const char* e = std::getenv("VTQ_DIFFNET_CLUSTER");
    return e && e[0] == '1';
  }();
  if (off || !on) return false;
  // 128-bit rows everywhere; every layer's slice must fit one pass of 8 warps x 6 channels; hidden-wide vectors must be
  // readable with 128-bit loads in the sliced layout (chunk % 4 == 0); the squeeze vector fits the row buffer (<= 128)
  if (hidden % 64 || ca_hidden % 4 || head_hidden % 64) return false;
  if ((hidden + DC_CL - 1) / DC_CL > DC_MAX_CHUNK) return false;
  if (ca_hidden > DC_NARROW_MAX || head_hidden > hidden) return false;
  return dc_smem_bytes(hidden) <= ctx->smem_optin;
}

int launch_diffnet_cluster(vtq_ctx* ctx, const float* diff, const void* const* params, int num_rgs, int num_rcabs,
                           int hidden, int ca_hidden, int head_hidden, int B, float* q, cudaStream_t st) {
  auto P = [&](int i) { return static_cast<const float*>(params[i]); };
  CLayerList L;
  L.n = 0;
  int pi = 0;
  for (int g = 0; g < num_rgs; ++g) {
    for (int r = 0; r < num_rcabs; ++r) {
      const float *a = P(pi), *W1 = P(pi + 1), *b1 = P(pi + 2), *Wd = P(pi + 3), *bd = P(pi + 4), *Wu = P(pi + 5),
                  *bu = P(pi + 6);
      pi += 7;
      cadd(L, W1, b1, hidden, hidden, a, CE_SAVE_Y);       // y = W1 prelu(x) + b1
      cadd(L, Wd, bd, hidden, ca_hidden, nullptr, CE_RELU);  // h = relu(Wd y + bd)
      cadd(L, Wu, bu, ca_hidden, hidden, nullptr, CE_GATE);  // x' = x + y * sigmoid(Wu h + bu)
    }
    cadd(L, P(pi), P(pi + 1), hidden, hidden, nullptr, CE_ADD);  // g' = g + (Wg x + bg)
    pi += 2;
  }
  if (num_rgs > 0) cadd(L, P(pi), P(pi + 1), hidden, hidden, nullptr, CE_NONE);  // z = Wf g + bf
  pi += 2;
  cadd(L, P(pi), P(pi + 1), hidden, head_hidden, nullptr, CE_PRELU, P(pi + 2));
  cadd(L, P(pi + 3), P(pi + 4), head_hidden, 1, nullptr, CE_OUT);
  if (L.n > DC_MAX_LAYERS) return fail(ctx, VTQ_ERR_INVALID, "diffnet: too many layers for one launch");

  const int smem = dc_smem_bytes(hidden);
  const void* key = reinterpret_cast<const void*>(diffnet_cluster_kernel);
  if (!ctx->smem_configured.count(key)) {
    cudaError_t e = cudaFuncSetAttribute(diffnet_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return check_cuda(ctx, e, "diffnet cluster: cudaFuncSetAttribute(smem)");
    e = cudaFuncSetAttribute(diffnet_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return check_cuda(ctx, e, "diffnet cluster: cudaFuncSetAttribute(cluster size)");
    ctx->smem_configured.insert(key);
  }
  const int n_tiles = (B + DC_PAIRS - 1) / DC_PAIRS;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(DC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = DC_CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // as many clusters as can be co-resident (each loops over its tiles); asked once per device
  if (ctx->diffnet_max_clusters == 0) {
    cfg.gridDim = dim3(DC_CL, 1, 1);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, diffnet_cluster_kernel, &cfg);
    if (std::getenv("VTQ_DEBUG"))
      fprintf(stderr, "diffnet cluster: cudaOccupancyMaxActiveClusters -> %s, %d clusters (smem %d B)\n",
              cudaGetErrorString(e), n, smem);
    if (e != cudaSuccess || n < 1) {
      cudaGetLastError();
      ctx->diffnet_max_clusters = -1;
    } else {
      ctx->diffnet_max_clusters = n;
    }
  }
  if (ctx->diffnet_max_clusters < 1) return 1;  // caller falls back to the cooperative kernel
  const int n_clusters = n_tiles < ctx->diffnet_max_clusters ? n_tiles : ctx->diffnet_max_clusters;
  cfg.gridDim = dim3(DC_CL, n_clusters, 1);
  static const int dbg = [] {  // timing diagnostics only (results are wrong with any bit set): 1 no FMA loop,
    const char* e = std::getenv("VTQ_DC_DBG");  // 2 no slice exchange, 4 no butterflies, 8 no weight stream
    return e ? std::atoi(e) : 0;
  }();
  cudaError_t e = cudaLaunchKernelEx(&cfg, diffnet_cluster_kernel, L, diff, q, B, hidden, dbg);
  if (e != cudaSuccess) return check_cuda(ctx, e, "diffnet cluster launch");
  VTQ_CHECK_LAUNCH(ctx, "diffnet cluster launch");
  return VTQ_OK;
}

}  // namespace vtq
