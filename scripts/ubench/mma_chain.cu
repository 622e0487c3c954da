// Micro-benchmark: what does one tcgen05.mma cost when it accumulates into the previous instruction's tile, and does
// the tensor pipe overlap INDEPENDENT accumulation chains?  One CTA per SM, one issuing thread; every pattern is issued
// REP times back to back and closed by one commit, so the figure is cycles of tensor-pipe time per pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../vtamiq_b200/csrc -o mma_chain mma_chain.cu
#include <cstdio>
#include <cstdint>
#include "common.cuh"
using namespace vtq;

// pattern ids
enum { P_TS64_DEP = 0, P_SS64_DEP, P_SS128_DEP, P_SS256_DEP, P_TS64_IL2, P_TS64_IL4, P_TS64_BURST2, P_SS128_IL2,
       P_MIX_SERIAL, P_MIX_IL, P_SS128_IL4, P_TS64_DEP_K1, N_PAT };
const char* names[N_PAT] = {
  "8 x TS M128 N64 K16, one accumulator (P V chain)",
  "8 x SS M128 N64 K16, one accumulator",
  "4 x SS M128 N128 K16, one accumulator (Q K^T chain)",
  "4 x SS M128 N256 K16, one accumulator (GEMM k-block)",
  "8+8 TS N64, two accumulators, interleaved A1 B1 A2 B2",
  "4 x 8 TS N64, four accumulators, interleaved",
  "8+8 TS N64, two accumulators, bursts A1..A8 B1..B8",
  "4+4 SS N128, two accumulators, interleaved",
  "key-tile pair as shipped: QK_A(4) PV_B(8) QK_B(4) PV_A(8), bursts",
  "key-tile pair interleaved: PV_A PV_B QK_A PV_A PV_B QK_B ... (4 chains round robin)",
  "4 x 4 SS N128, four accumulators, interleaved",
  "8 x TS N64 into 8 different accumulators (no dependence)",
};

__global__ void __launch_bounds__(128, 1) k(int pat, int rep, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&holder);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = holder;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem);
    const uint64_t dA = umma_smem_desc(a0, 16, 1024);            // K-major 128 rows
    const uint64_t dB = umma_smem_desc(a0 + 16384, 16, 1024);    // K-major up to 256 rows (32 KB)
    const uint64_t dV = umma_smem_desc(a0 + 49152, 1024, 1024);  // MN-major V tile
    constexpr uint32_t i_ts64 = umma_idesc_f16(0, 128, 64, 0, 1);
    constexpr uint32_t i_ss64 = umma_idesc_f16(0, 128, 64, 0, 0);
    constexpr uint32_t i_ss128 = umma_idesc_f16(0, 128, 128, 0, 0);
    constexpr uint32_t i_ss256 = umma_idesc_f16(0, 128, 256, 0, 0);
    auto ts64 = [&](uint32_t d, int kk) { umma_f16_ts(tm + d, tm + 448 + (kk & 7) * 8, dV + uint64_t((kk & 7) * 128), i_ts64, 1u); };
    auto ss128 = [&](uint32_t d, int k) { umma_f16_ss(tm + d, dA + uint64_t((k & 3) * 2), dB + uint64_t((k & 3) * 2), i_ss128, 1u); };
    const long long t0 = clock64();
    for (int r = 0; r < rep; ++r) {
      switch (pat) {
        case P_TS64_DEP: for (int kk = 0; kk < 8; ++kk) ts64(256, kk); break;
        case P_SS64_DEP: for (int kk = 0; kk < 8; ++kk) umma_f16_ss(tm + 256, dA + uint64_t((kk & 3) * 2), dB + uint64_t((kk & 3) * 2), i_ss64, 1u); break;
        case P_SS128_DEP: for (int k2 = 0; k2 < 4; ++k2) ss128(0, k2); break;
        case P_SS256_DEP: for (int k2 = 0; k2 < 4; ++k2) umma_f16_ss(tm, dA + uint64_t(k2 * 2), dB + uint64_t(k2 * 2), i_ss256, 1u); break;
        case P_TS64_IL2: for (int kk = 0; kk < 8; ++kk) { ts64(256, kk); ts64(320, kk); } break;
        case P_TS64_IL4: for (int kk = 0; kk < 8; ++kk) { ts64(0, kk); ts64(64, kk); ts64(128, kk); ts64(192, kk); } break;
        case P_TS64_BURST2: for (int kk = 0; kk < 8; ++kk) ts64(256, kk); for (int kk = 0; kk < 8; ++kk) ts64(320, kk); break;
        case P_SS128_IL2: for (int k2 = 0; k2 < 4; ++k2) { ss128(0, k2); ss128(128, k2); } break;
        case P_MIX_SERIAL:
          for (int k2 = 0; k2 < 4; ++k2) ss128(0, k2);
          for (int kk = 0; kk < 8; ++kk) ts64(320, kk);
          for (int k2 = 0; k2 < 4; ++k2) ss128(128, k2);
          for (int kk = 0; kk < 8; ++kk) ts64(256, kk);
          break;
        case P_MIX_IL:
          for (int k2 = 0; k2 < 4; ++k2) { ts64(256, 2 * k2); ts64(320, 2 * k2); ss128(0, k2); ts64(256, 2 * k2 + 1); ts64(320, 2 * k2 + 1); ss128(128, k2); }
          break;
        case P_SS128_IL4: for (int k2 = 0; k2 < 4; ++k2) { ss128(0, k2); ss128(128, k2); ss128(256, k2); ss128(384, k2); } break;
        case P_TS64_DEP_K1: for (int kk = 0; kk < 8; ++kk) ts64((kk % 7) * 64, kk); break;
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  for (int pat = 0; pat < N_PAT; ++pat) {
    double res[2];
    int reps[2] = {8, 72};
    for (int i = 0; i < 2; ++i) {
      k<<<148, 128, 65536 + 1024>>>(pat, reps[i], cyc);
      k<<<148, 128, 65536 + 1024>>>(pat, reps[i], cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", names[pat], cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0; for (long long v : h) s += double(v);
      res[i] = s / 148;
    }
    // slope between the two repetition counts removes the fixed issue->commit->wake latency
    printf("%-86s cycles/pattern = %7.1f   (fixed latency %5.0f)\n", names[pat], (res[1] - res[0]) / (reps[1] - reps[0]),
           res[0] - 8 * (res[1] - res[0]) / (reps[1] - reps[0]));
  }
  return 0;
}
