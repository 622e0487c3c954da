// Host-side context shared by the C-ABI entry points (api.cu) and the per-kernel launchers.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <string>

#include "../../include/vtamiq_b200.h"

struct vtq_ctx {
  int device = 0;
  int num_sms = 0;
  int smem_optin = 0;
  // driver entry point, resolved at vtq_create (no link-time dependency on libcuda)
  CUresult (*encode_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                           const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
  std::string last_error;
  unsigned long long launches = 0;  // kernels launched through this handle (bench.py reports it)
};

namespace vtq {

int fail(vtq_ctx* ctx, int code, const std::string& msg);
int check_cuda(vtq_ctx* ctx, cudaError_t e, const char* what);

// dims/strides innermost-first; strides in BYTES for dims 1..rank-1; all tiles use 128B swizzle.
int make_tensor_map(vtq_ctx* ctx, CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                    bool swizzle_64b = false);  // default: 128-byte swizzle (inner box extent of 128 bytes)

inline CUtensorMapDataType tm_dtype16(int dtype) {
  return dtype == VTQ_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

#define VTQ_CHECK_ARG(ctx, cond, msg) \
  do {                                \
    if (!(cond)) return ::vtq::fail((ctx), VTQ_ERR_INVALID, std::string(__func__) + ": " + (msg)); \
  } while (0)

#define VTQ_CHECK_LAUNCH(ctx, what)                                        \
  do {                                                                     \
    (ctx)->launches++;                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) return ::vtq::check_cuda((ctx), e__, (what));  \
  } while (0)

// Launch with Programmatic Dependent Launch enabled: the kernel may become resident while its predecessor in the
// stream drains, run its setup (barrier init, TMEM allocation, descriptor prefetch) and then blocks in
// griddepcontrol.wait until the predecessor has completed and its writes are visible.  Opt-in: VTQ_PDL=1.
bool pdl_enabled();
bool l2_hints_enabled();  // VTQ_L2_HINTS=1 turns the L2 eviction-priority hints on (A/B switch, default off)

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// per-kernel launchers (defined next to their kernels)
// LayerNorm folding around a GEMM (vtq_gemm_ln): exactly one of ln_in (consume) / ln_out (produce) is set
struct GemmLnArgs {
  const float* ln_in;      // [ln_in_slots][M][2] partial (sum, sum of squares) of the A rows
  int ln_in_slots;
  const float* ln_colsum;  // [N]
  float ln_eps;
  void* raw16_out;         // [M][N]
  float* ln_out;           // [gemm_ln_slots(N)][M][2]
};
int gemm_ln_slots(int N);
int launch_gemm(vtq_ctx* ctx, const void* A, int64_t lda, const void* W, const float* bias, int M, int N, int K,
                int dtype, int epilogue, void* out, int64_t ldo, const float* gamma, cudaStream_t st,
                const GemmLnArgs* lnargs = nullptr);
int launch_attention(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads, int dtype,
                     int q_rows, cudaStream_t st, long long* trace = nullptr);

}  // namespace vtq
