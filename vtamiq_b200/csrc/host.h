// Host-side context shared by the C-ABI entry points (api.cu) and the per-kernel launchers.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>

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
  // Per-DEVICE state (one handle per device): kernels whose dynamic shared-memory opt-in has been raised on this
  // device (cudaFuncSetAttribute is per device, not per process), and the encoded TMA descriptors keyed by their
  // full encoding arguments (base pointer, shape, strides, box, type, swizzle): steady-state launches re-use them.
  std::unordered_set<const void*> smem_configured;
  std::unordered_map<std::string, CUtensorMap> tensor_maps;
  unsigned long long tensor_map_hits = 0, tensor_map_misses = 0;
  int reverse_next = 0;           // traversal direction of the next row-/tile-walking launches (vtq_set_reverse)
  int diffnet_max_clusters = 0;   // co-resident 16-CTA clusters of the DiffNet kernel (0 = not asked yet, -1 = none)
  int* oob_flag_host = nullptr;  // pinned + mapped: gather kernels set it when a coordinate is out of range
  int* oob_flag_dev = nullptr;
};

namespace vtq {

int fail(vtq_ctx* ctx, int code, const std::string& msg);
int check_cuda(vtq_ctx* ctx, cudaError_t e, const char* what);

// Raise a kernel's dynamic shared-memory limit once per (device, kernel).
template <typename K>
int ensure_dyn_smem(vtq_ctx* ctx, K kern, int bytes, const char* what) {
  const void* key = reinterpret_cast<const void*>(kern);
  if (ctx->smem_configured.count(key)) return VTQ_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return check_cuda(ctx, e, what);
  ctx->smem_configured.insert(key);
  return VTQ_OK;
}

// Every entry point runs on the handle's device whatever the caller's current device is, and leaves the caller's
// current device untouched.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const vtq_ctx* ctx) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != ctx->device) switched = cudaSetDevice(ctx->device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define VTQ_ENTER(ctx)                     \
  if (!(ctx)) return VTQ_ERR_INVALID;      \
  ::vtq::DeviceGuard vtq_device_guard__(ctx)

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
// Slots (offsets in floats) of the activations the training forward of the tail keeps for its backward pass.
struct TailSaved {
  size_t d;             // gamma * d0                                   [B][hidden]
  size_t rcab0;         // first RCAB record; RCAB (g, r) at rcab0 + g*group_stride + r*rcab_stride:
  size_t rcab_stride;   //   y [B][hidden], sg (sigmoid gate) [B][hidden], xo (block output) [B][hidden], hc [B][ca]
  size_t group_stride;  //   ... followed, per group, by gout [B][hidden]
  size_t z;             // final conv output                           [B][hidden]   (num_rgs > 0)
  size_t u;             // head hidden pre-activation                  [B][head_hidden]
  size_t hh;            // head hidden after PReLU                     [B][head_hidden]
  size_t total;
};
int check_tail_args(vtq_ctx* ctx, const void* const* params, int n_params, int num_rgs, int num_rcabs, int hidden,
                    int ca_hidden, int head_hidden, int B);
TailSaved tail_saved_layout(int B, int num_rgs, int num_rcabs, int hidden, int ca_hidden, int head_hidden);
int launch_tail_train_forward(vtq_ctx* ctx, const float* d_scaled_in_saved, const void* const* params, int n_params,
                              int num_rgs, int num_rcabs, int hidden, int ca_hidden, int head_hidden, int B,
                              const float* drop_scale, float* saved, float* q, unsigned* counters, cudaStream_t st);
// DiffNet + head on 16-CTA clusters (diffnet_cluster.cu); launch returns 1 when clusters of that size cannot be
// scheduled on this device (the caller then uses the cooperative kernel)
bool diffnet_cluster_eligible(const vtq_ctx* ctx, int hidden, int ca_hidden, int head_hidden);
int launch_diffnet_cluster(vtq_ctx* ctx, const float* diff, const void* const* params, int num_rgs, int num_rcabs,
                           int hidden, int ca_hidden, int head_hidden, int B, float* q, cudaStream_t st);
int launch_attention(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads, int dtype,
                     int q_rows, cudaStream_t st, long long* trace = nullptr);

}  // namespace vtq
