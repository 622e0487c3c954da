// C-ABI plumbing: handle lifetime, error strings, tensor-map encoding, and the thin extern "C" wrappers around
// the tcgen05 launchers.  No torch types cross this boundary (include/vtamiq_b200.h).
#include "host.h"

#include <cstdio>
#include <mutex>

namespace vtq {

static std::string g_create_error;  // error of the last failed vtq_create (no handle exists yet)

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("VTQ_PDL");
    return e != nullptr && e[0] == '1';   // measured slightly slower on this path: opt-in
  }();
  return on;
}

bool l2_hints_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("VTQ_L2_HINTS");
    return e != nullptr && e[0] == '1';   // measured no effect at cfg2: opt-in
  }();
  return on;
}

int fail(vtq_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  else g_create_error = msg;
  return code;
}

int check_cuda(vtq_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return VTQ_OK;
  return fail(ctx, VTQ_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

int make_tensor_map(vtq_ctx* ctx, CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base,
                    const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, bool swizzle_64b) {
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  // cache key = every argument of the encoding (the descriptor is a pure function of them)
  struct {
    const void* base;
    uint64_t dims[5], strides[4];
    uint32_t box[5];
    int dt, rank, swz;
  } key;
  std::memset(&key, 0, sizeof(key));
  key.base = base;
  key.dt = static_cast<int>(dt);
  key.rank = rank;
  key.swz = swizzle_64b ? 1 : 0;
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i > 0) {
      gstrides[i - 1] = strides_bytes[i - 1];
      key.strides[i - 1] = strides_bytes[i - 1];
    }
  }
  const std::string skey(reinterpret_cast<const char*>(&key), sizeof(key));
  auto hit = ctx->tensor_maps.find(skey);
  if (hit != ctx->tensor_maps.end()) {
    *out = hit->second;
    ctx->tensor_map_hits++;
    return VTQ_OK;
  }
  ctx->tensor_map_misses++;
  CUresult r = ctx->encode_tiled(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims, gstrides,
                                 gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 swizzle_64b ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu] stride0 %llu box [%u,%u,%u]",
             static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 1 ? strides_bytes[0] : 0),
             box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return fail(ctx, VTQ_ERR_CUDA, buf);
  }
  if (ctx->tensor_maps.size() >= 8192) ctx->tensor_maps.clear();  // bound the cache (workspaces come and go)
  ctx->tensor_maps.emplace(skey, *out);
  return VTQ_OK;
}

}  // namespace vtq

using namespace vtq;

extern "C" int vtq_abi_version(void) { return VTQ_ABI_VERSION; }

extern "C" int vtq_set_reverse(vtq_ctx* ctx, int reverse) {
  if (!ctx) return VTQ_ERR_INVALID;
  ctx->reverse_next = reverse ? 1 : 0;
  return VTQ_OK;
}

extern "C" int vtq_create(vtq_ctx** out, int device) {
  if (!out) return VTQ_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return fail(nullptr, VTQ_ERR_NO_DEVICE, "vtq_create: no CUDA device visible; this library has no CPU fallback");
  }
  if (device < 0 || device >= count) return fail(nullptr, VTQ_ERR_INVALID, "vtq_create: device index out of range");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return check_cuda(nullptr, e, "vtq_create");
  if (prop.major != 10) {
    return fail(nullptr, VTQ_ERR_NO_DEVICE,
                std::string("vtq_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                    std::to_string(prop.minor) + "; the kernels are compiled for sm_100a only");
  }
  int prev_device = -1;
  cudaGetDevice(&prev_device);
  if ((e = cudaSetDevice(device)) != cudaSuccess) return check_cuda(nullptr, e, "vtq_create: cudaSetDevice");
  struct Restore {  // the caller's current device is left as it was
    int d;
    ~Restore() { if (d >= 0) cudaSetDevice(d); }
  } restore{prev_device};
  vtq_ctx* ctx = new vtq_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete ctx;
    return fail(nullptr, VTQ_ERR_CUDA, "vtq_create: cuTensorMapEncodeTiled is not available from this driver");
  }
  ctx->encode_tiled = reinterpret_cast<decltype(ctx->encode_tiled)>(fn);
  // coordinate-range flag of the gather kernels: pinned, mapped host word (the host reads it without a sync)
  e = cudaHostAlloc(reinterpret_cast<void**>(&ctx->oob_flag_host), sizeof(int), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *ctx->oob_flag_host = 0;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->oob_flag_dev), ctx->oob_flag_host, 0);
  }
  if (e != cudaSuccess) {
    if (ctx->oob_flag_host) cudaFreeHost(ctx->oob_flag_host);
    delete ctx;
    return check_cuda(nullptr, e, "vtq_create: pinned flag allocation");
  }
  *out = ctx;
  return VTQ_OK;
}

extern "C" int vtq_destroy(vtq_ctx* ctx) {
  if (ctx && ctx->oob_flag_host) cudaFreeHost(ctx->oob_flag_host);
  delete ctx;
  return VTQ_OK;
}

extern "C" int vtq_coord_status(vtq_ctx* ctx, int reset) {
  if (!ctx) return VTQ_ERR_INVALID;
  const int v = *static_cast<volatile int*>(ctx->oob_flag_host);
  if (reset) *static_cast<volatile int*>(ctx->oob_flag_host) = 0;
  return v;
}

extern "C" int vtq_tensor_map_stats(const vtq_ctx* ctx, unsigned long long* hits, unsigned long long* misses) {
  if (!ctx) return VTQ_ERR_INVALID;
  if (hits) *hits = ctx->tensor_map_hits;
  if (misses) *misses = ctx->tensor_map_misses;
  return VTQ_OK;
}

extern "C" const char* vtq_last_error_string(const vtq_ctx* ctx) {
  return ctx ? ctx->last_error.c_str() : g_create_error.c_str();
}

extern "C" unsigned long long vtq_launch_count(const vtq_ctx* ctx) { return ctx ? ctx->launches : 0ull; }

extern "C" int vtq_gemm(vtq_ctx* ctx, const void* A, int64_t lda, const void* W, const float* bias, int M, int N,
                        int K, int dtype, int epilogue, void* out, int64_t ldo, const float* gamma, void* stream) {
  VTQ_ENTER(ctx);
  return launch_gemm(ctx, A, lda, W, bias, M, N, K, dtype, epilogue, out, ldo, gamma,
                     static_cast<cudaStream_t>(stream));
}

extern "C" int vtq_gemm_ln(vtq_ctx* ctx, const void* A, int64_t lda, const void* W, const float* bias, int M, int N,
                           int K, int dtype, int epilogue, void* out, int64_t ldo, const float* gamma,
                           const float* ln_in, int ln_in_slots, const float* ln_colsum, float ln_eps,
                           void* raw16_out, float* ln_out, void* stream) {
  VTQ_ENTER(ctx);
  GemmLnArgs ln = {ln_in, ln_in_slots, ln_colsum, ln_eps, raw16_out, ln_out};
  return launch_gemm(ctx, A, lda, W, bias, M, N, K, dtype, epilogue, out, ldo, gamma,
                     static_cast<cudaStream_t>(stream), &ln);
}

extern "C" int vtq_gemm_ln_slots(int N) { return N >= 64 ? gemm_ln_slots(N) : 0; }

extern "C" int vtq_attention_fwd(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads, int dtype,
                                 int q_rows, void* stream) {
  VTQ_ENTER(ctx);
  return launch_attention(ctx, qkv, out, n_seq, S, heads, dtype, q_rows, static_cast<cudaStream_t>(stream));
}

extern "C" int vtq_attention_fwd_trace(vtq_ctx* ctx, const void* qkv, void* out, int n_seq, int S, int heads,
                                       int dtype, long long* trace, void* stream) {
  VTQ_ENTER(ctx);
  return launch_attention(ctx, qkv, out, n_seq, S, heads, dtype, 0, static_cast<cudaStream_t>(stream), trace);
}
