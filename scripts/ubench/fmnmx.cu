// Micro-benchmark: throughput of the row-max pass (128 fp32 per thread): FMNMX3 vs two-input FMNMX vs integer VIMNMX,
// 1 and 2 warps per SM sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fmnmx fmnmx.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fmax2(float a, float b) { float r; asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ int imax2(int a, int b) { int r; asm volatile("max.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(const float* in, float* out, long long* cyc, int rep) {
  float r[128];
#pragma unroll
  for (int e = 0; e < 128; ++e) r[e] = in[(threadIdx.x * 131 + e * 7) & 4095];
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < rep; ++it) {
    float m0 = acc, m1 = acc, m2 = acc, m3 = acc;
    if (MODE == 0) {
#pragma unroll
      for (int e = 0; e < 128; e += 8) { m0 = fmax3(m0, r[e], r[e+1]); m1 = fmax3(m1, r[e+2], r[e+3]); m2 = fmax3(m2, r[e+4], r[e+5]); m3 = fmax3(m3, r[e+6], r[e+7]); }
    } else if (MODE == 1) {
#pragma unroll
      for (int e = 0; e < 128; e += 4) { m0 = fmax2(m0, r[e]); m1 = fmax2(m1, r[e+1]); m2 = fmax2(m2, r[e+2]); m3 = fmax2(m3, r[e+3]); }
    } else {
      int i0 = __float_as_int(m0), i1 = i0, i2 = i0, i3 = i0;
#pragma unroll
      for (int e = 0; e < 128; e += 4) { i0 = imax2(i0, __float_as_int(r[e])); i1 = imax2(i1, __float_as_int(r[e+1])); i2 = imax2(i2, __float_as_int(r[e+2])); i3 = imax2(i3, __float_as_int(r[e+3])); }
      m0 = __int_as_float(i0); m1 = __int_as_float(i1); m2 = __int_as_float(i2); m3 = __int_as_float(i3);
    }
    acc = fmax2(fmax2(m0, m1), fmax2(m2, m3)) * 0.999f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, const float* in, float* out, long long* cyc) {
  for (int threads : {128, 256}) {
    k<MODE><<<148, threads>>>(in, out, cyc, 200); k<MODE><<<148, threads>>>(in, out, cyc, 200);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (long long v : h) s += double(v);
    printf("%-28s warps/SMSP=%d  cycles per 128-value max = %.0f\n", name, threads / 128, s / 148 / 200);
  }
}
int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = float(i % 97) * 0.11f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("FMNMX3 (64 instr)", in, out, cyc);
  run<1>("FMNMX (128 instr)", in, out, cyc);
  run<2>("VIMNMX s32 (128 instr)", in, out, cyc);
  return 0;
}
