// Micro-benchmark: what bounds the softmax exponential phase of vtq::attention_kernel on sm_100a?
// Each thread runs the per-tile body (128 keys) REP times; reports cycles per body for one warp per SM sub-partition
// (128 threads / CTA) and two (256 threads / CTA), one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o sfu_pipe sfu_pipe.cu && ./sfu_pipe
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint32_t cvt_pack(float a, float b) {
  uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r;
}
__device__ __forceinline__ uint32_t lea_pack(float a, float b) {
  uint32_t x = __float_as_uint(a) * 8u + 0x8000u, y = __float_as_uint(b) * 8u + 0x8000u;
  return __byte_perm(x, y, 0x7632);
}
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void sts4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// MODE: 0 ffma2 + ex2 + fadd2 (no pack, no store)      1 + cvt.f16x2 pack + STS (the kernel's body)
//       2 + lea/prmt pack + STS                          3 half cvt, half lea/prmt
//       4 body of 1 with 1/4 of the exps as a cubic on the FMA pipe
//       5 body of 1 with scalar FADD sums                6 body of 1 with scalar FFMA arguments
template <int MODE>
__global__ void __launch_bounds__(256, 1) body(const float* __restrict__ in, float* __restrict__ out, long long* cyc, int rep, int skew) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float r[128];
#pragma unroll
  for (int e = 0; e < 128; ++e) r[e] = in[(threadIdx.x * 131 + e * 7) & 4095];
  const float c = 0.18033688f;
  float nmc = -3.0f;
  const uint64_t c2 = pk(c, c);
  const uint32_t row = static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + threadIdx.x * 128;
  const uint32_t swz = threadIdx.x & 7;
  float acc0 = 0.f, acc1 = 0.f;
  uint64_t sa = 0, sb = 0;
  __syncthreads();
  if (threadIdx.x >= 128 && skew > 0) {  // de-phase the second warp of every sub-partition: the two warps then run
    const long long s0 = clock64();      // different parts of the unrolled body at any moment (as in the kernel)
    while (clock64() - s0 < skew) {}
  }
  const long long t0 = clock64();
  for (int it = 0; it < rep; ++it) {
    const uint64_t nmc2 = pk(nmc, nmc);
#pragma unroll
    for (int cc = 0; cc < 16; ++cc) {
      float p[8];
#pragma unroll
      for (int q = 0; q < 8; q += 2) {
        const int e = cc * 8 + q;
        float a, b;
        if (MODE == 6) { a = fmaf(r[e], c, nmc); b = fmaf(r[e + 1], c, nmc); }
        else upk(fma2(pk(r[e], r[e + 1]), c2, nmc2), a, b);
        if (MODE == 4 && (q == 6)) {
          float fa = a + 12582912.f, fb = b + 12582912.f;
          float xa = a - (fa - 12582912.f), xb = b - (fb - 12582912.f);
          float pa = fmaf(fmaf(fmaf(0.0555f, xa, 0.2402f), xa, 0.6931f), xa, 1.0f);
          float pb = fmaf(fmaf(fmaf(0.0555f, xb, 0.2402f), xb, 0.6931f), xb, 1.0f);
          p[q] = __uint_as_float(__float_as_uint(pa) + (__float_as_uint(fa) << 23));
          p[q + 1] = __uint_as_float(__float_as_uint(pb) + (__float_as_uint(fb) << 23));
        } else {
          p[q] = ex2(a); p[q + 1] = ex2(b);
        }
      }
      if (MODE == 5) {
        acc0 += (p[0] + p[1]) + (p[2] + p[3]); acc1 += (p[4] + p[5]) + (p[6] + p[7]);
      } else {
        sa = add2(sa, pk(p[0], p[1])); sb = add2(sb, pk(p[2], p[3]));
        sa = add2(sa, pk(p[4], p[5])); sb = add2(sb, pk(p[6], p[7]));
      }
      if (MODE == 0) continue;
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool lea = (MODE == 2 || (MODE == 3 && (q & 1)));
        w[q] = lea ? lea_pack(p[2 * q], p[2 * q + 1]) : cvt_pack(p[2 * q], p[2 * q + 1]);
      }
      sts4(row + ((static_cast<uint32_t>(cc & 7) ^ swz) << 4) + (cc >> 3) * 32768, w[0], w[1], w[2], w[3]);
    }
    nmc -= 1e-3f;   // loop-carried: every exp2 argument changes each iteration, nothing can be hoisted
  }
  const long long t1 = clock64();
  float s0, s1; upk(add2(sa, sb), s0, s1);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + s0 + s1;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, const float* in, float* out, long long* cyc, int skew = 0) {
  const int rep = 200;
  cudaFuncSetAttribute(body<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  for (int threads : {128, 256}) {
    body<MODE><<<148, threads, 65536 + 1024>>>(in, out, cyc, rep, skew);
    body<MODE><<<148, threads, 65536 + 1024>>>(in, out, cyc, rep, skew);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (long long v : h) s += double(v);
    printf("%-44s warps/SMSP=%d skew=%4d  cycles per 128-key body = %.0f\n", name, threads / 128, skew, s / 148 / rep);
  }
}

int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = -float(i % 97) * 0.11f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("0 ffma2 + ex2 + fadd2", in, out, cyc);
  run<1>("1 ffma2 + ex2 + fadd2 + cvt + STS", in, out, cyc);
  run<2>("2 ffma2 + ex2 + fadd2 + lea/prmt + STS", in, out, cyc);
  run<3>("3 half cvt, half lea/prmt", in, out, cyc);
  run<4>("4 as 1, 1/4 of exps as FMA cubic", in, out, cyc);
  run<5>("5 as 1, scalar FADD sums", in, out, cyc);
  run<6>("6 as 1, scalar FFMA arguments", in, out, cyc);
  for (int skew : {150, 300, 600, 900, 1100}) run<1>("1 with the second warp de-phased", in, out, cyc, skew);
  for (int skew : {300, 600}) run<0>("0 with the second warp de-phased", in, out, cyc, skew);
  return 0;
}
