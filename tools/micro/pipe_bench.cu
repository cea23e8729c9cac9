// Micro-benchmark: issue rates of the FMA-pipe instruction forms on sm_100a, alone and mixed - the depthwise
// time-convolution epilogues of the conv GEMMs (gemm_convt.cuh) are bound by exactly these.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe tools/micro/pipe_bench.cu && /tmp/pipe
// Modes: 0 FFMA | 1 FFMA2 (fma.rn.f32x2) | 2 HFMA2 f16x2 | 3 HFMA2 bf16x2 | 4 FFMA2+FFMA 1:1 | 5 HFMA2+FFMA 1:1 |
//        6 HFMA2+FFMA2 1:1 | 7 FFMA2+FFMA 1:2 | 8 HFMA2 f16x2 with fp32-pair conversion (cvt.rn.f16x2.f32) per 4
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r)
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return reinterpret_cast<float2&>(r);
}
__device__ __forceinline__ unsigned hfma2(unsigned a, unsigned b, unsigned c) {
  unsigned r;
  asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ unsigned bfma2(unsigned a, unsigned b, unsigned c) {
  unsigned r;
  asm volatile("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
  float r;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ unsigned cvt_f16x2(float lo, float hi) {
  unsigned r;
  asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// fmas[] = FMAs per thread and iteration of each mode
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float s, int iters) {
  float2 a[8];
  float f[16];
  unsigned h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    h[i] = 0x3c003c00u + threadIdx.x + i;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 0.002f + i;
  const float2 w = make_float2(s, s * 0.5f);
  const unsigned hw = 0x3c013c01u, hc = 0x00010001u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { f[i] = ffma(f[i], w.x, w.y); f[8 + i] = ffma(f[8 + i], w.x, w.y); }
      if (MODE == 1) a[i] = fma2(a[i], w, w);
      if (MODE == 2) h[i] = hfma2(h[i], hw, hc);
      if (MODE == 3) h[i] = bfma2(h[i], hw, hc);
      if (MODE == 4) { a[i] = fma2(a[i], w, w); f[i] = ffma(f[i], w.x, w.y); }
      if (MODE == 5) { h[i] = hfma2(h[i], hw, hc); f[i] = ffma(f[i], w.x, w.y); }
      if (MODE == 6) { h[i] = hfma2(h[i], hw, hc); a[i] = fma2(a[i], w, w); }
      if (MODE == 7) { a[i] = fma2(a[i], w, w); f[i] = ffma(f[i], w.x, w.y); f[8 + i] = ffma(f[8 + i], w.x, w.y); }
      if (MODE == 8) {
        h[i] = hfma2(h[i], hw, hc);
        if ((i & 3) == 0) h[i] ^= cvt_f16x2(f[i], f[i + 1]);
      }
    }
  }
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + __uint_as_float(h[i]) + f[i] + f[8 + i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
void run(float* out, const char* name, double fma_per_iter, double instr_per_iter, int sms, double mhz) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<sms * 8, 256>>>(out, 1.0001f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double thr = double(sms) * 8 * 256 * iters;
  printf("%-28s %8.3f ms  %7.2f TFMA/s  %6.2f Tinstr/s (thread)  %5.1f FMA/clk/SM @%.0f MHz nominal\n", name, ms,
         thr * fma_per_iter / ms / 1e9, thr * instr_per_iter / ms / 1e9, thr * fma_per_iter / (ms * 1e-3) / sms / (mhz * 1e6),
         mhz);
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1e3;
  float* out;
  cudaMalloc(&out, sms * 8 * 256 * 4);
  run<0>(out, "FFMA", 16, 16, sms, mhz);
  run<1>(out, "FFMA2", 16, 8, sms, mhz);
  run<2>(out, "HFMA2.f16", 16, 8, sms, mhz);
  run<3>(out, "HFMA2.bf16", 16, 8, sms, mhz);
  run<4>(out, "FFMA2+FFMA 1:1", 24, 16, sms, mhz);
  run<5>(out, "HFMA2+FFMA 1:1", 24, 16, sms, mhz);
  run<6>(out, "HFMA2+FFMA2 1:1", 32, 16, sms, mhz);
  run<7>(out, "FFMA2+FFMA 1:2", 32, 24, sms, mhz);
  run<8>(out, "HFMA2 + cvt.f16x2 per 4", 16, 10, sms, mhz);
  return 0;
}
