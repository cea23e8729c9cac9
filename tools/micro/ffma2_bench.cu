// Micro-benchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 tools/micro/ffma2_bench.cu && /tmp/ffma2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r)
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return reinterpret_cast<float2&>(r);
}
template <int MODE>
__global__ void k(float* out, float s, int iters) {
  float2 a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 w = make_float2(s, s * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        a[i].x = fmaf(a[i].x, w.x, w.y);
        a[i].y = fmaf(a[i].y, w.x, w.y);
      } else {
        a[i] = fma2(a[i], w, w);
      }
    }
  }
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, 1.0001f, iters);
      else k<1><<<148 * 8, 256>>>(out, 1.0001f, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 148.0 * 8 * 256 * iters * 16.0;
    printf("%s: %.3f ms, %.2f TFMA/s (%.1f TFLOP/s)\n", mode == 0 ? "FFMA " : "FFMA2", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
  }
  return 0;
}
