// Micro-benchmark of the k=17 depthwise time-convolution inner loop of gemm_convt.cuh in isolation (registers only:
// no TMEM, no MUFU, no stores), at the epilogue's occupancy (3 warps per scheduler = 384 threads per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/conv_bench tools/micro/conv_bench.cu && build/conv_bench
// One "step" = 16 outputs per thread from a 32-frame window (272 FMAs).  Variants:
//   0  packed even/odd accumulators (FFMA2) + FADD combine, window slid by register moves (rolled loop)  [round 1]
//   1  same arithmetic, three-block register ring with static indices (steps unrolled by 3: no moves)
//   2  even taps FFMA2, odd taps scalar FFMA into the same accumulators (no FADD, no second accumulator set), ring
//   3  all scalar FFMA, ring
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r)
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return reinterpret_cast<float2&>(r);
}

// 16 outputs from pairs P[0..15] (frames 0..31): out[j] = sum_k w[k] * frame[j + k], j = 0..15
template <int V>
__device__ __forceinline__ void conv16(const float2* Pa, const float2* Pb, const float (&w)[17], float (&out)[16]) {
  // window pairs 0..7 = Pa[0..7], 8..15 = Pb[0..7]
  auto P = [&](int i) -> float2 { return i < 8 ? Pa[i] : Pb[i - 8]; };
  if constexpr (V == 0 || V == 1) {
    float2 accA[8], accB[9];
#pragma unroll
    for (int m = 0; m < 8; ++m) accA[m] = make_float2(0.f, 0.f);
#pragma unroll
    for (int m = 0; m < 9; ++m) accB[m] = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 9; ++q) {
#pragma unroll
      for (int m = 0; m < 8; ++m) accA[m] = fma2(make_float2(w[2 * q], w[2 * q]), P(m + q), accA[m]);
      if (q < 8) {
#pragma unroll
        for (int m = 0; m < 9; ++m)
          if (m + q < 16) accB[m] = fma2(make_float2(w[2 * q + 1], w[2 * q + 1]), P(m + q), accB[m]);
      }
    }
    // odd tap 2q+1 on pair (2(m+q), 2(m+q)+1) feeds outputs 2m-1 (from .x) and 2m (from .y)
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      out[2 * m] = accA[m].x + accB[m].y;
      out[2 * m + 1] = accA[m].y + accB[m + 1].x;
    }
  } else if constexpr (V == 2) {
    float2 acc[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m] = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 9; ++q) {
#pragma unroll
      for (int m = 0; m < 8; ++m) acc[m] = fma2(make_float2(w[2 * q], w[2 * q]), P(m + q), acc[m]);
      if (q < 8) {
        // odd tap k = 2q+1: out[j] += w[k] * frame[j + k];  frame f = (f & 1) ? P(f/2).y : P(f/2).x
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          acc[m].x = fmaf(w[2 * q + 1], P(m + q).y, acc[m].x);          // j = 2m   -> frame 2m + 2q + 1
          acc[m].y = fmaf(w[2 * q + 1], P(m + q + 1).x, acc[m].y);      // j = 2m+1 -> frame 2m + 2q + 2
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      out[2 * m] = acc[m].x;
      out[2 * m + 1] = acc[m].y;
    }
  } else {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 17; ++k)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int f = j + k;
        const float2 p = P(f >> 1);
        acc[j] = fmaf(w[k], (f & 1) ? p.y : p.x, acc[j]);
      }
#pragma unroll
    for (int j = 0; j < 16; ++j) out[j] = acc[j];
  }
}

template <int V>
__global__ void __launch_bounds__(448, 1) k(float* out, const float* taps, int steps) {
  float w[17];
#pragma unroll
  for (int i = 0; i < 17; ++i) w[i] = taps[i] + threadIdx.x * 1e-9f;
  float2 B0[8], B1[8], B2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    B0[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    B1[i] = make_float2(threadIdx.x * 0.002f - i, i * 0.25f);
    B2[i] = make_float2(0.f, 0.f);
  }
  float sink = 0.f;
  float o[16];
  if constexpr (V == 0) {
#pragma unroll 1
    for (int s = 0; s < steps; ++s) {
      conv16<0>(B0, B1, w, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // slide by register moves, new block from the outputs (keeps a dependency)
        B0[i] = B1[i];
        B1[i] = make_float2(o[2 * i] * 0.01f, o[2 * i + 1] * 0.01f);
      }
      sink += o[3];
    }
  } else {
#pragma unroll 1
    for (int s = 0; s < steps; s += 3) {
      conv16<V>(B0, B1, w, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) B2[i] = make_float2(o[2 * i] * 0.01f, o[2 * i + 1] * 0.01f);
      sink += o[3];
      conv16<V>(B1, B2, w, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) B0[i] = make_float2(o[2 * i] * 0.01f, o[2 * i + 1] * 0.01f);
      sink += o[3];
      conv16<V>(B2, B0, w, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) B1[i] = make_float2(o[2 * i] * 0.01f, o[2 * i + 1] * 0.01f);
      sink += o[3];
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
}

template <int V>
void run(float* out, const float* taps, int sms, int threads, const char* name) {
  const int steps = 3000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<V><<<sms, threads>>>(out, taps, steps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double cyc = ms * 1e-3 * khz * 1e3 / steps;            // cycles per step at the nominal clock
  const double warps_per_smsp = threads / 32.0 / 4.0;
  printf("%-44s threads %3d: %7.3f ms  %6.1f cyc/step (nominal clk)  %5.1f cyc per warp-step per scheduler  "
         "FMA/clk/SM %.1f\n", name, threads, ms, cyc, cyc / warps_per_smsp, 272.0 * threads / cyc);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  float *out, *taps;
  cudaMalloc(&out, sms * 448 * 4);
  cudaMalloc(&taps, 17 * 4);
  float h[17];
  for (int i = 0; i < 17; ++i) h[i] = 0.05f * (i - 8);
  cudaMemcpy(taps, h, sizeof h, cudaMemcpyHostToDevice);
  for (int threads : {384, 448, 512 - 64}) {
    run<0>(out, taps, sms, threads, "0 FFMA2 even/odd + FADD, moves (round 1)");
    run<1>(out, taps, sms, threads, "1 FFMA2 even/odd + FADD, ring");
    run<2>(out, taps, sms, threads, "2 FFMA2 even + FFMA odd, ring");
    run<3>(out, taps, sms, threads, "3 FFMA all taps, ring");
  }
  return 0;
}
