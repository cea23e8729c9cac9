// Memory-bound kernels of the ERes2NetV2 speaker embedder (architecture restated in
// oracle/eres2netv2_port.py).  Feature maps are NHWC "pixel-major" bf16: [N*H*W][C], H = mel axis, W = time axis,
// so every 1x1 convolution is a plain GEMM over pixels and every stride-1 3x3 convolution an implicit GEMM
// (gemm_conv3.cuh); only the stride-2 layer3_ds convolution goes through the im2col matrix written here.  All maps are
// tensor-core operands of the next convolution anyway, so they are stored once, in bf16 (accumulation, BN, gates
// and the residual add happen in fp32 inside the GEMM epilogues).
#pragma once
#include "ptx.cuh"

namespace tdz {

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __bfloat162float(h[i].x);
    v[2 * i + 1] = __bfloat162float(h[i].y);
  }
}

// Stem: Conv2d(1,64,3,pad 1) + BN (folded) + ReLU on feat [N][frames][80] -> [N][80][frames][64] bf16.
// Thread = one pixel x 8 channels.
__global__ void __launch_bounds__(256) sv_stem_kernel(const float* __restrict__ feat, const float* __restrict__ w,
                                                      const float* __restrict__ bias,
                                                      __nv_bfloat16* __restrict__ out_bf, int N, int H, int W) {
  pdl_enter();
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(N) * H * W * 8;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx & 7);
  const int64_t pix = idx >> 3;
  const int wq = static_cast<int>(pix % W);
  const int hq = static_cast<int>((pix / W) % H);
  const int n = static_cast<int>(pix / (static_cast<int64_t>(W) * H));
  float x[9];
#pragma unroll
  for (int dh = 0; dh < 3; ++dh)
#pragma unroll
    for (int dw = 0; dw < 3; ++dw) {
      const int hh = hq + dh - 1, ww = wq + dw - 1;
      // feat is [n][time][mel]: pixel (h = mel, w = time)
      x[dh * 3 + dw] = (hh >= 0 && hh < H && ww >= 0 && ww < W)
                           ? feat[(static_cast<int64_t>(n) * W + ww) * H + hh]
                           : 0.f;
    }
  float o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ch = cg * 8 + c;
    float a = bias[ch];
#pragma unroll
    for (int k = 0; k < 9; ++k) a = fmaf(w[ch * 9 + k], x[k], a);
    o[c] = fmaxf(a, 0.f);
  }
  *reinterpret_cast<uint4*>(out_bf + pix * 64 + cg * 8) =
      make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
}

// Stride-2 pixel subsampling (the stride of a 1x1 conv): out[n,h,w,:] = in[n,2h,2w,:].
__global__ void __launch_bounds__(256) sv_subsample_kernel(const __nv_bfloat16* __restrict__ in,
                                                           __nv_bfloat16* __restrict__ out, int N, int Hin, int Win,
                                                           int Hout, int Wout, int C) {
  pdl_enter();
  const int c8 = C / 8;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(N) * Hout * Wout * c8;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % c8);
  const int64_t pix = idx / c8;
  const int wq = static_cast<int>(pix % Wout);
  const int hq = static_cast<int>((pix / Wout) % Hout);
  const int n = static_cast<int>(pix / (static_cast<int64_t>(Wout) * Hout));
  *reinterpret_cast<uint4*>(out + pix * C + cg * 8) = *reinterpret_cast<const uint4*>(
      in + ((static_cast<int64_t>(n) * Hin + 2 * hq) * Win + 2 * wq) * C + cg * 8);
}

// im2col for a 3x3 convolution (pad 1, stride s): col[p][tap*C + c] = (a + b)[pixel(h*s+dh-1, w*s+dw-1)][c]
// with zeros outside the image; a, b bf16 with their own leading dimension / channel offset (b optional:
// the Res2Net hierarchical add sp_{i-1} + x_i, ERes2NetV2 block forward; the sum is formed in fp32).  K = 9*C.
// Grid y = output image row (n, h); thread = one output pixel of the row x one tap x 8 channels, so consecutive
// threads write consecutive 16 B pieces of the col row (all index arithmetic is 32-bit).
__global__ void __launch_bounds__(256) sv_im2col_kernel(const __nv_bfloat16* __restrict__ a, int lda, int offa,
                                                        const __nv_bfloat16* __restrict__ b, int ldb, int offb,
                                                        __nv_bfloat16* __restrict__ col, int N, int Hin, int Win,
                                                        int Hout, int Wout, int C, int stride) {
  pdl_enter();
  const int c8 = C >> 3;
  const int per_pix = 9 * c8;
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;  // position inside the output row
  if (j >= static_cast<unsigned>(Wout * per_pix)) return;
  const int wq = j / per_pix;
  const int r = j - wq * per_pix;
  const int tap = r / c8;
  const int cg = r - tap * c8;
  const int n = blockIdx.y / Hout;
  const int hq = blockIdx.y - n * Hout;
  const int dh = tap / 3;
  const int hh = hq * stride + dh - 1, ww = wq * stride + (tap - 3 * dh) - 1;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (hh >= 0 && hh < Hin && ww >= 0 && ww < Win) {
    const size_t sp = (static_cast<size_t>(n) * Hin + hh) * Win + ww;
    o = *reinterpret_cast<const uint4*>(a + sp * lda + offa + cg * 8);
    if (b != nullptr) {
      const uint4 y = *reinterpret_cast<const uint4*>(b + sp * ldb + offb + cg * 8);
      float xa[8], xb[8];
      bf16x8_to_f32(o, xa);
      bf16x8_to_f32(y, xb);
      o = make_uint4(pack_bf16(xa[0] + xb[0], xa[1] + xb[1]), pack_bf16(xa[2] + xb[2], xa[3] + xb[3]),
                     pack_bf16(xa[4] + xb[4], xa[5] + xb[5]), pack_bf16(xa[6] + xb[6], xa[7] + xb[7]));
    }
  }
  const size_t pix = (static_cast<size_t>(n) * Hout + hq) * Wout + wq;
  *reinterpret_cast<uint4*>(col + pix * (9 * C) + r * 8) = o;
}

// AFF input: cat(x, y) along channels -> [P][2C].
__global__ void __launch_bounds__(256) sv_cat2_kernel(const __nv_bfloat16* __restrict__ a, int lda, int offa,
                                                      const __nv_bfloat16* __restrict__ b, int ldb, int offb,
                                                      __nv_bfloat16* __restrict__ out, int64_t P, int C) {
  pdl_enter();
  const int c8 = C / 8;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= P * 2 * c8) return;
  const int cg = static_cast<int>(idx % (2 * c8));
  const int64_t pix = idx / (2 * c8);
  const __nv_bfloat16* src = cg < c8 ? a + pix * lda + offa + cg * 8 : b + pix * ldb + offb + (cg - c8) * 8;
  *reinterpret_cast<uint4*>(out + pix * (2 * C) + cg * 8) = *reinterpret_cast<const uint4*>(src);
}

// TSTP pooling: mean and sqrt(unbiased var + 1e-8) over time for every (channel, mel) pair of
// fuse [N][H][W][C] (fp32); output stats[n][c*H + h] (mean block) and stats[n][C*H + c*H + h] (std block) as bf16
// operand of the embedding Linear, matching reshape(N, C*F, T) of an NCHW tensor.  Thread = (n, h, c).
__global__ void __launch_bounds__(256) sv_tstp_kernel(const float* __restrict__ fuse, __nv_bfloat16* __restrict__ stats,
                                                      int N, int H, int W, int C, int ld_stats) {
  pdl_enter();
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(N) * H * C) return;
  const int c = static_cast<int>(idx % C);
  const int h = static_cast<int>((idx / C) % H);
  const int n = static_cast<int>(idx / (static_cast<int64_t>(C) * H));
  const float* base = fuse + ((static_cast<int64_t>(n) * H + h) * W) * C + c;
  float s = 0.f;
  for (int w = 0; w < W; ++w) s += base[static_cast<int64_t>(w) * C];
  const float mean = s / static_cast<float>(W);
  float q = 0.f;
  for (int w = 0; w < W; ++w) {
    const float d = base[static_cast<int64_t>(w) * C] - mean;
    q += d * d;
  }
  const float var = W > 1 ? q / static_cast<float>(W - 1) : 0.f;
  __nv_bfloat16* dst = stats + static_cast<int64_t>(n) * ld_stats;
  dst[c * H + h] = __float2bfloat16(mean);
  dst[C * H + c * H + h] = __float2bfloat16(sqrtf(var + 1e-8f));
}

__global__ void sv_fill_kernel(float* __restrict__ dst, int64_t n, float value) {
  pdl_enter();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = value;
}

}  // namespace tdz
