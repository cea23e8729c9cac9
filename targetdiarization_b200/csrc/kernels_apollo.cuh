// SIMT kernels of the Apollo restorer (look2hear/models/apollo.py).  Activations are
// token-major: token = (row r, frame t, band) with the 256 features contiguous, so a Roformer sequence (the 80 bands
// of one frame) is 80 consecutive rows and the neighbours of a token in time are 80 rows apart.  The dense 1x1
// convolutions run as tcgen05 GEMMs (gemm_cfgs.cuh); what is here is the band split / merge around the per-band
// layers, the 80-token rotary attention, and the depthwise k7 convolution + RMSNorm.
#pragma once
#include "ptx.cuh"

namespace tdz {

constexpr int AP_NBAND = 80;
constexpr int AP_N = 256;        // feature_dim
constexpr int AP_BINS = 442;     // win / 2 + 1, win = 882
constexpr int AP_BW = 5;         // bins per band, bands 0..78
constexpr int AP_BW_LAST = AP_BINS - 79 * AP_BW;   // 47
constexpr int AP_FEAT = 2 * AP_BINS + AP_NBAND;    // 964 = sum over bands of (2 BW + 1)
constexpr int AP_PAIRS = 2 * AP_BINS;              // 884 (value, gate) output pairs of the band merge
constexpr int AP_HEADS = 8;
constexpr int AP_HD = 32;
constexpr float AP_EPS = 1e-5f;

__device__ __forceinline__ int ap_band_k0(int band) { return band * AP_BW; }
__device__ __forceinline__ int ap_band_bw(int band) { return band < AP_NBAND - 1 ? AP_BW : AP_BW_LAST; }
__device__ __forceinline__ int ap_feat_off(int band) { return band * (2 * AP_BW + 1); }   // prefix of (2 BW + 1)

// apollo.py:256-276 + BN[i] (RMSNorm(2 BW + 1) -> Conv1d(2 BW + 1, 256, 1)), one CTA of 256 threads per frame:
// spec [frames][442] complex -> x [frames * 80][256] fp32 + bf16 copy + the row's sum of squares (for the RMSNorm in
// front of the first Roformer GEMM; entry 0 of 4 partials).
//   g  [964]        RMSNorm gains, bands concatenated
//   w  [964][256]   conv weights, (band, k)-major, output channel contiguous
//   b  [80][256]
__global__ void __launch_bounds__(256) ap_bandsplit_kernel(const float2* __restrict__ spec, const float* __restrict__ g,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           float* __restrict__ x, __nv_bfloat16* __restrict__ xbf,
                                                           float* __restrict__ ss, int T, int t_lo, int Tl) {
  pdl_enter();
  __shared__ float2 s_spec[AP_BINS];
  __shared__ float s_feat[AP_FEAT];
  __shared__ float s_ss[8][AP_NBAND];   // per-warp partial sums of squares (added in warp order: no atomics)
  // block = local frame (row r, frame t_lo + tl of the T frames of the row): the token buffers hold Tl frames per row
  const int64_t frame = blockIdx.x;
  const int64_t src_frame = (frame / Tl) * T + t_lo + (frame % Tl);
  const int tid = threadIdx.x;
  for (int k = tid; k < AP_BINS; k += 256) s_spec[k] = spec[src_frame * AP_BINS + k];
  __syncthreads();
  // per band: power, normalised (re | im | log power), RMSNorm over the 2 BW + 1 features - a warp per band
  const int warp = tid >> 5, lane = tid & 31;
  for (int band = warp; band < AP_NBAND; band += 8) {
    const int k0 = ap_band_k0(band), bw = ap_band_bw(band), f0 = ap_feat_off(band);
    float p = 0.f;
    for (int k = lane; k < bw; k += 32) {
      const float2 s = s_spec[k0 + k];
      p += s.x * s.x + s.y * s.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    const float power = sqrtf(p + 1.1920928955078125e-07f);   // torch.finfo(float32).eps
    const float inv = 1.f / power, lp = logf(power);
    float q = 0.f;
    for (int k = lane; k < bw; k += 32) {
      const float2 s = s_spec[k0 + k];
      const float re = s.x * inv, im = s.y * inv;
      s_feat[f0 + k] = re;
      s_feat[f0 + bw + k] = im;
      q += re * re + im * im;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = rsqrtf((q + lp * lp) / static_cast<float>(2 * bw + 1) + AP_EPS);
    __syncwarp();
    for (int k = lane; k < 2 * bw + 1; k += 32) {
      const float v = k < 2 * bw ? s_feat[f0 + k] : lp;
      s_feat[f0 + k] = v * rs * __ldg(g + f0 + k);
    }
  }
  __syncthreads();
  // thread = output channel
  for (int band = 0; band < AP_NBAND; ++band) {
    const int nf = 2 * ap_band_bw(band) + 1, f0 = ap_feat_off(band);
    float acc = __ldg(b + band * AP_N + tid);
    const float* wp = w + static_cast<size_t>(f0) * AP_N + tid;
    for (int k = 0; k < nf; ++k) acc = fmaf(__ldg(wp + k * AP_N), s_feat[f0 + k], acc);
    const size_t row = static_cast<size_t>(frame) * AP_NBAND + band;
    x[row * AP_N + tid] = acc;
    xbf[row * AP_N + tid] = __float2bfloat16(acc);
    float q = acc * acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) s_ss[warp][band] = q;
  }
  __syncthreads();
  if (tid < AP_NBAND) {
    float* d = ss + (static_cast<size_t>(frame) * AP_NBAND + tid) * 4;
    *reinterpret_cast<float4*>(d) = make_float4(s_ss[0][tid] + s_ss[1][tid], s_ss[2][tid] + s_ss[3][tid],
                                                s_ss[4][tid] + s_ss[5][tid], s_ss[6][tid] + s_ss[7][tid]);
  }
}

// Roformer attention (apollo.py:120-135): one CTA per frame (a sequence of 80 band tokens), one warp per head.
// qkv [tokens][768] bf16, channel = head * 96 + (q 0..31 | k 32..63 | v 64..95); rotary embedding on q, k with the
// position = band index (interleaved pairs, cos / sin tables [100][32] of the checkpoint); softmax(q k^T / sqrt(32)) v;
// output [tokens][256] bf16, channel = head * 32 + d.
// 80 x 80 x 32 per head is far too small for a tcgen05 tile (M = 128, a TMEM round trip per head), so the two products
// run on warp-level mma.sync.m16n8k16 (bf16 operands, fp32 accumulate) entirely in registers, FlashAttention-2 style:
// 80 = 5 x 16 query rows = 10 x 8 keys, no padding anywhere.  A fragment register holds two ADJACENT features of one
// row - exactly a rotary pair - so the rotation is applied to the Q / K fragments in place as they are loaded from
// global memory; the S accumulator fragments of two neighbouring key tiles are the A fragment of the P V product.
// (History: K / V in shared memory and scalar FMAs - LSU bound, 845 us per layer for 10 s of audio; K / V in registers -
// 379 us, 62 % of 189 M warp instructions FFMA with two warps per scheduler and three dependent chains each; this form
// issues 200 MMAs + 200 exponentials per head instead of 14 000 FMAs.)
__device__ __forceinline__ void mma_bf16_16x8x16(float* c, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// two adjacent bf16 features (a, b) of position `pos`, first feature index d (even): rotated, scaled, re-packed
__device__ __forceinline__ uint32_t ap_rot_pair(uint32_t w, const float* __restrict__ rot_cos,
                                                const float* __restrict__ rot_sin, int pos, int d, float scale) {
  const float a = __uint_as_float(w << 16), b = __uint_as_float(w & 0xffff0000u);
  const float2 c = __ldg(reinterpret_cast<const float2*>(rot_cos + pos * AP_HD + d));
  const float2 sn = __ldg(reinterpret_cast<const float2*>(rot_sin + pos * AP_HD + d));
  return pack_bf16((a * c.x - b * sn.x) * scale, (b * c.y + a * sn.y) * scale);
}
__global__ void __launch_bounds__(256) ap_attn_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                      const float* __restrict__ rot_cos,
                                                      const float* __restrict__ rot_sin,
                                                      __nv_bfloat16* __restrict__ out) {
  pdl_enter();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  const size_t tok0 = static_cast<size_t>(blockIdx.x) * AP_NBAND;
  const __nv_bfloat16* base = qkv + tok0 * 768 + h * 96;
  auto word = [&](int row, int col) {   // two adjacent bf16 values (col even)
    return *reinterpret_cast<const uint32_t*>(base + static_cast<size_t>(row) * 768 + col);
  };
  // ---- B fragments of Q K^T: key tile nt (keys 8 nt + g), k-step ks (features 16 ks + 2 tg (+1), + 8)
  uint32_t kf[10][2][2];
#pragma unroll
  for (int nt = 0; nt < 10; ++nt) {
    const int key = nt * 8 + g;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int d = ks * 16 + 2 * tg;
      kf[nt][ks][0] = ap_rot_pair(word(key, 32 + d), rot_cos, rot_sin, key, d, 1.f);
      kf[nt][ks][1] = ap_rot_pair(word(key, 32 + d + 8), rot_cos, rot_sin, key, d + 8, 1.f);
    }
  }
  // ---- B fragments of P V: feature tile dt (feature 8 dt + g), key step kk (keys 16 kk + 2 tg (+1), + 8 (+9))
  uint32_t vf[4][5][2];
  {
    const unsigned short* vb = reinterpret_cast<const unsigned short*>(base) + 64;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      const int n = dt * 8 + g;
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) {
        const int k = kk * 16 + 2 * tg;
        const uint32_t v0 = vb[static_cast<size_t>(k) * 768 + n], v1 = vb[static_cast<size_t>(k + 1) * 768 + n];
        const uint32_t v2 = vb[static_cast<size_t>(k + 8) * 768 + n], v3 = vb[static_cast<size_t>(k + 9) * 768 + n];
        vf[dt][kk][0] = v0 | (v1 << 16);
        vf[dt][kk][1] = v2 | (v3 << 16);
      }
    }
  }
  const float scale = 0.17677669529663687f;   // 1 / sqrt(32), folded into q
#pragma unroll 1
  for (int mt = 0; mt < 5; ++mt) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    uint32_t qf[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int d = ks * 16 + 2 * tg;
      qf[ks][0] = ap_rot_pair(word(r0, d), rot_cos, rot_sin, r0, d, scale);
      qf[ks][1] = ap_rot_pair(word(r1, d), rot_cos, rot_sin, r1, d, scale);
      qf[ks][2] = ap_rot_pair(word(r0, d + 8), rot_cos, rot_sin, r0, d + 8, scale);
      qf[ks][3] = ap_rot_pair(word(r1, d + 8), rot_cos, rot_sin, r1, d + 8, scale);
    }
    float sc[10][4];
#pragma unroll
    for (int nt = 0; nt < 10; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
      mma_bf16_16x8x16(sc[nt], qf[0], kf[nt][0]);
      mma_bf16_16x8x16(sc[nt], qf[1], kf[nt][1]);
    }
    // softmax of rows r0 (elements 0, 1 of every tile) and r1 (elements 2, 3): the 4 threads of a group share a row
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 10; ++nt) {
      m0 = fmaxf(m0, fmaxf(sc[nt][0], sc[nt][1]));
      m1 = fmaxf(m1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
    uint32_t pf[5][4];
#pragma unroll
    for (int nt = 0; nt < 10; ++nt) {
      const float e0 = __expf(sc[nt][0] - m0), e1 = __expf(sc[nt][1] - m0);
      const float e2 = __expf(sc[nt][2] - m1), e3 = __expf(sc[nt][3] - m1);
      s0 += e0 + e1;
      s1 += e2 + e3;
      // key tiles 2 kk, 2 kk + 1 -> A fragment registers (0, 1) and (2, 3) of key step kk
      pf[nt >> 1][(nt & 1) * 2] = pack_bf16(e0, e1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(e2, e3);
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float i0 = 1.f / s0, i1 = 1.f / s1;
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < 5; ++kk) mma_bf16_16x8x16(o, pf[kk], vf[dt][kk]);
      __nv_bfloat16* op = out + (tok0 + r0) * AP_N + h * AP_HD + dt * 8 + 2 * tg;
      *reinterpret_cast<uint32_t*>(op) = pack_bf16(o[0] * i0, o[1] * i0);
      *reinterpret_cast<uint32_t*>(op + 8 * AP_N) = pack_bf16(o[2] * i1, o[3] * i1);
    }
  }
}

// ConvActNorm1d head (apollo.py:143-158): y = dwconv7(x) + bias along time (zero padding 3), then RMSNorm over the
// 256 features (the gain is folded into the following 1x1 conv) -> bf16 GEMM operand.
// A warp produces AP_DW_RUN consecutive frames of one (row, band), lane = 8 channels: all AP_DW_RUN + 6 input rows are
// requested up front (14 independent 32-byte loads per lane in flight) and the window is indexed statically.
// (First version: a 32-frame run with a shifted register window and one row load per step - latency bound, 16 % of the
// DRAM rate.)
//   taps [7][256], bias [256]
constexpr int AP_DW_RUN = 8;
__global__ void __launch_bounds__(256, 1) ap_dwconv_rms_kernel(const float* __restrict__ x, const float* __restrict__ taps,
                                                               const float* __restrict__ bias, int T, int runs_per_seq,
                                                               int64_t n_warps, __nv_bfloat16* __restrict__ u) {
  pdl_enter();
  const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= n_warps) return;
  // wid = (row * runs_per_seq + run) * 80 + band: neighbouring warps read neighbouring 1 KB rows
  const int band = static_cast<int>(wid % AP_NBAND);
  const int64_t rr = wid / AP_NBAND;
  const int run = static_cast<int>(rr % runs_per_seq);
  const int64_t r = rr / runs_per_seq;
  const int t0 = run * AP_DW_RUN;
  const int c0 = lane * 8;
  float win[AP_DW_RUN + 6][8];   // rows t0 - 3 .. t0 + AP_DW_RUN + 2
#pragma unroll
  for (int k = 0; k < AP_DW_RUN + 6; ++k) {
    const int t = t0 - 3 + k;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (t >= 0 && t < T) {
      const float* p = x + ((r * T + t) * AP_NBAND + band) * AP_N + c0;
      a = *reinterpret_cast<const float4*>(p);
      b = *reinterpret_cast<const float4*>(p + 4);
    }
    win[k][0] = a.x; win[k][1] = a.y; win[k][2] = a.z; win[k][3] = a.w;
    win[k][4] = b.x; win[k][5] = b.y; win[k][6] = b.z; win[k][7] = b.w;
  }
  float w[7][8], bv[8];
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(taps + j * AP_N + c0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(taps + j * AP_N + c0 + 4));
    w[j][0] = a.x; w[j][1] = a.y; w[j][2] = a.z; w[j][3] = a.w;
    w[j][4] = b.x; w[j][5] = b.y; w[j][6] = b.z; w[j][7] = b.w;
  }
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(bias + c0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
    bv[0] = a.x; bv[1] = a.y; bv[2] = a.z; bv[3] = a.w; bv[4] = b.x; bv[5] = b.y; bv[6] = b.z; bv[7] = b.w;
  }
#pragma unroll
  for (int k = 0; k < AP_DW_RUN; ++k) {
    const int t = t0 + k;
    float y[8];
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = bv[i];
#pragma unroll
      for (int j = 0; j < 7; ++j) a = fmaf(w[j][i], win[k + j][i], a);
      y[i] = a;
      q = fmaf(a, a, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = rsqrtf(q * (1.f / AP_N) + AP_EPS);
    if (t < T) {
      const uint4 o4 = make_uint4(pack_bf16(y[0] * rs, y[1] * rs), pack_bf16(y[2] * rs, y[3] * rs),
                                  pack_bf16(y[4] * rs, y[5] * rs), pack_bf16(y[6] * rs, y[7] * rs));
      *reinterpret_cast<uint4*>(u + ((r * T + t) * AP_NBAND + band) * AP_N + c0) = o4;
    }
  }
}

// Band merge (apollo.py:287-293): per band RMSNorm(256) -> Conv1d(256, 4 BW, 1) -> GLU -> (real BW | imag BW), one CTA
// per frame: the 80 normalised rows go to shared memory, thread = one (value, gate) output pair.
//   g  [80][256]         RMSNorm gains
//   wv, wg [256][884]    value / gate weights transposed (pair index contiguous: coalesced over threads)
//   bv, bg [884]
// pair p of band i: p = 2 k0(i) + j, j < 2 BW; j < BW -> real part of bin k0 + j, else imaginary part of bin k0 + j - BW
constexpr int AP_XLD = AP_N + 1;   // padded rows: threads of a warp may address different bands
__global__ void __launch_bounds__(256) ap_bandmerge_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                           const float* __restrict__ wv, const float* __restrict__ wg,
                                                           const float* __restrict__ bv, const float* __restrict__ bg,
                                                           float2* __restrict__ est, int T, int t_lo, int Tl, int keep_lo,
                                                           int keep_n) {
  pdl_enter();
  extern __shared__ uint8_t ap_smem_raw[];
  float* xs = reinterpret_cast<float*>(ap_smem_raw);   // [80][AP_XLD]
  __shared__ float s_out[AP_PAIRS];
  // block = one of the keep_n frames per row whose result is final (local frames [keep_lo, keep_lo + keep_n))
  const int64_t r = blockIdx.x / keep_n, kf = blockIdx.x % keep_n;
  const int64_t frame = r * Tl + keep_lo + kf;              // in the token buffers
  const int64_t dst_frame = r * T + t_lo + keep_lo + kf;    // in the spectrogram
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int band = warp; band < AP_NBAND; band += 8) {
    const float* xr = x + (static_cast<size_t>(frame) * AP_NBAND + band) * AP_N;
    float v[8];
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = xr[lane + 32 * i];
      q = fmaf(v[i], v[i], q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = rsqrtf(q * (1.f / AP_N) + AP_EPS);
#pragma unroll
    for (int i = 0; i < 8; ++i) xs[band * AP_XLD + lane + 32 * i] = v[i] * rs * __ldg(g + band * AP_N + lane + 32 * i);
  }
  __syncthreads();
  for (int p = tid; p < AP_PAIRS; p += 256) {
    const int band = min(p / (2 * AP_BW), AP_NBAND - 1);
    const float* xr = xs + band * AP_XLD;
    float a = __ldg(bv + p), gt = __ldg(bg + p);
#pragma unroll 16
    for (int k = 0; k < AP_N; ++k) {
      const float xv = xr[k];
      a = fmaf(__ldg(wv + static_cast<size_t>(k) * AP_PAIRS + p), xv, a);
      gt = fmaf(__ldg(wg + static_cast<size_t>(k) * AP_PAIRS + p), xv, gt);
    }
    s_out[p] = a / (1.f + __expf(-gt));
  }
  __syncthreads();
  for (int k = tid; k < AP_BINS; k += 256) {
    const int band = min(k / AP_BW, AP_NBAND - 1);
    const int k0 = ap_band_k0(band), bw = ap_band_bw(band);
    const int j = k - k0;
    est[dst_frame * AP_BINS + k] = make_float2(s_out[2 * k0 + j], s_out[2 * k0 + bw + j]);
  }
}

}  // namespace tdz
