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
                                                           float* __restrict__ ss) {
  __shared__ float2 s_spec[AP_BINS];
  __shared__ float s_feat[AP_FEAT];
  __shared__ float s_ss[8][AP_NBAND];   // per-warp partial sums of squares (added in warp order: no atomics)
  const int64_t frame = blockIdx.x;
  const int tid = threadIdx.x;
  for (int k = tid; k < AP_BINS; k += 256) s_spec[k] = spec[frame * AP_BINS + k];
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
// K and V of the head live in REGISTERS: lane l holds the rotated keys l, l + 32, l + 64 (3 x 32 values) and column l
// of V (80 values); per query only q (32 floats) and the 80 softmax weights go through shared memory as broadcast
// 128-bit reads.  (The first version kept K / V in shared memory: 350 LDS per query and warp, LSU bound at 845 us per
// layer for 10 s of audio - 42 % of the forward.)
struct ApAttnSmem {
  float q[AP_HEADS][AP_HD];
  float p[AP_HEADS][96];
};
__global__ void __launch_bounds__(256, 1) ap_attn_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                         const float* __restrict__ rot_cos,
                                                         const float* __restrict__ rot_sin,
                                                         __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) ApAttnSmem S;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t tok0 = static_cast<size_t>(blockIdx.x) * AP_NBAND;
  const __nv_bfloat16* base = qkv + tok0 * 768 + h * 96;
  // ---- keys lane, lane + 32, lane + 64, rotated: pairs (a, b) -> (a cos - b sin, b cos + a sin)
  float kreg[3][AP_HD];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int j = lane + 32 * r;
    if (j < AP_NBAND) {
      const uint4* kp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(j) * 768 + 32);
      const float4* cp = reinterpret_cast<const float4*>(rot_cos + j * AP_HD);
      const float4* sp = reinterpret_cast<const float4*>(rot_sin + j * AP_HD);
#pragma unroll
      for (int v8 = 0; v8 < 4; ++v8) {
        const uint4 raw = kp[v8];
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        float c[8], sn[8];
        const float4 c0 = __ldg(cp + 2 * v8), c1 = __ldg(cp + 2 * v8 + 1), s0 = __ldg(sp + 2 * v8), s1 = __ldg(sp + 2 * v8 + 1);
        c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
        sn[0] = s0.x; sn[1] = s0.y; sn[2] = s0.z; sn[3] = s0.w; sn[4] = s1.x; sn[5] = s1.y; sn[6] = s1.z; sn[7] = s1.w;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = __uint_as_float(w[i] << 16), b = __uint_as_float(w[i] & 0xffff0000u);
          kreg[r][8 * v8 + 2 * i] = a * c[2 * i] - b * sn[2 * i];
          kreg[r][8 * v8 + 2 * i + 1] = b * c[2 * i + 1] + a * sn[2 * i + 1];
        }
      }
    } else {
#pragma unroll
      for (int d = 0; d < AP_HD; ++d) kreg[r][d] = 0.f;
    }
  }
  // ---- column `lane` of V
  float vreg[AP_NBAND];
#pragma unroll
  for (int j = 0; j < AP_NBAND; ++j) vreg[j] = __bfloat162float(base[static_cast<size_t>(j) * 768 + 64 + lane]);
  const float sgn = (lane & 1) ? 1.f : -1.f;
  const float scale = 0.17677669529663687f;   // 1 / sqrt(32)
  const float rc = 0.f;
  (void)rc;
#pragma unroll 1
  for (int i = 0; i < AP_NBAND; ++i) {
    const float qv = __bfloat162float(base[static_cast<size_t>(i) * 768 + lane]);
    const float qp = __shfl_xor_sync(0xffffffffu, qv, 1);
    const float c = __ldg(rot_cos + i * AP_HD + lane), s = __ldg(rot_sin + i * AP_HD + lane);
    S.q[h][lane] = (qv * c + sgn * qp * s) * scale;
    __syncwarp();
    float sc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d4 = 0; d4 < AP_HD / 4; ++d4) {
      const float4 q4 = *reinterpret_cast<const float4*>(&S.q[h][4 * d4]);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        sc[r] = fmaf(q4.x, kreg[r][4 * d4], sc[r]);
        sc[r] = fmaf(q4.y, kreg[r][4 * d4 + 1], sc[r]);
        sc[r] = fmaf(q4.z, kreg[r][4 * d4 + 2], sc[r]);
        sc[r] = fmaf(q4.w, kreg[r][4 * d4 + 3], sc[r]);
      }
    }
    if (lane + 64 >= AP_NBAND) sc[2] = -INFINITY;
    float mx = fmaxf(fmaxf(sc[0], sc[1]), sc[2]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float e = __expf(sc[r] - mx);   // exp(-inf) = 0 for the keys that do not exist
      S.p[h][lane + 32 * r] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int j4 = 0; j4 < AP_NBAND / 4; ++j4) {
      const float4 p4 = *reinterpret_cast<const float4*>(&S.p[h][4 * j4]);
      acc0 = fmaf(p4.x, vreg[4 * j4], acc0);
      acc1 = fmaf(p4.y, vreg[4 * j4 + 1], acc1);
      acc0 = fmaf(p4.z, vreg[4 * j4 + 2], acc0);
      acc1 = fmaf(p4.w, vreg[4 * j4 + 3], acc1);
    }
    out[(tok0 + i) * AP_N + h * AP_HD + lane] = __float2bfloat16((acc0 + acc1) / sum);
    __syncwarp();
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
                                                           float2* __restrict__ est) {
  extern __shared__ uint8_t ap_smem_raw[];
  float* xs = reinterpret_cast<float*>(ap_smem_raw);   // [80][AP_XLD]
  __shared__ float s_out[AP_PAIRS];
  const int64_t frame = blockIdx.x;
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
#pragma unroll 4
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
    est[frame * AP_BINS + k] = make_float2(s_out[2 * k0 + j], s_out[2 * k0 + bw + j]);
  }
}

}  // namespace tdz
