// Memory-bound kernels of the MossFormer2 separator (everything that is not a dense contraction).
// Activations are token-major fp32 [B*Sp][C]; Sp = frames per sample rounded up to the 256-frame attention
// group (mossformer_block.py:236-250), frames t >= S are padding.  Time convolutions keep a register
// sliding window per thread so every input element is loaded once; lanes run along channels so every
// global access is a fully coalesced 256 B row segment.
#pragma once
#include "ptx.cuh"

namespace tdz {

// ---------------------------------------------------------------- encoder  (mossformer2.py:178-208)
// enc[b,t,c] = relu(sum_j w[c,j] * mix[b, 8t+j]),  + per-sample sum / sum^2 for GroupNorm(1,512).
constexpr int ENC_FRAMES = 64;
__global__ void __launch_bounds__(512) encoder_kernel(const float* __restrict__ mix, int T, const float* __restrict__ w,
                                                      float* __restrict__ enc, double* __restrict__ gn_stats, int B,
                                                      int Sp, int S) {
  pdl_enter();
  __shared__ __align__(16) float xs[ENC_FRAMES * 8 + 8];
  __shared__ float red[2][16];
  const int strips = Sp / ENC_FRAMES;
  const int b = blockIdx.x / strips;
  const int t0 = (blockIdx.x - b * strips) * ENC_FRAMES;
  const int c = threadIdx.x;
  for (int i = threadIdx.x; i < ENC_FRAMES * 8 + 8; i += 512) {
    const int n = t0 * 8 + i;
    xs[i] = (n < T) ? mix[static_cast<size_t>(b) * T + n] : 0.f;
  }
  float wr[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) wr[j] = w[c * 16 + j];
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  // the 16-sample input window of a frame lives in registers and slides by 8 samples per frame: two 128-bit
  // broadcast loads per frame instead of sixteen scalar ones (the loop was shared-memory-issue bound)
  float xw[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) xw[8 + j] = xs[j];
#pragma unroll 4
  for (int f = 0; f < ENC_FRAMES; ++f) {
    const int t = t0 + f;
#pragma unroll
    for (int j = 0; j < 8; ++j) xw[j] = xw[8 + j];
    const float4 n0 = *reinterpret_cast<const float4*>(xs + f * 8 + 8);
    const float4 n1 = *reinterpret_cast<const float4*>(xs + f * 8 + 12);
    xw[8] = n0.x; xw[9] = n0.y; xw[10] = n0.z; xw[11] = n0.w;
    xw[12] = n1.x; xw[13] = n1.y; xw[14] = n1.z; xw[15] = n1.w;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc = fmaf(wr[j], xw[j], acc);
    acc = (t < S) ? fmaxf(acc, 0.f) : 0.f;
    enc[(static_cast<size_t>(b) * Sp + t) * 512 + c] = acc;
    s1 += acc;
    s2 += acc * acc;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s1;
    red[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float a = threadIdx.x < 16 ? red[0][threadIdx.x] : 0.f;
    float q = threadIdx.x < 16 ? red[1][threadIdx.x] : 0.f;
    a = warp_sum(a);
    q = warp_sum(q);
    if (threadIdx.x == 0) {
      atomicAdd(gn_stats + 2 * b, static_cast<double>(a));
      atomicAdd(gn_stats + 2 * b + 1, static_cast<double>(q));
    }
  }
}

// GroupNorm(1,C,eps) statistics -> per-sample scale/shift (mossformer2.py:152).  A = rstd, Bv = -rstd*mean.
__global__ void gn_finalize_kernel(const double* __restrict__ stats, float* __restrict__ sampA,
                                   float* __restrict__ sampB, int B, double count, double eps) {
  pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double mean = stats[2 * b] / count;
  double var = stats[2 * b + 1] / count - mean * mean;
  if (var < 0) var = 0;
  const double rstd = 1.0 / sqrt(var + eps);
  sampA[b] = static_cast<float>(rstd);
  sampB[b] = static_cast<float>(-rstd * mean);
}

// Rotary angle table (rotary_embedding_torch: angle[t,j] = t * freqs[j]); fp32 positions.
__global__ void rotary_table_kernel(const float* __restrict__ freqs, float2* __restrict__ tab, int Sp) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Sp * 16) return;
  const int t = i >> 4, j = i & 15;
  const float a = static_cast<float>(t) * freqs[j];
  tab[i] = make_float2(cosf(a), sinf(a));
}

// ScaledSinuEmbedding table (mossformer_block.py:60-73): tab[t][c] = scale * sin(t f_c) for c < N/2, scale * cos(t
// f_{c-N/2}) above; fp32 positions.  Built once per forward (the rows are shared by every sample of the batch), so the
// GEMM epilogue that adds it reads a float4 instead of evaluating four sinf / cosf per output.
__global__ void posenc_table_kernel(const float* __restrict__ inv_freq, const float* __restrict__ scale,
                                    float* __restrict__ tab, int Sp, int N) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Sp * N) return;
  const int t = i / N, c = i - t * N, hlf = N >> 1;
  const float a = static_cast<float>(t) * inv_freq[c < hlf ? c : c - hlf];
  tab[i] = scale[0] * (c < hlf ? sinf(a) : cosf(a));
}

// OffsetScale (4 heads) + rotary on dims 0..31 (interleaved pairs, rotary_embedding_torch) of the to_qk output
// (mossformer_block.py:76-86,214,230-233) -> qk4 [Mtot][512] 16-bit = quad_q | lin_q | quad_k | lin_k.  Three heads are
// bf16; lin_q is stored as FP16: the product lin_q @ lin_kv carries the largest rounding error of the network when both
// operands have 8 mantissa bits (the time-averaged lin_kv has a large common part), with 11 bits it is as exact as
// the three-term bf16 split used before (oracle emulation: 45.7 vs 46.0 dB on speech, 41.2 dB with plain bf16) at a
// third of the MMA work.
// Thread = 2 adjacent channels x QKH_FRAMES frames (the OffsetScale constants of its channels stay in registers);
// block = 4 x QKH_FRAMES consecutive frames.
constexpr int QKH_FRAMES = 8;
__global__ void __launch_bounds__(256) qk_heads_kernel(const float* __restrict__ qkf, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float2* __restrict__ rot,
                                                       __nv_bfloat16* __restrict__ qk4, int Sp, int S, size_t rows) {
  pdl_enter();
  const int c = (threadIdx.x & 63) * 2;
  float2 g[4], bt[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    g[h] = *reinterpret_cast<const float2*>(gamma + h * 128 + c);
    bt[h] = *reinterpret_cast<const float2*>(beta + h * 128 + c);
  }
  const size_t row0 = static_cast<size_t>(blockIdx.x) * (4 * QKH_FRAMES) + (threadIdx.x >> 6);
#pragma unroll 2
  for (int f = 0; f < QKH_FRAMES; ++f) {
    const size_t row = row0 + 4 * f;
    if (row >= rows) return;
    const int t = static_cast<int>(row % Sp);
    if (t >= S) continue;  // padded frames stay zero (zeroed once per forward)
    const float2 v = *reinterpret_cast<const float2*>(qkf + row * 128 + c);
    float2 cs = make_float2(1.f, 0.f);
    if (c < 32) cs = rot[t * 16 + (c >> 1)];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float x0 = fmaf(v.x, g[h].x, bt[h].x), x1 = fmaf(v.y, g[h].y, bt[h].y);
      const float r0 = x0 * cs.x - x1 * cs.y;
      const float r1 = x1 * cs.x + x0 * cs.y;
      const uint32_t packed = (h == 1) ? pack_f16(r0, r1) : pack_bf16(r0, r1);
      *reinterpret_cast<uint32_t*>(qk4 + row * 512 + h * 128 + c) = packed;
    }
  }
}

// ScaleNorm (mossformer_block.py:44-54) as one scale per frame: out[row] = 0.5 / clamp(||x_row|| dim^-0.5, 1e-5),
// from the partial sums of squares the producing GEMM epilogue left behind.  The factor 0.5 belongs to the
// tanh form of SiLU the consumer uses.  SHIFT: the row is the token-shifted frame (channels 0..255 of the
// previous frame | channels 256..511 of this frame, :204-207) and `parts` holds 8 sums of 64 channels per frame;
// otherwise `parts` holds 32 sums per frame, part-major ([32][rows]).
template <bool SHIFT>
__global__ void rowscale_kernel(const float* __restrict__ parts, float* __restrict__ out, int Sp, int S, size_t rows,
                                float dim_rsqrt) {
  pdl_enter();
  const size_t row = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const int t = static_cast<int>(row % Sp);
  if (t >= S) return;
  float ss;
  if (SHIFT) {
    const float4 cur = *reinterpret_cast<const float4*>(parts + row * 8 + 4);   // channels 256..511 of this frame
    ss = (cur.x + cur.y) + (cur.z + cur.w);
    if (t > 0) {
      const float4 prv = *reinterpret_cast<const float4*>(parts + (row - 1) * 8);  // channels 0..255 of the previous
      ss += (prv.x + prv.y) + (prv.z + prv.w);
    }
  } else {
    ss = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 4)
      ss += (parts[i * rows + row] + parts[(i + 1) * rows + row]) +
            (parts[(i + 2) * rows + row] + parts[(i + 3) * rows + row]);
  }
  out[row] = 0.5f / fmaxf(sqrtf(ss) * dim_rsqrt, 1e-5f);
}

// ---------------------------------------------------------------- DilatedDenseNet  (fsmn.py:76-111)
// stage 1: y1 = depthwise conv (39 taps, pad 19) of p; statistics for InstanceNorm over all S frames.
// stage 2: y2[c] = sum_{j<2} conv39_dilation2( cat[2c+j] ), cat = [PReLU(IN(y1)) ; p], pad 38.
//
// Streaming kernel: a CTA owns (sample, group of 128 input channels, time segment) and walks along time.
// Thread 0 keeps TMA loads of 64-row x 512 B chunks one step ahead in a 4-slot shared-memory ring, so every
// input element is read from HBM once and the loads overlap the FMAs; a fix-up pass applies InstanceNorm +
// PReLU (stage 2, y1 half) and the zero padding in place in shared memory, after which the inner loop is pure
// FMA: thread = one channel (pair) x 16 outputs, 54-deep register window, taps from shared memory.
constexpr int DD_CHUNK = 64;                 // rows per ring slot / outputs per step
constexpr int DD_SLOTS = 4;                  // 256 ring rows: a power of two, so the wrap is a mask
constexpr int DD_RING_ROWS = DD_CHUNK * DD_SLOTS;
// DD_CH input channels per CTA.  64 (two channel-pair warps x DD_NQ time groups = 128 threads, 80 KB of shared memory)
// lets TWO independent CTAs share an SM: a CTA's warps run in lock step through its phases (chunk wait / fix-up /
// window loads / FMAs / stores, one __syncthreads per step), so with ONE 256-thread CTA per SM the FMA pipe idled
// during every non-FMA phase (45 % busy); two CTAs drift apart and fill each other's gaps.
#ifndef TDZ_DD_CH
#define TDZ_DD_CH 64
#endif
constexpr int DD_CH = TDZ_DD_CH;
constexpr int DD_CP = DD_CH / 2;             // channel pairs (stage 1) / output channels (stage 2) per CTA
constexpr int DD_ROW_BYTES = DD_CH * 4;
// DD_NQ time groups of DD_OUT outputs per 64-row step.  4 x 16 (256 threads, two warps per scheduler) is the default:
// the 8 x 8 form (512 threads, four warps per scheduler) was measured 15 % SLOWER - its windows overlap more, so it
// issues 85 shared-memory loads per 312 packed FMAs instead of 93 per 624 and becomes shared-memory bound.
#ifndef TDZ_DD_NQ
#define TDZ_DD_NQ 4
#endif
constexpr int DD_NQ = TDZ_DD_NQ;
constexpr int DD_OUT = DD_CHUNK / DD_NQ;
constexpr int DD_WIN = DD_OUT + 38;
constexpr int DD_THREADS = DD_CP * DD_NQ;
static_assert(DD_CP % 32 == 0, "a warp must not straddle two time groups");
// ring | taps | barriers (the fp64 statistics scratch of the epilogue reuses the ring)
constexpr int DD_SMEM_BYTES = DD_RING_ROWS * DD_ROW_BYTES + 39 * DD_CP * 8 + 64 /*barriers*/ + 128;
#ifndef TDZ_DD_CTAS
#define TDZ_DD_CTAS 3
#endif
constexpr int DD_CTAS_PER_SM = TDZ_DD_CTAS;  // 3 x 75 KB of shared memory, 3 x 128 threads x <= 168 registers
static_assert(DD_NQ * DD_CP * 4 * 8 <= DD_RING_ROWS * DD_ROW_BYTES, "statistics scratch must fit in the ring");

struct DdParams {
  CUtensorMap tmA;       // stage 1: p;  stage 2: y1          3-D {256, Sp, B}, box {128, 64, 1}, no swizzle
  CUtensorMap tmB;       // stage 2: p
  const float* taps;     // stage 1: [256][39];  stage 2: [256][2][39]
  const float2* in_ss;   // stage 2: InstanceNorm-1 (scale, shift) [B][256]
  const float* prelu;    // stage 2: PReLU-1 slopes [256]
  float* out;            // y1 / y2 [Mtot][256]
  double* stats;         // [B][256][2] sum, sum of squares of the outputs
  int B, Sp, S, nseg, seg_len;
};

template <int STAGE>
__global__ void __launch_bounds__(DD_THREADS, DD_CTAS_PER_SM) dd_stream_kernel(const __grid_constant__ DdParams P) {
  pdl_enter();
  constexpr int NCG = (STAGE == 1 ? 256 : 512) / DD_CH;  // channel groups: stage 2 reads y1 (first half) and p
  constexpr int HALO = STAGE == 1 ? 19 : 38;
  constexpr int STEP = STAGE == 1 ? 1 : 2;   // dilation
  extern __shared__ uint8_t dd_smem_raw[];
  uint8_t* sm = dd_smem_raw + ((128u - (smem_u32(dd_smem_raw) & 127u)) & 127u);
  float* ring = reinterpret_cast<float*>(sm);
  float2* ws = reinterpret_cast<float2*>(sm + DD_RING_ROWS * DD_ROW_BYTES);
  const uint32_t bar0 = smem_u32(sm + DD_RING_ROWS * DD_ROW_BYTES + 39 * DD_CP * 8);
  double* red = reinterpret_cast<double*>(sm);  // reuses the ring after the last step

  const int tid = threadIdx.x;
  int bid = blockIdx.x;
  const int cg = bid % NCG;
  bid /= NCG;
  const int seg = bid % P.nseg;
  const int b = bid / P.nseg;
  const int seg_lo = seg * P.seg_len;
  if (seg_lo >= P.S) return;
  const int seg_hi = min(seg_lo + P.seg_len, P.S);
  const int nsteps = (seg_hi - seg_lo + DD_CHUNK - 1) / DD_CHUNK;
  const bool from_y1 = (STAGE == 2) && cg < NCG / 2;
  const int cin0 = (STAGE == 1) ? cg * DD_CH : (cg % (NCG / 2)) * DD_CH;
  const CUtensorMap* map = (STAGE == 2 && !from_y1) ? &P.tmB : &P.tmA;

  const int cp = tid % DD_CP;   // channel pair (stage 1) / output channel (stage 2) inside the group
  const int q = tid / DD_CP;    // which DD_OUT outputs of the step
  for (int k = 0; k < 39; ++k) {
    if (tid < DD_CP) {
      if (STAGE == 1) {
        const int c = cg * DD_CH + 2 * cp;
        ws[k * DD_CP + cp] = make_float2(P.taps[c * 39 + k], P.taps[(c + 1) * 39 + k]);
      } else {
        const int oc = cg * DD_CP + cp;
        ws[k * DD_CP + cp] = make_float2(P.taps[(oc * 2 + 0) * 39 + k], P.taps[(oc * 2 + 1) * 39 + k]);
      }
    }
  }
  if (tid == 0) {
    tma_prefetch_desc(map);
    for (int s = 0; s < DD_SLOTS; ++s) mbar_init(bar0 + 8u * s, 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int m) {  // chunk m = rows [seg_lo + 64 m, +64) of the 128 channels -> slot (m+1) % 4
    const int slot = (m + 1) % DD_SLOTS;
    const uint32_t bar = bar0 + 8u * slot;
    mbar_arrive_expect_tx(bar, DD_CHUNK * DD_ROW_BYTES);
    tma_load_3d(smem_u32(ring) + slot * DD_CHUNK * DD_ROW_BYTES, map, bar, cin0, seg_lo + DD_CHUNK * m, b);
  };
  auto wait_chunk = [&](int m) { mbar_wait(bar0 + 8u * ((m + 1) % DD_SLOTS), ((m + 1) / DD_SLOTS) & 1); };
  if (tid == 0) {
    for (int m = -1; m <= 1 && m <= nsteps; ++m) issue(m);
  }

  // fix-up pass constants: this thread touches 4 fixed channels (DD_THREADS is a multiple of the DD_CH / 4 vectors
  // of a row, so a thread's vector column never changes)
  constexpr int VPR = DD_CH / 4;                     // float4 vectors per ring row
  constexpr int FIX_ROWS = DD_THREADS / VPR;         // rows covered by one pass of the CTA
  static_assert(DD_THREADS % VPR == 0 && DD_CHUNK % FIX_ROWS == 0, "fix-up tiling");
  const int fc4 = (tid % VPR) * 4;
  float fsc[4] = {1.f, 1.f, 1.f, 1.f}, fsh[4] = {0.f, 0.f, 0.f, 0.f}, fal[4] = {1.f, 1.f, 1.f, 1.f};
  if (from_y1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 ss = P.in_ss[b * 256 + cin0 + fc4 + i];
      fsc[i] = ss.x;
      fsh[i] = ss.y;
      fal[i] = P.prelu[cin0 + fc4 + i];
    }
  }
  auto fixup = [&](int m) {  // InstanceNorm + PReLU (y1 half) and zero padding outside [0, S), in place
    const int t0 = seg_lo + DD_CHUNK * m;
    const bool all_valid = t0 >= 0 && t0 + DD_CHUNK <= P.S;
    if (!from_y1 && all_valid) return;
    float* base = ring + ((m + 1) % DD_SLOTS) * DD_CHUNK * DD_CH + fc4;
#pragma unroll
    for (int i = 0; i < DD_CHUNK / FIX_ROWS; ++i) {
      const int rr = tid / VPR + FIX_ROWS * i;
      const int t = t0 + rr;
      float4* ptr = reinterpret_cast<float4*>(base + rr * DD_CH);
      float4 v = *ptr;
      if (t < 0 || t >= P.S) {
        v = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (from_y1) {
        v.x = fmaf(v.x, fsc[0], fsh[0]); v.x = v.x >= 0.f ? v.x : fal[0] * v.x;
        v.y = fmaf(v.y, fsc[1], fsh[1]); v.y = v.y >= 0.f ? v.y : fal[1] * v.y;
        v.z = fmaf(v.z, fsc[2], fsh[2]); v.z = v.z >= 0.f ? v.z : fal[2] * v.z;
        v.w = fmaf(v.w, fsc[3], fsh[3]); v.w = v.w >= 0.f ? v.w : fal[3] * v.w;
      }
      *ptr = v;
    }
  };

  float s1x = 0.f, s2x = 0.f, s1y = 0.f, s2y = 0.f;
  for (int n = 0; n < nsteps; ++n) {
    if (n == 0) {
      wait_chunk(-1);
      fixup(-1);
      wait_chunk(0);
      fixup(0);
    }
    wait_chunk(n + 1);
    fixup(n + 1);
    __syncthreads();  // fix-ups visible; everybody is done with step n-1, so the slot of chunk n-2 is free
    if (tid == 0 && n + 2 <= nsteps) {
      fence_proxy_async();
      issue(n + 2);  // into the slot of chunk n-2; needed at the top of step n+1: a whole step of latency hiding
    }
    // first input row of this thread's window, relative to the ring origin (row -64 of the segment = ring row 0)
    int r_out0, r_in0;
    if (STAGE == 1) {
      r_out0 = DD_CHUNK * n + DD_OUT * q;
      r_in0 = r_out0 - HALO;
    } else {
      r_out0 = DD_CHUNK * n + (q & 1) + 2 * DD_OUT * (q >> 1);
      r_in0 = r_out0 - HALO;
    }
    static_assert((DD_RING_ROWS & (DD_RING_ROWS - 1)) == 0, "ring rows must be a power of two");
    const int R = (r_in0 + DD_CHUNK) & (DD_RING_ROWS - 1);
    float2 buf[DD_WIN];
    const float* col = ring + 2 * cp;
    if (R + (DD_WIN - 1) * STEP < DD_RING_ROWS) {  // warp-uniform: the window does not wrap -> immediate offsets
      const float* base = col + R * DD_CH;
#pragma unroll
      for (int i = 0; i < DD_WIN; ++i) buf[i] = *reinterpret_cast<const float2*>(base + i * STEP * DD_CH);
    } else {
#pragma unroll
      for (int i = 0; i < DD_WIN; ++i)
        buf[i] = *reinterpret_cast<const float2*>(col + ((R + i * STEP) & (DD_RING_ROWS - 1)) * DD_CH);
    }
    if (STAGE == 1) {
      float2 acc[DD_OUT];
#pragma unroll
      for (int j = 0; j < DD_OUT; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 39; ++k) {
        const float2 wk = ws[k * DD_CP + cp];
#pragma unroll
        for (int j = 0; j < DD_OUT; ++j) acc[j] = fma2(wk, buf[j + k], acc[j]);
      }
      const int t0 = seg_lo + r_out0;
      float* dst = P.out + (static_cast<size_t>(b) * P.Sp + t0) * 256 + cg * DD_CH + 2 * cp;
#pragma unroll
      for (int j = 0; j < DD_OUT; ++j) {
        if (t0 + j < seg_hi) {
          *reinterpret_cast<float2*>(dst + static_cast<size_t>(j) * 256) = acc[j];
          s1x += acc[j].x;
          s2x = fmaf(acc[j].x, acc[j].x, s2x);
          s1y += acc[j].y;
          s2y = fmaf(acc[j].y, acc[j].y, s2y);
        }
      }
    } else {
      // the two input channels of an output accumulate side by side (packed FMA) and are added at the end
      float2 acc2[DD_OUT];
#pragma unroll
      for (int j = 0; j < DD_OUT; ++j) acc2[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 39; ++k) {
        const float2 wk = ws[k * DD_CP + cp];
#pragma unroll
        for (int j = 0; j < DD_OUT; ++j) acc2[j] = fma2(wk, buf[j + k], acc2[j]);
      }
      float acc[DD_OUT];
#pragma unroll
      for (int j = 0; j < DD_OUT; ++j) acc[j] = acc2[j].x + acc2[j].y;
      const int t0 = seg_lo + r_out0;
      float* dst = P.out + (static_cast<size_t>(b) * P.Sp + t0) * 256 + cg * DD_CP + cp;
#pragma unroll
      for (int j = 0; j < DD_OUT; ++j) {
        if (t0 + 2 * j < seg_hi) {
          dst[static_cast<size_t>(2 * j) * 256] = acc[j];
          s1x += acc[j];
          s2x = fmaf(acc[j], acc[j], s2x);
        }
      }
    }
  }
  // statistics: reduce the time groups of a channel, then one fp64 atomic pair per channel
  __syncthreads();  // everybody is done reading the ring
  red[(q * DD_CP + cp) * 2 + 0] = static_cast<double>(s1x);
  red[(q * DD_CP + cp) * 2 + 1] = static_cast<double>(s2x);
  if (STAGE == 1) {
    red[DD_NQ * DD_CP * 2 + (q * DD_CP + cp) * 2 + 0] = static_cast<double>(s1y);
    red[DD_NQ * DD_CP * 2 + (q * DD_CP + cp) * 2 + 1] = static_cast<double>(s2y);
  }
  __syncthreads();
  if (tid < DD_CP) {
    double a = 0, c2 = 0, ay = 0, cy = 0;
    for (int g = 0; g < DD_NQ; ++g) {
      a += red[(g * DD_CP + tid) * 2];
      c2 += red[(g * DD_CP + tid) * 2 + 1];
      if (STAGE == 1) {
        ay += red[DD_NQ * DD_CP * 2 + (g * DD_CP + tid) * 2];
        cy += red[DD_NQ * DD_CP * 2 + (g * DD_CP + tid) * 2 + 1];
      }
    }
    if (STAGE == 1) {
      double* st = P.stats + (static_cast<size_t>(b) * 256 + cg * DD_CH + 2 * tid) * 2;
      atomicAdd(st + 0, a);
      atomicAdd(st + 1, c2);
      atomicAdd(st + 2, ay);
      atomicAdd(st + 3, cy);
    } else {
      double* st = P.stats + (static_cast<size_t>(b) * 256 + cg * DD_CP + tid) * 2;
      atomicAdd(st + 0, a);
      atomicAdd(st + 1, c2);
    }
  }
}

// InstanceNorm2d(affine) scale / shift per (sample, channel) from the accumulated fp64 sums: biased variance,
// eps 1e-5 (fsmn.py:93,103).  out[b*256+c] = (rstd*g, beta - mean*rstd*g).
__global__ void in_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ g,
                                   const float* __restrict__ bt, float2* __restrict__ out, int n, double count) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = i & 255;
  const double mean = stats[2 * i] / count;
  double var = stats[2 * i + 1] / count - mean * mean;
  if (var < 0) var = 0;
  const float rstd = static_cast<float>(1.0 / sqrt(var + 1e-5));
  out[i] = make_float2(rstd * g[c], bt[c] - static_cast<float>(mean) * rstd * g[c]);
}

// FSMN tail: o2 = PReLU(IN(y2)); f = x_u + o2 (fsmn.py:144); g = x_v*f + c (mossformer_block.py:324);
// CLayerNorm(256) (norm2, :423) with its affine folded into conv2 -> tf32 operand.  One warp per group of
// TAIL_FRAMES consecutive frames of one sample (the per-channel InstanceNorm / PReLU constants of the lane's 8
// channels are loaded once per group, as vectors).
// TAIL_FRAMES = 8 amortises the constants at large batches; small calls (the streaming shape: 1 280 frames) run one
// frame per warp - with 8 the batch-1 call was 160 warps that each walked 8 dependent rounds of loads (13 us).
template <int TAIL_FRAMES>
__global__ void __launch_bounds__(256) fsmn_tail_kernel(const float* __restrict__ y2, const float2* __restrict__ in2_ss,
                                                        const float* __restrict__ prelu2,
                                                        const float* __restrict__ xuv, const float* __restrict__ cres,
                                                        float* __restrict__ gout, int B, int Sp, int S) {
  pdl_enter();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int groups = Sp / TAIL_FRAMES;  // Sp is a multiple of 256
  if (warp >= B * groups) return;
  const int b = warp / groups;
  const int t0 = (warp - b * groups) * TAIL_FRAMES;
  const int c0 = lane * 8;
  float sc[8], sh[8], al[8];
  {
    const float4* sp = reinterpret_cast<const float4*>(in2_ss + static_cast<size_t>(b) * 256 + c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = sp[i];
      sc[2 * i] = v.x;
      sh[2 * i] = v.y;
      sc[2 * i + 1] = v.z;
      sh[2 * i + 1] = v.w;
    }
    *reinterpret_cast<float4*>(al) = *reinterpret_cast<const float4*>(prelu2 + c0);
    *reinterpret_cast<float4*>(al + 4) = *reinterpret_cast<const float4*>(prelu2 + c0 + 4);
  }
#pragma unroll 2
  for (int f = 0; f < TAIL_FRAMES; ++f) {
    const int t = t0 + f;
    const size_t grow = static_cast<size_t>(b) * Sp + t;
    float4* op = reinterpret_cast<float4*>(gout + grow * 256 + c0);
    if (t >= S) {  // warp-uniform
      op[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      op[1] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float4* yp = reinterpret_cast<const float4*>(y2 + grow * 256 + c0);
    const float4* up = reinterpret_cast<const float4*>(xuv + grow * 512 + c0);
    const float4* vp = reinterpret_cast<const float4*>(xuv + grow * 512 + 256 + c0);
    const float4* cp = reinterpret_cast<const float4*>(cres + grow * 256 + c0);
    float y[8], u[8], v[8], cr[8], g[8];
    *reinterpret_cast<float4*>(y) = yp[0];
    *reinterpret_cast<float4*>(y + 4) = yp[1];
    *reinterpret_cast<float4*>(u) = up[0];
    *reinterpret_cast<float4*>(u + 4) = up[1];
    *reinterpret_cast<float4*>(v) = vp[0];
    *reinterpret_cast<float4*>(v + 4) = vp[1];
    *reinterpret_cast<float4*>(cr) = cp[0];
    *reinterpret_cast<float4*>(cr + 4) = cp[1];
    float s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float o = fmaf(y[i], sc[i], sh[i]);
      o = o >= 0.f ? o : al[i] * o;
      g[i] = fmaf(v[i], u[i] + o, cr[i]);
      s1 += g[i];
    }
    const float mean = warp_sum(s1) * (1.f / 256.f);
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float d = g[i] - mean;
      s2 += d * d;
    }
    const float rstd = rsqrtf(warp_sum(s2) * (1.f / 256.f) + 1e-5f);
    // g is only the tf32 operand of conv2: round to nearest here (the MMA would truncate)
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = round_tf32_rn((g[i] - mean) * rstd);
    op[0] = make_float4(g[0], g[1], g[2], g[3]);
    op[1] = make_float4(g[4], g[5], g[6], g[7]);
  }
}

// lin_kv reduce: KV[b][d][e] = (sum_s part[b][s][d][e]) / S  (mossformer_block.py:286,289), stored as FP16 (see
// qk_heads_kernel).
__global__ void kv_reduce_kernel(const float* __restrict__ part, __half* __restrict__ kv, int nsplit,
                                 float inv_n, size_t per_sample /*128*2048*/, size_t total4) {
  pdl_enter();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const size_t e = i * 4;
  const size_t b = e / per_sample;
  const size_t r = e - b * per_sample;
  const float* src = part + b * nsplit * per_sample + r;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < nsplit; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(src + s * per_sample);
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
  }
  const float x0 = acc.x * inv_n, x1 = acc.y * inv_n, x2 = acc.z * inv_n, x3 = acc.w * inv_n;
  *reinterpret_cast<uint2*>(kv + b * per_sample + r) = make_uint2(pack_f16(x0, x1), pack_f16(x2, x3));
}

// ---------------------------------------------------------------- after the 24 layers
// LayerNorm(512, eps 1e-6) per frame (mossformer2.py:307,320) + per-sample statistics of its output
// for the GroupNorm(1,512) that follows (mossformer2.py:388-390).  One warp per frame.
__global__ void __launch_bounds__(256) final_ln_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                       const float* __restrict__ bta, float* __restrict__ out,
                                                       double* __restrict__ gn_stats, int B, int Sp, int S) {
  pdl_enter();
  __shared__ float red[2][8];
  const int warp_in_block = threadIdx.x >> 5;
  const int warp = blockIdx.x * 8 + warp_in_block;
  const int lane = threadIdx.x & 31;
  // all 8 frames of a block belong to one sample: Sp is a multiple of 8
  const int b = (blockIdx.x * 8) / Sp;
  const int t = warp - b * Sp;
  float q1 = 0.f, q2 = 0.f;
  if (t < S) {
    const size_t grow = warp;
    float v[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(v + 4 * i) = *reinterpret_cast<const float4*>(x + grow * 512 + i * 128 + lane * 4);
    float s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s1 += v[i];
    const float mean = warp_sum(s1) * (1.f / 512.f);
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = v[i] - mean;
      s2 += d * d;
    }
    const float rstd = rsqrtf(warp_sum(s2) * (1.f / 512.f) + 1e-6f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 128 + lane * 4;
      float4 o;
      o.x = (v[4 * i + 0] - mean) * rstd * g[c + 0] + bta[c + 0];
      o.y = (v[4 * i + 1] - mean) * rstd * g[c + 1] + bta[c + 1];
      o.z = (v[4 * i + 2] - mean) * rstd * g[c + 2] + bta[c + 2];
      o.w = (v[4 * i + 3] - mean) * rstd * g[c + 3] + bta[c + 3];
      *reinterpret_cast<float4*>(out + grow * 512 + c) = o;
      q1 += o.x + o.y + o.z + o.w;
      q2 += o.x * o.x + o.y * o.y + o.z * o.z + o.w * o.w;
    }
  }
  q1 = warp_sum(q1);
  q2 = warp_sum(q2);
  if (lane == 0) {
    red[0][warp_in_block] = q1;
    red[1][warp_in_block] = q2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, q = 0.f;
    for (int i = 0; i < 8; ++i) {
      a += red[0][i];
      q += red[1][i];
    }
    atomicAdd(gn_stats + 2 * b, static_cast<double>(a));
    atomicAdd(gn_stats + 2 * b + 1, static_cast<double>(q));
  }
}

// GroupNorm apply (per-channel affine) + skip around the block (mossformer2.py:393-394) + mask-net PReLU
// (mossformer2.py:500).
__global__ void final_gn_kernel(const float* __restrict__ ln, const float* __restrict__ sampA,
                                const float* __restrict__ sampB, const float* __restrict__ g,
                                const float* __restrict__ bta, const float* __restrict__ x0,
                                const float* __restrict__ alpha, float* __restrict__ out, int Sp, int S,
                                size_t total4) {
  pdl_enter();
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const size_t e = i * 4;
  const size_t grow = e / 512;
  const int c = static_cast<int>(e - grow * 512);
  const int b = static_cast<int>(grow / Sp);
  const int t = static_cast<int>(grow - static_cast<size_t>(b) * Sp);
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < S) {
    const float4 v = *reinterpret_cast<const float4*>(ln + e);
    const float4 r = *reinterpret_cast<const float4*>(x0 + e);
    const float A = sampA[b], Bv = sampB[b], al = alpha[0];
    float z;
    // the result is only the tf32 operand of conv1d_out: round to nearest (the MMA would truncate)
    z = (v.x * A + Bv) * g[c + 0] + bta[c + 0] + r.x;
    o.x = round_tf32_rn(z >= 0.f ? z : al * z);
    z = (v.y * A + Bv) * g[c + 1] + bta[c + 1] + r.y;
    o.y = round_tf32_rn(z >= 0.f ? z : al * z);
    z = (v.z * A + Bv) * g[c + 2] + bta[c + 2] + r.z;
    o.z = round_tf32_rn(z >= 0.f ? z : al * z);
    z = (v.w * A + Bv) * g[c + 3] + bta[c + 3] + r.w;
    o.w = round_tf32_rn(z >= 0.f ? z : al * z);
  }
  *reinterpret_cast<float4*>(out + e) = o;
}

// (decoder: ConvTranspose1d as a skinny tf32 GEMM, struct DecoderGemm in gemm_cfgs.cuh)

}  // namespace tdz
