// Kaldi-compatible log-mel filterbank front-end of the speaker embedder, fused into one kernel per frame:
// framing (400/160, snip_edges) -> DC removal -> pre-emphasis 0.97 -> povey window -> 512-point FFT ->
// power spectrum -> 80 triangular mel filters -> log(max(., eps)); then per-utterance mean subtraction.
// Follows torchaudio/compliance/kaldi.py:514-646 (fbank) with the defaults the modelscope ERes2NetV2
// pipeline uses.  One warp owns one frame; the FFT runs in shared memory.
#pragma once
#include "ptx.cuh"

namespace tdz {

constexpr int FB_WIN = 400, FB_SHIFT = 160, FB_NFFT = 512, FB_NMEL = 80, FB_NBIN = 257;

struct FbankTables {
  const float* window;   // [400] povey window
  const float2* twiddle; // [256] exp(-2*pi*i*k/512)
  const float* mel;      // [80][257] dense filter weights
  const int* mel_lo;     // [80] first bin with non-zero weight
  const int* mel_hi;     // [80] last bin with non-zero weight
};

__global__ void __launch_bounds__(128) fbank_kernel(const float* __restrict__ wav, int64_t T, int64_t frames,
                                                    int64_t total_frames, FbankTables tb, float* __restrict__ feat) {
  pdl_enter();
  __shared__ float2 fft[4][FB_NFFT];
  __shared__ float raw[4][FB_WIN];
  __shared__ float2 tw[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 256; i += 128) tw[i] = tb.twiddle[i];
  __syncthreads();
  const int64_t fidx = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  if (fidx >= total_frames) return;
  const int64_t n = fidx / frames;
  const int64_t f = fidx - n * frames;
  const float* src = wav + n * T + f * FB_SHIFT;
  float2* s = fft[warp];
  float* r = raw[warp];
  float sum = 0.f;
  for (int i = lane; i < FB_WIN; i += 32) {
    const float v = src[i];
    r[i] = v;
    sum += v;
  }
  const float mean = warp_sum(sum) * (1.f / FB_WIN);
  __syncwarp();
  for (int i = lane; i < FB_NFFT; i += 32) {
    float v = 0.f;
    if (i < FB_WIN) {
      const float cur = r[i] - mean;
      const float prev = r[i > 0 ? i - 1 : 0] - mean;
      v = (cur - 0.97f * prev) * tb.window[i];
    }
    const int rev = __brev(static_cast<unsigned>(i)) >> 23;  // 9-bit reversal
    s[rev] = make_float2(v, 0.f);
  }
  __syncwarp();
#pragma unroll 1
  for (int half = 1; half < FB_NFFT; half <<= 1) {
    const int tstep = 256 / half;
    for (int j = lane; j < 256; j += 32) {
      const int grp = j / half, pos = j - grp * half;
      const int i0 = grp * 2 * half + pos, i1 = i0 + half;
      const float2 w = tw[pos * tstep];
      const float2 a = s[i0], b = s[i1];
      const float2 bw = make_float2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
      s[i0] = make_float2(a.x + bw.x, a.y + bw.y);
      s[i1] = make_float2(a.x - bw.x, a.y - bw.y);
    }
    __syncwarp();
  }
  // power spectrum into r[0..256] (raw frame no longer needed)
  for (int i = lane; i < FB_NBIN; i += 32) {
    const float2 c = s[i];
    r[i] = c.x * c.x + c.y * c.y;
  }
  __syncwarp();
  for (int mbin = lane; mbin < FB_NMEL; mbin += 32) {
    const int lo = tb.mel_lo[mbin], hi = tb.mel_hi[mbin];
    float e = 0.f;
    for (int k = lo; k <= hi; ++k) e = fmaf(r[k], tb.mel[mbin * FB_NBIN + k], e);
    feat[fidx * FB_NMEL + mbin] = logf(fmaxf(e, 1.1920928955078125e-07f));
  }
}

// feature - feature.mean(dim=0): one block per utterance, 240 threads = 80 bins x 3 frame phases.
__global__ void __launch_bounds__(240) fbank_meannorm_kernel(float* __restrict__ feat, int64_t frames) {
  pdl_enter();
  __shared__ float part[3][FB_NMEL];
  const int bin = threadIdx.x % FB_NMEL, ph = threadIdx.x / FB_NMEL;
  float* base = feat + static_cast<int64_t>(blockIdx.x) * frames * FB_NMEL;
  float s = 0.f;
  for (int64_t f = ph; f < frames; f += 3) s += base[f * FB_NMEL + bin];
  part[ph][bin] = s;
  __syncthreads();
  const float mean = (part[0][bin] + part[1][bin] + part[2][bin]) / static_cast<float>(frames);
  for (int64_t f = ph; f < frames; f += 3) base[f * FB_NMEL + bin] -= mean;
}

}  // namespace tdz
