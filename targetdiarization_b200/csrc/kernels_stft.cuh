// STFT / inverse STFT engine: torch.stft / torch.istft as the reference calls them -
// center = True with reflect padding, periodic Hann window, one-sided, un-normalised - for the MDX-Net front / back end
// (AudioProcessor.py:82-120: n_fft 6144 = 2^11 * 3) and for the Apollo restorer (look2hear/models/apollo.py:261-262,
// 294-295: n_fft 882 = 2 * 3^2 * 7^2, hop 441).
//
// One CTA transforms TWO real frames with ONE complex FFT (z = a + i b; X_a[k] = (Z[k] + conj Z[N-k]) / 2,
// X_b[k] = -i (Z[k] - conj Z[N-k]) / 2), as a mixed-radix (4, 2, 3, 5, 7) Stockham autosort FFT in shared memory
// (two ping-pong buffers of n_fft complex values).  Twiddles come from a W_N table computed in double precision
// by the caller (exact fp32 roundings - sincosf in the kernel would cost 20 dB of SNR at n_fft 6144).
//
// The spectrogram is addressed with four strides (row, bin, frame, re/im), so the same kernels produce the MDX layout
// [B, 4 = ch * 2 + re/im, dim_f, dim_t] (frame-minor) and Apollo's frame-major complex [rows, T, bins].
#pragma once
#include "ptx.cuh"

namespace tdz {

constexpr int FFT_MAX_FACTORS = 12;
constexpr int FFT_THREADS = 256;

struct FftPlan {
  int n;                        // n_fft
  int hop;
  int nfac;
  int fac[FFT_MAX_FACTORS];     // radices, product = n
  const float* window;          // [n]
  const float2* tw;             // [n]: (cos, -sin)(2 pi k / n), i.e. W_n^k of the forward transform
};

struct SpecStrides {
  int64_t row, bin, frame, reim;  // element strides of the spectrogram
};

// host: n = product of radices from {4, 2, 3, 5, 7}; returns false if n has another prime factor
inline bool fft_factorize(int n, FftPlan* p) {
  p->n = n;
  p->nfac = 0;
  const int radices[5] = {4, 2, 3, 5, 7};
  for (int r : radices)
    while (n % r == 0) {
      if (p->nfac == FFT_MAX_FACTORS) return false;
      p->fac[p->nfac++] = r;
      n /= r;
    }
  return n == 1;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// W_n^k for the forward (INV = false) or inverse (conjugated) transform
template <bool INV>
__device__ __forceinline__ float2 tw_at(const float2* __restrict__ tw, int k) {
  const float2 w = __ldg(tw + k);
  return INV ? make_float2(w.x, -w.y) : w;
}

// One Stockham pass of radix R (Govindaraju et al.): butterfly j reads src[j + i n/R], multiplies by
// W_{Ns R}^{i (j mod Ns)}, transforms, and writes dst[expand(j) + i Ns].
template <int R, bool INV>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ src, float2* __restrict__ dst, int n, int Ns,
                                         const float2* __restrict__ tw) {
  const int nb = n / R;
  const int tstep = n / (Ns * R);   // W_{Ns R}^m = W_n^{m tstep}
  for (int j = threadIdx.x; j < nb; j += FFT_THREADS) {
    const int k = j % Ns;
    float2 v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      v[i] = src[j + i * nb];
      if (i > 0 && Ns > 1) v[i] = cmul(v[i], tw_at<INV>(tw, i * k * tstep));
    }
    const int d0 = (j - k) * R + k;
    if constexpr (R == 2) {
      dst[d0] = cadd(v[0], v[1]);
      dst[d0 + Ns] = csub(v[0], v[1]);
    } else if constexpr (R == 4) {
      const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
      // forward: -i d = (d.y, -d.x); inverse: +i d = (-d.y, d.x)
      const float2 jd = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
      dst[d0] = cadd(a, c);
      dst[d0 + Ns] = cadd(b, jd);
      dst[d0 + 2 * Ns] = csub(a, c);
      dst[d0 + 3 * Ns] = csub(b, jd);
    } else {
      // generic odd radix: out[m] = sum_i v[i] W_R^{i m}
      float2 wr[R];
#pragma unroll
      for (int q = 0; q < R; ++q) wr[q] = tw_at<INV>(tw, q * (n / R));
#pragma unroll
      for (int m = 0; m < R; ++m) {
        float2 acc = v[0];
#pragma unroll
        for (int i = 1; i < R; ++i) acc = cadd(acc, cmul(v[i], wr[(i * m) % R]));
        dst[d0 + m * Ns] = acc;
      }
    }
  }
}

// In-CTA complex FFT of `buf0` (n values); returns the buffer holding the result.  All threads of the CTA call it.
template <bool INV>
__device__ float2* fft_cta(const FftPlan& P, float2* buf0, float2* buf1) {
  float2 *src = buf0, *dst = buf1;
  int Ns = 1;
  for (int f = 0; f < P.nfac; ++f) {
    const int R = P.fac[f];
    __syncthreads();
    if (R == 4) fft_pass<4, INV>(src, dst, P.n, Ns, P.tw);
    else if (R == 2) fft_pass<2, INV>(src, dst, P.n, Ns, P.tw);
    else if (R == 3) fft_pass<3, INV>(src, dst, P.n, Ns, P.tw);
    else if (R == 5) fft_pass<5, INV>(src, dst, P.n, Ns, P.tw);
    else fft_pass<7, INV>(src, dst, P.n, Ns, P.tw);
    Ns *= R;
    float2* t = src;
    src = dst;
    dst = t;
  }
  __syncthreads();
  return src;
}

// torch.stft(x, n_fft, hop, window, center=True, pad_mode='reflect', return_complex=True) of rows [rows][L]:
// frames t = 0 .. L / hop, bins 0 .. n_keep - 1 are written.  grid = (ceil(T / 2), rows).
__global__ void __launch_bounds__(FFT_THREADS) stft_kernel(const __grid_constant__ FftPlan P, const float* __restrict__ x,
                                                           int64_t L, int T, int n_keep, float* __restrict__ out,
                                                           SpecStrides S) {
  pdl_enter();
  extern __shared__ float2 fft_smem[];
  float2* buf0 = fft_smem;
  float2* buf1 = fft_smem + P.n;
  const int n = P.n, half = n / 2;
  const int64_t row = blockIdx.y;
  const int ta = 2 * blockIdx.x, tb = ta + 1;
  const float* xr = x + row * L;
  auto sample = [&](int t, int i) -> float {
    int64_t pos = static_cast<int64_t>(t) * P.hop + i - half;
    if (pos < 0) pos = -pos;                   // reflect (no edge repeat), needs half < L
    if (pos >= L) pos = 2 * (L - 1) - pos;
    return xr[pos];
  };
  for (int i = threadIdx.x; i < n; i += FFT_THREADS) {
    const float w = __ldg(P.window + i);
    buf0[i] = make_float2(w * sample(ta, i), tb < T ? w * sample(tb, i) : 0.f);
  }
  const float2* Z = fft_cta<false>(P, buf0, buf1);
  float* oa = out + row * S.row + static_cast<int64_t>(ta) * S.frame;
  float* ob = oa + S.frame;
  for (int k = threadIdx.x; k < n_keep; k += FFT_THREADS) {
    const float2 zk = Z[k];
    const float2 zc = Z[k == 0 ? 0 : n - k];   // conj taken below
    const float2 xa = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
    const float2 d = make_float2(zk.x - zc.x, zk.y + zc.y);        // Z[k] - conj Z[n-k]
    const float2 xb = make_float2(0.5f * d.y, -0.5f * d.x);        // -i d / 2
    oa[k * S.bin] = xa.x;
    oa[k * S.bin + S.reim] = xa.y;
    if (tb < T) {
      ob[k * S.bin] = xb.x;
      ob[k * S.bin + S.reim] = xb.y;
    }
  }
}

// First half of torch.istft: frames[row][t][i] = window[i] * irfft(spec[row][:, t])[i]; bins >= n_keep are zero
// (the freq_pad of ConvTDFNet.istft); the imaginary parts of the DC and Nyquist bins are ignored like in a c2r FFT.
__global__ void __launch_bounds__(FFT_THREADS) istft_frames_kernel(const __grid_constant__ FftPlan P,
                                                                   const float* __restrict__ spec, SpecStrides S, int T,
                                                                   int n_keep, float* __restrict__ frames) {
  pdl_enter();
  extern __shared__ float2 fft_smem[];
  float2* buf0 = fft_smem;
  float2* buf1 = fft_smem + P.n;
  const int n = P.n, half = n / 2;
  const int64_t row = blockIdx.y;
  const int ta = 2 * blockIdx.x, tb = ta + 1;
  const float* sa = spec + row * S.row + static_cast<int64_t>(ta) * S.frame;
  const float* sb = sa + S.frame;
  for (int k = threadIdx.x; k <= half; k += FFT_THREADS) {
    float2 xa = make_float2(0.f, 0.f), xb = xa;
    if (k < n_keep) {
      xa = make_float2(sa[k * S.bin], sa[k * S.bin + S.reim]);
      if (tb < T) xb = make_float2(sb[k * S.bin], sb[k * S.bin + S.reim]);
    }
    if (k == 0 || 2 * k == n) {
      xa.y = 0.f;
      xb.y = 0.f;
    }
    // Z[k] = Xa[k] + i Xb[k];  Z[n-k] = conj(Xa[k]) + i conj(Xb[k])
    buf0[k] = make_float2(xa.x - xb.y, xa.y + xb.x);
    if (k > 0 && 2 * k != n) buf0[n - k] = make_float2(xa.x + xb.y, -xa.y + xb.x);
  }
  const float2* z = fft_cta<true>(P, buf0, buf1);
  const float inv_n = 1.f / static_cast<float>(n);
  float* fa = frames + (row * T + ta) * static_cast<int64_t>(n);
  float* fb = fa + n;
  for (int i = threadIdx.x; i < n; i += FFT_THREADS) {
    const float w = __ldg(P.window + i) * inv_n;
    fa[i] = w * z[i].x;
    if (tb < T) fb[i] = w * z[i].y;
  }
}

// Second half: y[row][m] = sum_t frames[row][t][p - t hop] / sum_t window^2[p - t hop], p = m + n_fft / 2, frames in
// ascending order (fixed summation order); positions beyond the last frame are zero (torch pads when `length` is
// longer than the transform's support).
__global__ void istft_ola_kernel(const float* __restrict__ frames, const float* __restrict__ window, int n, int hop,
                                 int T, int64_t out_len, float* __restrict__ out) {
  pdl_enter();
  const int64_t m = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t row = blockIdx.y;
  if (m >= out_len) return;
  const int64_t p = m + n / 2;
  float acc = 0.f, env = 0.f;
  int64_t t_hi = p / hop;
  int64_t t_lo = p - n + 1 <= 0 ? 0 : (p - n + hop) / hop;   // ceil((p - n + 1) / hop)
  if (t_hi > T - 1) t_hi = T - 1;
  const float* fr = frames + row * T * static_cast<int64_t>(n);
  for (int64_t t = t_lo; t <= t_hi; ++t) {
    const int i = static_cast<int>(p - t * hop);
    const float w = __ldg(window + i);
    acc += fr[t * n + i];
    env = fmaf(w, w, env);
  }
  out[row * out_len + m] = env > 1e-11f ? acc / env : 0.f;
}

}  // namespace tdz
