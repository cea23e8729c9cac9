// Chunk stitching (overlap-add / segment gather) and cosine scoring kernels.
#pragma once
#include "ptx.cuh"

namespace tdz {

// look2hear/utils/separator.py:95-112 -- segment i of the zero-padded mixture.
__global__ void gather_segments_kernel(const float* __restrict__ mix, int64_t L, int64_t session, int64_t hop,
                                       int64_t seg_begin, int64_t n_seg, float* __restrict__ seg) {
  pdl_enter();
  const int64_t i4 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n_seg * session) return;
  const int64_t pad = session - hop;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t idx = i4 + k;
    if (idx >= n_seg * session) break;
    const int64_t si = idx / session;
    const int64_t s = idx - si * session;
    const int64_t n = (seg_begin + si) * hop + s - pad;
    seg[idx] = (n >= 0 && n < L) ? mix[n] : 0.f;
  }
}

// look2hear/utils/separator.py:126-130 -- rectangular overlap-add, ascending segment order, then / ratio.
__global__ void stitch_ola_kernel(const float* __restrict__ est, int64_t session, int64_t hop, int64_t seg_begin,
                                  int64_t n_seg, int64_t L, int64_t out_begin, int64_t n_out, float ratio,
                                  float* __restrict__ out) {
  pdl_enter();
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= 2 * n_out) return;
  const int64_t trk = idx / n_out;
  const int64_t k = idx - trk * n_out;
  const int64_t n = out_begin + k;
  float acc = 0.f;
  if (n < L) {
    const int64_t p = n + (session - hop);  // position in the padded stream
    int64_t i_hi = p / hop;
    int64_t i_lo = (p - session + hop) / hop;  // ceil((p - session + 1) / hop) for p - session + 1 > -hop
    if (p - session + 1 <= 0) i_lo = 0;
    for (int64_t i = i_lo; i <= i_hi; ++i) {
      const int64_t li = i - seg_begin;
      const int64_t s = p - i * hop;
      if (li < 0 || li >= n_seg || s < 0 || s >= session) continue;
      acc += est[(li * 2 + trk) * session + s];
    }
    acc = acc / ratio;
  }
  out[trk * n_out + k] = acc;
}

// TargetASR.cosine_similarity (TargetASR.py:144-152): zero vector -> 1.0, clamp to [0,1]. Warp per row.
__global__ void cosine_scores_kernel(const float* __restrict__ emb, const float* __restrict__ target, int N, int dim,
                                     float* __restrict__ scores) {
  pdl_enter();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  float dot = 0.f, na = 0.f, nb = 0.f;
  for (int i = lane; i < dim; i += 32) {
    const float a = emb[static_cast<size_t>(row) * dim + i];
    const float b = target[i];
    dot = fmaf(a, b, dot);
    na = fmaf(a, a, na);
    nb = fmaf(b, b, nb);
  }
  dot = warp_sum(dot);
  na = warp_sum(na);
  nb = warp_sum(nb);
  if (lane == 0) {
    float s;
    if (na == 0.f || nb == 0.f) {
      s = 1.f;
    } else {
      s = dot / (sqrtf(na) * sqrtf(nb));
      s = fminf(fmaxf(s, 0.f), 1.f);
    }
    scores[row] = s;
  }
}

}  // namespace tdz

namespace tdz {

// ---------------------------------------------------------------- BS.1770 loudness (AudioProcessor.meter_loudness)
// K-weighting = two cascaded biquads (pyloudnorm: high shelf + high pass), evaluated in fp64 like scipy.lfilter on
// the host.  The recursion is cut into segments of LK_SEG samples that each start LK_WARM samples early from a
// zero state: the slowest pole pair of the 38 Hz high pass decays by < 1e-13 over the warm-up, so the segments are
// independent to fp64 round-off.  Output: squared filtered samples (fp64), input of the gated block means.
constexpr int LK_SEG = 2048, LK_WARM = 2048;
struct KWeight {
  double b[2][3];
  double a[2][3];  // a[.][0] == 1
};
__global__ void __launch_bounds__(128) kweight_sq_kernel(const float* __restrict__ x, int64_t L, int64_t n_streams,
                                                         KWeight kw, double* __restrict__ ysq) {
  pdl_enter();
  const int64_t seg = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t segs_per_stream = (L + LK_SEG - 1) / LK_SEG;
  if (seg >= segs_per_stream * n_streams) return;
  const int64_t s = seg / segs_per_stream;
  const int64_t n0 = (seg - s * segs_per_stream) * LK_SEG;
  const float* xs = x + s * L;
  double* ys = ysq + s * L;
  double z1a = 0, z2a = 0, z1b = 0, z2b = 0;
  const int64_t start = n0 - LK_WARM < 0 ? 0 : n0 - LK_WARM;
  const int64_t end = n0 + LK_SEG < L ? n0 + LK_SEG : L;
  for (int64_t n = start; n < end; ++n) {
    const double v = static_cast<double>(xs[n]);
    const double y1 = kw.b[0][0] * v + z1a;           // direct form II transposed, as scipy.signal.lfilter
    z1a = kw.b[0][1] * v - kw.a[0][1] * y1 + z2a;
    z2a = kw.b[0][2] * v - kw.a[0][2] * y1;
    const double y2 = kw.b[1][0] * y1 + z1b;
    z1b = kw.b[1][1] * y1 - kw.a[1][1] * y2 + z2b;
    z2b = kw.b[1][2] * y1 - kw.a[1][2] * y2;
    if (n >= n0) ys[n] = y2 * y2;
  }
}

// z[s][j] = sum(ysq[s][lo[j] : hi[j]]) * inv_len   (400 ms blocks; the bounds come from the host so that they are
// the same integers pyloudnorm's float arithmetic produces).  One warp per block.
__global__ void __launch_bounds__(256) loudness_blocks_kernel(const double* __restrict__ ysq, int64_t L,
                                                              const int64_t* __restrict__ lo,
                                                              const int64_t* __restrict__ hi, int64_t nblk,
                                                              int64_t n_streams, double inv_len,
                                                              double* __restrict__ z) {
  pdl_enter();
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nblk * n_streams) return;
  const int64_t s = w / nblk, j = w - s * nblk;
  const double* ys = ysq + s * L;
  double acc = 0;
  for (int64_t n = lo[j] + lane; n < hi[j]; n += 32) acc += ys[n];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) z[w] = acc * inv_len;
}

}  // namespace tdz
