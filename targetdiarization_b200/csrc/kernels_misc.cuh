// Chunk stitching (overlap-add / segment gather) and cosine scoring kernels.
#pragma once
#include "ptx.cuh"

namespace tdz {

// look2hear/utils/separator.py:95-112 -- segment i of the zero-padded mixture.
__global__ void gather_segments_kernel(const float* __restrict__ mix, int64_t L, int64_t session, int64_t hop,
                                       int64_t seg_begin, int64_t n_seg, float* __restrict__ seg) {
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
