// Host-side launch sequence of the Apollo restorer (tdz_apollo_restore) and the STFT entry points.  Included by
// tdz_api.cu.  Reference: look2hear/models/apollo.py (Apollo.forward :278-297, BSNet :186-212, Roformer :120-141,
// ConvActNorm1d :143-170); the dataflow is restated in token-major form by oracle/apollo_port.py.
#pragma once
#include "kernels_apollo.cuh"
#include "kernels_stft.cuh"

struct ApModel {
  bool ready = false;
  tdz_apollo_weights w;
  FftPlan plan;
  struct LayerMaps {
    CUtensorMap qkv, out, mlp1, mlp2, w1[3], w2[3];
    CUtensorMap w1_128[3];   // 128-row boxes: hidden chunks of the back-to-back ConvActNorm1d GEMM
    CUtensorMap w1_64[3], w2_128[3];   // its cta_group::2 form: each CTA stages half of every weight tile
  } lm[TDZ_AP_LAYERS];
};

static int stft_plan_init(tdz_ctx* ctx, const tdz_stft_plan* p, FftPlan* f) {
  if (!p || p->n_fft < 4 || (p->n_fft & 1) || p->hop <= 0 || !p->window_dev || !p->twiddle_dev)
    return fail(ctx, "stft plan: need an even n_fft, a positive hop and the window / twiddle tables");
  if (!fft_factorize(p->n_fft, f)) return fail(ctx, "stft plan: n_fft %d has a prime factor other than 2, 3, 5, 7", p->n_fft);
  f->hop = p->hop;
  f->window = p->window_dev;
  f->tw = reinterpret_cast<const float2*>(p->twiddle_dev);
  if (2 * p->n_fft * 8 > 200 * 1024) return fail(ctx, "stft plan: n_fft %d does not fit the shared-memory FFT", p->n_fft);
  return 0;
}

static int stft_launch(tdz_ctx* ctx, const FftPlan& f, const float* x, int64_t rows, int64_t L, int64_t n_keep,
                       float* spec, const SpecStrides& S, cudaStream_t st) {
  if (rows <= 0 || L <= f.n / 2) return fail(ctx, "tdz_stft: need rows > 0 and more than n_fft / 2 samples per row");
  if (n_keep <= 0 || n_keep > f.n / 2 + 1) return fail(ctx, "tdz_stft: n_keep outside 1 .. n_fft / 2 + 1");
  if (rows > 65535) return fail(ctx, "tdz_stft: more than 65535 rows in one call");
  const int64_t T = 1 + L / f.hop;
  const int smem = 2 * f.n * 8;
  static std::atomic<unsigned long long> configured{0};
  CUDA_OK(set_max_smem_once(reinterpret_cast<const void*>(stft_kernel), 200 * 1024, configured));
  pdl(stft_kernel, dim3(static_cast<unsigned>((T + 1) / 2), static_cast<unsigned>(rows)), FFT_THREADS, smem, st)(
      f, x, L, static_cast<int>(T), static_cast<int>(n_keep), spec, S);
  CUDA_OK(cudaGetLastError());
  return 0;
}

static int istft_launch(tdz_ctx* ctx, const FftPlan& f, const float* spec, int64_t rows, int64_t T, int64_t n_keep,
                        const SpecStrides& S, float* frames, float* out, int64_t out_len, cudaStream_t st) {
  if (rows <= 0 || T <= 0 || out_len <= 0) return fail(ctx, "tdz_istft: empty input");
  if (n_keep <= 0 || n_keep > f.n / 2 + 1) return fail(ctx, "tdz_istft: n_keep outside 1 .. n_fft / 2 + 1");
  if (rows > 65535) return fail(ctx, "tdz_istft: more than 65535 rows in one call");
  const int smem = 2 * f.n * 8;
  static std::atomic<unsigned long long> configured{0};
  CUDA_OK(set_max_smem_once(reinterpret_cast<const void*>(istft_frames_kernel), 200 * 1024, configured));
  pdl(istft_frames_kernel, dim3(static_cast<unsigned>((T + 1) / 2), static_cast<unsigned>(rows)), FFT_THREADS, smem, st)(
      f, spec, S, static_cast<int>(T), static_cast<int>(n_keep), frames);
  pdl(istft_ola_kernel, dim3(static_cast<unsigned>((out_len + 255) / 256), static_cast<unsigned>(rows)), 256, 0, st)(
      frames, f.window, f.n, f.hop, static_cast<int>(T), out_len, out);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------- weights
static int ap_set_weights(tdz_ctx* ctx, ApModel* M, const tdz_apollo_weights* w) {
  M->w = *w;
  if (stft_plan_init(ctx, &w->plan, &M->plan)) return 1;
  if (M->plan.n != 882 || M->plan.hop != 441) return fail(ctx, "tdz_set_apollo_weights: the plan must be n_fft 882 / hop 441");
  for (int l = 0; l < TDZ_AP_LAYERS; ++l) {
    const tdz_apollo_layer& L = w->layers[l];
    auto& m = M->lm[l];
    if (w_map(ctx, &m.qkv, L.w_qkv, false, 768, 256, 256)) return 1;
    if (w_map(ctx, &m.out, L.w_out, false, 256, 256, 256)) return 1;
    if (w_map(ctx, &m.mlp1, L.w_mlp1, false, 2048, 256, 128)) return 1;   // split-N: two 128-row boxes per tile
    if (w_map(ctx, &m.mlp2, L.w_mlp2, false, 256, 1024, 256)) return 1;
    for (int b = 0; b < 3; ++b) {
      if (w_map(ctx, &m.w1[b], L.icb[b].w1, false, 1024, 256, 256)) return 1;
      if (w_map(ctx, &m.w2[b], L.icb[b].w2, false, 256, 1024, 256)) return 1;
      if (w_map(ctx, &m.w1_128[b], L.icb[b].w1, false, 1024, 256, 128)) return 1;
      if (w_map(ctx, &m.w1_64[b], L.icb[b].w1, false, 1024, 256, 64)) return 1;
      if (w_map(ctx, &m.w2_128[b], L.icb[b].w2, false, 256, 1024, 128)) return 1;
    }
  }
  M->ready = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------- workspace
// Fixed part: spectrogram, estimated spectrogram, inverse-STFT frames (10.6 KB per frame).  Token part (6 160 B per
// token, 80 tokens per frame): sized for `Tl` frames per row.  Only the depthwise k7 convolutions mix frames - 3 taps
// each side x 3 blocks x 6 layers = 54 frames of receptive field - so a long input is processed in frame chunks with a
// 54-frame halo on either side, the halo results discarded: same bits as one pass (every token's arithmetic is the
// same), bounded memory (an hour of 44.1 kHz audio would need 180 GB of token buffers in one pass).
constexpr int AP_HALO = 54;
constexpr size_t AP_TOKEN_BYTES = 256 * 4 + 256 * 2 + 4 * 4 + 768 * 2 + 256 * 2 + 1024 * 2 + 256 * 2;
struct ApLayout {
  size_t spec, est, frames, x, xbf, ss, qkv, att, h, u, total;
  int64_t T, Tl, Mp;   // frames per row, frames per row the token buffers hold, token rows padded to the GEMM tile
};
static size_t ap_fixed_bytes(int64_t rows, int64_t T) {
  const size_t fr = static_cast<size_t>(rows * T);
  auto up = [](size_t b) { return (b + 1023) / 1024 * 1024; };
  return 2 * up(fr * AP_BINS * 8) + up(fr * 882 * 4);
}
static size_t ap_token_bytes(int64_t rows, int64_t Tl) {
  const size_t m = static_cast<size_t>((rows * Tl * AP_NBAND + 127) / 128 * 128);
  return m * AP_TOKEN_BYTES + 7 * 1024;
}
static void ap_layout(int64_t rows, int64_t nsample, int64_t Tl, ApLayout* L) {
  const int64_t T = 1 + nsample / 441;
  L->T = T;
  L->Tl = Tl;
  L->Mp = (rows * Tl * AP_NBAND + 127) / 128 * 128;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 1023) / 1024 * 1024;
    return o;
  };
  const size_t fr = static_cast<size_t>(rows * T), m = static_cast<size_t>(L->Mp);
  L->spec = take(fr * AP_BINS * 8);
  L->est = take(fr * AP_BINS * 8);
  L->frames = take(fr * 882 * 4);
  L->x = take(m * 256 * 4);      // residual stream, fp32
  L->xbf = take(m * 256 * 2);    // its bf16 copy (GEMM operand)
  L->ss = take(m * 4 * 4);       // per-row partial sums of squares of x (RMSNorm)
  L->qkv = take(m * 768 * 2);
  L->att = take(m * 256 * 2);
  L->h = take(m * 1024 * 2);     // gated-MLP hidden activations
  L->u = take(m * 256 * 2);      // dwconv7 + RMSNorm output
  L->total = off;
}
// Largest number of frames per row the token buffers may hold in `ws_bytes` (0: not even the smallest chunk fits)
static int64_t ap_frames_that_fit(int64_t rows, int64_t T, size_t ws_bytes) {
  const size_t fixed = ap_fixed_bytes(rows, T);
  if (ws_bytes <= fixed) return 0;
  int64_t Tl = static_cast<int64_t>((ws_bytes - fixed) / (static_cast<size_t>(rows) * AP_NBAND * AP_TOKEN_BYTES));
  while (Tl > 0 && fixed + ap_token_bytes(rows, Tl) > ws_bytes) --Tl;
  if (Tl >= T) return T;
  return Tl > 2 * AP_HALO ? Tl : 0;
}

// ---------------------------------------------------------------------------------------------- GEMMs
static void ap_lin(tdz_ctx* ctx, LinearParams& P, const CUtensorMap& tmA, const CUtensorMap& tmB, int64_t tokens,
                   int64_t Mp, int N, int K) {
  memset(&P, 0, sizeof P);
  P.tmA = tmA;
  P.tmB = tmB;
  P.B = 1;
  P.Sp = static_cast<int>(Mp);
  P.S = static_cast<int>(tokens);
  P.N = N;
  P.K = K;
  P.n_tiles = (N + 255) / 256;
  (void)ctx;
}

enum ApTap { AP_TAP_SPEC = 0, AP_TAP_FEAT = 1, AP_TAP_ATT0 = 2, AP_TAP_BAND0 = 3, AP_TAP_LAYER0 = 4, AP_TAP_EST = 10,
             AP_RUN_ALL = 1000 };

__global__ void ap_widen_kernel(const __nv_bfloat16* __restrict__ a, float* __restrict__ o, int64_t n) {
  pdl_enter();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) o[i] = __bfloat162float(a[i]);
}

static int ap_forward(tdz_ctx* ctx, const ApModel& M, const float* wav, int64_t rows, int64_t nsample, float* out,
                      void* ws, size_t ws_bytes, cudaStream_t st, int tap = AP_RUN_ALL) {
  if (!M.ready) return fail(ctx, "tdz_apollo_restore: weights not set");
  if (rows <= 0 || nsample <= 441) return fail(ctx, "tdz_apollo_restore: need more than 441 samples per row (reflect padding)");
  const int64_t T = 1 + nsample / 441;
  const int64_t Tl_max = ap_frames_that_fit(rows, T, ws_bytes);
  if (Tl_max == 0)
    return fail(ctx, "tdz_apollo_restore: workspace too small (%zu bytes; the smallest chunked form needs %zu)", ws_bytes,
                ap_fixed_bytes(rows, T) + ap_token_bytes(rows, 2 * AP_HALO + 1));
  if (Tl_max < T && tap != AP_RUN_ALL) return fail(ctx, "tdz_apollo_debug: the taps need a workspace for the whole input");
  ApLayout L;
  ap_layout(rows, nsample, Tl_max, &L);
  if (L.Mp > 0x7fffff00ll) return fail(ctx, "tdz_apollo_restore: chunk too long for one call");
  uint8_t* base = static_cast<uint8_t*>(ws);
  float2* spec = reinterpret_cast<float2*>(base + L.spec);
  float2* est = reinterpret_cast<float2*>(base + L.est);
  float* frames = reinterpret_cast<float*>(base + L.frames);
  float* x = reinterpret_cast<float*>(base + L.x);
  __nv_bfloat16* xbf = reinterpret_cast<__nv_bfloat16*>(base + L.xbf);
  float* ss = reinterpret_cast<float*>(base + L.ss);
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(base + L.qkv);
  __nv_bfloat16* att = reinterpret_cast<__nv_bfloat16*>(base + L.att);
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(base + L.h);
  __nv_bfloat16* u = reinterpret_cast<__nv_bfloat16*>(base + L.u);
  const tdz_apollo_weights& W = M.w;
  const int64_t nframes = rows * T;
  const int sms = ctx->num_sms;
  auto copy_out = [&](const void* src, size_t bytes) -> int {
    CUDA_OK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToDevice, st));
    return 0;
  };

  // spectrogram, frame-major complex
  const SpecStrides SS{T * AP_BINS * 2, 2, AP_BINS * 2, 1};
  if (stft_launch(ctx, M.plan, wav, rows, nsample, AP_BINS, reinterpret_cast<float*>(spec), SS, st)) return 1;
  if (tap == AP_TAP_SPEC) return copy_out(spec, static_cast<size_t>(nframes) * AP_BINS * 8);
  static std::atomic<unsigned long long> merge_cfg{0};
  constexpr int merge_smem = AP_NBAND * AP_XLD * 4;
  CUDA_OK(set_max_smem_once(reinterpret_cast<const void*>(ap_bandmerge_kernel), merge_smem, merge_cfg));
  // frame chunks: results are final for frames [k0, k1), computed from frames [t_lo, t_hi) = the chunk + its halo
  const int64_t keep_max = L.Tl >= T ? T : L.Tl - 2 * AP_HALO;
  for (int64_t k0 = 0; k0 < T; k0 += keep_max) {
  const int64_t k1 = std::min(T, k0 + keep_max);
  const int64_t t_lo = std::max<int64_t>(0, k0 - AP_HALO), t_hi = std::min(T, k1 + AP_HALO);
  const int64_t Tc = t_hi - t_lo;                 // frames per row in the token buffers for this chunk
  const int64_t tokens = rows * Tc * AP_NBAND, Mp = (tokens + 127) / 128 * 128;
  const int mtiles = static_cast<int>(Mp / 128);
  // band split + per-band bottleneck
  pdl(ap_bandsplit_kernel, static_cast<unsigned>(rows * Tc), 256, 0, st)(spec, W.bn_g, W.bn_w, W.bn_b, x, xbf, ss,
                                                                       static_cast<int>(T), static_cast<int>(t_lo),
                                                                       static_cast<int>(Tc));
  CUDA_OK(cudaGetLastError());
  if (tap == AP_TAP_FEAT) return copy_out(x, static_cast<size_t>(tokens) * 256 * 4);

  CUtensorMap m_xbf, m_att, m_h, m_u;
  if (act_map(ctx, &m_xbf, xbf, false, 256, Mp, 1, 64, 128)) return 1;
  if (act_map(ctx, &m_att, att, false, 256, Mp, 1, 64, 128)) return 1;
  if (act_map(ctx, &m_h, h, false, 1024, Mp, 1, 64, 128)) return 1;
  if (act_map(ctx, &m_u, u, false, 256, Mp, 1, 64, 128)) return 1;
  const int runs = static_cast<int>((Tc + AP_DW_RUN - 1) / AP_DW_RUN);
  const int64_t dw_warps = rows * runs * AP_NBAND;

  for (int l = 0; l < TDZ_AP_LAYERS; ++l) {
    const auto& m = M.lm[l];
    const tdz_apollo_layer& LW = W.layers[l];
    LinearParams P;
    // ---- band_net: Roformer over the 80 bands of every frame
    ap_lin(ctx, P, m_xbf, m.qkv, tokens, Mp, 768, 256);
    P.e.ss_in = ss;
    P.e.ss_dim_rsqrt = 1.f / 256.f;
    P.e.out_bf16 = qkv;
    P.e.out_bf_ld = 768;
    CUDA_OK((launch_gemm<LinearPanel<1, 256, 3, EF_RMS4 | EF_OUT_BF16, ACT_NONE, 4>>(P, mtiles * P.n_tiles, sms, st)));
    pdl(ap_attn_kernel, static_cast<unsigned>(rows * Tc), 256, 0, st)(qkv, W.rot_cos, W.rot_sin, att);
    CUDA_OK(cudaGetLastError());
    if (l == 0 && tap == AP_TAP_ATT0) {
      pdl(ap_widen_kernel, static_cast<unsigned>((tokens * 256 + 255) / 256), 256, 0, st)(att, out, tokens * 256);
      CUDA_OK(cudaGetLastError());
      return 0;
    }
    ap_lin(ctx, P, m_att, m.out, tokens, Mp, 256, 256);      // output conv + residual (in place)
    P.e.resid = x;
    P.e.resid_ld = 256;
    P.e.out_f32 = x;
    P.e.out_ld = 256;
    P.e.out_bf16 = xbf;
    P.e.out_bf_ld = 256;
    P.e.ss_out = ss;
    P.e.ss_out_ld = 4;
    CUDA_OK((launch_gemm<LinearPanel<1, 256, 3, EF_RESID | EF_OUT_F32 | EF_OUT_BF16 | EF_SS_OUT, ACT_NONE, 4>>(
        P, mtiles * P.n_tiles, sms, st)));
    ap_lin(ctx, P, m_xbf, m.mlp1, tokens, Mp, 2048, 256);    // gated MLP, first conv
    P.split_n = 1024;
    P.e.ss_in = ss;
    P.e.ss_dim_rsqrt = 1.f / 256.f;
    P.e.out_bf16 = h;
    P.e.out_bf_ld = 1024;
    CUDA_OK((launch_gemm<LinearGLU<4>>(P, mtiles * P.n_tiles, sms, st)));
    ap_lin(ctx, P, m_h, m.mlp2, tokens, Mp, 256, 1024);      // second conv + residual
    P.e.resid = x;
    P.e.resid_ld = 256;
    P.e.out_f32 = x;
    P.e.out_ld = 256;
    CUDA_OK((launch_gemm<LinearPanel<1, 256, 3, EF_RESID | EF_OUT_F32, ACT_NONE, 4>>(P, mtiles * P.n_tiles, sms, st)));
    if (l == 0 && tap == AP_TAP_BAND0) return copy_out(x, static_cast<size_t>(tokens) * 256 * 4);
    // ---- seq_net: three ConvActNorm1d blocks along time
    for (int b = 0; b < 3; ++b) {
      const tdz_apollo_icb& I = LW.icb[b];
      pdl(ap_dwconv_rms_kernel, static_cast<unsigned>((dw_warps * 32 + 255) / 256), 256, 0, st)(
          x, I.dw, I.dw_b, static_cast<int>(Tc), runs, dw_warps, u);
      CUDA_OK(cudaGetLastError());
      if (!ctx->no_b2b) {
        // Conv1d(256, 1024) -> SiLU -> Conv1d(1024, 256) + residual as one back-to-back GEMM (gemm_b2b.cuh): the
        // 1 024-wide hidden activations never leave the SM
        B2bParams Q;
        memset(&Q, 0, sizeof Q);
        Q.tmX = m_u;
        Q.tmW1 = m.w1_128[b];
        Q.tmW2 = m.w2[b];
        Q.B = 1;
        Q.Sp = static_cast<int>(Mp);
        Q.S = static_cast<int>(tokens);
        Q.H = 1024;
        Q.bias1 = I.b1;
        Q.e.bias = I.b2;
        Q.e.resid = x;
        Q.e.resid_ld = 256;
        Q.e.out_f32 = x;
        Q.e.out_ld = 256;
        constexpr unsigned EF_MID = EF_BIAS | EF_RESID | EF_OUT_F32, EF_LAST = EF_MID | EF_OUT_BF16 | EF_SS_OUT;
        if (b == 2) {  // the layer output also feeds the next Roformer: bf16 copy + RMSNorm sums
          Q.e.out_bf16 = xbf;
          Q.e.out_bf_ld = 256;
          Q.e.ss_out = ss;
          Q.e.ss_out_ld = 4;
        }
        if (!ctx->b2b_cg2) {
          if (b < 2) CUDA_OK((launch_gemm_b2b<ACT_SILU, EF_MID, 1, 1024, false>(Q, sms, st)));
          else CUDA_OK((launch_gemm_b2b<ACT_SILU, EF_LAST, 1, 1024, false>(Q, sms, st)));
        } else {
          Q.tmW1 = m.w1_64[b];
          Q.tmW2 = m.w2_128[b];
          if (b < 2) CUDA_OK((launch_gemm_b2b<ACT_SILU, EF_MID, 1, 1024, true>(Q, sms, st)));
          else CUDA_OK((launch_gemm_b2b<ACT_SILU, EF_LAST, 1, 1024, true>(Q, sms, st)));
        }
        continue;
      }
      ap_lin(ctx, P, m_u, m.w1[b], tokens, Mp, 1024, 256);
      P.e.bias = I.b1;
      P.e.out_bf16 = h;
      P.e.out_bf_ld = 1024;
      CUDA_OK((launch_gemm<LinearPanel<1, 256, 3, EF_BIAS | EF_OUT_BF16, ACT_SILU, 4>>(P, mtiles * P.n_tiles, sms, st)));
      ap_lin(ctx, P, m_h, m.w2[b], tokens, Mp, 256, 1024);
      P.e.bias = I.b2;
      P.e.resid = x;
      P.e.resid_ld = 256;
      P.e.out_f32 = x;
      P.e.out_ld = 256;
      if (b < 2) {
        CUDA_OK((launch_gemm<LinearPanel<1, 256, 3, EF_BIAS | EF_RESID | EF_OUT_F32, ACT_NONE, 4>>(
            P, mtiles * P.n_tiles, sms, st)));
      } else {  // the layer output also feeds the next Roformer: bf16 copy + RMSNorm sums
        P.e.out_bf16 = xbf;
        P.e.out_bf_ld = 256;
        P.e.ss_out = ss;
        P.e.ss_out_ld = 4;
        CUDA_OK((launch_gemm<LinearPanel<1, 256, 3, EF_BIAS | EF_RESID | EF_OUT_F32 | EF_OUT_BF16 | EF_SS_OUT, ACT_NONE, 4>>(
            P, mtiles * P.n_tiles, sms, st)));
      }
    }
    if (tap == AP_TAP_LAYER0 + l) return copy_out(x, static_cast<size_t>(tokens) * 256 * 4);
  }
  // band merge -> estimated spectrogram (the chunk's own frames only)
  pdl(ap_bandmerge_kernel, static_cast<unsigned>(rows * (k1 - k0)), 256, merge_smem, st)(
      x, W.out_g, W.out_wv, W.out_wg, W.out_bv, W.out_bg, est, static_cast<int>(T), static_cast<int>(t_lo),
      static_cast<int>(Tc), static_cast<int>(k0 - t_lo), static_cast<int>(k1 - k0));
  CUDA_OK(cudaGetLastError());
  }  // frame chunks
  if (tap == AP_TAP_EST) return copy_out(est, static_cast<size_t>(nframes) * AP_BINS * 8);
  return istft_launch(ctx, M.plan, reinterpret_cast<const float*>(est), rows, T, AP_BINS, SS, frames, out, nsample, st);
}
