// Host-side launch sequence of the ERes2NetV2-Large speaker embedder (tdz_embed).  Included by tdz_api.cu.
// Architecture (restated in oracle/eres2netv2_port.py): stem conv -> 16 Res2Net-style
// blocks (1x1 conv, four chained 3x3 convs on channel groups - joined by addition in layers 1-2 and by the
// AFF gate in layers 3-4 -, 1x1 conv, shortcut) -> layer3 downsample + AFF fusion -> TSTP pooling -> Linear.
// BatchNorm (eval) is folded into every convolution by the packer; all convolutions run as tcgen05 GEMMs over
// NHWC pixel-major activations (3x3 convolutions as implicit GEMMs, gemm_conv3.cuh).
#pragma once
#include "gemm_conv3.cuh"

struct SvConv {
  CUtensorMap map;
  const float* bias = nullptr;
  int N = 0, K = 0, BN = 0;
};

struct SvBlockSpec {
  int in_planes, planes, stride, aff, width;
};

struct SvModel {
  bool ready = false;
  const float* stem_w = nullptr;
  const float* stem_b = nullptr;
  SvBlockSpec spec[TDZ_SV_NUM_BLOCKS];
  SvConv conv1[TDZ_SV_NUM_BLOCKS], convs[TDZ_SV_NUM_BLOCKS][4], aff_a[TDZ_SV_NUM_BLOCKS][3],
      aff_b[TDZ_SV_NUM_BLOCKS][3], conv3[TDZ_SV_NUM_BLOCKS], shortcut[TDZ_SV_NUM_BLOCKS];
  bool has_shortcut[TDZ_SV_NUM_BLOCKS];
  SvConv layer3_ds, fuse_a, fuse_b, seg1;
};

static int sv_block_n(int N) { return N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256; }

static void sv_fill_specs(SvBlockSpec* spec) {
  const int nb[4] = {3, 4, 6, 3}, mult[4] = {1, 2, 4, 8}, stride[4] = {1, 2, 2, 2};
  int in_planes = 64, k = 0;
  for (int l = 0; l < 4; ++l) {
    const int planes = 64 * mult[l];
    for (int b = 0; b < nb[l]; ++b, ++k) {
      spec[k].in_planes = in_planes;
      spec[k].planes = planes;
      spec[k].stride = b == 0 ? stride[l] : 1;
      spec[k].aff = l >= 2;
      spec[k].width = planes * 24 / 64;
      in_planes = planes * 4;
    }
  }
}

// W is stored [round_up(N, BN)][round_up(K, 64)] bf16 (zero padded by the packer), bias [round_up(N, BN)].
static int sv_conv_init(tdz_ctx* ctx, SvConv* c, const tdz_conv* w, int N, int K) {
  c->N = N;
  c->K = K;
  c->BN = sv_block_n(N);
  c->bias = w->b;
  const int Np = (N + c->BN - 1) / c->BN * c->BN, Kp = (K + 63) / 64 * 64;
  const uint64_t dims[2] = {static_cast<uint64_t>(Kp), static_cast<uint64_t>(Np)};
  const uint32_t box[2] = {64u, static_cast<uint32_t>(c->BN)};
  return make_tmap(ctx, &c->map, w->w, false, 2, dims, box);
}

static int sv_set_weights(tdz_ctx* ctx, SvModel* M, const tdz_eres2netv2_weights* w) {
  sv_fill_specs(M->spec);
  M->stem_w = w->stem_w;
  M->stem_b = w->stem_b;
  for (int k = 0; k < TDZ_SV_NUM_BLOCKS; ++k) {
    const SvBlockSpec& s = M->spec[k];
    const tdz_eres_block& b = w->blocks[k];
    const int wd = s.width;
    if (sv_conv_init(ctx, &M->conv1[k], &b.conv1, 4 * wd, s.in_planes)) return 1;
    for (int i = 0; i < 4; ++i)
      if (sv_conv_init(ctx, &M->convs[k][i], &b.convs[i], wd, 9 * wd)) return 1;
    if (s.aff)
      for (int i = 0; i < 3; ++i) {
        if (sv_conv_init(ctx, &M->aff_a[k][i], &b.aff_a[i], wd / 4, 2 * wd)) return 1;
        if (sv_conv_init(ctx, &M->aff_b[k][i], &b.aff_b[i], wd, wd / 4)) return 1;
      }
    if (sv_conv_init(ctx, &M->conv3[k], &b.conv3, 4 * s.planes, 4 * wd)) return 1;
    M->has_shortcut[k] = (s.stride != 1 || s.in_planes != 4 * s.planes);
    if (M->has_shortcut[k]) {
      if (b.shortcut.w == nullptr) return fail(ctx, "tdz_set_eres2netv2_weights: block %d needs a shortcut conv", k);
      if (sv_conv_init(ctx, &M->shortcut[k], &b.shortcut, 4 * s.planes, s.in_planes)) return 1;
    }
  }
  if (sv_conv_init(ctx, &M->layer3_ds, &w->layer3_ds, 2048, 9 * 1024)) return 1;
  if (sv_conv_init(ctx, &M->fuse_a, &w->fuse_a, 512, 4096)) return 1;
  if (sv_conv_init(ctx, &M->fuse_b, &w->fuse_b, 2048, 512)) return 1;
  if (sv_conv_init(ctx, &M->seg1, &w->seg1, 192, 40960)) return 1;
  M->ready = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------- workspace
struct SvDims {
  int64_t H[4], W[4], P[4], Pp[4];  // per layer: mel bins, frames, pixels, pixels padded to the GEMM tile
};
static void sv_dims(int64_t N, int64_t frames, SvDims* d) {
  int64_t h = 80, w = frames;
  for (int l = 0; l < 4; ++l) {
    if (l > 0) {
      h = (h + 1) / 2;
      w = (w + 1) / 2;
    }
    d->H[l] = h;
    d->W[l] = w;
    d->P[l] = N * h * w;
    d->Pp[l] = (d->P[l] + 127) / 128 * 128;
  }
}

constexpr int SV_SEG1_SLICES = 64;   // Linear(40960 -> 192): 640 k-blocks in 64 slices of 10 (fixed: batch invariant)
struct SvLayout {
  size_t xa, xb, xs, c1, fused, cat2, mid, col, cat4, res, ds, fcat, fmid, fuse, stats, part, total;
};
static void sv_layout(int64_t N, int64_t frames, SvLayout* L) {
  SvDims d;
  sv_dims(N, frames, &d);
  const int planes[4] = {64, 128, 256, 512};
  size_t act = 0, c1 = 0, sp = 0, col = 0, cat4 = 0, xs = 0, cat2 = 0;
  for (int l = 0; l < 4; ++l) {
    const size_t pp = static_cast<size_t>(d.Pp[l]);
    const size_t wd = planes[l] * 24 / 64;
    act = std::max(act, pp * planes[l] * 4);
    c1 = std::max(c1, pp * wd * 4);
    sp = std::max(sp, pp * wd);
    cat4 = std::max(cat4, pp * wd * 4);
    xs = std::max(xs, pp * (l == 0 ? 64 : planes[l - 1] * 4));
    cat2 = std::max(cat2, pp * wd * 2);
  }
  act = std::max(act, static_cast<size_t>(d.Pp[0]) * 64);
  col = static_cast<size_t>(d.Pp[3]) * 9216;  // only the stride-2 layer3_ds convolution uses an im2col matrix
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 1023) / 1024 * 1024;
    return o;
  };
  L->xa = take(act * 2);     // block input / output (bf16, ping-pong)
  L->xb = take(act * 2);
  L->xs = take(xs * 2);      // stride-2 subsampled block input
  L->c1 = take(c1 * 2);      // conv1 output, 4 groups of `width` channels
  L->fused = take(sp * 2);   // AFF(sp_{i-1}, x_i)
  L->cat2 = take(cat2 * 2);
  L->mid = take(static_cast<size_t>(d.Pp[2]) * 64 * 2);
  L->col = take(col * 2);    // im2col matrix of layer3_ds
  L->cat4 = take(cat4 * 2);  // the four 3x3 conv outputs, concatenated (input of conv3)
  L->res = take(act * 2);    // shortcut conv output
  L->ds = take(static_cast<size_t>(d.Pp[3]) * 2048 * 2);
  L->fcat = take(static_cast<size_t>(d.Pp[3]) * 4096 * 2);
  L->fmid = take(static_cast<size_t>(d.Pp[3]) * 512 * 2);
  L->fuse = take(static_cast<size_t>(d.Pp[3]) * 2048 * 4);
  L->stats = take(static_cast<size_t>((N + 127) / 128 * 128) * 40960 * 2);
  L->part = take(static_cast<size_t>(SV_SEG1_SLICES) * ((N + 127) / 128 * 128) * 256 * 4);  // split-K partial tiles
  L->total = off;
}

// ---------------------------------------------------------------------------------------------- GEMM dispatch
// epilogue kinds: Hardtanh(0,20) -> bf16; SiLU -> bf16; AFF gate -> bf16 / fp32; + residual, Hardtanh -> bf16;
// plain linear -> bf16 / fp32
enum SvEpi { SV_HT20, SV_SILU, SV_AFF, SV_AFF_F32, SV_RES_HT20, SV_LINEAR, SV_LINEAR_F32 };

template <int BN, unsigned EF, int ACT>
static cudaError_t sv_launch(const LinearParams& P, int sms, cudaStream_t st) {
  if constexpr (BN >= 128) {  // wide outputs go through the coalescing panel epilogue
    constexpr int STAGES = BN == 256 ? 3 : 4;
    return launch_gemm<LinearPanel<1, BN, STAGES, EF, ACT, 4>>(P, (P.B * P.Sp / 128) * P.n_tiles, sms, st);
  } else {
    return launch_gemm<LinearGeneric<1, BN, 6, EF, ACT>>(P, (P.B * P.Sp / 128) * P.n_tiles, sms, st);
  }
}

// out = epilogue(A[P][K] @ W^T): A bf16 with leading dimension lda (>= K), P pixels in a Pp-row buffer.
static int sv_gemm(tdz_ctx* ctx, cudaStream_t st, const SvConv& c, const void* A, int lda, int64_t P, int64_t Pp,
                   SvEpi kind, const EpiGeneric& e) {
  LinearParams L;
  memset(&L, 0, sizeof L);
  if (act_map(ctx, &L.tmA, A, false, lda, Pp, 1, 64, 128)) return 1;
  L.tmB = c.map;
  L.B = 1;
  L.Sp = static_cast<int>(Pp);
  L.S = static_cast<int>(P);
  L.N = c.N;
  L.K = c.K;
  L.n_tiles = (c.N + c.BN - 1) / c.BN;
  L.e = e;
  L.e.bias = c.bias;
  const int sms = ctx->num_sms;
  cudaError_t r = cudaErrorInvalidValue;
  constexpr unsigned BF = EF_BIAS | EF_OUT_BF16, F32 = EF_BIAS | EF_OUT_F32, OPS = EF_OPS_BF16;
  switch (kind) {
    case SV_HT20:
      if (c.BN == 32) r = sv_launch<32, BF, ACT_HARDTANH20>(L, sms, st);
      else if (c.BN == 64) r = sv_launch<64, BF, ACT_HARDTANH20>(L, sms, st);
      else if (c.BN == 128) r = sv_launch<128, BF, ACT_HARDTANH20>(L, sms, st);
      else r = sv_launch<256, BF, ACT_HARDTANH20>(L, sms, st);
      break;
    case SV_SILU:
      if (c.BN == 32) r = sv_launch<32, BF, ACT_SILU>(L, sms, st);
      else if (c.BN == 64) r = sv_launch<64, BF, ACT_SILU>(L, sms, st);
      else if (c.BN == 256) r = sv_launch<256, BF, ACT_SILU>(L, sms, st);
      break;
    case SV_AFF:
      if (c.BN == 128) r = sv_launch<128, BF | OPS, ACT_AFF>(L, sms, st);
      else if (c.BN == 256) r = sv_launch<256, BF | OPS, ACT_AFF>(L, sms, st);
      break;
    case SV_AFF_F32:
      if (c.BN == 256) r = sv_launch<256, F32 | OPS, ACT_AFF>(L, sms, st);
      break;
    case SV_RES_HT20:
      if (c.BN == 256) r = sv_launch<256, BF | OPS | EF_RESID_PRE, ACT_HARDTANH20>(L, sms, st);
      break;
    case SV_LINEAR:
      if (c.BN == 256) r = sv_launch<256, BF, ACT_NONE>(L, sms, st);
      break;
    case SV_LINEAR_F32:
      if (c.BN == 256) r = sv_launch<256, F32, ACT_NONE>(L, sms, st);
      break;
  }
  if (r != cudaSuccess) return fail(ctx, "tdz_embed: GEMM launch failed (%s), N=%d K=%d BN=%d kind=%d",
                                    cudaGetErrorString(r), c.N, c.K, c.BN, static_cast<int>(kind));
  return 0;
}

// 3x3 convolution (stride 1, pad 1) + BN + Hardtanh(0,20) as an implicit GEMM (gemm_conv3.cuh); only the stride-2
// layer3_ds convolution still goes through the im2col matrix.
static bool sv_conv3_supported(const SvConv& c) { return c.BN <= 256; }
static int sv_conv3(tdz_ctx* ctx, cudaStream_t st, const SvConv& c, const __nv_bfloat16* a, int lda, int offa, int H,
                    int W, int C, int64_t P, int64_t Pp, __nv_bfloat16* out, int out_ld, int out_col0,
                    const __nv_bfloat16* nx, int nx_ld, int nx_off, __nv_bfloat16* out2, int out2_ld) {
  Conv3Params CP;
  memset(&CP, 0, sizeof CP);
  LinearParams& L = CP.L;
  L.tmB = c.map;
  L.B = 1;
  L.Sp = static_cast<int>(Pp);
  L.S = static_cast<int>(P);
  L.N = c.N;
  L.K = c.K;
  L.n_tiles = (c.N + c.BN - 1) / c.BN;
  L.e.bias = c.bias;
  CP.a = a;
  CP.lda = lda;
  CP.offa = offa;
  CP.H = H;
  CP.W = W;
  CP.C = C;
  CP.out = out;
  CP.out_ld = out_ld;
  CP.out_col0 = out_col0;
  CP.nx = nx;
  CP.nx_ld = nx_ld;
  CP.nx_off = nx_off;
  CP.out2 = out2;
  CP.out2_ld = out2_ld;
  if (c.K != 9 * C || (C & 7) != 0 || c.K / 8 > C3_MAX_KCHUNKS) return fail(ctx, "tdz_embed: conv3 shape K=%d C=%d", c.K, C);
  const int ntiles = static_cast<int>(Pp / 128) * L.n_tiles;
  cudaError_t r;
  if (c.BN == 32) r = launch_gemm_conv3<LinearGeneric<1, 32, 6, 0u, ACT_NONE>>(CP, ntiles, ctx->num_sms, st);
  else if (c.BN == 64) r = launch_gemm_conv3<LinearGeneric<1, 64, 6, 0u, ACT_NONE>>(CP, ntiles, ctx->num_sms, st);
  else if (c.BN == 128) r = launch_gemm_conv3<LinearGeneric<1, 128, 6, 0u, ACT_NONE>>(CP, ntiles, ctx->num_sms, st);
  else r = launch_gemm_conv3<LinearGeneric<1, 256, 4, 0u, ACT_NONE>>(CP, ntiles, ctx->num_sms, st);
  if (r != cudaSuccess) return fail(ctx, "tdz_embed: conv3 launch failed (%s), N=%d K=%d", cudaGetErrorString(r), c.N, c.K);
  return 0;
}

static unsigned sv_grid(int64_t total) { return static_cast<unsigned>((total + 255) / 256); }

// ---------------------------------------------------------------------------------------------- forward
// stop_block >= -1: test hook, copies the bf16 NHWC output of the stem (-1) or of residual block `stop_block`
// (0..15), or the fp32 fuse34 map (16), into `emb` and returns.  stop_block == SV_RUN_ALL runs the whole model.
constexpr int SV_RUN_ALL = 1000;
static int sv_embed(tdz_ctx* ctx, const SvModel& M, const float* feat, int64_t N, int64_t frames, float* emb,
                    void* ws, size_t ws_bytes, cudaStream_t st, int stop_block = SV_RUN_ALL) {
  if (!M.ready) return fail(ctx, "tdz_embed: weights not set");
  if (N <= 0 || frames < 1) return fail(ctx, "tdz_embed: need at least one feature frame per utterance");
  SvDims d;
  sv_dims(N, frames, &d);
  if (d.W[3] < 2 && stop_block == SV_RUN_ALL) {
    // Fewer than 9 fbank frames leave ONE time step after the three stride-2 stages; TSTP's unbiased variance over
    // one step is 0/0, so the reference model returns an all-NaN embedding (dropped by the enrolment path,
    // TargetASR.py:235-236; scored 0.0 by cosine_similarity, :151).  Same result here, without running the network.
    pdl(sv_fill_kernel, sv_grid(N * TDZ_SV_EMBED_DIM), 256, 0, st)(emb, N * TDZ_SV_EMBED_DIM, nanf(""));
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  SvLayout L;
  sv_layout(N, frames, &L);
  if (ws_bytes < L.total) return fail(ctx, "tdz_embed: workspace too small (%zu < %zu)", ws_bytes, L.total);
  if (d.Pp[0] * 256 > 0x7fffffffll) return fail(ctx, "tdz_embed: batch too large for one call");
  uint8_t* base = static_cast<uint8_t*>(ws);
  auto Hh = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(base + o); };
  __nv_bfloat16 *x = Hh(L.xa), *y = Hh(L.xb), *xs = Hh(L.xs), *c1 = Hh(L.c1), *fused = Hh(L.fused),
                *cat2 = Hh(L.cat2), *mid = Hh(L.mid), *col = Hh(L.col), *cat4 = Hh(L.cat4), *res = Hh(L.res),
                *ds = Hh(L.ds), *fcat = Hh(L.fcat), *fmid = Hh(L.fmid), *stats = Hh(L.stats);
  float* fuse = reinterpret_cast<float*>(base + L.fuse);
  const int n = static_cast<int>(N);
  // epilogue operands that are bf16 travel through the float* fields of EpiGeneric (EF_OPS_BF16)
  auto as_f = [](const __nv_bfloat16* p) { return reinterpret_cast<const float*>(p); };

  // columns N..63 of `mid` are never written but are read (against zero weights) by the second AFF conv
  CUDA_OK(cudaMemsetAsync(mid, 0, static_cast<size_t>(d.Pp[2]) * 64 * 2, st));
  // stem
  pdl(sv_stem_kernel, sv_grid(d.P[0] * 8), 256, 0, st)(feat, M.stem_w, M.stem_b, x, n, static_cast<int>(d.H[0]),
                                                      static_cast<int>(d.W[0]));
  if (stop_block == -1) {
    CUDA_OK(cudaMemcpyAsync(emb, x, static_cast<size_t>(d.P[0]) * 64 * 2, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  int layer = 0;
  for (int k = 0; k < TDZ_SV_NUM_BLOCKS; ++k) {
    const SvBlockSpec& s = M.spec[k];
    const int wd = s.width;
    const int lin = layer;
    if (s.stride == 2) ++layer;
    const int64_t P = d.P[layer], Pp = d.Pp[layer];
    const int Hc = static_cast<int>(d.H[layer]), Wc = static_cast<int>(d.W[layer]);
    const __nv_bfloat16* a_in = x;
    if (s.stride == 2) {
      pdl(sv_subsample_kernel, sv_grid(P * (s.in_planes / 8)), 256, 0, st)(
          x, xs, n, static_cast<int>(d.H[lin]), static_cast<int>(d.W[lin]), Hc, Wc, s.in_planes);
      a_in = xs;
    }
    EpiGeneric e;
    // conv1 (1x1) + BN + Hardtanh(0,20)
    memset(&e, 0, sizeof e);
    e.out_bf16 = c1;
    e.out_bf_ld = 4 * wd;
    if (sv_gemm(ctx, st, M.conv1[k], a_in, s.in_planes, P, Pp, SV_HT20, e)) return 1;
    // shortcut
    const __nv_bfloat16* resid = a_in;  // identity (never strided: stride-2 blocks always have a shortcut conv)
    if (M.has_shortcut[k]) {
      memset(&e, 0, sizeof e);
      e.out_bf16 = res;
      e.out_bf_ld = 4 * s.planes;
      if (sv_gemm(ctx, st, M.shortcut[k], a_in, s.in_planes, P, Pp, SV_LINEAR, e)) return 1;
      resid = res;
    }
    // four chained 3x3 convs over the channel groups; conv i writes block i of cat4, which is also sp_i
    const bool implicit = sv_conv3_supported(M.convs[k][0]);
    for (int i = 0; i < 4; ++i) {
      // input of conv i: x_0 | sp_{i-1} + x_i | AFF(sp_{i-1}, x_i)
      const __nv_bfloat16 *ia = c1, *ib = nullptr;
      int ilda = 4 * wd, ioffa = 0, ildb = 0, ioffb = 0;
      if (i > 0 && !s.aff) {
        ia = cat4;
        ioffa = (i - 1) * wd;
        ib = c1;
        ildb = 4 * wd;
        ioffb = i * wd;
      } else if (i > 0) {
        // AFF(sp, x_i): two 1x1 convs on cat(sp, x_i), then the gate in the second epilogue
        pdl(sv_cat2_kernel, sv_grid(P * 2 * (wd / 8)), 256, 0, st)(cat4, 4 * wd, (i - 1) * wd, c1, 4 * wd, i * wd, cat2,
                                                                 P, wd);
        memset(&e, 0, sizeof e);
        e.out_bf16 = mid;
        e.out_bf_ld = 64;
        if (sv_gemm(ctx, st, M.aff_a[k][i - 1], cat2, 2 * wd, P, Pp, SV_SILU, e)) return 1;
        memset(&e, 0, sizeof e);
        e.mul = as_f(cat4 + (i - 1) * wd);
        e.mul_ld = 4 * wd;
        e.resid = as_f(c1 + i * wd);
        e.resid_ld = 4 * wd;
        e.out_bf16 = fused;
        e.out_bf_ld = wd;
        if (sv_gemm(ctx, st, M.aff_b[k][i - 1], mid, 64, P, Pp, SV_AFF, e)) return 1;
        ia = fused;
        ilda = wd;
      }
      memset(&e, 0, sizeof e);
      e.out_bf16 = cat4;
      e.out_bf_ld = 4 * wd;
      e.out_bf_col0 = i * wd;
      if (implicit) {
        // the conv gathers ONE map: x_0, the AFF output, or the sum sp_{i-1} + x_i that conv i-1 left in `fused`
        const bool chain = !s.aff && i < 3;  // this conv also writes sp_i + x_{i+1} for the next one
        // (ping-pong between `fused` and `cat2`, both unused otherwise in a block without AFF: a conv must not
        // overwrite the map its neighbours' halos are still being gathered from)
        __nv_bfloat16* sum_buf[2] = {fused, cat2};
        if (i > 0 && !s.aff) {
          ia = sum_buf[(i - 1) & 1];
          ilda = wd;
          ioffa = 0;
        }
        if (sv_conv3(ctx, st, M.convs[k][i], ia, ilda, ioffa, Hc, Wc, wd, P, Pp, cat4, 4 * wd, i * wd,
                     chain ? c1 : nullptr, 4 * wd, (i + 1) * wd, chain ? sum_buf[i & 1] : nullptr, wd))
          return 1;
      } else {
        pdl(sv_im2col_kernel, dim3(sv_grid(static_cast<int64_t>(Wc) * 9 * (wd / 8)), n * Hc), 256, 0, st)(
            ia, ilda, ioffa, ib, ildb, ioffb, col, n, Hc, Wc, Hc, Wc, wd, 1);
        if (sv_gemm(ctx, st, M.convs[k][i], col, 9 * wd, P, Pp, SV_HT20, e)) return 1;
      }
    }
    // conv3 (1x1) + BN + shortcut + Hardtanh
    memset(&e, 0, sizeof e);
    e.resid = as_f(resid);
    e.resid_ld = 4 * s.planes;
    e.out_bf16 = y;
    e.out_bf_ld = 4 * s.planes;
    if (sv_gemm(ctx, st, M.conv3[k], cat4, 4 * wd, P, Pp, SV_RES_HT20, e)) return 1;
    std::swap(x, y);
    if (k == stop_block) {
      CUDA_OK(cudaMemcpyAsync(emb, x, static_cast<size_t>(P) * 4 * s.planes * 2, cudaMemcpyDeviceToDevice, st));
      return 0;
    }
    if (k == 12) {
      // end of layer3: out3_ds = Conv2d(1024, 2048, 3, stride 2, pad 1)(out3), needed after layer4
      const int64_t P4 = d.P[3], Pp4 = d.Pp[3];
      pdl(sv_im2col_kernel, dim3(sv_grid(d.W[3] * 9 * (1024 / 8)), n * static_cast<int>(d.H[3])), 256, 0, st)(x, 1024, 0, nullptr, 0, 0, col, n, Hc, Wc,
                                                                    static_cast<int>(d.H[3]),
                                                                    static_cast<int>(d.W[3]), 1024, 2);
      memset(&e, 0, sizeof e);
      e.out_bf16 = ds;
      e.out_bf_ld = 2048;
      if (sv_gemm(ctx, st, M.layer3_ds, col, 9216, P4, Pp4, SV_LINEAR, e)) return 1;
    }
  }
  // fuse34 = AFF(out4, out3_ds)
  {
    const int64_t P4 = d.P[3], Pp4 = d.Pp[3];
    EpiGeneric e;
    pdl(sv_cat2_kernel, sv_grid(P4 * 2 * (2048 / 8)), 256, 0, st)(x, 2048, 0, ds, 2048, 0, fcat, P4, 2048);
    memset(&e, 0, sizeof e);
    e.out_bf16 = fmid;
    e.out_bf_ld = 512;
    if (sv_gemm(ctx, st, M.fuse_a, fcat, 4096, P4, Pp4, SV_SILU, e)) return 1;
    memset(&e, 0, sizeof e);
    e.mul = as_f(x);
    e.mul_ld = 2048;
    e.resid = as_f(ds);
    e.resid_ld = 2048;
    e.out_f32 = fuse;
    e.out_ld = 2048;
    if (sv_gemm(ctx, st, M.fuse_b, fmid, 512, P4, Pp4, SV_AFF_F32, e)) return 1;
    if (stop_block == TDZ_SV_NUM_BLOCKS) {
      CUDA_OK(cudaMemcpyAsync(emb, fuse, static_cast<size_t>(P4) * 2048 * 4, cudaMemcpyDeviceToDevice, st));
      return 0;
    }
    // TSTP + embedding Linear
    const int64_t Np = (N + 127) / 128 * 128;
    CUDA_OK(cudaMemsetAsync(stats, 0, static_cast<size_t>(Np) * 40960 * 2, st));
    pdl(sv_tstp_kernel, sv_grid(N * d.H[3] * 2048), 256, 0, st)(fuse, stats, n, static_cast<int>(d.H[3]),
                                                               static_cast<int>(d.W[3]), 2048, 40960);
    {
      float* part = reinterpret_cast<float*>(base + L.part);
      LinearParams LP;
      memset(&LP, 0, sizeof LP);
      if (act_map(ctx, &LP.tmA, stats, false, 40960, Np, 1, 64, 128)) return 1;
      LP.tmB = M.seg1.map;
      LP.B = 1;
      LP.Sp = static_cast<int>(Np);
      LP.S = static_cast<int>(N);
      LP.N = 192;
      LP.K = 40960;
      LP.n_tiles = SV_SEG1_SLICES;
      LP.tps = (40960 / 64 + SV_SEG1_SLICES - 1) / SV_SEG1_SLICES;
      LP.e.out_f32 = part;
      cudaError_t r = launch_gemm<LinearSplitK<4>>(LP, static_cast<int>(Np / 128) * SV_SEG1_SLICES, ctx->num_sms, st);
      if (r != cudaSuccess) return fail(ctx, "tdz_embed: split-K launch failed (%s)", cudaGetErrorString(r));
      pdl(splitk_reduce_kernel, sv_grid(N * 192), 256, 0, st)(part, M.seg1.bias, SV_SEG1_SLICES, Np, n, 192, emb, 192);
    }
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}
