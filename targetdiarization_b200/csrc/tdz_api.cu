// libtdz.so: C-ABI entry points (include/tdz.h) and the host-side launch sequence of the separator.
// One tdz_separate() call = the whole MossFormer2 forward (look2hear/models/mossformer2.py:563-589 of the
// reference) on B chunks, launched on the caller's stream, no allocation, no host synchronisation.
#include "../../include/tdz.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>

#include "gemm_cfgs.cuh"
#include "gemm_b2b.cuh"
#include "gemm_convt.cuh"
#include "kernels_fbank.cuh"
#include "kernels_misc.cuh"
#include "kernels_sep.cuh"
#include "kernels_sv.cuh"

#include <algorithm>

using namespace tdz;

// ------------------------------------------------------------------------------------------------ ctx
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct tdz_ctx {
  int device = 0;
  int num_sms = 0;
  EncodeTiledFn encode = nullptr;
  std::string err;
  std::mutex mu;
  bool have_sep = false;
  // development switch, read once from the environment when the handle is created: TDZ_CONVT_SINGLE runs the
  // to_out / to_u|to_v conv GEMMs without cta_group::2 (A/B measurements)
  bool convt_single = false;
  bool no_b2b = false;  // TDZ_NO_B2B: fsmn.linear / fsmn.project as two kernels (A/B measurements)
  bool b2b_cg2 = false; // TDZ_B2B_CG2: the back-to-back GEMMs in their cta_group::2 form (measured 6 % slower: off)
  int dd_seg_len = 0;   // TDZ_DD_SEG_LEN: forces the DilatedDenseNet segment length (A/B measurements)
  tdz_mossformer2_weights sep;
  // weight tensor maps (built once per tdz_set_mossformer2_weights)
  struct LayerMaps {
    CUtensorMap w_in, w_out, w_c1, w_uv, w_lin, w_proj, w_c2;
    CUtensorMap w_in128, w_out128, w_uv128;  // 128-row boxes: M operand of the channel-major conv GEMMs
    CUtensorMap w_lin128;                    // 128-row boxes: hidden chunks of the back-to-back linear -> project GEMM
    CUtensorMap w_lin64, w_proj128;          // the cta_group::2 form of it: each CTA stages half of every weight tile
  } lm[TDZ_NUM_LAYERS];
  CUtensorMap m_enc1x1, m_out1, m_tg, m_dec1, m_dec;
  bool have_fbank = false;
  FbankTables fb;
  struct SvModel* sv = nullptr;
  struct ApModel* ap = nullptr;
};

static int fail(tdz_ctx* c, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return 1;
}
#define CUDA_OK(call)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                        \
  } while (0)

// A handle is bound to the device it was created on: the compute entry points make that device current for the
// duration of the call (and restore the caller's), so handles on several GPUs can be driven from one thread.
struct DeviceScope {
  int prev = -1;
  bool switched = false;
  explicit DeviceScope(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceScope() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};

extern "C" const char* tdz_version(void) { return "tdz 0.1 (sm_100a)"; }

extern "C" int tdz_create(int device, tdz_ctx** out) {
  if (!out) return 1;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return 2;
  DeviceScope dev_scope(device);  // the caller's current device is restored on return
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != device) return 3;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 4;
  if (prop.major != 10) return 5;  // hand-written for sm_100a; no fallback
  tdz_ctx* c = new tdz_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    delete c;
    return 6;
  }
  c->encode = reinterpret_cast<EncodeTiledFn>(fn);
  c->convt_single = getenv("TDZ_CONVT_SINGLE") != nullptr;
  c->no_b2b = getenv("TDZ_NO_B2B") != nullptr;
  c->b2b_cg2 = getenv("TDZ_B2B_CG2") != nullptr;
  if (const char* e = getenv("TDZ_DD_SEG_LEN")) c->dd_seg_len = atoi(e);
  *out = c;
  return 0;
}
static void sv_free(tdz_ctx* ctx);
static void ap_free(tdz_ctx* ctx);
extern "C" void tdz_destroy(tdz_ctx* ctx) {
  if (ctx) sv_free(ctx);
  if (ctx) ap_free(ctx);
  delete ctx;
}
extern "C" const char* tdz_last_error(tdz_ctx* ctx) { return ctx ? ctx->err.c_str() : "null handle"; }
extern "C" int tdz_num_sms(tdz_ctx* ctx) { return ctx ? ctx->num_sms : 0; }

// ------------------------------------------------------------------------------------------------ tensor maps
// rank-2/3 row-major tensor, innermost dimension contiguous, 128 B swizzle.
static int make_tmap(tdz_ctx* ctx, CUtensorMap* m, const void* ptr, bool f32, int rank, const uint64_t* dims,
                     const uint32_t* box, bool swizzle = true) {
  const uint64_t es = f32 ? 4 : 2;
  cuuint64_t gdim[3] = {1, 1, 1};
  cuuint64_t gstr[2] = {0, 0};
  cuuint32_t bx[3] = {1, 1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  uint64_t stride = es;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    stride *= dims[i];
    if (i < rank - 1) gstr[i] = stride;
  }
  if (swizzle && bx[0] * es > 128) return fail(ctx, "tensor map inner box exceeds the 128 B swizzle span");
  CUresult r = ctx->encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                           const_cast<void*>(ptr), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return 0;
}
// activation [B][Sp][C] -> 3-D map {C, Sp, B}
static int act_map(tdz_ctx* ctx, CUtensorMap* m, const void* ptr, bool f32, int C, int64_t Sp, int64_t B,
                   uint32_t box_c, uint32_t box_rows) {
  const uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(Sp), static_cast<uint64_t>(B)};
  const uint32_t box[3] = {box_c, box_rows, 1};
  return make_tmap(ctx, m, ptr, f32, 3, dims, box);
}
// fp32 activation [B][Sp][256] read in un-swizzled {128 channels, 64 rows} chunks (DilatedDenseNet ring)
static int dd_map(tdz_ctx* ctx, CUtensorMap* m, const void* ptr, int64_t Sp, int64_t B) {
  const uint64_t dims[3] = {256, static_cast<uint64_t>(Sp), static_cast<uint64_t>(B)};
  const uint32_t box[3] = {DD_CH, DD_CHUNK, 1};
  return make_tmap(ctx, m, ptr, true, 3, dims, box, false);
}
// weight [N][K] -> 2-D map {K, N}
static int w_map(tdz_ctx* ctx, CUtensorMap* m, const void* ptr, bool f32, int N, int K, uint32_t box_rows) {
  const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
  const uint32_t box[2] = {f32 ? 32u : 64u, box_rows};
  return make_tmap(ctx, m, ptr, f32, 2, dims, box);
}

extern "C" int tdz_set_mossformer2_weights(tdz_ctx* ctx, const tdz_mossformer2_weights* w) {
  if (!ctx || !w) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->sep = *w;
  for (int i = 0; i < TDZ_NUM_LAYERS; ++i) {
    const tdz_layer_weights& L = w->layers[i];
    auto& M = ctx->lm[i];
    if (w_map(ctx, &M.w_in, L.w_in, false, 2176, 512, 256)) return 1;
    if (w_map(ctx, &M.w_out, L.w_out, false, 512, 1024, 256)) return 1;
    if (w_map(ctx, &M.w_c1, L.w_c1, true, 256, 512, 256)) return 1;
    if (w_map(ctx, &M.w_uv, L.w_uv, false, 512, 256, 256)) return 1;
    if (w_map(ctx, &M.w_lin, L.w_lin, false, 256, 256, 256)) return 1;
    if (w_map(ctx, &M.w_proj, L.w_proj, false, 256, 256, 256)) return 1;
    if (w_map(ctx, &M.w_c2, L.w_c2, true, 512, 256, 256)) return 1;
    if (w_map(ctx, &M.w_in128, L.w_in, false, 2176, 512, 128)) return 1;
    if (w_map(ctx, &M.w_out128, L.w_out, false, 512, 1024, 128)) return 1;
    if (w_map(ctx, &M.w_uv128, L.w_uv, false, 512, 256, 128)) return 1;
    if (w_map(ctx, &M.w_lin128, L.w_lin, false, 256, 256, 128)) return 1;
    if (w_map(ctx, &M.w_lin64, L.w_lin, false, 256, 256, 64)) return 1;
    if (w_map(ctx, &M.w_proj128, L.w_proj, false, 256, 256, 128)) return 1;
  }
  if (w_map(ctx, &ctx->m_enc1x1, w->w_enc1x1, true, 512, 512, 256)) return 1;
  if (w_map(ctx, &ctx->m_out1, w->w_out1, true, 1024, 512, 256)) return 1;
  if (w_map(ctx, &ctx->m_tg, w->w_tg, true, 1024, 512, 128)) return 1;  // split-N: two 128-row boxes per tile
  if (w_map(ctx, &ctx->m_dec1, w->w_dec1, true, 512, 512, 256)) return 1;
  if (w_map(ctx, &ctx->m_dec, w->dec_wt, true, 16, 512, 16)) return 1;
  ctx->have_sep = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------ layout
constexpr size_t TDZ_HRS_FRONT = 8;
extern "C" int64_t tdz_num_frames(int64_t T) { return T < 16 ? 0 : (T - 16) / 8 + 1; }
extern "C" int64_t tdz_padded_frames(int64_t T) {
  const int64_t S = tdz_num_frames(T);
  return (S + 255) / 256 * 256;
}

extern "C" int tdz_separate_layout(int64_t B, int64_t T, int num_sms, tdz_sep_layout* L) {
  if (!L || B <= 0 || T < 16) return 1;
  memset(L, 0, sizeof *L);
  const int64_t S = tdz_num_frames(T), Sp = tdz_padded_frames(T), M = B * Sp;
  L->S = S;
  L->Sp = Sp;
  L->Mtot = M;
  // lin_kv GEMM: the frame axis is cut into fixed spans of 2048 frames (32 k-blocks) whatever the batch size, so
  // that the summation order - and with it every output bit - does not depend on how chunks are batched.
  (void)num_sms;
  const int total_kb = static_cast<int>(Sp / 64);
  // (chunks of up to 2048 frames - the streaming shape - are cut into 512-frame spans: with one span the batch-1 call
  // was 8 CTAs walking 20 k-blocks each)
  const int kbps = Sp <= 2048 ? 8 : 32;
  const int nsplit = (total_kb + kbps - 1) / kbps;
  L->kv_nsplit = nsplit;
  L->kv_kb_per_split = kbps;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 1023) / 1024 * 1024;
    return o;
  };
  const size_t m = static_cast<size_t>(M);
  L->enc = take(m * 512 * 4);
  L->x0 = take(m * 512 * 4);
  L->x = take(m * 512 * 4);
  L->xbf = take(m * 512 * 2);
  L->ss = take(m * 8 * 4);     // ScaleNorm partial sums: 8 x 64 channels per frame
  // one region, two views: the per-layer intermediates, and (after the 24 layers) the mask-head buffers
  const size_t region = off;
  L->vu = take(m * 2048 * 2);
  L->qk4 = take(m * 512 * 2);
  L->lq_lo = off;  // (no longer used: lin_q is stored as fp16 inside qk4)
  L->qkf = take(m * 128 * 4);
  L->P = take(m * 256 * 2);
  L->o = take(m * 1024 * 2);
  L->o_ss = take(m * 32 * 4);
  L->c = take(m * 256 * 4);
  L->nhat = take(m * 256 * 2);
  L->xuv = take(m * 512 * 4);
  L->xubf = take(m * 256 * 2);
  L->f1 = take(m * 256 * 2);
  L->p = take(m * 256 * 4);
  L->y1 = take(m * 256 * 4);
  L->y2 = take(m * 256 * 4);
  L->g = take(m * 256 * 4);
  const size_t layer_end = off;
  off = region;
  L->lnb = take(m * 512 * 4);
  L->ab = take(m * 512 * 4);
  L->mb = take(m * 1024 * 4);
  L->gated = take(m * 1024 * 4);
  L->sep = take(m * 1024 * 4);
  off = std::max(off, layer_end);
  L->kv_part = take(static_cast<size_t>(B) * nsplit * 128 * 2048 * 4);
  L->kv = take(static_cast<size_t>(B) * 128 * 2048 * 2);           // lin_kv | lin_ku, fp16
  L->gn_stats = take(static_cast<size_t>(B) * 2 * 8 * 2);       // two GroupNorms
  L->in_stats = take(static_cast<size_t>(B) * 256 * 2 * 8 * 2); // two InstanceNorms (re-zeroed per layer)
  L->in_ss = take(static_cast<size_t>(B) * 256 * 8 * 2);        // their (scale, shift) tables
  L->samp = take(static_cast<size_t>(B) * 4 * 4);               // A,B for each GroupNorm
  L->rot = take(static_cast<size_t>(Sp) * 16 * 8);
  L->pos = take(static_cast<size_t>(Sp) * 512 * 4);                // ScaledSinuEmbedding rows (shared by the batch)
  // per-frame ScaleNorm scale of the next conv GEMM; its epilogue reads whole 96-frame runs around a tile without
  // clamping (frames outside a sample are masked afterwards): TDZ_HRS_FRONT floats before and 256 after stay readable
  L->hrs = take((m + TDZ_HRS_FRONT + 256) * 4);
  L->total = off;
  return 0;
}

extern "C" size_t tdz_separate_workspace_bytes(int64_t B, int64_t T) {
  tdz_sep_layout L;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (tdz_separate_layout(B, T, sms, &L)) return 0;
  return L.total;
}

// ------------------------------------------------------------------------------------------------ forward

// Launch steps, selectable through tdz_separate_debug (tests drive single kernels with oracle inputs).
enum Step : int {
  ST_ENCODER = 0, ST_ENC1X1, ST_FLASH_IN, ST_SIM, ST_KV, ST_ATT_OUT, ST_TO_OUT, ST_FSMN_C1, ST_FSMN_UV, ST_FSMN_LIN,
  ST_FSMN_PROJ, ST_DD1, ST_DD2, ST_FSMN_TAIL, ST_FSMN_C2, ST_FINAL_LN, ST_FINAL_GN, ST_OUT1, ST_TANHSIG, ST_DEC1,
  ST_DECODER, ST_COUNT
};

// Calls of at most this many padded frames (the shapes the host side replays as CUDA graphs) launch their kernels
// with the programmatic-dependent-launch attribute (ptx.cuh); 4 s of fbank frames per utterance for the embedder.
constexpr size_t TDZ_PDL_MAX_FRAMES = 32768;
constexpr int64_t TDZ_PDL_MAX_SV_FRAMES = 4096;

static int run_separate(tdz_ctx* ctx, const float* mix, int64_t B64, int64_t T64, float* out, int64_t out_cs,
                        int64_t out_ss, void* ws, size_t ws_bytes, cudaStream_t st, int num_layers, int step_lo,
                        int step_hi) {
#define STEP(k) if ((k) >= step_lo && (k) <= step_hi)
  if (!ctx->have_sep) return fail(ctx, "tdz_separate: weights not set");
  if (B64 <= 0 || T64 < 16 || B64 > 65535) return fail(ctx, "tdz_separate: bad shape B=%lld T=%lld", (long long)B64, (long long)T64);
  tdz_sep_layout L;
  tdz_separate_layout(B64, T64, ctx->num_sms, &L);
  if (ws_bytes < L.total) return fail(ctx, "tdz_separate: workspace too small (%zu < %zu)", ws_bytes, L.total);
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return fail(ctx, "tdz_separate: workspace must be 1024 B aligned");
  const int B = static_cast<int>(B64), T = static_cast<int>(T64);
  const int S = static_cast<int>(L.S), Sp = static_cast<int>(L.Sp);
  const size_t M = static_cast<size_t>(L.Mtot);
  if (M > (1ull << 24)) return fail(ctx, "tdz_separate: batch too large for one call (more than 2^24 padded frames)");
  PdlScope pdl_scope(M <= TDZ_PDL_MAX_FRAMES);  // small (latency-bound) calls: programmatic dependent launches
  const int sms = ctx->num_sms;
  const tdz_mossformer2_weights& W = ctx->sep;
  uint8_t* base = static_cast<uint8_t*>(ws);
  auto F = [&](size_t o) { return reinterpret_cast<float*>(base + o); };
  auto H = [&](size_t o) { return reinterpret_cast<__nv_bfloat16*>(base + o); };
  float *enc = F(L.enc), *x0 = F(L.x0), *x = F(L.x), *ss = F(L.ss), *o_ss = F(L.o_ss), *c = F(L.c), *xuv = F(L.xuv),
        *p = F(L.p), *y1 = F(L.y1), *y2 = F(L.y2), *g = F(L.g), *kv_part = F(L.kv_part), *samp = F(L.samp), *qkf = F(L.qkf);
  __nv_bfloat16 *xbf = H(L.xbf), *vu = H(L.vu), *qk4 = H(L.qk4), *Pm = H(L.P), *o = H(L.o), *nhat = H(L.nhat),
                *xubf = H(L.xubf), *f1 = H(L.f1);
  __half* kv = reinterpret_cast<__half*>(base + L.kv);
  double* gn_stats = reinterpret_cast<double*>(base + L.gn_stats);
  double* in_stats = reinterpret_cast<double*>(base + L.in_stats);
  float2* in_ss = reinterpret_cast<float2*>(base + L.in_ss);
  float2* rot = reinterpret_cast<float2*>(base + L.rot);
  float* pos_tab = F(L.pos);
  float* hrs = F(L.hrs) + TDZ_HRS_FRONT;
  // mask-head buffers (second view of the layer region, used after the layer loop)
  float *lnb = F(L.lnb), *ab = F(L.ab), *mb = F(L.mb), *gated = F(L.gated), *sep = F(L.sep);

  // ---- activation tensor maps (buffers are reused by every layer)
  CUtensorMap m_enc, m_xbf, m_x, m_o, m_nhat, m_xubf, m_f1, m_g, m_ab, m_mb, m_gated0, m_gated1;
  CUtensorMap m_xbf256, m_o256, m_nhat256;  // 256-frame boxes: N operand of the channel-major conv GEMMs
  if (act_map(ctx, &m_xbf256, xbf, false, 512, Sp, B, 64, 256)) return 1;
  if (act_map(ctx, &m_o256, o, false, 1024, Sp, B, 64, 256)) return 1;
  if (act_map(ctx, &m_nhat256, nhat, false, 256, Sp, B, 64, 256)) return 1;
  const int tps_t = (S + CT_ROWS - 1) / CT_ROWS;
  if (act_map(ctx, &m_enc, enc, true, 512, Sp, B, 32, 128)) return 1;
  if (act_map(ctx, &m_xbf, xbf, false, 512, Sp, B, 64, 128)) return 1;
  if (act_map(ctx, &m_x, x, true, 512, Sp, B, 32, 128)) return 1;
  if (act_map(ctx, &m_o, o, false, 1024, Sp, B, 64, 128)) return 1;
  if (act_map(ctx, &m_nhat, nhat, false, 256, Sp, B, 64, 128)) return 1;
  if (act_map(ctx, &m_xubf, xubf, false, 256, Sp, B, 64, 128)) return 1;
  if (act_map(ctx, &m_f1, f1, false, 256, Sp, B, 64, 128)) return 1;
  if (act_map(ctx, &m_g, g, true, 256, Sp, B, 32, 128)) return 1;
  if (act_map(ctx, &m_ab, ab, true, 512, Sp, B, 32, 128)) return 1;
  if (act_map(ctx, &m_mb, mb, true, 1024, Sp, B, 32, 128)) return 1;
  if (act_map(ctx, &m_gated0, gated, true, 512, Sp, B, 32, 128)) return 1;
  if (act_map(ctx, &m_gated1, gated + M * 512, true, 512, Sp, B, 32, 128)) return 1;
  AttnParams AP;
  memset(&AP, 0, sizeof AP);
  if (act_map(ctx, &AP.tmQK, qk4, false, 512, Sp, B, 64, 128)) return 1;
  if (act_map(ctx, &AP.tmQKb, qk4, false, 512, Sp, B, 64, 256)) return 1;
  if (act_map(ctx, &AP.tmQKmn, qk4, false, 512, Sp, B, 64, 64)) return 1;
  if (act_map(ctx, &AP.tmVUmn, vu, false, 2048, Sp, B, 64, 64)) return 1;
  if (act_map(ctx, &AP.tmP, Pm, false, 256, Sp, B, 64, 128)) return 1;
  if (act_map(ctx, &AP.tmKVmn, kv, false, 2048, 128, B, 64, 64)) return 1;  // (fp16: same 16-bit element size)
  AP.B = B;
  AP.Sp = Sp;
  AP.S = S;
  AP.nsplit = L.kv_nsplit;
  AP.kb_per_split = L.kv_kb_per_split;
  AP.P = Pm;
  AP.kv_part = kv_part;
  AP.vu = vu;
  AP.o = o;
  AP.o_ss = o_ss;
  // DilatedDenseNet streaming kernels: time segments long enough to amortise the 76-row halo (7 %), short enough
  // to give every SM a few CTAs even for a single 10 s chunk
  CUtensorMap m_p, m_y1;
  if (dd_map(ctx, &m_p, p, Sp, B)) return 1;
  if (dd_map(ctx, &m_y1, y1, Sp, B)) return 1;
  DdParams dd;
  memset(&dd, 0, sizeof dd);
  dd.B = B;
  dd.Sp = Sp;
  dd.S = S;
  // a function of the chunk length alone: per-segment fp32 partial sums must not depend on the batch (bit-identical
  // batching).  Short chunks (streaming: 600 ms = 1 199 frames) are cut finer - with 1024-frame segments a batch-1 call
  // ran 8 / 16 CTAs that each walked up to 16 steps in sequence (30 / 37 us per launch, 19 % of the batch-1 step).
  dd.seg_len = S <= 2048 ? 256 : 1024;
  if (ctx->dd_seg_len >= 64) dd.seg_len = ctx->dd_seg_len / 64 * 64;
  dd.nseg = (S + dd.seg_len - 1) / dd.seg_len;
  {
    static std::atomic<unsigned long long> dd1_configured{0}, dd2_configured{0};
    CUDA_OK(set_max_smem_once(reinterpret_cast<const void*>(dd_stream_kernel<1>), DD_SMEM_BYTES, dd1_configured));
    CUDA_OK(set_max_smem_once(reinterpret_cast<const void*>(dd_stream_kernel<2>), DD_SMEM_BYTES, dd2_configured));
  }

  const int mtiles = static_cast<int>(M / 128);
  const int tps = (S + CONV_ROWS - 1) / CONV_ROWS;
  auto lin_base = [&](LinearParams& P, const CUtensorMap& a, const CUtensorMap& w, int N, int K, int block_n) {
    memset(&P, 0, sizeof P);
    P.tmA = a;
    P.tmB = w;
    P.B = B;
    P.Sp = Sp;
    P.S = S;
    P.N = N;
    P.K = K;
    P.n_tiles = (N + block_n - 1) / block_n;
    P.tps = tps;
  };

  // ---- front: encoder -> GroupNorm -> conv1d_encoder (+pos enc)   (mossformer2.py:573,487-496)
  STEP(ST_ENCODER) {
    CUDA_OK(cudaMemsetAsync(gn_stats, 0, static_cast<size_t>(B) * 4 * 8, st));
    if (Sp > S) {
      // padded frames of the attention operands stay zero for the whole forward (nobody writes them later)
      CUDA_OK(cudaMemset2DAsync(vu + static_cast<size_t>(S) * 2048, static_cast<size_t>(Sp) * 2048 * 2, 0,
                                static_cast<size_t>(Sp - S) * 2048 * 2, B, st));
      CUDA_OK(cudaMemset2DAsync(qk4 + static_cast<size_t>(S) * 512, static_cast<size_t>(Sp) * 512 * 2, 0,
                                static_cast<size_t>(Sp - S) * 512 * 2, B, st));
    }
    pdl(encoder_kernel, B * (Sp / ENC_FRAMES), 512, 0, st)(mix, T, W.enc_w, enc, gn_stats, B, Sp, S);
    pdl(gn_finalize_kernel, (B + 127) / 128, 128, 0, st)(gn_stats, samp, samp + B, B, 512.0 * S, 1e-8);
    pdl(rotary_table_kernel, (Sp * 16 + 255) / 256, 256, 0, st)(W.rot_freqs, rot, Sp);
  }
  STEP(ST_ENC1X1) {
    pdl(posenc_table_kernel, (Sp * 512 + 255) / 256, 256, 0, st)(W.pos_inv_freq, W.pos_scale, pos_tab, Sp, 512);
    LinearParams P;
    lin_base(P, m_enc, ctx->m_enc1x1, 512, 512, 256);
    P.e.sampA = samp;
    P.e.sampB = samp + B;
    P.e.colsum = W.enc1x1_colsum;
    P.e.bias = W.enc1x1_bias;
    P.e.pos_tab = pos_tab;
    P.e.out_f32 = x0;
    P.e.out_ld = 512;
    P.e.out_bf16 = xbf;
    P.e.out_bf_ld = 512;
    P.e.ss_out = ss;
    P.e.ss_out_ld = 8;
    CUDA_OK((launch_gemm<LinearPanel<2, 256, 3,
                                       EF_SAMP | EF_BIAS | EF_POS | EF_OUT_F32 | EF_OUT_BF16 | EF_SS_OUT | EF_ZERO_PAD,
                                       ACT_NONE, 4>>(P, mtiles * P.n_tiles, sms, st)));
  }

  for (int li = 0; li < num_layers; ++li) {
    const tdz_layer_weights& LW = W.layers[li];
    const auto& LM = ctx->lm[li];
    const float* x_in = (li == 0) ? x0 : x;
    // ---------------- FLASH_ShareA_FFConvM (mossformer_block.py:191-220)
    STEP(ST_FLASH_IN) {  // token shift + ScaleNorm + to_hidden|to_qk Linear + SiLU + ConvModule + OffsetScale/rotary
      LinearParams P;
      lin_base(P, m_xbf, LM.w_in, 2176, 512, 256);
      P.shift_kblocks = 4;
      P.e.ss_in = ss;
      P.e.ss_dim_rsqrt = 0.044194173824159216f;  // 512^-0.5
      P.e.bias = LW.b_in;
      P.cv.dw_t = LW.dw_in;
      P.cv.ldw = 2176;
      P.cv.vu = vu;
      P.cv.qkf = qkf;
      pdl(rowscale_kernel<true>, static_cast<unsigned>((M + 255) / 256), 256, 0, st)(ss, hrs, Sp, S, M,
                                                                                   0.044194173824159216f);
      P.e.ss_in = hrs;
      P.tmA = m_xbf256;
      P.tmB = LM.w_in128;
      P.n_tiles = 17;
      P.tps = tps_t;
      CUDA_OK((launch_gemm_convt<CONV_VUQK>(P, B * tps_t * P.n_tiles, sms, st)));
      pdl(qk_heads_kernel, static_cast<unsigned>((M + 4 * QKH_FRAMES - 1) / (4 * QKH_FRAMES)), 256, 0, st)(qkf, LW.os_gamma, LW.os_beta, rot, qk4, Sp, S, M);
    }
    STEP(ST_SIM) CUDA_OK((launch_gemm<AttnSim>(AP, mtiles, sms, st)));
    STEP(ST_KV) {
      CUDA_OK((launch_gemm<AttnKV>(AP, B * AP.nsplit * 8, sms, st)));
      const size_t total4 = static_cast<size_t>(B) * 128 * 2048 / 4;
      pdl(kv_reduce_kernel, static_cast<unsigned>((total4 + 255) / 256), 256, 0, st)(
          kv_part, kv, AP.nsplit, 1.f / static_cast<float>(S), static_cast<size_t>(128) * 2048, total4);
    }
    STEP(ST_ATT_OUT) {
      CUDA_OK((launch_gemm_cg2<AttnOut>(AP, (mtiles / 2) * 8, sms, st)));  // one cta_group::2 MMA per CTA pair
    }
    STEP(ST_TO_OUT) {  // ScaleNorm(1024) + to_out Linear + SiLU + ConvModule + FLASH residual (mossformer_block.py:219)
      LinearParams P;
      lin_base(P, m_o, LM.w_out, 512, 1024, 256);
      P.e.ss_in = o_ss;
      P.e.ss_dim_rsqrt = 0.03125f;  // 1024^-0.5
      P.e.bias = LW.b_out;
      P.cv.dw_t = LW.dw_out;
      P.cv.ldw = 512;
      P.cv.x_in = x_in;
      P.cv.x_out = x;
      pdl(rowscale_kernel<false>, static_cast<unsigned>((M + 255) / 256), 256, 0, st)(o_ss, hrs, Sp, S, M, 0.03125f);
      P.e.ss_in = hrs;
      P.tmA = m_o256;
      P.tmAh = m_o;
      P.tmB = LM.w_out128;
      P.n_tiles = 4;
      P.tps = tps_t;
      // K = 1024: two thirds of a tile's L2 -> SM bytes are operands, so the CTA pair shares the X tile
      // (cta_group::2, M = 256 channels); the convt_single development switch selects the single-CTA form
      if (ctx->convt_single) {
        CUDA_OK((launch_gemm_convt<CONV_RESX>(P, B * tps_t * 4, sms, st)));
      } else {
        CUDA_OK((launch_gemm_convt_cg2<CONV_RESX>(P, B * tps_t * 2, sms, st)));
      }
    }
    // ---------------- GatedFSMNBlockDilated (mossformer_block.py:419-425)
    STEP(ST_FSMN_C1) {  // conv1 + PReLU + norm1 + inner LayerNorm statistics
      LinearParams P;
      lin_base(P, m_x, LM.w_c1, 256, 512, 256);
      P.e.bias = LW.b_c1;
      P.alpha = LW.prelu_c1;
      P.ln_g1 = LW.ln1_g;
      P.ln_b1 = LW.ln1_b;
      P.e.out_f32 = c;
      P.e.out_bf16 = nhat;
      CUDA_OK((launch_gemm<LinearLN256<2, 3>>(P, mtiles, sms, st)));
    }
    STEP(ST_FSMN_UV) {  // to_u | to_v: Linear + SiLU + ConvModule
      LinearParams P;
      lin_base(P, m_nhat, LM.w_uv, 512, 256, 256);
      P.e.bias = LW.b_uv;
      P.cv.dw_t = LW.dw_uv;
      P.cv.ldw = 512;
      P.cv.xuv = xuv;
      P.cv.xubf = xubf;
      P.tmA = m_nhat256;
      P.tmAh = m_nhat;
      P.tmB = LM.w_uv128;
      P.n_tiles = 4;
      P.tps = tps_t;
      if (ctx->convt_single) {
        CUDA_OK((launch_gemm_convt<CONV_UV>(P, B * tps_t * 4, sms, st)));
      } else {
        CUDA_OK((launch_gemm_convt_cg2<CONV_UV>(P, B * tps_t * 2, sms, st)));
      }
    }
    // fsmn.linear -> ReLU -> fsmn.project (fsmn.py:131-139) as ONE back-to-back GEMM whose hidden activations stay on
    // chip; the two-kernel form below runs when a test asks for one of the two steps alone (and under TDZ_NO_B2B) and
    // gives the same bits
    const bool b2b = step_lo <= ST_FSMN_LIN && step_hi >= ST_FSMN_PROJ && !ctx->no_b2b;
    if (b2b) {
      B2bParams Q;
      memset(&Q, 0, sizeof Q);
      Q.tmX = m_xubf;
      Q.tmW1 = LM.w_lin128;
      Q.tmW2 = LM.w_proj;
      Q.B = B;
      Q.Sp = Sp;
      Q.S = S;
      Q.H = 256;
      Q.bias1 = LW.b_lin;
      Q.e.out_f32 = p;
      Q.e.out_ld = 256;
      if (!ctx->b2b_cg2) {
        CUDA_OK((launch_gemm_b2b<ACT_RELU, EF_OUT_F32, 2, 256, false>(Q, sms, st)));
      } else {
        Q.tmW1 = LM.w_lin64;
        Q.tmW2 = LM.w_proj128;
        CUDA_OK((launch_gemm_b2b<ACT_RELU, EF_OUT_F32, 2, 256, true>(Q, sms, st)));
      }
    }
    if (!b2b) STEP(ST_FSMN_LIN) {  // fsmn.linear + ReLU
      LinearParams P;
      lin_base(P, m_xubf, LM.w_lin, 256, 256, 256);
      P.e.bias = LW.b_lin;
      P.e.out_bf16 = f1;
      P.e.out_bf_ld = 256;
      CUDA_OK((launch_gemm<LinearPanel<1, 256, 3, EF_BIAS | EF_OUT_BF16 | EF_ZERO_PAD, ACT_RELU, 4>>(P, mtiles, sms,
                                                                                                    st)));
    }
    if (!b2b) STEP(ST_FSMN_PROJ) {  // fsmn.project
      LinearParams P;
      lin_base(P, m_f1, LM.w_proj, 256, 256, 256);
      P.e.out_f32 = p;
      P.e.out_ld = 256;
      CUDA_OK((launch_gemm<LinearPanel<1, 256, 3, EF_OUT_F32, ACT_NONE, 4>>(P, mtiles, sms, st)));
    }
    double* st1 = in_stats;
    double* st2 = in_stats + static_cast<size_t>(B) * 512;
    STEP(ST_DD1) {
      CUDA_OK(cudaMemsetAsync(in_stats, 0, static_cast<size_t>(B) * 256 * 2 * 8 * 2, st));
      DdParams D = dd;
      D.taps = LW.dd_w1;
      D.out = y1;
      D.stats = st1;
      D.tmA = m_p;
      pdl(dd_stream_kernel<1>, B * dd.nseg * (256 / DD_CH), DD_THREADS, DD_SMEM_BYTES, st)(D);
    }
    float2* in_ss1 = in_ss;
    float2* in_ss2 = in_ss + static_cast<size_t>(B) * 256;
    STEP(ST_DD2) {
      pdl(in_finalize_kernel, B, 256, 0, st)(st1, LW.in1_g, LW.in1_b, in_ss1, B * 256, static_cast<double>(S));
      DdParams D = dd;
      D.taps = LW.dd_w2;
      D.in_ss = in_ss1;
      D.prelu = LW.dd_prelu1;
      D.out = y2;
      D.stats = st2;
      D.tmA = m_y1;
      D.tmB = m_p;
      pdl(dd_stream_kernel<2>, B * dd.nseg * (512 / DD_CH), DD_THREADS, DD_SMEM_BYTES, st)(D);
    }
    STEP(ST_FSMN_TAIL) {
      pdl(in_finalize_kernel, B, 256, 0, st)(st2, LW.in2_g, LW.in2_b, in_ss2, B * 256, static_cast<double>(S));
      if (M <= 32768)
        pdl(fsmn_tail_kernel<1>, static_cast<unsigned>((M * 32 + 255) / 256), 256, 0, st)(y2, in_ss2, LW.dd_prelu2, xuv, c, g, B,
                                                                                      Sp, S);
      else
        pdl(fsmn_tail_kernel<8>, static_cast<unsigned>((M / 8 * 32 + 255) / 256), 256, 0, st)(y2, in_ss2, LW.dd_prelu2, xuv, c,
                                                                                          g, B, Sp, S);
    }
    STEP(ST_FSMN_C2) {  // conv2 + residual; also the bf16 copy and ScaleNorm sums the next FLASH layer needs
      LinearParams P;
      lin_base(P, m_g, LM.w_c2, 512, 256, 256);
      P.e.bias = LW.b_c2;
      P.e.resid = x;
      P.e.resid_ld = 512;
      P.e.out_f32 = x;
      P.e.out_ld = 512;
      P.e.out_bf16 = xbf;
      P.e.out_bf_ld = 512;
      P.e.ss_out = ss;
      P.e.ss_out_ld = 8;
      CUDA_OK((launch_gemm<LinearPanel<2, 256, 3,
                                         EF_BIAS | EF_RESID | EF_OUT_F32 | EF_OUT_BF16 | EF_SS_OUT | EF_ZERO_PAD,
                                         ACT_NONE, 4>>(P, mtiles * P.n_tiles, sms, st)));
    }
    CUDA_OK(cudaGetLastError());
  }
  // ---- back: LayerNorm -> GroupNorm + skip -> PReLU -> mask head -> decoder (mossformer2.py:320,388-396,500-523,575-589)
  const float* x_fin = (num_layers == 0) ? x0 : x;
  STEP(ST_FINAL_LN) {
    CUDA_OK(cudaMemsetAsync(gn_stats + 2 * B, 0, static_cast<size_t>(B) * 2 * 8, st));
    pdl(final_ln_kernel, static_cast<unsigned>(M / 8), 256, 0, st)(x_fin, W.fln_g, W.fln_b, lnb, gn_stats + 2 * B, B, Sp,
                                                                  S);
    pdl(gn_finalize_kernel, (B + 127) / 128, 128, 0, st)(gn_stats + 2 * B, samp + 2 * B, samp + 3 * B, B, 512.0 * S,
                                                        1e-8);
  }
  STEP(ST_FINAL_GN) {
    const size_t total4 = M * 512 / 4;
    pdl(final_gn_kernel, static_cast<unsigned>((total4 + 255) / 256), 256, 0, st)(
        lnb, samp + 2 * B, samp + 3 * B, W.fgn_g, W.fgn_b, x0, W.mask_prelu, ab, Sp, S, total4);
  }
  STEP(ST_OUT1) {  // conv1d_out 512 -> 1024 (+bias)
    LinearParams P;
    lin_base(P, m_ab, ctx->m_out1, 1024, 512, 256);
    P.e.bias = W.b_out1;
    P.e.out_f32 = mb;
    P.e.out_ld = 1024;
    // m is only the tf32 operand of the two gate convs: rounded to nearest tf32 on store
    CUDA_OK((launch_gemm<LinearPanel<2, 256, 3, EF_BIAS | EF_OUT_F32 | EF_ZERO_PAD | EF_ROUND_TF32, ACT_NONE, 4>>(
        P, mtiles * P.n_tiles, sms, st)));
  }
  for (int spk = 0; spk < 2; ++spk) {
    STEP(ST_TANHSIG) {  // tanh(output) * sigmoid(output_gate)
      LinearParams P;
      lin_base(P, m_mb, ctx->m_tg, 1024, 512, 256);
      P.a_k0 = spk * 512;
      P.split_n = 512;
      P.e.bias = W.b_tg;
      P.e.out_f32 = gated + static_cast<size_t>(spk) * M * 512;
      P.e.out_ld = 512;
      CUDA_OK((launch_gemm<LinearTanhSig<2, 4>>(P, mtiles * P.n_tiles, sms, st)));
    }
    STEP(ST_DEC1) {  // conv1_decoder + ReLU, times the encoder output
      LinearParams P;
      lin_base(P, spk == 0 ? m_gated0 : m_gated1, ctx->m_dec1, 512, 512, 256);
      P.e.mul = enc;
      P.e.mul_ld = 512;
      P.e.out_f32 = sep + static_cast<size_t>(spk) * M * 512;
      P.e.out_ld = 512;
      // relu(mask) * encoder output: only the tf32 operand of the decoder GEMM -> rounded to nearest tf32 on store
      CUDA_OK((launch_gemm<LinearPanel<2, 256, 3, EF_MUL | EF_OUT_F32 | EF_ZERO_PAD | EF_ROUND_TF32, ACT_RELU, 4>>(
          P, mtiles * P.n_tiles, sms, st)));
    }
  }
  STEP(ST_DECODER) {
    DecParams D;
    memset(&D, 0, sizeof D);
    if (act_map(ctx, &D.tmA, sep, true, 512, Sp, 2 * B64, 32, 128)) return 1;
    D.tmB = ctx->m_dec;
    D.out = out;
    D.out_cs = out_cs;
    D.out_ss = out_ss;
    D.B = B;
    D.Sp = Sp;
    D.S = S;
    D.T = T;
    D.tps = (S + 1 + DEC_TILE_FRAMES - 1) / DEC_TILE_FRAMES;
    // samples from 8 (S + 1) on (at most 7: T < 8 S + 16) are the zero pad of mossformer2.py:585-586
    if (static_cast<int64_t>(S + 1) * 8 < T64) {
      for (int spk = 0; spk < 2; ++spk)
        CUDA_OK(cudaMemset2DAsync(out + spk * out_ss + static_cast<int64_t>(S + 1) * 8, static_cast<size_t>(out_cs) * 4, 0,
                                  static_cast<size_t>(T64 - static_cast<int64_t>(S + 1) * 8) * 4, B, st));
    }
    CUDA_OK((launch_gemm<DecoderGemm>(D, 2 * B * D.tps, sms, st)));
  }
  CUDA_OK(cudaGetLastError());
#undef STEP
  return 0;
}

extern "C" int tdz_separate(tdz_ctx* ctx, const float* mix_dev, int64_t B, int64_t T, float* out_dev, void* ws,
                            size_t ws_bytes, void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  return run_separate(ctx, mix_dev, B, T, out_dev, 2 * T, T, ws, ws_bytes, static_cast<cudaStream_t>(stream),
                      TDZ_NUM_LAYERS, 0, ST_COUNT);
}
extern "C" int tdz_separate_strided(tdz_ctx* ctx, const float* mix_dev, int64_t B, int64_t T, float* out_dev,
                                    int64_t out_chunk_stride, int64_t out_spk_stride, void* ws, size_t ws_bytes,
                                    void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (out_chunk_stride < T || out_spk_stride < T)
    return fail(ctx, "tdz_separate_strided: strides shorter than a chunk (%lld, %lld < %lld)",
                (long long)out_chunk_stride, (long long)out_spk_stride, (long long)T);
  return run_separate(ctx, mix_dev, B, T, out_dev, out_chunk_stride, out_spk_stride, ws, ws_bytes,
                      static_cast<cudaStream_t>(stream), TDZ_NUM_LAYERS, 0, ST_COUNT);
}
extern "C" int tdz_separate_debug(tdz_ctx* ctx, const float* mix_dev, int64_t B, int64_t T, float* out_dev, void* ws,
                                  size_t ws_bytes, void* stream, int num_layers, int step_lo, int step_hi) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  if (num_layers < 0 || num_layers > TDZ_NUM_LAYERS) return fail(ctx, "bad num_layers");
  std::lock_guard<std::mutex> lk(ctx->mu);
  return run_separate(ctx, mix_dev, B, T, out_dev, 2 * T, T, ws, ws_bytes, static_cast<cudaStream_t>(stream),
                      num_layers, step_lo, step_hi);
}

// ------------------------------------------------------------------------------------------------ stitching / scoring
extern "C" int tdz_gather_segments_span(tdz_ctx* ctx, const float* mix_dev, int64_t mix_origin, int64_t mix_len,
                                        int64_t L, int64_t session, int64_t hop, int64_t seg_begin, int64_t n_seg,
                                        float* seg_dev, void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  if (n_seg <= 0) return 0;
  if (session <= 0 || hop <= 0 || hop > session) return fail(ctx, "tdz_gather_segments: bad session / hop");
  // samples of [0, L) the requested segments read must be resident
  const int64_t pad = session - hop;
  const int64_t need_lo = std::max<int64_t>(seg_begin * hop - pad, 0);
  const int64_t need_hi = std::min<int64_t>((seg_begin + n_seg - 1) * hop - pad + session, L);
  if (need_hi > need_lo && (need_lo < mix_origin || need_hi > mix_origin + mix_len))
    return fail(ctx, "tdz_gather_segments: segments read samples [%lld, %lld), resident are [%lld, %lld)",
                (long long)need_lo, (long long)need_hi, (long long)mix_origin, (long long)(mix_origin + mix_len));
  const size_t total = static_cast<size_t>(n_seg) * session;
  const size_t threads = (total + 3) / 4;  // one thread per 4 outputs, the last one possibly partial
  pdl(gather_segments_kernel, static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream))(
      mix_dev - mix_origin, L, session, hop, seg_begin, n_seg, seg_dev);
  CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int tdz_gather_segments(tdz_ctx* ctx, const float* mix_dev, int64_t L, int64_t session, int64_t hop,
                                   int64_t seg_begin, int64_t n_seg, float* seg_dev, void* stream) {
  return tdz_gather_segments_span(ctx, mix_dev, 0, L, L, session, hop, seg_begin, n_seg, seg_dev, stream);
}
extern "C" int tdz_stitch_ola(tdz_ctx* ctx, const float* est_dev, int64_t session, int64_t hop, int64_t seg_begin,
                              int64_t n_seg, int64_t L, int64_t out_begin, int64_t n_out, float ratio, float* out_dev,
                              void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  if (n_out <= 0) return 0;
  const size_t total = static_cast<size_t>(n_out) * 2;
  pdl(stitch_ola_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream))(
      est_dev, session, hop, seg_begin, n_seg, L, out_begin, n_out, ratio, out_dev);
  CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int tdz_stitch_concat(tdz_ctx* ctx, const float* est_dev, int64_t len, int64_t start, int64_t L,
                                 float* out_dev, void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  if (start < 0 || start + len > L) return fail(ctx, "tdz_stitch_concat: chunk outside the output");
  CUDA_OK(cudaMemcpy2DAsync(out_dev + start, static_cast<size_t>(L) * 4, est_dev, static_cast<size_t>(len) * 4,
                            static_cast<size_t>(len) * 4, 2, cudaMemcpyDeviceToDevice,
                            static_cast<cudaStream_t>(stream)));
  return 0;
}
extern "C" int tdz_cosine_scores(tdz_ctx* ctx, const float* emb_dev, const float* target_dev, int64_t N, int64_t dim,
                                 float* scores_dev, void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  if (N <= 0) return 0;
  PdlScope pdl_scope(N <= 64);
  pdl(cosine_scores_kernel, static_cast<unsigned>((N * 32 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream))(
      emb_dev, target_dev, static_cast<int>(N), static_cast<int>(dim), scores_dev);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ loudness
extern "C" int tdz_loudness_blocks(tdz_ctx* ctx, const float* x_dev, int64_t n_streams, int64_t L, const double* coef,
                                   const int64_t* lo_dev, const int64_t* hi_dev, int64_t nblk, double inv_len,
                                   double* ysq_dev, double* z_dev, void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  if (n_streams <= 0 || L <= 0 || nblk <= 0) return fail(ctx, "tdz_loudness_blocks: empty input");
  KWeight kw;
  for (int f = 0; f < 2; ++f)
    for (int i = 0; i < 3; ++i) {
      kw.b[f][i] = coef[f * 6 + i];
      kw.a[f][i] = coef[f * 6 + 3 + i];
    }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t segs = (L + LK_SEG - 1) / LK_SEG * n_streams;
  pdl(kweight_sq_kernel, static_cast<unsigned>((segs + 127) / 128), 128, 0, st)(x_dev, L, n_streams, kw, ysq_dev);
  pdl(loudness_blocks_kernel, static_cast<unsigned>((nblk * n_streams * 32 + 255) / 256), 256, 0, st)(
      ysq_dev, L, lo_dev, hi_dev, nblk, n_streams, inv_len, z_dev);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ fbank
extern "C" int64_t tdz_fbank_frames(int64_t T) { return T < FB_WIN ? 0 : 1 + (T - FB_WIN) / FB_SHIFT; }
extern "C" int tdz_set_fbank_tables(tdz_ctx* ctx, const float* window_dev, const float* twiddle_dev,
                                    const float* mel_dev, const int32_t* mel_lo_dev, const int32_t* mel_hi_dev) {
  if (!ctx) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->fb.window = window_dev;
  ctx->fb.twiddle = reinterpret_cast<const float2*>(twiddle_dev);
  ctx->fb.mel = mel_dev;
  ctx->fb.mel_lo = mel_lo_dev;
  ctx->fb.mel_hi = mel_hi_dev;
  ctx->have_fbank = true;
  return 0;
}
extern "C" int tdz_fbank(tdz_ctx* ctx, const float* wav_dev, int64_t N, int64_t T, float* feat_dev, void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->have_fbank) return fail(ctx, "tdz_fbank: tables not set");
  const int64_t frames = tdz_fbank_frames(T);
  if (N <= 0 || frames <= 0) return fail(ctx, "tdz_fbank: input shorter than one 25 ms window");
  const int64_t total = N * frames;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PdlScope pdl_scope(total <= TDZ_PDL_MAX_SV_FRAMES);
  pdl(fbank_kernel, static_cast<unsigned>((total + 3) / 4), 128, 0, st)(wav_dev, T, frames, total, ctx->fb, feat_dev);
  pdl(fbank_meannorm_kernel, static_cast<unsigned>(N), 240, 0, st)(feat_dev, frames);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ embedder
#include "sv_api.cuh"

static void sv_free(tdz_ctx* ctx) {
  delete ctx->sv;
  ctx->sv = nullptr;
}
extern "C" int tdz_set_eres2netv2_weights(tdz_ctx* ctx, const tdz_eres2netv2_weights* w) {
  if (!ctx || !w) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->sv) ctx->sv = new SvModel();
  ctx->sv->ready = false;
  return sv_set_weights(ctx, ctx->sv, w);
}
extern "C" size_t tdz_embed_workspace_bytes(int64_t N, int64_t frames) {
  if (N <= 0 || frames < 1) return 0;
  SvLayout L;
  sv_layout(N, frames, &L);
  return L.total;
}
extern "C" int tdz_embed(tdz_ctx* ctx, const float* feat_dev, int64_t N, int64_t frames, float* emb_dev, void* ws,
                         size_t ws_bytes, void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->sv) return fail(ctx, "tdz_embed: weights not set");
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return fail(ctx, "tdz_embed: workspace must be 1024 B aligned");
  PdlScope pdl_scope(N * frames <= TDZ_PDL_MAX_SV_FRAMES);
  return sv_embed(ctx, *ctx->sv, feat_dev, N, frames, emb_dev, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
extern "C" int tdz_embed_debug(tdz_ctx* ctx, const float* feat_dev, int64_t N, int64_t frames, float* out_dev, void* ws,
                               size_t ws_bytes, void* stream, int stop_block) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->sv) return fail(ctx, "tdz_embed: weights not set");
  return sv_embed(ctx, *ctx->sv, feat_dev, N, frames, out_dev, ws, ws_bytes, static_cast<cudaStream_t>(stream),
                  stop_block);
}

// ------------------------------------------------------------------------------------------------ STFT + Apollo restorer
#include "ap_api.cuh"  // (uses gemm_b2b.cuh, included above)

static void ap_free(tdz_ctx* ctx) {
  delete ctx->ap;
  ctx->ap = nullptr;
}
extern "C" int64_t tdz_stft_frames(int64_t L, int64_t hop) { return hop > 0 && L >= 0 ? 1 + L / hop : 0; }
extern "C" int tdz_stft(tdz_ctx* ctx, const tdz_stft_plan* plan, const float* x_dev, int64_t rows, int64_t L,
                        int64_t n_keep, float* spec_dev, int64_t s_row, int64_t s_bin, int64_t s_frame, int64_t s_reim,
                        void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  FftPlan f;
  if (stft_plan_init(ctx, plan, &f)) return 1;
  return stft_launch(ctx, f, x_dev, rows, L, n_keep, spec_dev, SpecStrides{s_row, s_bin, s_frame, s_reim},
                     static_cast<cudaStream_t>(stream));
}
extern "C" int tdz_istft(tdz_ctx* ctx, const tdz_stft_plan* plan, const float* spec_dev, int64_t rows, int64_t T,
                         int64_t n_keep, int64_t s_row, int64_t s_bin, int64_t s_frame, int64_t s_reim,
                         float* frames_ws_dev, float* out_dev, int64_t out_len, void* stream) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  FftPlan f;
  if (stft_plan_init(ctx, plan, &f)) return 1;
  return istft_launch(ctx, f, spec_dev, rows, T, n_keep, SpecStrides{s_row, s_bin, s_frame, s_reim}, frames_ws_dev,
                      out_dev, out_len, static_cast<cudaStream_t>(stream));
}
extern "C" int tdz_set_apollo_weights(tdz_ctx* ctx, const tdz_apollo_weights* w) {
  if (!ctx || !w) return 1;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->ap) ctx->ap = new ApModel();
  ctx->ap->ready = false;
  return ap_set_weights(ctx, ctx->ap, w);
}
extern "C" size_t tdz_apollo_workspace_bytes(int64_t rows, int64_t nsample) {
  if (rows <= 0 || nsample <= 441) return 0;
  const int64_t T = 1 + nsample / 441;
  return ap_fixed_bytes(rows, T) + ap_token_bytes(rows, T);
}
extern "C" size_t tdz_apollo_min_workspace_bytes(int64_t rows, int64_t nsample) {
  if (rows <= 0 || nsample <= 441) return 0;
  const int64_t T = 1 + nsample / 441;
  return ap_fixed_bytes(rows, T) + ap_token_bytes(rows, std::min<int64_t>(T, 4 * AP_HALO));
}
extern "C" int tdz_apollo_debug(tdz_ctx* ctx, const float* wav_dev, int64_t rows, int64_t nsample, float* out_dev,
                                void* ws, size_t ws_bytes, void* stream, int tap) {
  if (!ctx) return 1;
  DeviceScope dev_scope(ctx->device);
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (!ctx->ap) return fail(ctx, "tdz_apollo_restore: weights not set");
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0) return fail(ctx, "tdz_apollo_restore: workspace must be 1024 B aligned");
  return ap_forward(ctx, *ctx->ap, wav_dev, rows, nsample, out_dev, ws, ws_bytes, static_cast<cudaStream_t>(stream), tap);
}
extern "C" int tdz_apollo_restore(tdz_ctx* ctx, const float* wav_dev, int64_t rows, int64_t nsample, float* out_dev,
                                  void* ws, size_t ws_bytes, void* stream) {
  return tdz_apollo_debug(ctx, wav_dev, rows, nsample, out_dev, ws, ws_bytes, stream, AP_RUN_ALL);
}
