// Instances of the tcgen05 GEMM pipeline (gemm_core.cuh) for the dense ops of the MossFormer2 separator
// and the ERes2NetV2 embedder.  Each Cfg says (a) how a work item maps to tile coordinates, (b) which TMA
// boxes make up one k-block and (c) what the epilogue fuses.  Reference op for each is cited inline
// (files under look2hear/models/ of the reference).
#pragma once
#include "gemm_core.cuh"

namespace tdz {

enum Act : int { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2, ACT_HARDTANH20 = 3, ACT_AFF = 4 };

// Compile-time switches of the generic fused epilogue (a bit mask template argument): the epilogue warps
// run one warp per scheduler, so runtime-uniform branches per element would leave them latency bound.
enum EpiFlags : unsigned {
  EF_SS_SHIFT = 1u << 0,    // ScaleNorm row scale from token-shifted half sums (mossformer_block.py:44-54,204-207)
  EF_SS_PARTS = 1u << 1,    // ScaleNorm row scale from 16 partial sums (dim 1024)
  EF_SAMP = 1u << 2,        // GroupNorm(1,C) folded behind the GEMM: v*A[b] + B[b]*colsum[col]
  EF_BIAS = 1u << 3,
  EF_RESID = 1u << 4,       // + resid[row,col] after the activation
  EF_RESID_PRE = 1u << 5,   // + resid[row,col] before the activation (ResNet block end)
  EF_MUL = 1u << 6,         // * mul[row,col]
  EF_POS = 1u << 7,         // + ScaledSinuEmbedding (mossformer_block.py:60-73)
  EF_OUT_F32 = 1u << 8,
  EF_OUT_BF16 = 1u << 9,
  EF_SS_OUT = 1u << 10,     // per-row sum of squares of the stored values, one partial per 128 columns
  EF_ZERO_PAD = 1u << 11,   // padded frames (t >= S) are written as zeros instead of skipped
  EF_ROUND_TF32 = 1u << 12, // fp32 output rounded to nearest tf32 (buffer is only a tf32 MMA operand)
  EF_OPS_BF16 = 1u << 13,   // resid / mul point to bf16 data (LinearPanel only)
  EF_RMS4 = 1u << 14,       // RMSNorm row scale rsqrt(sum of 4 partial sums * ss_dim_rsqrt [= 1 / dim] + 1e-5) (apollo.py:7-23)
};

struct EpiGeneric {
  const float* ss_in;      // EF_SS_SHIFT: [Mtot][4] (128-column partial sums); EF_SS_PARTS: [Mtot][16]
  float ss_dim_rsqrt;      // dim^-0.5 of the ScaleNorm
  const float* sampA;      // [B]
  const float* sampB;      // [B]
  const float* colsum;     // [N]
  const float* bias;       // [N]
  const float* resid;      // [Mtot][resid_ld]
  int resid_ld;
  const float* mul;        // [Mtot][mul_ld]
  int mul_ld;
  const float* pos_inv_freq;  // [N/2]   (LinearGeneric: computed in the epilogue)
  const float* pos_scale;     // [1]
  const float* pos_tab;       // [Sp][N] scale * (sin | cos)(t f_j), built once per forward (LinearPanel)
  float* out_f32;
  int out_ld;
  int out_col0;            // column offset added when storing
  __nv_bfloat16* out_bf16;
  int out_bf_ld;
  int out_bf_col0;         // column offset of the bf16 store
  float* ss_out;           // [Mtot][ss_out_ld], entry = global column / 128
  int ss_out_ld;
};

// Outputs of the GEMMs whose epilogue also runs the depthwise k=17 time convolution of the ConvModule that
// follows the Linear in every FFConvM (mossformer_block.py:89-102, conv_module.py:209-220).
struct EpiConv {
  const float* dw_t;   // per output channel 20 floats: 17 taps (tap 8 + 1 = the residual of y + conv(y)), bias / 2, 0, 0
  int ldw;             // (unused)
  // CONV_VUQK: to_hidden|to_qk  -> (v|u) bf16 and the to_qk activations fp32
  __nv_bfloat16* vu;   // [Mtot][2048]
  float* qkf;          // [Mtot][128] to_qk output (fp32; the four heads are made by qk_heads_kernel)
  // CONV_RESX: to_out -> x_out = x_in + y + dwconv(y)
  const float* x_in;   // [Mtot][512]
  float* x_out;
  // CONV_UV: to_u|to_v -> xuv fp32 [Mtot][512] and bf16 copy of x_u [Mtot][256]
  float* xuv;
  __nv_bfloat16* xubf;
};

struct LinearParams {
  CUtensorMap tmA;  // 3-D {K, Sp, B}, box {KB, 128, 1}
  CUtensorMap tmB;  // 2-D {K, N},     box {KB, BLOCK_N} (split_n: {KB, BLOCK_N/2})
  CUtensorMap tmAh; // conv GEMM, cta_group::2 form only: the activation with 128-row boxes (half an X tile per CTA)
  int B, Sp, S, N, K;
  int n_tiles;
  int shift_kblocks;  // leading k-blocks read one frame earlier (token shift, mossformer_block.py:204-207)
  int a_k0;           // element offset along K inside A
  int split_n;        // >0: tile columns [0,BN/2) come from W rows n0/2.., [BN/2,BN) from rows split_n+n0/2..
  int tps;            // LinearConv: 112-row output tiles per sample = ceil(S / 112)
  EpiGeneric e;
  EpiConv cv;
  // epilogue specific extras
  const float* alpha;  // PReLU slope (LN256 epilogue)
  const float* ln_g1;
  const float* ln_b1;
};

template <int FMT_, int BLOCK_N_, int STAGES_>
struct LinearBase {
  using Params = LinearParams;
  static constexpr int FMT = FMT_;
  static constexpr int BLOCK_N = BLOCK_N_;
  static constexpr int STAGES = STAGES_;
  static constexpr int A_MN = 0;
  static constexpr int B_MN = 0;
  static constexpr int PANEL_BYTES = 0;
  static constexpr int KB = (FMT_ == 2) ? 32 : 64;

  __device__ static void prefetch(const Params& P) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
  }
  __device__ static int num_tiles(const Params& P) { return (P.B * P.Sp / GEMM_BLOCK_M) * P.n_tiles; }
  __device__ static void tile_info(const Params& P, int tile, TileInfo& ti) {
    const int mt = tile / P.n_tiles;
    const int nt = tile - mt * P.n_tiles;
    ti.m0 = mt * GEMM_BLOCK_M;
    ti.n0 = nt * BLOCK_N;
    ti.b = ti.m0 / P.Sp;
    ti.t0 = ti.m0 - ti.b * P.Sp;
    ti.nkb = (P.K + KB - 1) / KB;  // a partial last k-block is zero-filled by TMA
    ti.aux = nt;
  }
  // The B operand is a weight matrix: its k-blocks may be requested before griddepcontrol.wait (gemm_kernel asks for
  // the first ring of them while the producer kernel is still running).  Configurations that override load() or
  // read another kernel's output through B set this to false.
  static constexpr bool PREFETCH_B = true;
  __device__ static void load_a(const Params& P, const TileInfo& ti, int kb, uint32_t sa, uint32_t bar) {
    const int trow = ti.t0 - (kb < P.shift_kblocks ? 1 : 0);
    tma_load_3d(sa, &P.tmA, bar, P.a_k0 + kb * KB, trow, ti.b);
  }
  __device__ static void load_b(const Params& P, const TileInfo& ti, int kb, uint32_t sb, uint32_t bar) {
    if (P.split_n == 0) {
      tma_load_2d(sb, &P.tmB, bar, kb * KB, ti.n0);
    } else {
      tma_load_2d(sb, &P.tmB, bar, kb * KB, ti.n0 / 2);
      tma_load_2d(sb + (BLOCK_N / 2) * 128, &P.tmB, bar, kb * KB, P.split_n + ti.n0 / 2);
    }
  }
  __device__ static void load(const Params& P, const TileInfo& ti, int kb, uint32_t sa, uint32_t sb, uint32_t bar) {
    load_a(P, ti, kb, sa, bar);
    load_b(P, ti, kb, sb, bar);
  }
};

__device__ __forceinline__ float scalenorm_rscale(float ss, float dim_rsqrt) {
  // x / clamp(||x|| * dim^-0.5, 1e-5)   (mossformer_block.py:52-54)
  return 1.f / fmaxf(sqrtf(ss) * dim_rsqrt, 1e-5f);
}

__device__ __forceinline__ float rms4_rscale(const float* ss4, float inv_dim) {
  const float4 a = *reinterpret_cast<const float4*>(ss4);
  return rsqrtf(((a.x + a.y) + (a.z + a.w)) * inv_dim + 1e-5f);
}

template <int ACT>
__device__ __forceinline__ float act_apply(float v) {
  if constexpr (ACT == ACT_SILU) return silu_f(v);
  if constexpr (ACT == ACT_RELU) return fmaxf(v, 0.f);
  if constexpr (ACT == ACT_HARDTANH20) return fminf(fmaxf(v, 0.f), 20.f);
  return v;
}

__device__ __forceinline__ void ld_f32x16(const float* p, float* v) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = __ldg(q + j);
    v[4 * j] = t.x;
    v[4 * j + 1] = t.y;
    v[4 * j + 2] = t.z;
    v[4 * j + 3] = t.w;
  }
}
__device__ __forceinline__ float4 ld_bf16x4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__bfloat162float(a.x), __bfloat162float(a.y), __bfloat162float(b.x), __bfloat162float(b.y));
}
// same for data written earlier in the same stream (plain loads, no read-only path assumption needed)
__device__ __forceinline__ void ld_f32x16_rw(const float* p, float* v, bool full) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < 2 || full) t = q[j];
    v[4 * j] = t.x;
    v[4 * j + 1] = t.y;
    v[4 * j + 2] = t.z;
    v[4 * j + 3] = t.w;
  }
}

// Generic linear layer: out = epilogue(A @ W^T).  Order of operations on an accumulator value v:
//   v *= rowscale; v = v*sampA + sampB*colsum; v += bias; [v += resid]; v = act(v); [v += resid]; v *= mul;
//   v += posenc; stores; row sum of squares.
// Work is done in chunks of 16 columns; N must be a multiple of 8 (a partial chunk has exactly 8 columns;
// per-column vectors (bias, colsum) must be readable up to the next multiple of 16).
template <int FMT_, int BLOCK_N_, int STAGES_, unsigned EF, int ACT>
struct LinearGeneric : LinearBase<FMT_, BLOCK_N_, STAGES_> {
  using Params = LinearParams;
  static constexpr int BLOCK_N = BLOCK_N_;
  static constexpr int EPI_SPLIT = (BLOCK_N_ >= 64) ? 2 : 1;
  static constexpr int COLS = BLOCK_N_ / EPI_SPLIT;  // columns handled by one epilogue warp

  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row, int half,
                                  const EpiCtx&) {
    const EpiGeneric& e = P.e;
    const int t = ti.t0 + row;
    const bool valid = t < P.S;
    const size_t grow = static_cast<size_t>(ti.m0) + row;
    float rs = 1.f;
    if constexpr ((EF & EF_SS_SHIFT) != 0) {
      const float4 cur = *reinterpret_cast<const float4*>(e.ss_in + grow * 4);
      float ss = cur.z + cur.w;  // channels 256..511 of this frame
      if (t > 0) {
        const float4 prv = *reinterpret_cast<const float4*>(e.ss_in + (grow - 1) * 4);
        ss += prv.x + prv.y;     // channels 0..255 of the previous frame
      }
      rs = scalenorm_rscale(ss, e.ss_dim_rsqrt);
    }
    if constexpr ((EF & EF_SS_PARTS) != 0) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(e.ss_in + grow * 16 + 4 * i);
        ss += (a.x + a.y) + (a.z + a.w);
      }
      rs = scalenorm_rscale(ss, e.ss_dim_rsqrt);
    }
    float sA = 1.f, sB = 0.f;
    if constexpr ((EF & EF_SAMP) != 0) {
      sA = e.sampA[ti.b];
      sB = e.sampB[ti.b];
    }
    float pscale = 0.f;
    if constexpr ((EF & EF_POS) != 0) pscale = e.pos_scale[0];
    float ssq = 0.f;
    const int cbase = half * COLS;
#pragma unroll 1
    for (int cc = 0; cc < COLS; cc += 16) {
      const int c0 = cbase + cc;
      const int col0 = ti.n0 + c0;
      if (col0 >= P.N) break;  // warp-uniform
      const bool full = col0 + 16 <= P.N;  // else exactly 8 columns exist
      float v[16];
      tmem_ld16(tacc + c0, v);
      float bias[16], cs[16], rsd[16], ml[16];
      if constexpr ((EF & EF_BIAS) != 0) ld_f32x16(e.bias + col0, bias);
      if constexpr ((EF & EF_SAMP) != 0) ld_f32x16(e.colsum + col0, cs);
      if (valid) {
        if constexpr ((EF & (EF_RESID | EF_RESID_PRE)) != 0 || ACT == ACT_AFF)
          ld_f32x16_rw(e.resid + grow * e.resid_ld + col0, rsd, full);
        if constexpr ((EF & EF_MUL) != 0 || ACT == ACT_AFF)
          ld_f32x16_rw(e.mul + grow * e.mul_ld + col0, ml, full);
      }
      tmem_ld_wait();
      if (!valid) {
        if constexpr ((EF & EF_ZERO_PAD) == 0) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float x = v[j] * rs;
          if constexpr ((EF & EF_SAMP) != 0) x = x * sA + sB * cs[j];
          if constexpr ((EF & EF_BIAS) != 0) x += bias[j];
          if constexpr (ACT == ACT_AFF) {
            // AFF gate: att = 1 + tanh(x); out = a*att + b*(2-att)   (ERes2NetV2 AFF)
            const float att = 1.f + tanhf(x);
            x = ml[j] * att + rsd[j] * (2.f - att);
          } else {
            if constexpr ((EF & EF_RESID_PRE) != 0) x += rsd[j];
            x = act_apply<ACT>(x);
            if constexpr ((EF & EF_RESID) != 0) x += rsd[j];
            if constexpr ((EF & EF_MUL) != 0) x *= ml[j];
          }
          if constexpr ((EF & EF_POS) != 0) {
            const int col = col0 + j;
            const int hlf = P.N >> 1;
            const float f = __ldg(e.pos_inv_freq + (col < hlf ? col : col - hlf));
            const float a = static_cast<float>(t) * f;
            x += pscale * (col < hlf ? sinf(a) : cosf(a));
          }
          if constexpr ((EF & EF_SS_OUT) != 0) ssq += x * x;
          if constexpr ((EF & EF_ROUND_TF32) != 0) x = round_tf32_rn(x);
          v[j] = x;
        }
      }
      if constexpr ((EF & EF_OUT_F32) != 0) {
        float4* o = reinterpret_cast<float4*>(e.out_f32 + grow * e.out_ld + e.out_col0 + col0);
        o[0] = make_float4(v[0], v[1], v[2], v[3]);
        o[1] = make_float4(v[4], v[5], v[6], v[7]);
        if (full) {
          o[2] = make_float4(v[8], v[9], v[10], v[11]);
          o[3] = make_float4(v[12], v[13], v[14], v[15]);
        }
      }
      if constexpr ((EF & EF_OUT_BF16) != 0) {
        uint4* o = reinterpret_cast<uint4*>(e.out_bf16 + grow * e.out_bf_ld + e.out_bf_col0 + col0);
        o[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        if (full)
          o[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]),
                            pack_bf16(v[14], v[15]));
      }
    }
    if constexpr ((EF & EF_SS_OUT) != 0) {
      // one partial per 128 output columns (COLS == 128 for the instances that use it)
      e.ss_out[grow * e.ss_out_ld + (ti.n0 + cbase) / 128] = valid ? ssq : 0.f;
    }
  }
};

constexpr int PANEL_LD = 132;  // floats per panel row: 128 + 4 keeps the row-wise float4 stores conflict free (LN256)
constexpr int HP_COLS = 64;    // LinearPanel: half panels of 64 columns ...
constexpr int HP_LD = 68;      // ... 64 + 4 floats per row (conflict-free row-wise float4 stores)

// LinearGeneric with a coalescing, self-overlapping epilogue for BLOCK_N = 128 / 256.  With tcgen05.ld a thread owns
// a tile ROW, so a direct epilogue reads / writes global memory in pieces that are a whole row pitch apart (32
// different lines per warp instruction).  Here the accumulator goes through shared memory in half panels of 128 rows
// x 64 columns:
//   phase 1  thread = row : tcgen05.ld -> row scale / GroupNorm fold / bias -> half panel
//   phase 2  half warp = row, lane = 4 consecutive columns: residual / gate operands are read and all outputs written
//            as 256 B (fp32) or 128 B (bf16) contiguous row segments; activation, AFF gate, positional encoding, tf32
//            rounding and the ScaleNorm partial sums (one half-warp reduction per row) happen here.
// The 16 epilogue warps form TWO groups of 8 with a half panel and a named barrier each; group g takes the half
// panels g, g + 2 of a tile.  The groups are not synchronised with each other, so one group's phase 1 (TMEM ->
// shared memory, no global traffic) runs under the other group's phase 2 (global loads / stores): with one 128-column
// panel and all 16 warps in lock step the memory system idled during every phase 1 (FSMN conv2 at 63 % of the HBM
// peak, the embedder's 1x1 convolutions at 45 %).
template <int FMT_, int BLOCK_N_, int STAGES_, unsigned EF, int ACT, int ES = 4>
struct LinearPanel : LinearBase<FMT_, BLOCK_N_, STAGES_> {
  using Params = LinearParams;
  static_assert(BLOCK_N_ == 128 || BLOCK_N_ == 256, "panel epilogue handles 128 / 256-column tiles");
  static_assert(ES == 4, "16 epilogue warps: two groups of 8");
  static constexpr int BLOCK_N = BLOCK_N_;
  static constexpr int EPI_SPLIT = ES;
  static constexpr int PANEL_BYTES = 2 * 128 * HP_LD * 4;
  // row pairs a warp has in flight in phase 2: the global operand loads of 2 RB rows are issued together.  All 16
  // rows of the warp at once where one operand is read (memory-level parallelism is what bounds these kernels: one
  // CTA per SM must keep ~35 KB of reads in flight to cover the DRAM latency at 44 GB/s per SM), 8 rows with two.
#ifndef TDZ_PANEL_RB
#define TDZ_PANEL_RB 8
#endif
  static constexpr int RB = ((EF & EF_MUL) != 0 || ACT == ACT_AFF) ? TDZ_PANEL_RB / 2 : TDZ_PANEL_RB;
  static constexpr bool HAS_R = (EF & (EF_RESID | EF_RESID_PRE)) != 0 || ACT == ACT_AFF;
  static constexpr bool HAS_M = (EF & EF_MUL) != 0 || ACT == ACT_AFF;
  static constexpr bool OPS16 = (EF & EF_OPS_BF16) != 0;
  // The global operands (residual / gate rows) of a half panel are requested ONE HALF PANEL AHEAD and travel in
  // registers across the phase-1 work, the group barrier and - for the last half panel of a tile - the tile
  // boundary: with the loads issued at the start of phase 2 every half panel paid a full DRAM latency (ncu: 51 % of
  // the stall samples on the first use of a loaded row), which bounded the kernel, not bandwidth.
  // (Measured and rejected: requesting these operands one half panel AHEAD - registers carried across phase 1, the group
  // barrier and the tile boundary - made FSMN conv2 35 % slower (0.77 -> 1.04 ms) and the embedder 5 % slower: the
  // 96-register cap of 18 warps forces spills, and the kernel is not bound by the latency of these loads.)
  static constexpr bool PIPELINED = false;
  struct EpiState {
    float4 rsd[OPS16 || !HAS_R ? 1 : RB], ml[OPS16 || !HAS_M ? 1 : RB];
    uint2 rraw[OPS16 && HAS_R ? RB : 1], mraw[OPS16 && HAS_M ? RB : 1];
  };
  struct Coord {  // warp-level constants of the thread
    int grp, sub, wg, lane, rsub, c4;
  };
  __device__ static Coord coord(int row, int half, const EpiCtx& cx) {
    Coord c;
    c.grp = half >> 1;               // epilogue group (8 warps) = half-panel buffer
    c.sub = half & 1;                // which 32 of the half panel's 64 columns this thread fills in phase 1
    c.wg = (row >> 5) + 4 * c.sub;   // warp index inside the group, 0..7
    c.lane = cx.tid & 31;
    c.rsub = c.lane >> 4;            // phase 2: a half warp per row
    c.c4 = (c.lane & 15) * 4;        //          4 consecutive columns per lane
    return c;
  }
  // operands of rows [r0, r0 + 2 RB) of half panel pn of tile ti -> st (unconditional loads, issued back to back:
  // padded rows / columns read an allocated, ignored location; bf16 data stays raw until the compute loop)
  __device__ static void load_ops(const Params& P, const TileInfo& ti, int pn, int r0, const Coord& c, EpiState& st) {
    const EpiGeneric& e = P.e;
    const int col = ti.n0 + pn * HP_COLS + c.c4;
    const int colc = col < P.N ? col : 0;
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const size_t grow = static_cast<size_t>(ti.m0) + r0 + 2 * i + c.rsub;
      if constexpr (HAS_R) {
        if constexpr (OPS16)
          st.rraw[i] = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(e.resid) +
                                                       grow * e.resid_ld + colc);
        else
          st.rsd[i] = *reinterpret_cast<const float4*>(e.resid + grow * e.resid_ld + colc);
      }
      if constexpr (HAS_M) {
        if constexpr (OPS16)
          st.mraw[i] = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(e.mul) +
                                                       grow * e.mul_ld + colc);
        else
          st.ml[i] = *reinterpret_cast<const float4*>(e.mul + grow * e.mul_ld + colc);
      }
    }
  }
  __device__ static void epi_begin(const Params& P, const TileInfo& ti, int row, int half, const EpiCtx& cx,
                                   EpiState& st) {
    const Coord c = coord(row, half, cx);
    if (ti.n0 + c.grp * HP_COLS < P.N) load_ops(P, ti, c.grp, c.wg * 16, c, st);
  }
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row, int half,
                                  const EpiCtx& cx) {
    EpiState st;
    epilogue(P, ti, ti, false, tacc, row, half, cx, st);
  }

  __device__ static void epilogue(const Params& P, const TileInfo& ti, const TileInfo& nx, bool has_next, uint32_t tacc,
                                  int row, int half, const EpiCtx& cx, EpiState& st) {
    const EpiGeneric& e = P.e;
    const size_t grow_t = static_cast<size_t>(ti.m0) + row;  // this thread's row (phase 1)
    float rs = 1.f;
    if constexpr ((EF & EF_SS_SHIFT) != 0) {
      const int t = ti.t0 + row;
      const float4 cur = *reinterpret_cast<const float4*>(e.ss_in + grow_t * 4);
      float ss = cur.z + cur.w;
      if (t > 0) {
        const float4 prv = *reinterpret_cast<const float4*>(e.ss_in + (grow_t - 1) * 4);
        ss += prv.x + prv.y;
      }
      rs = scalenorm_rscale(ss, e.ss_dim_rsqrt);
    }
    if constexpr ((EF & EF_SS_PARTS) != 0) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(e.ss_in + grow_t * 16 + 4 * i);
        ss += (a.x + a.y) + (a.z + a.w);
      }
      rs = scalenorm_rscale(ss, e.ss_dim_rsqrt);
    }
    if constexpr ((EF & EF_RMS4) != 0) rs = rms4_rscale(e.ss_in + grow_t * 4, e.ss_dim_rsqrt);
    float sA = 1.f, sB = 0.f;
    if constexpr ((EF & EF_SAMP) != 0) {
      sA = e.sampA[ti.b];
      sB = e.sampB[ti.b];
    }
    const Coord c = coord(row, half, cx);
    const int grp = c.grp, sub = c.sub, wg = c.wg, lane = c.lane, rsub = c.rsub, c4 = c.c4;
    float* panel = cx.panel + grp * (128 * HP_LD);
    float* prow = panel + row * HP_LD + sub * 32;
    auto group_sync = [&]() {
      if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync 2, 256;" ::: "memory");
    };
#pragma unroll 1
    for (int pn = grp; pn < BLOCK_N / HP_COLS; pn += 2) {
      const int pc0 = ti.n0 + pn * HP_COLS;  // first output column of the half panel
      if (pc0 >= P.N) break;                 // uniform inside the group
      // ---- phase 1
#pragma unroll
      for (int cc = 0; cc < 32; cc += 16) {
        const int c0 = pn * HP_COLS + sub * 32 + cc;
        if (ti.n0 + c0 < P.N) {
          float v[16], bias[16], cs[16];
          tmem_ld16(tacc + c0, v);
          if constexpr ((EF & EF_BIAS) != 0) ld_f32x16(e.bias + ti.n0 + c0, bias);
          if constexpr ((EF & EF_SAMP) != 0) ld_f32x16(e.colsum + ti.n0 + c0, cs);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float x = v[j] * rs;
            if constexpr ((EF & EF_SAMP) != 0) x = x * sA + sB * cs[j];
            if constexpr ((EF & EF_BIAS) != 0) x += bias[j];
            v[j] = x;
          }
          float4* dst = reinterpret_cast<float4*>(prow + cc);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      group_sync();
      // ---- phase 2: warp wg owns rows 16 wg .. 16 wg + 15, two rows per instruction
      {
        const int col = pc0 + c4;
        const bool col_ok = col < P.N;  // N is a multiple of 4
        const int colc = col_ok ? col : 0;
        // where this warp's NEXT half panel is: same tile (pn + 2) or the first one of the next tile
        const bool next_here = pn + 2 < BLOCK_N / HP_COLS && ti.n0 + (pn + 2) * HP_COLS < P.N;
        const bool next_any = next_here || (has_next && nx.n0 + grp * HP_COLS < P.N);
#pragma unroll 1
        for (int r0 = wg * 16; r0 < wg * 16 + 16; r0 += 2 * RB) {
          float4 x4[RB];
          bool valid[RB];
          if constexpr (PIPELINED) {   // batch 0 arrives in `st`; later batches (two-operand forms) load here
            if (r0 != wg * 16) load_ops(P, ti, pn, r0, c, st);
          } else if constexpr (HAS_R || HAS_M) {
            load_ops(P, ti, pn, r0, c, st);
          }
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int r = r0 + 2 * i + rsub;
            valid[i] = (ti.t0 + r) < P.S;
            x4[i] = *reinterpret_cast<const float4*>(panel + r * HP_LD + c4);
          }
          auto bf4 = [](const uint2& u) {
            const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
            const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
            return make_float4(__bfloat162float(a.x), __bfloat162float(a.y), __bfloat162float(b.x),
                               __bfloat162float(b.y));
          };
          float4 rq[RB], mq[RB];
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            rq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            mq[i] = rq[i];
            if constexpr (HAS_R) rq[i] = OPS16 ? bf4(st.rraw[OPS16 ? i : 0]) : st.rsd[OPS16 ? 0 : i];
            if constexpr (HAS_M) mq[i] = OPS16 ? bf4(st.mraw[OPS16 ? i : 0]) : st.ml[OPS16 ? 0 : i];
          }
          // the registers are free again: request batch 0 of the next half panel before the arithmetic and the stores
          if constexpr (PIPELINED) {
            if (r0 + 2 * RB >= wg * 16 + 16 && next_any) {
              if (next_here) load_ops(P, ti, pn + 2, wg * 16, c, st);
              else load_ops(P, nx, grp, wg * 16, c, st);
            }
          }
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int r = r0 + 2 * i + rsub;
            const int t = ti.t0 + r;
            const size_t grow = static_cast<size_t>(ti.m0) + r;
            float xv[4] = {x4[i].x, x4[i].y, x4[i].z, x4[i].w};
            const float rv[4] = {rq[i].x, rq[i].y, rq[i].z, rq[i].w};
            const float mv[4] = {mq[i].x, mq[i].y, mq[i].z, mq[i].w};
            [[maybe_unused]] float pv[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr ((EF & EF_POS) != 0) {  // ScaledSinuEmbedding row of frame t (table shared by all samples)
              const float4 pt = __ldg(reinterpret_cast<const float4*>(
                  e.pos_tab + static_cast<size_t>(min(t, P.Sp - 1)) * P.N + colc));
              pv[0] = pt.x; pv[1] = pt.y; pv[2] = pt.z; pv[3] = pt.w;
            }
            float ssq = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float x = xv[j];
              if constexpr (ACT == ACT_AFF) {
                const float att = 1.f + tanhf(x);
                x = mv[j] * att + rv[j] * (2.f - att);
              } else {
                if constexpr ((EF & EF_RESID_PRE) != 0) x += rv[j];
                x = act_apply<ACT>(x);
                if constexpr ((EF & EF_RESID) != 0) x += rv[j];
                if constexpr ((EF & EF_MUL) != 0) x *= mv[j];
              }
              if constexpr ((EF & EF_POS) != 0) x += pv[j];
              if (!valid[i]) x = 0.f;
              if constexpr ((EF & EF_SS_OUT) != 0) ssq += x * x;
              if constexpr ((EF & EF_ROUND_TF32) != 0) x = round_tf32_rn(x);
              xv[j] = x;
            }
            if constexpr ((EF & EF_SS_OUT) != 0) {  // one partial per 64 output columns: sum over the row's 16 lanes
#pragma unroll
              for (int o = 8; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
              if ((lane & 15) == 0) e.ss_out[grow * e.ss_out_ld + pc0 / HP_COLS] = ssq;
            }
            if ((valid[i] || (EF & EF_ZERO_PAD) != 0) && col_ok) {
              if constexpr ((EF & EF_OUT_F32) != 0)
                *reinterpret_cast<float4*>(e.out_f32 + grow * e.out_ld + e.out_col0 + col) =
                    make_float4(xv[0], xv[1], xv[2], xv[3]);
              if constexpr ((EF & EF_OUT_BF16) != 0)
                *reinterpret_cast<uint2*>(e.out_bf16 + grow * e.out_bf_ld + e.out_bf_col0 + col) =
                    make_uint2(pack_bf16(xv[0], xv[1]), pack_bf16(xv[2], xv[3]));
            }
          }
        }
      }
      group_sync();  // the half panel is free again
    }
  }
};

// Linear + SiLU + ConvModule in one kernel: see gemm_conv.cuh.  Modes of its tail:
enum ConvMode : int { CONV_VUQK = 0, CONV_RESX = 1, CONV_UV = 2 };
constexpr int CONV_ROWS = 112;  // output frames per 128-row accumulator tile (8-row halo on each side)

// FSMN block entry: Conv1d(512->256,k1)+bias -> PReLU(1) -> CLayerNorm(256) -> (to_u|to_v) LayerNorm(256)
// statistics (mossformer_block.py:405-409,419-421,301-312; layer_norm.py:9-30).  One tile holds the whole
// 256-wide row, so both LayerNorms run in the epilogue.  Outputs: c = norm1 output (fp32, later residual)
// and nhat = (c-mean(c))*rstd(c) in bf16 (the affine of the two inner LayerNorms is folded into W_u|W_v).
// 16 epilogue warps: thread = (tile row, quarter of the 256 columns); the two LayerNorm statistics are combined
// across the four threads of a row through shared memory (two-pass: mean, then centred sum of squares), then
// the values go through the 128-column panel so that c (fp32) and nhat (bf16) are written as whole row segments.
template <int FMT_, int STAGES_>
struct LinearLN256 : LinearBase<FMT_, 256, STAGES_> {
  using Params = LinearParams;
  static constexpr int EPI_SPLIT = 4;
  static constexpr int PANEL_BYTES = 128 * PANEL_LD * 4 + 128 * 8 * 4;  // panel + per-row scratch [128][8]
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row, int cq,
                                  const EpiCtx& cx) {
    const EpiGeneric& e = P.e;
    float* rowst = cx.panel + 128 * PANEL_LD;  // [128][8]: 4 partials | results
    const float alpha = P.alpha[0];
    const int w = cx.tid >> 5, lane = cx.tid & 31;
    const int cb = cq * 64;  // this thread's columns
    // activation of the accumulator: bias + PReLU
    auto act = [&](int c0, float* v) {
      float bias[16];
      tmem_ld16(tacc + c0, v);
      ld_f32x16(e.bias + c0, bias);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x = v[j] + bias[j];
        v[j] = x >= 0.f ? x : alpha * x;
      }
    };
    // ---- statistics of norm1 (two-pass over the row: mean, then centred second moment)
    float v[16];
    float s = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 16) {
      act(cb + cc, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) s += v[j];
    }
    rowst[row * 8 + cq] = s;
    epi_bar_sync<512>();
    const float mean1 = (rowst[row * 8] + rowst[row * 8 + 1] + rowst[row * 8 + 2] + rowst[row * 8 + 3]) * (1.f / 256.f);
    s = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 16) {
      act(cb + cc, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float d = v[j] - mean1;
        s = fmaf(d, d, s);
      }
    }
    rowst[row * 8 + 4 + cq] = s;
    epi_bar_sync<512>();
    const float rstd1 =
        rsqrtf((rowst[row * 8 + 4] + rowst[row * 8 + 5] + rowst[row * 8 + 6] + rowst[row * 8 + 7]) * (1.f / 256.f) + 1e-5f);
    epi_bar_sync<512>();  // everybody has read the partials before they are reused
    // ---- statistics of the inner LayerNorms on c = norm1 output
    auto norm1 = [&](int c0, float* x) {  // x: activated values in, c out
      float g1[16], b1[16];
      ld_f32x16(P.ln_g1 + c0, g1);
      ld_f32x16(P.ln_b1 + c0, b1);
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = fmaf((x[j] - mean1) * rstd1, g1[j], b1[j]);
    };
    s = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 16) {
      act(cb + cc, v);
      norm1(cb + cc, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) s += v[j];
    }
    rowst[row * 8 + cq] = s;
    epi_bar_sync<512>();
    const float mean2 = (rowst[row * 8] + rowst[row * 8 + 1] + rowst[row * 8 + 2] + rowst[row * 8 + 3]) * (1.f / 256.f);
    s = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 16) {
      act(cb + cc, v);
      norm1(cb + cc, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float d = v[j] - mean2;
        s = fmaf(d, d, s);
      }
    }
    rowst[row * 8 + 4 + cq] = s;
    epi_bar_sync<512>();
    const float rstd2 =
        rsqrtf((rowst[row * 8 + 4] + rowst[row * 8 + 5] + rowst[row * 8 + 6] + rowst[row * 8 + 7]) * (1.f / 256.f) + 1e-5f);
    epi_bar_sync<512>();
    if (cq == 0) {  // results for phase 2
      rowst[row * 8] = mean2;
      rowst[row * 8 + 1] = rstd2;
    }
    // ---- values through the panel, 128 columns at a time
#pragma unroll 1
    for (int pn = 0; pn < 2; ++pn) {
      if ((cq >> 1) == pn) {
        float* prow = cx.panel + row * PANEL_LD + (cq & 1) * 64;
#pragma unroll 1
        for (int cc = 0; cc < 64; cc += 16) {
          act(cb + cc, v);
          norm1(cb + cc, v);
          float4* dst = reinterpret_cast<float4*>(prow + cc);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      epi_bar_sync<512>();
      // phase 2: warp = 8 rows, lane = 4 columns
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = w * 8 + i;
        const bool valid = (ti.t0 + r) < P.S;
        const size_t grow = static_cast<size_t>(ti.m0) + r;
        const float4 c4 = *reinterpret_cast<const float4*>(cx.panel + r * PANEL_LD + 4 * lane);
        const float m2 = rowst[r * 8], r2 = rowst[r * 8 + 1];
        const int col = pn * 128 + 4 * lane;
        if (valid) *reinterpret_cast<float4*>(e.out_f32 + grow * 256 + col) = c4;
        const uint2 nb = valid ? make_uint2(pack_bf16((c4.x - m2) * r2, (c4.y - m2) * r2),
                                            pack_bf16((c4.z - m2) * r2, (c4.w - m2) * r2))
                               : make_uint2(0u, 0u);
        *reinterpret_cast<uint2*>(e.out_bf16 + grow * 256 + col) = nb;
      }
      epi_bar_sync<512>();
    }
  }
};

// Mask-net gated output: tanh(W_t m + b_t) * sigmoid(W_g m + b_g) (mossformer2.py:465-468,510).
// W = [W_t; W_g] stacked (1024 x 512); split_n = 512 puts the matching tanh / sigmoid columns in one tile.
template <int FMT_, int STAGES_>
struct LinearTanhSig : LinearBase<FMT_, 256, STAGES_> {
  using Params = LinearParams;
  static constexpr int EPI_SPLIT = 2;
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row, int half,
                                  const EpiCtx&) {
    const EpiGeneric& e = P.e;
    const int t = ti.t0 + row;
    const bool valid = t < P.S;  // no early return: tcgen05.ld is warp-collective
    const size_t grow = static_cast<size_t>(ti.m0) + row;
    const int oc0 = ti.n0 / 2;  // first output column of this tile
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 16) {
      const int c0 = half * 64 + cc;
      float a[16], g[16], ba[16], bg[16];
      tmem_ld16(tacc + c0, a);
      tmem_ld16(tacc + 128 + c0, g);
      ld_f32x16(e.bias + oc0 + c0, ba);
      ld_f32x16(e.bias + P.split_n + oc0 + c0, bg);
      tmem_ld_wait();
      if (!valid) continue;
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float xa = a[j] + ba[j];
        const float xg = g[j] + bg[j];
        o[j] = round_tf32_rn(tanh_approx(xa) * sigmoid_f(xg));  // MUFU forms: 2^-11 on the tanh = the tf32 rounding
      }
      float4* dst = reinterpret_cast<float4*>(e.out_f32 + grow * e.out_ld + e.out_col0 + oc0 + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
  }
};

// Gated pair of linear maps in one tile (Apollo restorer, look2hear/models/apollo.py:137-139): W = [W_gate; W_z]
// stacked (2 * split_n rows), split_n puts gate column j next to z column j in one 256-column tile; the RMSNorm in
// front of the conv is a row scale (its gain is folded into W by the packer):
//   out[row][j] = silu(silu(rs * acc_gate[j])) * silu(rs * acc_z[j])  -> bf16
// (nn.Sequential applies SiLU to all 2 * split_n channels, then F.silu is applied to the gate half again.)
template <int STAGES_>
struct LinearGLU : LinearBase<1, 256, STAGES_> {
  using Params = LinearParams;
  static constexpr int EPI_SPLIT = 2;
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row, int half,
                                  const EpiCtx&) {
    const EpiGeneric& e = P.e;
    const bool valid = ti.t0 + row < P.S;  // no early return: tcgen05.ld is warp-collective
    const size_t grow = static_cast<size_t>(ti.m0) + row;
    const float rs = rms4_rscale(e.ss_in + grow * 4, e.ss_dim_rsqrt);
    const int oc0 = ti.n0 / 2;  // first output column of this tile
#pragma unroll 1
    for (int cc = 0; cc < 64; cc += 16) {
      const int c0 = half * 64 + cc;
      float a[16], z[16];
      tmem_ld16(tacc + c0, a);
      tmem_ld16(tacc + 128 + c0, z);
      tmem_ld_wait();
      if (!valid) continue;
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = silu_f(silu_f(a[j] * rs)) * silu_f(z[j] * rs);
      uint4* dst = reinterpret_cast<uint4*>(e.out_bf16 + grow * e.out_bf_ld + oc0 + c0);
      dst[0] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
      dst[1] = make_uint4(pack_bf16(o[8], o[9]), pack_bf16(o[10], o[11]), pack_bf16(o[12], o[13]),
                          pack_bf16(o[14], o[15]));
    }
  }
};

// Split-K form of a skinny linear layer (the embedder's Linear(40960 -> 192) on at most a few hundred rows: as one
// tile per 128 rows it was a single CTA walking 640 k-blocks, 277 us whatever the batch).  Work item = (row tile,
// K slice of P.tps k-blocks); the fp32 partial tile goes to out_f32[(slice * B * Sp + row) * 256 + col] and
// splitk_reduce_kernel adds the slices in ascending order (+ bias): the summation order is fixed by K alone.
// P.n_tiles = number of slices; N <= 256.
template <int STAGES_>
struct LinearSplitK : LinearBase<1, 256, STAGES_> {
  using Params = LinearParams;
  using Base = LinearBase<1, 256, STAGES_>;
  static constexpr int EPI_SPLIT = 2;
  static constexpr bool PREFETCH_B = false;  // (own load(): K-sliced coordinates)
  __device__ static void tile_info(const Params& P, int tile, TileInfo& ti) {
    const int mt = tile / P.n_tiles;
    ti.aux = tile - mt * P.n_tiles;  // K slice
    ti.m0 = mt * GEMM_BLOCK_M;
    ti.n0 = 0;
    ti.b = ti.m0 / P.Sp;
    ti.t0 = ti.m0 - ti.b * P.Sp;
    const int total = (P.K + Base::KB - 1) / Base::KB;
    const int left = total - ti.aux * P.tps;
    ti.nkb = left < P.tps ? left : P.tps;
  }
  __device__ static void load(const Params& P, const TileInfo& ti, int kb, uint32_t sa, uint32_t sb, uint32_t bar) {
    const int k = (ti.aux * P.tps + kb) * Base::KB;
    tma_load_3d(sa, &P.tmA, bar, k, ti.t0, ti.b);
    tma_load_2d(sb, &P.tmB, bar, k, 0);
  }
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row, int half,
                                  const EpiCtx&) {
    const size_t mtot = static_cast<size_t>(P.B) * P.Sp;
    float* dst = P.e.out_f32 + (static_cast<size_t>(ti.aux) * mtot + ti.m0 + row) * 256;
#pragma unroll 1
    for (int cc = 0; cc < 128; cc += 32) {
      const int c0 = half * 128 + cc;
      float v[32];
      tmem_ld16(tacc + c0, v);
      tmem_ld16(tacc + c0 + 16, v + 16);
      tmem_ld_wait();
      float4* o = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
};
// out[row][col] = bias[col] + sum over slices (ascending) of part[slice][row][col]; rows < rows_valid, col < N
__global__ void splitk_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, int nslice,
                                     int64_t mtot, int rows_valid, int N, float* __restrict__ out, int out_ld) {
  pdl_enter();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(rows_valid) * N) return;
  const int row = static_cast<int>(i / N), col = static_cast<int>(i - static_cast<int64_t>(row) * N);
  float acc = 0.f;
  for (int s = 0; s < nslice; ++s) acc += part[(static_cast<size_t>(s) * mtot + row) * 256 + col];
  out[static_cast<size_t>(row) * out_ld + col] = acc + bias[col];
}

// Decoder: ConvTranspose1d(512 -> 1, k 16, stride 8) + overlap-by-2 add + pad / trim to T (mossformer2.py:213-257,
// 579-589) as a skinny tf32 GEMM: D[t][j] = sum_c sep[t][c] w[c][j] (N = 16), out[8 t + j] = D[t][j] + D[t-1][8 + j].
// A tile holds 128 frames of which the first is the halo row t0 - 1 (tiles advance by 127 frames; TMA zero-fills rows
// outside the sample, the producer zero-fills frames >= S), so the overlap add needs no second pass: a thread owns a
// frame, takes the upper half of the previous frame's 16 taps from the neighbouring lane (warp shuffle; across warps
// through 8 floats of shared memory) and writes 8 consecutive samples - a warp writes 1 KB contiguous.
// "Batch" dimension of A = (speaker, sample): sep is [2][B][Sp][512].
struct DecParams {
  CUtensorMap tmA;  // 3-D {512, Sp, 2 B} fp32, box {32, 128, 1}
  CUtensorMap tmB;  // 2-D {512, 16} fp32 (w transposed: [16][512]), box {32, 16}
  float* out;       // stream spk of sample b at out + b * out_cs + spk * out_ss
  int64_t out_cs, out_ss;
  int B, Sp, S, T, tps;  // tps = tiles per (speaker, sample) = ceil((S + 1) / 127)
};
constexpr int DEC_TILE_FRAMES = 127;
struct DecoderGemm {
  using Params = DecParams;
  static constexpr int FMT = 2, BLOCK_N = 16, STAGES = 8, A_MN = 0, B_MN = 0, EPI_SPLIT = 1, PANEL_BYTES = 0;
  static constexpr bool PREFETCH_B = false;
  __device__ static void prefetch(const Params& P) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
  }
  __device__ static int num_tiles(const Params& P) { return 2 * P.B * P.tps; }
  __device__ static void tile_info(const Params& P, int tile, TileInfo& ti) {
    ti.b = tile / P.tps;                       // speaker * B + sample
    ti.aux = tile - ti.b * P.tps;
    ti.t0 = ti.aux * DEC_TILE_FRAMES - 1;      // frame of tile row 0 (the halo row)
    ti.m0 = 0;
    ti.n0 = 0;
    ti.nkb = 16;
  }
  __device__ static void load(const Params& P, const TileInfo& ti, int kb, uint32_t sa, uint32_t sb, uint32_t bar) {
    tma_load_3d(sa, &P.tmA, bar, kb * 32, ti.t0, ti.b);
    tma_load_2d(sb, &P.tmB, bar, kb * 32, 0);
  }
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row, int, const EpiCtx& cx) {
    __shared__ float exch[4][8];
    float v[16];
    tmem_ld16(tacc, v);
    tmem_ld_wait();
    const int w = row >> 5, lane = row & 31;  // TMEM lane quarter = 32 consecutive frames (not the warp's launch order)
    (void)cx;
    if (lane == 31) {
#pragma unroll
      for (int j = 0; j < 8; ++j) exch[w][j] = v[8 + j];
    }
    float prev[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) prev[j] = __shfl_up_sync(0xffffffffu, v[8 + j], 1);
    epi_bar_sync<128>();
    if (lane == 0 && w > 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) prev[j] = exch[w - 1][j];
    }
    const int t = ti.t0 + row;  // output frame: samples [8 t, 8 t + 8)
    const int spk = ti.b / P.B, b = ti.b - spk * P.B;
    if (row > 0 && t <= P.S) {
      float* dst = P.out + static_cast<int64_t>(b) * P.out_cs + static_cast<int64_t>(spk) * P.out_ss;
      const int64_t n0 = static_cast<int64_t>(t) * 8;
      if (n0 + 8 <= P.T && ((reinterpret_cast<uintptr_t>(dst + n0) & 15) == 0)) {
        float4* o = reinterpret_cast<float4*>(dst + n0);
        o[0] = make_float4(v[0] + prev[0], v[1] + prev[1], v[2] + prev[2], v[3] + prev[3]);
        o[1] = make_float4(v[4] + prev[4], v[5] + prev[5], v[6] + prev[6], v[7] + prev[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (n0 + j < P.T) dst[n0 + j] = v[j] + prev[j];
      }
    }
    epi_bar_sync<128>();  // exch is reused by the next tile
  }
};

// ------------------------------------------------------------------------------------------------
// FLASH attention pieces (mossformer_block.py:222-294).  qk4 = [Mtot][512] bf16 holds the four rotated
// OffsetScale heads (quad_q | lin_q | quad_k | lin_k); vu = [Mtot][2048] bf16 holds (v | u), zero in
// padded frames.
struct AttnParams {
  CUtensorMap tmQK;    // 3-D {512, Sp, B}  box {64, 128, 1}  K-major rows (A of sim, lin_q)
  CUtensorMap tmQKb;   // 3-D {512, Sp, B}  box {64, 256, 1}  K-major rows (B of sim: 256 keys of the group)
  CUtensorMap tmQKmn;  // 3-D {512, Sp, B}  box {64, 64, 1}   MN-major atoms (A of kv: lin_k^T)
  CUtensorMap tmVUmn;  // 3-D {2048, Sp, B} box {64, 64, 1}   MN-major atoms (B of quad*VU and of kv)
  CUtensorMap tmP;     // 3-D {256, Sp, B}  box {64, 128, 1}  relu^2 attention weights
  CUtensorMap tmKVmn;  // 3-D {2048, 128, B} box {64, 64, 1}  MN-major atoms of lin_kv | lin_ku (FP16, like lin_q)
  int B, Sp, S;
  int nsplit;          // kv: splits of the frame axis
  int kb_per_split;
  __nv_bfloat16* P;        // [Mtot][256]
  float* kv_part;          // [B][nsplit][128][2048]
  const __nv_bfloat16* vu; // [Mtot][2048]
  __nv_bfloat16* o;        // [Mtot][1024] gated output
  float* o_ss;             // [32][Mtot]  partial sums of squares of o (ScaleNorm(1024) of to_out)
};

// sim = quad_q quad_k^T / 256 ; attn = relu(sim)^2    (mossformer_block.py:256-258)
struct AttnSim {
  using Params = AttnParams;
  static constexpr int PANEL_BYTES = 0;
  static constexpr int FMT = 1, BLOCK_N = 256, STAGES = 4, A_MN = 0, B_MN = 0, EPI_SPLIT = 2;
  static constexpr bool PREFETCH_B = false;  // both operands are activations
  __device__ static void prefetch(const Params& P) {
    tma_prefetch_desc(&P.tmQK);
    tma_prefetch_desc(&P.tmQKb);
  }
  __device__ static int num_tiles(const Params& P) { return P.B * P.Sp / GEMM_BLOCK_M; }
  __device__ static void tile_info(const Params& P, int tile, TileInfo& ti) {
    ti.m0 = tile * GEMM_BLOCK_M;
    ti.n0 = 0;
    ti.b = ti.m0 / P.Sp;
    ti.t0 = ti.m0 - ti.b * P.Sp;
    ti.nkb = 2;
    ti.aux = (ti.t0 / 256) * 256;  // first frame of the group
  }
  __device__ static void load(const Params& P, const TileInfo& ti, int kb, uint32_t sa, uint32_t sb, uint32_t bar) {
    tma_load_3d(sa, &P.tmQK, bar, 0 + kb * 64, ti.t0, ti.b);      // quad_q
    tma_load_3d(sb, &P.tmQKb, bar, 256 + kb * 64, ti.aux, ti.b);  // quad_k of the whole group
  }
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row, int half,
                                  const EpiCtx&) {
    const size_t grow = static_cast<size_t>(ti.m0) + row;
#pragma unroll 1
    for (int cc = 0; cc < 128; cc += 32) {
      const int c0 = half * 128 + cc;
      float v[32];
      tmem_ld16(tacc + c0, v);
      tmem_ld16(tacc + c0 + 16, v + 16);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = fmaxf(v[j] * (1.f / 256.f), 0.f);
        v[j] = s * s;
      }
      uint4* o = reinterpret_cast<uint4*>(P.P + grow * 256 + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                          pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    }
  }
};

// lin_kv | lin_ku partial sums: kv_part[b][s][d][e] = sum_{t in split s} lin_k[t,d] * vu[t,e]
// (mossformer_block.py:286,289; the 1/n and the sum over splits happen in kv_reduce_kernel).
// Both operands are read MN-major straight from the token-major buffers.
struct AttnKV {
  using Params = AttnParams;
  static constexpr int PANEL_BYTES = 0;
  static constexpr int FMT = 1, BLOCK_N = 256, STAGES = 4, A_MN = 1, B_MN = 1, EPI_SPLIT = 2;
  static constexpr bool PREFETCH_B = false;
  __device__ static void prefetch(const Params& P) {
    tma_prefetch_desc(&P.tmQKmn);
    tma_prefetch_desc(&P.tmVUmn);
  }
  __device__ static int num_tiles(const Params& P) { return P.B * P.nsplit * 8; }
  __device__ static void tile_info(const Params& P, int tile, TileInfo& ti) {
    const int nt = tile & 7;
    const int bs = tile >> 3;
    ti.b = bs / P.nsplit;
    ti.aux = bs - ti.b * P.nsplit;  // split index
    ti.n0 = nt * 256;
    ti.m0 = 0;
    ti.t0 = ti.aux * P.kb_per_split * 64;
    const int total_kb = P.Sp / 64;
    int nkb = total_kb - ti.aux * P.kb_per_split;
    ti.nkb = nkb < P.kb_per_split ? nkb : P.kb_per_split;
  }
  __device__ static void load(const Params& P, const TileInfo& ti, int kb, uint32_t sa, uint32_t sb, uint32_t bar) {
    const int trow = ti.t0 + kb * 64;
#pragma unroll
    for (int j = 0; j < 2; ++j) tma_load_3d(sa + j * 8192, &P.tmQKmn, bar, 384 + j * 64, trow, ti.b);  // lin_k
#pragma unroll
    for (int j = 0; j < 4; ++j) tma_load_3d(sb + j * 8192, &P.tmVUmn, bar, ti.n0 + j * 64, trow, ti.b);
  }
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row, int half,
                                  const EpiCtx&) {
    float* dst = P.kv_part + ((static_cast<size_t>(ti.b) * P.nsplit + ti.aux) * 128 + row) * 2048 + ti.n0;
#pragma unroll 1
    for (int cc = 0; cc < 128; cc += 32) {
      const int c0 = half * 128 + cc;
      float v[32];
      tmem_ld16(tacc + c0, v);
      tmem_ld16(tacc + c0 + 16, v + 16);
      tmem_ld_wait();
      float4* o = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
};

// att = attn @ [v|u] (per 256-frame group) + lin_q @ [lin_kv|lin_ku]; out = (att_u*v)*sigmoid(att_v*u)
// (mossformer_block.py:269-270,287-294,217).  One accumulator tile = 128 v-columns next to the matching
// 128 u-columns, so the gate runs in the epilogue and the [.,2048] attention output never reaches HBM.
// Runs as gemm_cg2_kernel: the CTA pair that owns the two query tiles of a 256-frame group executes ONE
// tcgen05.mma.cta_group::2 (M = 256) per k-step; rank 0 stages the 128 v columns of B, rank 1 the 128 u columns, each
// its own A rows.  k-blocks 0..3: quadratic part (bf16, K = 256 keys); k-blocks 4..5: linear part lin_q @ lin_kv with
// FP16 operands (K = 128, see qk_heads_kernel) - two instruction descriptors accumulate into one TMEM tile.
struct AttnOut {
  using Params = AttnParams;
  static constexpr int PANEL_BYTES = 0;
  static constexpr int FMT = 1, BLOCK_N = 256, STAGES = 4, A_MN = 0, B_MN = 1, EPI_SPLIT = 2;
  static constexpr bool PREFETCH_B = false;
  static constexpr int F16_FROM_KB = 4;  // k-blocks from this one on use fp16 operands (instruction descriptor fmt 0)
  __device__ static void prefetch(const Params& P) {
    tma_prefetch_desc(&P.tmP);
    tma_prefetch_desc(&P.tmVUmn);
    tma_prefetch_desc(&P.tmQK);
    tma_prefetch_desc(&P.tmKVmn);
  }
  __device__ static void tile_info(const Params& P, int tile, TileInfo& ti) {
    const int mt = tile >> 3;
    const int nt = tile & 7;
    ti.m0 = mt * GEMM_BLOCK_M;
    ti.n0 = nt * 128;  // first v channel; the u channel is 1024 + n0
    ti.b = ti.m0 / P.Sp;
    ti.t0 = ti.m0 - ti.b * P.Sp;
    ti.nkb = 6;  // 4 quadratic (256 keys) + 2 linear (128 features)
    ti.aux = nt;
  }
  __device__ static int num_pair_tiles(const Params& P) { return (P.B * P.Sp / 256) * 8; }
  __device__ static void pair_tile_info(const Params& P, int w, uint32_t rank, TileInfo& ti) {
    tile_info(P, ((w >> 3) * 2 + static_cast<int>(rank)) * 8 + (w & 7), ti);
  }
  static constexpr int CG2_STAGES = 6;
  __device__ static void load_cg2(const Params& P, const TileInfo& ti, int kb, uint32_t sa, uint32_t sb, uint32_t bar,
                                  uint32_t rank) {
    const int ch0 = (rank == 0 ? 0 : 1024) + ti.n0;
    if (kb < 4) {
      const int g0 = (ti.t0 / 256) * 256;
      tma_load_3d_cg2(sa, &P.tmP, bar, kb * 64, ti.t0, ti.b);
      tma_load_3d_cg2(sb, &P.tmVUmn, bar, ch0, g0 + kb * 64, ti.b);
      tma_load_3d_cg2(sb + 8192, &P.tmVUmn, bar, ch0 + 64, g0 + kb * 64, ti.b);
    } else {
      const int k0 = (kb - 4) * 64;
      tma_load_3d_cg2(sa, &P.tmQK, bar, 128 + k0, ti.t0, ti.b);  // lin_q (fp16)
      tma_load_3d_cg2(sb, &P.tmKVmn, bar, ch0, k0, ti.b);
      tma_load_3d_cg2(sb + 8192, &P.tmKVmn, bar, ch0 + 64, k0, ti.b);
    }
  }
  // ---- epilogue: 16 warps, thread = one row x 32 v/u columns; v / u are fetched one tile ahead with 256-bit loads
  //      and the gated output leaves with 256-bit stores (the 128-bit row-per-thread form kept the LSU busy for
  //      longer than the tile's MMAs take: ncu, stall_lg / long scoreboard on LDTM)
  static constexpr int CG2_EPI_SPLIT = 4;
  struct EpiPre4 {
    U8 v[2], u[2];
  };
  __device__ static void epi_prefetch4(const Params& P, const TileInfo& ti, int row, int part, int cc, EpiPre4& pre) {
    if (ti.t0 + row < P.S) {
      const __nv_bfloat16* vp = P.vu + (static_cast<size_t>(ti.m0) + row) * 2048 + ti.n0 + part * 32 + cc * 16;
      pre.v[cc] = ld_global_256(vp);
      pre.u[cc] = ld_global_256(vp + 1024);
    }
  }
  // `pre` holds v / u of this tile on entry and of tile `nx` (if has_next) on exit: each half is re-requested as
  // soon as it has been consumed, so the loads of the next tile fly during the rest of this tile's epilogue.
  // o_ss holds 32 partial sums per row, part-major ([32][Mtot]: a warp stores 128 contiguous bytes): index =
  // 4 * n_tile + column quarter (the consumer adds them in index order)
  __device__ static void epilogue4(const Params& P, const TileInfo& ti, const TileInfo& nx, bool has_next,
                                   uint32_t tacc, int row, int part, EpiPre4& pre) {
    const bool valid = ti.t0 + row < P.S;
    const size_t grow = static_cast<size_t>(ti.m0) + row;
    float ssq = 0.f;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int c0 = part * 32 + cc * 16;
      float av[16], au[16];
      tmem_ld16(tacc + c0, av);
      tmem_ld16(tacc + 128 + c0, au);
      tmem_ld_wait();
      U8 o;
      if (valid) {
        const __nv_bfloat16* vb = reinterpret_cast<const __nv_bfloat16*>(&pre.v[cc]);
        const __nv_bfloat16* ub = reinterpret_cast<const __nv_bfloat16*>(&pre.u[cc]);
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float x0 = (au[j] * __bfloat162float(vb[j])) * sigmoid_f(av[j] * __bfloat162float(ub[j]));
          const float x1 =
              (au[j + 1] * __bfloat162float(vb[j + 1])) * sigmoid_f(av[j + 1] * __bfloat162float(ub[j + 1]));
          ssq = fmaf(x0, x0, ssq);
          ssq = fmaf(x1, x1, ssq);
          o.r[j / 2] = pack_bf16(x0, x1);
        }
      }
      if (has_next) epi_prefetch4(P, nx, row, part, cc, pre);
      if (valid) st_global_256(P.o + grow * 1024 + ti.n0 + c0, o);
    }
    P.o_ss[(ti.aux * 4 + part) * (static_cast<size_t>(P.B) * P.Sp) + grow] = valid ? ssq : 0.f;
  }
};

}  // namespace tdz
