// Instances of the tcgen05 GEMM pipeline (gemm_core.cuh) for the dense ops of the MossFormer2 separator
// and the ERes2NetV2 embedder.  Each Cfg says (a) how a work item maps to tile coordinates, (b) which TMA
// boxes make up one k-block and (c) what the epilogue fuses.  Reference op for each is cited inline
// (files under look2hear/models/ of the reference).
#pragma once
#include "gemm_core.cuh"

namespace tdz {

enum Act : int { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2, ACT_PRELU = 3, ACT_HARDTANH20 = 4 };

// ------------------------------------------------------------------------------------------------
// Generic fused elementwise epilogue.  Order of operations on an accumulator value v at (row, col):
//   v *= rowscale(row)                   ScaleNorm folded behind the GEMM (mossformer_block.py:44-54)
//   v  = v*sampA[b] + sampB[b]*colsum[col]   GroupNorm(1,C) folded behind the GEMM (mossformer2.py:487-490)
//   v += bias[col]; v = act(v); v += resid[row,col]; v *= mul[row,col]; v += posenc(t,col)
// then optional fp32 / bf16 stores and a per-row sum of squares over the tile's columns.
struct EpiGeneric {
  const float* ss_in;      // ss_mode 1: [Mtot][2] (lo,hi halves); ss_mode 2: [Mtot][ss_parts] partial sums
  int ss_mode;             // 0 none, 1 token-shifted halves (dim 512), 2 partial sums
  int ss_parts;
  float ss_dim_rsqrt;      // dim^-0.5 of the ScaleNorm
  const float* sampA;      // [B]
  const float* sampB;      // [B]
  const float* colsum;     // [N]
  const float* bias;       // [N]
  int act;
  const float* alpha;      // PReLU slope (1 value)
  const float* resid;      // [Mtot][resid_ld]
  int resid_ld;
  const float* mul;        // [Mtot][mul_ld]
  int mul_ld;
  const float* pos_inv_freq;  // [N/2]  ScaledSinuEmbedding (mossformer_block.py:60-73)
  const float* pos_scale;     // [1]
  float* out_f32;
  int out_ld;
  int out_col0;            // column offset added when storing
  __nv_bfloat16* out_bf16;
  int out_bf_ld;
  float* ss_out;           // [Mtot][ss_out_ld], entry n_tile
  int ss_out_ld;
  int zero_pad_rows;       // rows with t >= S are written as zeros instead of skipped
};

struct LinearParams {
  CUtensorMap tmA;  // 3-D {K, Sp, B}, box {KB, 128, 1}
  CUtensorMap tmB;  // 2-D {K, N},     box {KB, BLOCK_N} (split_n: {KB, BLOCK_N/2})
  int B, Sp, S, N, K;
  int n_tiles;
  int shift_kblocks;  // leading k-blocks read one frame earlier (token shift, mossformer_block.py:204-207)
  int a_k0;           // element offset along K inside A
  int split_n;        // >0: tile columns [0,BN/2) come from W rows n0/2.., [BN/2,BN) from rows split_n+n0/2..
  EpiGeneric e;
  // epilogue specific extras
  const float* ln_g1;  // LN256 epilogue
  const float* ln_b1;
  float* out2_f32;
};

template <int FMT_, int BLOCK_N_, int STAGES_>
struct LinearBase {
  using Params = LinearParams;
  static constexpr int FMT = FMT_;
  static constexpr int BLOCK_N = BLOCK_N_;
  static constexpr int STAGES = STAGES_;
  static constexpr int A_MN = 0;
  static constexpr int B_MN = 0;
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
    ti.nkb = P.K / KB;
    ti.aux = nt;
  }
  __device__ static void load(const Params& P, const TileInfo& ti, int kb, uint32_t sa, uint32_t sb, uint32_t bar) {
    const int trow = ti.t0 - (kb < P.shift_kblocks ? 1 : 0);
    tma_load_3d(sa, &P.tmA, bar, P.a_k0 + kb * KB, trow, ti.b);
    if (P.split_n == 0) {
      tma_load_2d(sb, &P.tmB, bar, kb * KB, ti.n0);
    } else {
      tma_load_2d(sb, &P.tmB, bar, kb * KB, ti.n0 / 2);
      tma_load_2d(sb + (BLOCK_N / 2) * 128, &P.tmB, bar, kb * KB, P.split_n + ti.n0 / 2);
    }
  }
};

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  switch (act) {
    case ACT_SILU: return silu_f(v);
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_PRELU: return v >= 0.f ? v : alpha * v;
    case ACT_HARDTANH20: return fminf(fmaxf(v, 0.f), 20.f);
    default: return v;
  }
}

__device__ __forceinline__ float scalenorm_rscale(float ss, float dim_rsqrt) {
  // x / clamp(||x|| * dim^-0.5, 1e-5)   (mossformer_block.py:52-54)
  return 1.f / fmaxf(sqrtf(ss) * dim_rsqrt, 1e-5f);
}

template <int FMT_, int BLOCK_N_, int STAGES_>
struct LinearGeneric : LinearBase<FMT_, BLOCK_N_, STAGES_> {
  using Params = LinearParams;
  static constexpr int BLOCK_N = BLOCK_N_;
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row) {
    const EpiGeneric& e = P.e;
    const int t = ti.t0 + row;
    const bool valid = t < P.S;
    const size_t grow = static_cast<size_t>(ti.m0) + row;
    float rs = 1.f;
    if (e.ss_mode == 1) {
      float ss = e.ss_in[grow * 2 + 1];
      if (t > 0) ss += e.ss_in[(grow - 1) * 2];
      rs = scalenorm_rscale(ss, e.ss_dim_rsqrt);
    } else if (e.ss_mode == 2) {
      float ss = 0.f;
      for (int i = 0; i < e.ss_parts; ++i) ss += e.ss_in[grow * e.ss_parts + i];
      rs = scalenorm_rscale(ss, e.ss_dim_rsqrt);
    }
    float sA = 1.f, sB = 0.f;
    if (e.sampA) {
      sA = e.sampA[ti.b];
      sB = e.sampB[ti.b];
    }
    const float alpha = e.alpha ? e.alpha[0] : 0.f;
    const float pscale = e.pos_scale ? e.pos_scale[0] : 0.f;
    float ssq = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      const int col0 = ti.n0 + c0;
      if (col0 >= P.N) break;  // warp-uniform: N is a multiple of 32
      float v[32];
      tmem_ld16(tacc + c0, v);
      tmem_ld16(tacc + c0 + 16, v + 16);
      tmem_ld_wait();
      if (!valid) {
        if (!e.zero_pad_rows) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = col0 + j;
          float x = v[j] * rs;
          if (e.sampA) x = x * sA + sB * __ldg(e.colsum + col);
          if (e.bias) x += __ldg(e.bias + col);
          x = apply_act(x, e.act, alpha);
          if (e.resid) x += e.resid[grow * e.resid_ld + col];
          if (e.mul) x *= e.mul[grow * e.mul_ld + col];
          if (e.pos_inv_freq) {
            const int half = P.N >> 1;
            const float f = __ldg(e.pos_inv_freq + (col < half ? col : col - half));
            const float a = static_cast<float>(t) * f;
            x += pscale * (col < half ? sinf(a) : cosf(a));
          }
          v[j] = x;
          ssq += x * x;
        }
      }
      if (e.out_f32) {
        float4* o = reinterpret_cast<float4*>(e.out_f32 + grow * e.out_ld + e.out_col0 + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (e.out_bf16) {
        uint4* o = reinterpret_cast<uint4*>(e.out_bf16 + grow * e.out_bf_ld + e.out_col0 + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          o[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                            pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
      }
    }
    if (e.ss_out) e.ss_out[grow * e.ss_out_ld + ti.aux] = valid ? ssq : 0.f;
  }
};

// FSMN block entry: Conv1d(512->256,k1)+bias -> PReLU(1) -> CLayerNorm(256) -> (to_u|to_v) LayerNorm(256)
// statistics (mossformer_block.py:405-409,419-421,301-312; layer_norm.py:9-30).  One tile holds the whole
// 256-wide row, so both LayerNorms run in the epilogue.  Outputs: c = norm1 output (fp32, later residual)
// and nhat = (c-mean(c))*rstd(c) in bf16 (the affine of the two inner LayerNorms is folded into W_u|W_v).
template <int FMT_, int STAGES_>
struct LinearLN256 : LinearBase<FMT_, 256, STAGES_> {
  using Params = LinearParams;
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row) {
    const EpiGeneric& e = P.e;
    const int t = ti.t0 + row;
    const bool valid = t < P.S;
    const size_t grow = static_cast<size_t>(ti.m0) + row;
    const float alpha = e.alpha[0];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      float v[32];
      tmem_ld16(tacc + c0, v);
      tmem_ld16(tacc + c0 + 16, v + 16);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = v[j] + __ldg(e.bias + c0 + j);
        x = x >= 0.f ? x : alpha * x;
        s1 += x;
        s2 += x * x;
      }
    }
    const float mean1 = s1 * (1.f / 256.f);
    const float rstd1 = rsqrtf(fmaxf(s2 * (1.f / 256.f) - mean1 * mean1, 0.f) + 1e-5f);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      float v[32];
      tmem_ld16(tacc + c0, v);
      tmem_ld16(tacc + c0 + 16, v + 16);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = v[j] + __ldg(e.bias + c0 + j);
        x = x >= 0.f ? x : alpha * x;
        x = (x - mean1) * rstd1 * __ldg(P.ln_g1 + c0 + j) + __ldg(P.ln_b1 + c0 + j);
        v[j] = x;
        c1 += x;
        c2 += x * x;
      }
      if (valid) {
        float4* o = reinterpret_cast<float4*>(e.out_f32 + grow * 256 + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
    const float mean2 = c1 * (1.f / 256.f);
    const float rstd2 = rsqrtf(fmaxf(c2 * (1.f / 256.f) - mean2 * mean2, 0.f) + 1e-5f);
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      float v[32];
      tmem_ld16(tacc + c0, v);
      tmem_ld16(tacc + c0 + 16, v + 16);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = v[j] + __ldg(e.bias + c0 + j);
        x = x >= 0.f ? x : alpha * x;
        x = (x - mean1) * rstd1 * __ldg(P.ln_g1 + c0 + j) + __ldg(P.ln_b1 + c0 + j);
        v[j] = valid ? (x - mean2) * rstd2 : 0.f;
      }
      uint4* o = reinterpret_cast<uint4*>(e.out_bf16 + grow * 256 + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o[j] = make_uint4(pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                          pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    }
  }
};

// Mask-net gated output: tanh(W_t m + b_t) * sigmoid(W_g m + b_g) (mossformer2.py:465-468,510).
// W = [W_t; W_g] stacked (1024 x 512); split_n = 512 puts the matching tanh / sigmoid columns in one tile.
template <int FMT_, int STAGES_>
struct LinearTanhSig : LinearBase<FMT_, 256, STAGES_> {
  using Params = LinearParams;
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row) {
    const EpiGeneric& e = P.e;
    const int t = ti.t0 + row;
    const bool valid = t < P.S;  // no early return: tcgen05.ld is warp-collective
    const size_t grow = static_cast<size_t>(ti.m0) + row;
    const int oc0 = ti.n0 / 2;  // first output column of this tile
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 16) {
      float a[16], g[16];
      tmem_ld16(tacc + c0, a);
      tmem_ld16(tacc + 128 + c0, g);
      tmem_ld_wait();
      if (!valid) continue;
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float xa = a[j] + __ldg(e.bias + oc0 + c0 + j);
        const float xg = g[j] + __ldg(e.bias + P.split_n + oc0 + c0 + j);
        o[j] = tanhf(xa) * (1.f / (1.f + expf(-xg)));
      }
      float4* dst = reinterpret_cast<float4*>(e.out_f32 + grow * e.out_ld + e.out_col0 + oc0 + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    }
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
  CUtensorMap tmKVmn;  // 3-D {2048, 128, B} box {64, 64, 1}  MN-major atoms of lin_kv | lin_ku
  int B, Sp, S;
  int nsplit;          // kv: splits of the frame axis
  int kb_per_split;
  __nv_bfloat16* P;        // [Mtot][256]
  float* kv_part;          // [B][nsplit][128][2048]
  const __nv_bfloat16* vu; // [Mtot][2048]
  __nv_bfloat16* o;        // [Mtot][1024] gated output
  float* o_ss;             // [Mtot][8]   partial sums of squares of o (ScaleNorm(1024) of to_out)
};

// sim = quad_q quad_k^T / 256 ; attn = relu(sim)^2    (mossformer_block.py:256-258)
struct AttnSim {
  using Params = AttnParams;
  static constexpr int FMT = 1, BLOCK_N = 256, STAGES = 4, A_MN = 0, B_MN = 0;
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
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row) {
    const size_t grow = static_cast<size_t>(ti.m0) + row;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
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
  static constexpr int FMT = 1, BLOCK_N = 256, STAGES = 4, A_MN = 1, B_MN = 1;
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
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row) {
    float* dst = P.kv_part + ((static_cast<size_t>(ti.b) * P.nsplit + ti.aux) * 128 + row) * 2048 + ti.n0;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
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
struct AttnOut {
  using Params = AttnParams;
  static constexpr int FMT = 1, BLOCK_N = 256, STAGES = 4, A_MN = 0, B_MN = 1;
  __device__ static void prefetch(const Params& P) {
    tma_prefetch_desc(&P.tmP);
    tma_prefetch_desc(&P.tmVUmn);
    tma_prefetch_desc(&P.tmQK);
    tma_prefetch_desc(&P.tmKVmn);
  }
  __device__ static int num_tiles(const Params& P) { return (P.B * P.Sp / GEMM_BLOCK_M) * 8; }
  __device__ static void tile_info(const Params& P, int tile, TileInfo& ti) {
    const int mt = tile >> 3;
    const int nt = tile & 7;
    ti.m0 = mt * GEMM_BLOCK_M;
    ti.n0 = nt * 128;  // first v channel; the u channel is 1024 + n0
    ti.b = ti.m0 / P.Sp;
    ti.t0 = ti.m0 - ti.b * P.Sp;
    ti.nkb = 6;
    ti.aux = nt;
  }
  __device__ static void load(const Params& P, const TileInfo& ti, int kb, uint32_t sa, uint32_t sb, uint32_t bar) {
    if (kb < 4) {
      const int g0 = (ti.t0 / 256) * 256;
      tma_load_3d(sa, &P.tmP, bar, kb * 64, ti.t0, ti.b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = (j < 2 ? ti.n0 + j * 64 : 1024 + ti.n0 + (j - 2) * 64);
        tma_load_3d(sb + j * 8192, &P.tmVUmn, bar, ch, g0 + kb * 64, ti.b);
      }
    } else {
      const int k0 = (kb - 4) * 64;
      tma_load_3d(sa, &P.tmQK, bar, 128 + k0, ti.t0, ti.b);  // lin_q
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = (j < 2 ? ti.n0 + j * 64 : 1024 + ti.n0 + (j - 2) * 64);
        tma_load_3d(sb + j * 8192, &P.tmKVmn, bar, ch, k0, ti.b);
      }
    }
  }
  __device__ static void epilogue(const Params& P, const TileInfo& ti, uint32_t tacc, int row) {
    const int t = ti.t0 + row;
    const bool valid = t < P.S;
    const size_t grow = static_cast<size_t>(ti.m0) + row;
    float ssq = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 16) {
      float av[16], au[16];
      tmem_ld16(tacc + c0, av);
      tmem_ld16(tacc + 128 + c0, au);
      tmem_ld_wait();
      if (!valid) continue;
      const uint4* vp = reinterpret_cast<const uint4*>(P.vu + grow * 2048 + ti.n0 + c0);
      const uint4* up = reinterpret_cast<const uint4*>(P.vu + grow * 2048 + 1024 + ti.n0 + c0);
      uint4 vr[2] = {vp[0], vp[1]};
      uint4 ur[2] = {up[0], up[1]};
      const __nv_bfloat16* vb = reinterpret_cast<const __nv_bfloat16*>(vr);
      const __nv_bfloat16* ub = reinterpret_cast<const __nv_bfloat16*>(ur);
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v = __bfloat162float(vb[j]);
        const float u = __bfloat162float(ub[j]);
        const float x = (au[j] * v) * sigmoid_f(av[j] * u);
        o[j] = x;
        ssq += x * x;
      }
      uint4* dst = reinterpret_cast<uint4*>(P.o + grow * 1024 + ti.n0 + c0);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        dst[j] = make_uint4(pack_bf16(o[8 * j], o[8 * j + 1]), pack_bf16(o[8 * j + 2], o[8 * j + 3]),
                            pack_bf16(o[8 * j + 4], o[8 * j + 5]), pack_bf16(o[8 * j + 6], o[8 * j + 7]));
    }
    P.o_ss[grow * 8 + ti.aux] = valid ? ssq : 0.f;
  }
};

}  // namespace tdz
