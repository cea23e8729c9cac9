// Linear + SiLU + ConvModule (depthwise k=17 over time + residual) in one kernel, with a two-stage pipelined
// epilogue (mossformer_block.py:89-102, conv_module.py:209-220).
//
//   warp 0        TMA producer          3-stage ring of {A 128 x 64, B 256 x 64} bf16 k-blocks
//   warp 1        tcgen05.mma issuer    fp32 accumulators in TMEM, two accumulator stages
//   warps 2..9    "activation" warps    tcgen05.ld -> ScaleNorm row scale + bias + SiLU -> fp32 panel in shared
//                                       memory (thread = tile row; MUFU-bound)
//   warps 10..17  "convolution" warps   register sliding window down the panel columns: y + dwconv17(y), then the
//                                       op-specific tail and the global stores (thread = 2 columns x 14 rows;
//                                       FMA-bound)
// The two epilogue stages work on different 64-column panels at the same time (two panel buffers, handed over
// with named barriers), so the special-function and FMA pipes overlap instead of alternating.
//
// An accumulator tile covers 128 consecutive frames of ONE sample, of which the inner 112 are outputs and 8 on
// each side are the halo of the convolution (tiles overlap by 16 rows; TMA zero-fills rows outside the sample;
// activation rows outside [0,S) are written as zeros = the convolution's zero padding).  The pre-convolution
// activation (8.7 KB per frame for to_hidden|to_qk) never goes to HBM.
#pragma once
#include "gemm_cfgs.cuh"

namespace tdz {

constexpr int CV_STAGES = 3;
constexpr int CV_STAGE_BYTES = GEMM_STAGE_A_BYTES + 256 * 128;  // 48 KB
constexpr int CV_PANEL_COLS = 64;
constexpr int CV_PANEL_LD = 68;                                 // floats per panel row (64 + 4: conflict-free rows)
constexpr int CV_PANEL_BYTES = 128 * CV_PANEL_LD * 4;           // 34 816 B
constexpr int CV_SMEM_BYTES = CV_STAGES * CV_STAGE_BYTES + 256 + 2 * CV_PANEL_BYTES + 1024;
constexpr int CV_THREADS = 64 + 256 + 256;

// named barriers: 1 + buffer = panel full (activation -> convolution), 3 + buffer = panel free
__device__ __forceinline__ void nbar_sync(int id) { asm volatile("bar.sync %0, 512;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id) { asm volatile("bar.arrive %0, 512;" ::"r"(id) : "memory"); }

template <int MODE>
struct ConvTile {
  __device__ static int num_tiles(const LinearParams& P) { return P.B * P.tps * P.n_tiles; }
  __device__ static void tile_info(const LinearParams& P, int tile, TileInfo& ti) {
    const int mt = tile / P.n_tiles;
    const int nt = tile - mt * P.n_tiles;
    ti.b = mt / P.tps;
    const int j = mt - ti.b * P.tps;
    ti.t0 = j * CONV_ROWS - 8;  // frame of tile row 0 (negative for the first tile of a sample)
    ti.m0 = ti.b * P.Sp;        // first row of the sample in the token space
    ti.n0 = nt * 256;
    ti.nkb = P.K / 64;
    ti.aux = nt;
  }
};

// tail of the op for 7 consecutive output frames of two adjacent columns; PRED = false when all rows are inside
template <int MODE, bool PRED>
__device__ __forceinline__ void conv_emit(const LinearParams& P, int c, int tt0, size_t grow0, int nrow,
                                          const float2 (&acc)[7]) {
  const EpiConv& cv = P.cv;
  if constexpr (MODE == CONV_VUQK) {
    if (c < 2048) {  // warp-uniform: a panel is either all (v|u) or all qk
      __nv_bfloat16* dst = cv.vu + grow0 * 2048 + c;
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (!PRED || j < nrow)
          *reinterpret_cast<uint32_t*>(dst + static_cast<size_t>(j) * 2048) = pack_bf16(acc[j].x, acc[j].y);
    } else {
      const int qc = c - 2048;
      __nv_bfloat16* dst = cv.qk4 + grow0 * 512 + qc;
      float2 cs[7];
#pragma unroll
      for (int j = 0; j < 7; ++j)
        cs[j] = (qc < 32 && (!PRED || j < nrow)) ? cv.rot[(tt0 + j) * 16 + (qc >> 1)] : make_float2(1.f, 0.f);
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {
        const float g0 = __ldg(cv.gamma + h * 128 + qc), g1 = __ldg(cv.gamma + h * 128 + qc + 1);
        const float b0 = __ldg(cv.beta + h * 128 + qc), b1 = __ldg(cv.beta + h * 128 + qc + 1);
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const float x0 = fmaf(acc[j].x, g0, b0), x1 = fmaf(acc[j].y, g1, b1);
          // rotary on interleaved pairs of dims 0..31 (identity rotation elsewhere)
          const float r0v = x0 * cs[j].x - x1 * cs[j].y;
          const float r1v = x1 * cs[j].x + x0 * cs[j].y;
          if (!PRED || j < nrow)
            *reinterpret_cast<uint32_t*>(dst + static_cast<size_t>(j) * 512 + h * 128) = pack_bf16(r0v, r1v);
        }
      }
    }
  }
  if constexpr (MODE == CONV_RESX) {
    const float* src = cv.x_in + grow0 * 512 + c;
    float* dst = cv.x_out + grow0 * 512 + c;
    float2 r[7];
#pragma unroll
    for (int j = 0; j < 7; ++j)
      r[j] = (!PRED || j < nrow) ? *reinterpret_cast<const float2*>(src + static_cast<size_t>(j) * 512)
                                 : make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (!PRED || j < nrow)
        *reinterpret_cast<float2*>(dst + static_cast<size_t>(j) * 512) = make_float2(r[j].x + acc[j].x, r[j].y + acc[j].y);
  }
  if constexpr (MODE == CONV_UV) {
    float* dst = cv.xuv + grow0 * 512 + c;
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (!PRED || j < nrow) *reinterpret_cast<float2*>(dst + static_cast<size_t>(j) * 512) = acc[j];
    if (c < 256) {
      __nv_bfloat16* db = cv.xubf + grow0 * 256 + c;
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (!PRED || j < nrow)
          *reinterpret_cast<uint32_t*>(db + static_cast<size_t>(j) * 256) = pack_bf16(acc[j].x, acc[j].y);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(CV_THREADS, 1) gemm_conv_kernel(const __grid_constant__ LinearParams P) {
  using Tile = ConvTile<MODE>;
  constexpr int BLOCK_N = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + CV_STAGES * CV_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (CV_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * CV_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * CV_STAGES + 2 + s); };
  float* panels = reinterpret_cast<float*>(smem_al + CV_STAGES * CV_STAGE_BYTES + 256);
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
    for (int s = 0; s < CV_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&s_tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const int ntiles = Tile::num_tiles(P);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        TileInfo ti;
        Tile::tile_info(P, tile, ti);
        for (int kb = 0; kb < ti.nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), CV_STAGE_BYTES);
          const uint32_t sa = smem_base + stage * CV_STAGE_BYTES;
          const int trow = ti.t0 - (kb < P.shift_kblocks ? 1 : 0);  // token shift (mossformer_block.py:204-207)
          tma_load_3d(sa, &P.tmA, full_bar(stage), kb * 64, trow, ti.b);
          tma_load_2d(sa + GEMM_STAGE_A_BYTES, &P.tmB, full_bar(stage), kb * 64, ti.n0);
          if (++stage == CV_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc(1, GEMM_BLOCK_M, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        TileInfo ti;
        Tile::tile_info(P, tile, ti);
        const int as = it & 1;
        mbar_wait(tempty_bar(as), ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < ti.nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * CV_STAGE_BYTES;
          const uint32_t sb = sa + GEMM_STAGE_A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * 32u, 16u, 1024u);
            const uint64_t db = umma_smem_desc(sb + k * 32u, 16u, 1024u);
            umma_f16(tacc, da, db, IDESC, (kb | k) != 0);
          }
          umma_commit(empty_bar(stage));
          if (kb == ti.nkb - 1) umma_commit(tfull_bar(as));
          if (++stage == CV_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp < 10) {
    // ---------------- activation warps: TMEM -> row scale + bias + SiLU -> panel
    const int q = warp & 3;              // TMEM lane quarter this warp may read
    const int ch = (warp - 2) >> 2;      // which 32 of the panel's 64 columns
    const int row = q * 32 + lane;
    const EpiGeneric& e = P.e;
    int it = 0;
    uint32_t np = 0;  // panels handed over so far (buffer = np & 1)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      TileInfo ti;
      Tile::tile_info(P, tile, ti);
      const int t = ti.t0 + row;
      const bool valid = t >= 0 && t < P.S;
      float rs = 1.f;
      if (valid) {
        const size_t grow = static_cast<size_t>(ti.m0) + t;
        if constexpr (MODE == CONV_VUQK) {
          const float4 cur = *reinterpret_cast<const float4*>(e.ss_in + grow * 4);
          float ss = cur.z + cur.w;  // channels 256..511 of this frame
          if (t > 0) {
            const float4 prv = *reinterpret_cast<const float4*>(e.ss_in + (grow - 1) * 4);
            ss += prv.x + prv.y;     // channels 0..255 of the previous frame (token shift)
          }
          rs = scalenorm_rscale(ss, e.ss_dim_rsqrt);
        }
        if constexpr (MODE == CONV_RESX) {
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(e.ss_in + grow * 16 + 4 * i);
            ss += (a.x + a.y) + (a.z + a.w);
          }
          rs = scalenorm_rscale(ss, e.ss_dim_rsqrt);
        }
      }
      const int as = it & 1;
      mbar_wait(tfull_bar(as), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int sp = 0; sp < 4; ++sp) {
        if (ti.n0 + sp * CV_PANEL_COLS >= P.N) break;
        const int c0 = sp * CV_PANEL_COLS + ch * 32;
        float v[32], bias[32];
        tmem_ld16(tacc + c0, v);
        tmem_ld16(tacc + c0 + 16, v + 16);
        ld_f32x16(e.bias + ti.n0 + c0, bias);
        ld_f32x16(e.bias + ti.n0 + c0 + 16, bias + 16);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = valid ? silu_f(fmaf(v[j], rs, bias[j])) : 0.f;
        const int buf = np & 1;
        nbar_sync(3 + buf);  // the convolution warps are done with this buffer
        float4* dst = reinterpret_cast<float4*>(panels + buf * (128 * CV_PANEL_LD) + row * CV_PANEL_LD + ch * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __threadfence_block();
        nbar_arrive(1 + buf);
        ++np;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
  } else {
    // ---------------- convolution warps: y + dwconv17(y) down the panel columns, tail, stores
    const int tid = threadIdx.x - 320;
    const int cp = tid & 31;   // column pair inside the panel
    const int rg = tid >> 5;   // 14 output rows: tile rows 8 + 14 rg ..
    const EpiConv& cv = P.cv;
    nbar_arrive(3);            // both panel buffers start free
    nbar_arrive(4);
    uint32_t np = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      TileInfo ti;
      Tile::tile_info(P, tile, ti);
      const size_t srow = static_cast<size_t>(ti.m0);
#pragma unroll 1
      for (int sp = 0; sp < 4; ++sp) {
        const int pc0 = ti.n0 + sp * CV_PANEL_COLS;
        if (pc0 >= P.N) break;
        const int c = pc0 + 2 * cp;
        float2 wt[17];
#pragma unroll
        for (int k = 0; k < 17; ++k) wt[k] = __ldg(reinterpret_cast<const float2*>(cv.dw_t + k * cv.ldw + c));
        const int buf = np & 1;
        nbar_sync(1 + buf);  // panel written
        const float* pcol = panels + buf * (128 * CV_PANEL_LD) + 2 * cp;
#pragma unroll 1
        for (int it = 0; it < 2; ++it) {
          const int r0 = 14 * rg + 7 * it;
          float2 win[23];
#pragma unroll
          for (int i = 0; i < 23; ++i) win[i] = *reinterpret_cast<const float2*>(pcol + (r0 + i) * CV_PANEL_LD);
          float2 acc[7];
#pragma unroll
          for (int j = 0; j < 7; ++j) acc[j] = win[j + 8];
#pragma unroll
          for (int k = 0; k < 17; ++k) {
#pragma unroll
            for (int j = 0; j < 7; ++j) acc[j] = fma2(wt[k], win[j + k], acc[j]);
          }
          if (it == 1) nbar_arrive(3 + buf);  // last read of the panel is done: hand the buffer back early
          const int tt0 = ti.t0 + r0 + 8;     // frame of acc[0]
          const int nrow = P.S - tt0;         // rows j < nrow are inside the sample
          const size_t grow0 = srow + tt0;
          if (nrow >= 7) {
            conv_emit<MODE, false>(P, c, tt0, grow0, nrow, acc);
          } else if (nrow > 0) {
            conv_emit<MODE, true>(P, c, tt0, grow0, nrow, acc);
          }
        }
        ++np;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int MODE>
cudaError_t launch_gemm_conv(const LinearParams& P, int ntiles, int num_sms, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_conv_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         CV_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (ntiles <= 0) return cudaSuccess;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  gemm_conv_kernel<MODE><<<grid, CV_THREADS, CV_SMEM_BYTES, st>>>(P);
  return cudaGetLastError();
}

}  // namespace tdz
