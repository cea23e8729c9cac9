// Linear + SiLU + ConvModule (depthwise k=17 over time + residual) in one kernel, channel-major accumulator
// (mossformer_block.py:89-102, conv_module.py:209-220).
//
// The product is computed transposed, D[channel][frame] = sum_k W[channel][k] X[frame][k]: the weight tile is the
// M=128 operand, 256 consecutive frames of ONE sample are the N operand.  In TMEM a lane is then an output channel
// and the columns are time, so an epilogue thread reads "its" channel's frames straight into a register sliding
// window with tcgen05.ld and runs the time convolution there: the pre-convolution activation (8.7 KB per frame
// for to_hidden|to_qk) goes neither to HBM nor to shared memory, bias / taps / OffsetScale are per-thread
// constants, and shared memory bandwidth is left to the MMA operands (4-stage ring).
//
//   warp 0        TMA producer          {W 128 x 64, X 256 x 64} bf16 k-blocks, token shift = X row offset -1
//   warp 1        tcgen05.mma issuer    fp32 accumulators in TMEM, two accumulator stages
//   warps 2..13   epilogue              lane quarter (32 channels) x time third (80 output frames + 16 halo)
//
// A tile covers frames [240 j - 8, 240 j + 248) of a sample, of which the inner 240 are outputs (TMA zero-fills
// rows outside the sample; activations of frames outside [0,S) are forced to zero = the convolution's padding).
#pragma once
#include <type_traits>

#include "gemm_cfgs.cuh"

namespace tdz {

constexpr int CT_STAGES = 4;
constexpr int CT_STAGE_BYTES = GEMM_STAGE_A_BYTES + 256 * 128;  // W tile 16 KB + X tile 32 KB
constexpr int CT_ROWS = 240;                                     // output frames per tile
constexpr int CT_EPI_WARPS = 12;                                 // 4 lane quarters x 3 time thirds
constexpr int CT_THREADS = 64 + 32 * CT_EPI_WARPS;
constexpr int CT_SCR = 96;                                       // frames per epilogue warp (80 outputs + 16 halo)
constexpr int CT_CONST_FLOATS = 18 * 128;                        // per channel of the tile: 17 taps + bias
// Both scratch tables are written (cp.async) one tile ahead into the other of two buffers.  Warps share table
// entries, and the two TMEM stages would let a fast warp run two tiles ahead of a slow one, so the epilogue warps
// meet at a named barrier at the top of every tile: nobody overwrites a buffer that a straggler has yet to read.
constexpr int CT_NBUF = 2;
constexpr int CT_HRS_FLOATS = 3 * CT_SCR;                        // one row of scales per time third
// operand ring | barriers | frame scales [2 buffers] | per-channel constants [2 buffers]
constexpr int CT_SMEM_BYTES =
    CT_STAGES * CT_STAGE_BYTES + 256 + CT_NBUF * CT_HRS_FLOATS * 4 + CT_NBUF * CT_CONST_FLOATS * 4 + 1024;

// 4-byte asynchronous global -> shared copy (the per-tile constants of the NEXT tile are fetched this way while
// the current tile is being processed, so no global-load latency sits on the per-tile critical path)
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  tmem_ld16(taddr, v);
  tmem_ld16(taddr + 16, v + 16);
}

// SiLU through one MUFU.TANH: x*sigmoid(x) = h + h*tanh(h), h = x/2 (tanh.approx: relative error 2^-11 on the
// tanh, i.e. an absolute error below 5e-4*|h| - smaller than the bf16 rounding of the operand copy that follows).
__device__ __forceinline__ float silu_half(float h) { return fmaf(h, tanh_approx(h), h); }

struct ConvTTile {
  int b, j, ct, t0, srow;
};
// Tile order: channel tile fastest, then time tile, then sample; CTA i takes tiles i, i + grid, ... (concurrent
// CTAs then share X tiles in L2 and walk through DRAM together; contiguous per-CTA ranges were measured 40 % slower
// on TO_OUT).  The coordinates of the next tile follow from the current ones by increments - the two integer
// divisions per tile and per role were a measurable part of the epilogue preamble.
__device__ __forceinline__ void convt_tile(const LinearParams& P, int nct, int tile, ConvTTile& ti) {
  const int mt = tile / nct;            // (sample, time tile)
  ti.ct = tile - mt * nct;
  ti.b = mt / P.tps;
  ti.j = mt - ti.b * P.tps;
  ti.t0 = ti.j * CT_ROWS - 8;           // frame of accumulator column 0
  ti.srow = ti.b * P.Sp;
}
struct ConvTStep {
  int nct, dct, dmt;  // channel tiles (or tile pairs) per time tile; stride = dmt * nct + dct
};
__device__ __forceinline__ ConvTStep convt_step(int nct, int stride) {
  ConvTStep st;
  st.nct = nct;
  st.dmt = stride / nct;
  st.dct = stride - st.dmt * nct;
  return st;
}
__device__ __forceinline__ void convt_next(const LinearParams& P, const ConvTStep& st, ConvTTile& ti) {
  ti.ct += st.dct;
  ti.j += st.dmt;
  if (ti.ct >= st.nct) {
    ti.ct -= st.nct;
    ++ti.j;
  }
  while (ti.j >= P.tps) {
    ti.j -= P.tps;
    ++ti.b;
    ti.srow += P.Sp;
  }
  ti.t0 = ti.j * CT_ROWS - 8;
}

// CG2 = true: the CTA pair of a cluster owns two neighbouring channel tiles of the same time tile and executes one
// tcgen05.mma.cta_group::2 (M = 256 channels) per k-step; each CTA stages its own weight tile and only half of the
// X rows (32 KB per k-block instead of 48).  Used where the operand traffic matters (to_out, K = 1024: 2/3 of the
// tile's L2 -> SM bytes were operands) and the channel-tile count is even; the epilogue is the same.
constexpr int CT2_STAGES = 6;
constexpr int CT2_STAGE_BYTES = GEMM_STAGE_A_BYTES + 128 * 128;  // W tile 16 KB + half an X tile 16 KB
constexpr int CT2_SMEM_BYTES =
    CT2_STAGES * CT2_STAGE_BYTES + 256 + CT_NBUF * CT_HRS_FLOATS * 4 + CT_NBUF * CT_CONST_FLOATS * 4 + 1024;

template <int MODE, bool CG2>
__device__ __forceinline__ void gemm_convt_body(const LinearParams& P) {
  constexpr int LDW = (MODE == CONV_VUQK) ? 2176 : 512;  // leading dimension of the tap-major tap table
  constexpr int CT_STAGES = CG2 ? CT2_STAGES : tdz::CT_STAGES;  // (shadow the single-CTA constants)
  constexpr int CT_STAGE_BYTES = CG2 ? CT2_STAGE_BYTES : tdz::CT_STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + CT_STAGES * CT_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (CT_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * CT_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * CT_STAGES + 2 + s); };
  float* scratch = reinterpret_cast<float*>(smem_al + CT_STAGES * CT_STAGE_BYTES + 256);
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG2 ? cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmA);
    tma_prefetch_desc(&P.tmB);
    if constexpr (CG2) tma_prefetch_desc(&P.tmAh);
    for (int s = 0; s < CT_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), CG2 ? 2 * CT_EPI_WARPS : CT_EPI_WARPS);  // CG2: the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG2) {
      tmem_alloc_cg2(smem_u32(&s_tmem_base), 512);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(smem_u32(&s_tmem_base), 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  // work items: (channel tile, time tile); CG2: (pair of channel tiles, time tile), one per cluster
  const int nct = CG2 ? P.n_tiles / 2 : P.n_tiles;
  const int ntiles = P.B * P.tps * nct;
  const int nkb = P.K / 64;
  const int first = CG2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int stride = CG2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const ConvTStep tstep = convt_step(nct, stride);
  auto chan_tile = [&](const ConvTTile& t) { return CG2 ? 2 * t.ct + static_cast<int>(rank) : t.ct; };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      ConvTTile ti;
      convt_tile(P, nct, first, ti);
      [[maybe_unused]] const uint32_t leader_bars = CG2 ? mapa_shared(bar_base, 0) : 0u;
      for (int tile = first; tile < ntiles; tile += stride, convt_next(P, tstep, ti)) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * CT_STAGE_BYTES;
          const int trow = ti.t0 - (kb < P.shift_kblocks ? 1 : 0);  // token shift (mossformer_block.py:204-207)
          if constexpr (CG2) {
            // both CTAs' bytes are counted on the leader's barrier; each CTA: its weight tile + its half of X
            if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * CT_STAGE_BYTES);
            const uint32_t fb = leader_bars + 8u * stage;
            tma_load_2d_cg2(sa, &P.tmB, fb, kb * 64, chan_tile(ti) * 128);
            tma_load_3d_cg2(sa + GEMM_STAGE_A_BYTES, &P.tmAh, fb, kb * 64, trow + 128 * static_cast<int>(rank), ti.b);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), CT_STAGE_BYTES);
            tma_load_2d(sa, &P.tmB, full_bar(stage), kb * 64, ti.ct * 128);                 // weights: M operand
            tma_load_3d(sa + GEMM_STAGE_A_BYTES, &P.tmA, full_bar(stage), kb * 64, trow, ti.b);  // frames: N operand
          }
          if (++stage == CT_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // CG2: only the leader issues MMAs
      constexpr uint32_t IDESC = umma_idesc(1, CG2 ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M, 256, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = first; tile < ntiles; tile += stride, ++it) {
        const int as = it & 1;
        mbar_wait(tempty_bar(as), ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * 256;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * CT_STAGE_BYTES;
          const uint32_t sb = sa + GEMM_STAGE_A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * 32u, 16u, 1024u);
            const uint64_t db = umma_smem_desc(sb + k * 32u, 16u, 1024u);
            if constexpr (CG2) umma_f16_cg2(tacc, da, db, IDESC, (kb | k) != 0);
            else umma_f16(tacc, da, db, IDESC, (kb | k) != 0);
          }
          if constexpr (CG2) {
            umma_commit_cg2_mc(empty_bar(stage), 0x3);
            if (kb == nkb - 1) umma_commit_cg2_mc(tfull_bar(as), 0x3);
          } else {
            umma_commit(empty_bar(stage));
            if (kb == nkb - 1) umma_commit(tfull_bar(as));
          }
          if (++stage == CT_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ---------------- epilogue: thread = one output channel, 80 output frames (5 x 16) + 16 halo frames
    const int q = warp & 3;           // TMEM lane quarter -> 32 channels
    const int tq = (warp - 2) >> 2;   // time third
    const int col0 = 80 * tq;         // first accumulator column this warp reads
    // shared scratch: the 96 per-frame scales of this warp's time third (0.5 / ScaleNorm denominator, precomputed
    // per frame by rowscale_kernel) and the per-channel constants of the tile (taps, bias); warps that share an
    // entry copy the same values (benign)
    float* hrs_buf = scratch + tq * CT_SCR;                                // + buf * CT_HRS_FLOATS
    float* cst_buf = scratch + CT_NBUF * CT_HRS_FLOATS + q * 32 + lane;    // + buf * CT_CONST_FLOATS + k * 128
    const EpiGeneric& e = P.e;
    const EpiConv& cv = P.cv;
    auto prefetch = [&](const ConvTTile& tn, int buf) {  // asynchronous: completes before the tile that uses `buf`
      const int c = chan_tile(tn) * 128 + q * 32 + lane;
      float* cs = cst_buf + buf * CT_CONST_FLOATS;
#pragma unroll
      for (int k = 0; k < 17; ++k) cp_async4(cs + k * 128, cv.dw_t + k * LDW + c);
      cp_async4(cs + 17 * 128, e.bias + c);
      if constexpr (MODE != CONV_UV) {
        float* hs = hrs_buf + buf * CT_HRS_FLOATS;
        const int tb = tn.t0 + col0;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int i = lane + 32 * r;
          const int t = min(max(tb + i, 0), P.S - 1);  // frames outside [0,S) are masked later; any valid address
          cp_async4(hs + i, e.ss_in + static_cast<size_t>(tn.srow) + t);
        }
      }
    };
    int it = 0;
    ConvTTile ti, tn;
    [[maybe_unused]] const uint32_t leader_tempty = CG2 ? mapa_shared(tempty_bar(0), 0) : 0u;
    convt_tile(P, nct, first, ti);
    if (first < ntiles) prefetch(ti, 0);
    for (int tile = first; tile < ntiles; tile += stride, ++it, ti = tn) {
      tn = ti;
      convt_next(P, tstep, tn);
      const int c = chan_tile(ti) * 128 + q * 32 + lane;  // output channel of this thread
      const int buf = it % CT_NBUF;
      cp_async_wait_all();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * CT_EPI_WARPS) : "memory");  // epilogue warps only
      float2 wt[17];  // (w_k, w_k): operand of the packed FMA
#pragma unroll
      for (int k = 0; k < 17; ++k) {
        const float wk = cst_buf[buf * CT_CONST_FLOATS + k * 128];
        wt[k] = make_float2(wk, wk);
      }
      const float hb = 0.5f * cst_buf[buf * CT_CONST_FLOATS + 17 * 128];
      const float2 hb2 = make_float2(hb, hb);
      const float* hrs_s = hrs_buf + buf * CT_HRS_FLOATS;
      if (tile + stride < ntiles) prefetch(tn, (it + 1) % CT_NBUF);
      const int tbase = ti.t0 + col0;  // frame of accumulator column col0
      const bool all_valid = tbase >= 0 && tbase + 96 <= P.S;

      const int as = it & 1;
      mbar_wait(tfull_bar(as), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * 256 + (static_cast<uint32_t>(q * 32) << 16) + col0;

      // SiLU(scale * acc + bias) of N freshly loaded accumulator columns, in place; sidx = index of the first one
      // in the per-frame scratch (a multiple of 4)
      // window of 32 consecutive frames as 16 register pairs (frames 2m, 2m+1): the operand form of FFMA2
      float2 P0[16];
      float* win = reinterpret_cast<float*>(P0);
      auto activate_t = [&](float* w, int sidx, auto n_tag, auto masked_tag) {
        constexpr int N = decltype(n_tag)::value;
        constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
        for (int i = 0; i < N; i += 4) {  // four frames at a time keeps the scale / mask operands short-lived
          float4 hs = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
          if constexpr (MODE != CONV_UV) hs = *reinterpret_cast<const float4*>(hrs_s + sidx + i);
          // h = scale * acc + bias/2 and h + h tanh(h) as packed FMAs on frame pairs
          const float2 h01 = fma2(make_float2(w[i + 0], w[i + 1]), make_float2(hs.x, hs.y), hb2);
          const float2 h23 = fma2(make_float2(w[i + 2], w[i + 3]), make_float2(hs.z, hs.w), hb2);
          const float2 a01 = fma2(h01, make_float2(tanh_approx(h01.x), tanh_approx(h01.y)), h01);
          const float2 a23 = fma2(h23, make_float2(tanh_approx(h23.x), tanh_approx(h23.y)), h23);
          if constexpr (!MASKED) {
            w[i + 0] = a01.x;
            w[i + 1] = a01.y;
            w[i + 2] = a23.x;
            w[i + 3] = a23.y;
          } else {
            // frames outside [0,S) are SELECTED to zero (their accumulators may hold anything, also NaN: the rows
            // of the N operand beyond S are never written by the producer kernels)
            const unsigned t = static_cast<unsigned>(tbase + sidx + i);  // negative frames wrap to huge values
            const unsigned Su = static_cast<unsigned>(P.S);
            w[i + 0] = (t + 0u < Su) ? a01.x : 0.f;
            w[i + 1] = (t + 1u < Su) ? a01.y : 0.f;
            w[i + 2] = (t + 2u < Su) ? a23.x : 0.f;
            w[i + 3] = (t + 3u < Su) ? a23.y : 0.f;
          }
        }
      };
      auto activate = [&](float* w, int sidx, auto n_tag) {  // warp-uniform: interior tiles carry no masks
        if (all_valid) {
          activate_t(w, sidx, n_tag, std::false_type{});
        } else {
          activate_t(w, sidx, n_tag, std::true_type{});
        }
      };
      tmem_ld32(tacc, win);
      tmem_ld_wait();
      activate(win, 0, std::integral_constant<int, 32>{});
#pragma unroll 1
      for (int itn = 0; itn < 5; ++itn) {
        [[maybe_unused]] float rin[16];
        if constexpr (MODE == CONV_RESX) {  // residual stream of these 16 frames: in flight during the FMAs
          const int tt0r = tbase + 8 + 16 * itn;
          const float* src = cv.x_in + (static_cast<size_t>(ti.srow) + tt0r) * 512 + c;
          if (tt0r + 16 <= P.S) {  // warp-uniform fast path: no per-row predicates
#pragma unroll
            for (int j = 0; j < 16; ++j) rin[j] = src[static_cast<size_t>(j) * 512];
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) rin[j] = (tt0r + j < P.S) ? src[static_cast<size_t>(j) * 512] : 0.f;
          }
        }
        // y + dwconv17(y) for 16 frames with packed FMAs (two fp32 FMAs per issue slot).  Even taps accumulate
        // into pairs (out[2m], out[2m+1]), odd taps into pairs (out[2m-1], out[2m]): both then read ALIGNED window
        // pairs P0[m+q], so no shifted copy of the window is needed; the two partial sums are added at the end.
        float2 accA[8], accB[9];
#pragma unroll
        for (int m = 0; m < 8; ++m) accA[m] = P0[m + 4];
#pragma unroll
        for (int m = 0; m < 9; ++m) accB[m] = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 9; ++q) {
#pragma unroll
          for (int m = 0; m < 8; ++m) accA[m] = fma2(wt[2 * q], P0[m + q], accA[m]);
          if (q < 8) {
#pragma unroll
            for (int m = 0; m < 9; ++m) accB[m] = fma2(wt[2 * q + 1], P0[m + q], accB[m]);
          }
        }
        float acc[16];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          acc[2 * m] = accA[m].x + accB[m].y;
          acc[2 * m + 1] = accA[m].y + accB[m + 1].x;
        }
        if (itn < 4) {  // slide the window: the next 16 accumulator columns
#pragma unroll
          for (int i = 0; i < 16; ++i) win[i] = win[16 + i];
          tmem_ld16(tacc + 32 + 16 * itn, win + 16);
          tmem_ld_wait();
          if (itn == 3) {  // last TMEM read of this tile is done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG2) mbar_arrive_cluster(leader_tempty + 8u * as);
              else mbar_arrive(tempty_bar(as));
            }
          }
          activate(win + 16, 32 + 16 * itn, std::integral_constant<int, 16>{});
        }
        const int tt0 = tbase + 8 + 16 * itn;  // frame of acc[0]
        const int nrow = min(P.S - tt0, 16);   // rows j < nrow are inside the sample (may be <= 0)
        const size_t grow0 = static_cast<size_t>(ti.srow) + tt0;
        // rows j < nrow are stored; interior tiles (nrow == 16, warp-uniform) take the unpredicated form
        auto store_rows = [&](auto* dst, size_t ld, auto conv) {
          if (nrow >= 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[static_cast<size_t>(j) * ld] = conv(j);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nrow) dst[static_cast<size_t>(j) * ld] = conv(j);
          }
        };
        if constexpr (MODE == CONV_VUQK) {
          if (c < 2048) {  // warp-uniform (the qk channels are one whole channel tile)
            store_rows(cv.vu + grow0 * 2048 + c, 2048, [&](int j) { return __float2bfloat16(acc[j]); });
          } else {
            // to_qk channels: fp32, OffsetScale + rotary + bf16 split happen in qk_heads_kernel (the 32 rotary
            // channels all sit in one TMEM lane quarter, i.e. on one scheduler: doing that work here serialises it)
            store_rows(cv.qkf + grow0 * 128 + (c - 2048), 128, [&](int j) { return acc[j]; });
          }
        }
        if constexpr (MODE == CONV_RESX) {
          store_rows(cv.x_out + grow0 * 512 + c, 512, [&](int j) { return rin[j] + acc[j]; });
        }
        if constexpr (MODE == CONV_UV) {
          store_rows(cv.xuv + grow0 * 512 + c, 512, [&](int j) { return acc[j]; });
          if (c < 256) store_rows(cv.xubf + grow0 * 256 + c, 256, [&](int j) { return __float2bfloat16(acc[j]); });
        }
      }
    }
  }

  tc_fence_before();
  if constexpr (CG2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    if constexpr (CG2) tmem_dealloc_cg2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

template <int MODE>
__global__ void __launch_bounds__(CT_THREADS, 1) gemm_convt_kernel(const __grid_constant__ LinearParams P) {
  gemm_convt_body<MODE, false>(P);
}
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CT_THREADS, 1)
    gemm_convt_cg2_kernel(const __grid_constant__ LinearParams P) {
  gemm_convt_body<MODE, true>(P);
}

template <int MODE>
cudaError_t launch_gemm_convt(const LinearParams& P, int ntiles, int num_sms, cudaStream_t st) {
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(gemm_convt_kernel<MODE>), CT_SMEM_BYTES, configured); e != cudaSuccess)
    return e;
  if (ntiles <= 0) return cudaSuccess;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  gemm_convt_kernel<MODE><<<grid, CT_THREADS, CT_SMEM_BYTES, st>>>(P);
  return cudaGetLastError();
}

// cta_group::2 form; needs an even number of channel tiles and the tmB3 / tmAh maps.  npairs = work items.
template <int MODE>
cudaError_t launch_gemm_convt_cg2(const LinearParams& P, int npairs, int num_sms, cudaStream_t st) {
  constexpr int smem = CT2_SMEM_BYTES;
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(gemm_convt_cg2_kernel<MODE>), smem, configured); e != cudaSuccess)
    return e;
  if (npairs <= 0) return cudaSuccess;
  const int grid = 2 * npairs < num_sms ? 2 * npairs : (num_sms & ~1);
  gemm_convt_cg2_kernel<MODE><<<grid, CT_THREADS, smem, st>>>(P);
  return cudaGetLastError();
}

}  // namespace tdz
