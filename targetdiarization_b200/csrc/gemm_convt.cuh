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
//   warps 2..13   epilogue              lane quarter (32 channels) x time third (80 output frames + 16 halo);
//                                       per warp a software pipeline tcgen05.ld | SiLU (MUFU) | k17 window (FMA pipe)
//                                       over a four-block register ring, see convt_epilogue_tile
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
constexpr int CT_CPC = 20;                                       // floats per channel: 17 taps, bias / 2, 2 pad (80 B)
// operand ring | barriers
constexpr int CT_SMEM_BYTES = CT_STAGES * CT_STAGE_BYTES + 256 + 1024;

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

// tcgen05.ld of 16 columns into 8 register pairs.  No "memory" clobber: tensor memory is not addressable memory, and
// with the clobber the compiler would pin every global store of the surrounding code to its side of the load.
__device__ __forceinline__ void tmem_ld16_regs(uint32_t taddr, float2* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15}, [%16];"
      : "=f"(r[0].x), "=f"(r[0].y), "=f"(r[1].x), "=f"(r[1].y), "=f"(r[2].x), "=f"(r[2].y), "=f"(r[3].x),
        "=f"(r[3].y), "=f"(r[4].x), "=f"(r[4].y), "=f"(r[5].x), "=f"(r[5].y), "=f"(r[6].x), "=f"(r[6].y),
        "=f"(r[7].x), "=f"(r[7].y)
      : "r"(taddr));
}
// tcgen05.wait::ld that also names the 16 registers of an earlier tcgen05.ld as read-write operands: uses of those
// registers cannot be scheduled above the wait, other arithmetic can (the load of the next block stays in flight
// behind the convolution of the current one)
__device__ __forceinline__ void tmem_ld_wait_dep(float2* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(r[0].x), "+f"(r[0].y), "+f"(r[1].x), "+f"(r[1].y), "+f"(r[2].x), "+f"(r[2].y), "+f"(r[3].x),
                 "+f"(r[3].y), "+f"(r[4].x), "+f"(r[4].y), "+f"(r[5].x), "+f"(r[5].y), "+f"(r[6].x), "+f"(r[6].y),
                 "+f"(r[7].x), "+f"(r[7].y));
}

// Epilogue of one tile for one thread (= one output channel): 96 accumulator columns = 6 blocks of 16 frames -> 80
// outputs in 5 steps; step k convolves blocks k, k+1 (32-frame window).  Software pipeline over a four-block register
// ring with static indices (everything is unrolled: no register moves, immediate TMEM / global offsets):
//     tcgen05.ld of block k+3   |   activation of block k+2 (MUFU)   |   convolution of blocks k, k+1 (FMA pipe)
// The activation and the convolution of a step form one straight-line region, so their instruction streams are
// interleaved by the scheduler and the two pipes work at the same time inside ONE warp (three epilogue warps per
// scheduler cannot hide a serial MUFU phase behind each other).
//   even taps: packed FMAs on the aligned frame pairs; odd taps: scalar FMAs into the same accumulators (as fast as
//   the packed even/odd form + combining adds in tools/micro/conv_bench.cu, and 18 registers lighter)
// IS_QK: the tile holds the to_qk channels of CONV_VUQK (fp32 output); a template argument so that the whole tile
// epilogue is ONE basic block and the scheduler may run the stores of a step under the FMAs of the next one.
template <int MODE, bool MASKED, bool IS_QK, class Release>
__device__ __forceinline__ void convt_epilogue_tile(const LinearParams& P, const ConvTTile& ti, int c,
                                                    const float (&w)[17], float hb, const float* hrs_s, uint32_t tacc,
                                                    int tbase, Release release) {
  const EpiConv& cv = P.cv;
  float2 R[32];  // ring: block b lives in R[(b & 3) * 8 .. +8), pair i = frames (2i, 2i+1) of the block
  const float2 hb2 = make_float2(hb, hb);
  const unsigned Su = static_cast<unsigned>(P.S);
  auto ld_block = [&](int b) { tmem_ld16_regs(tacc + 16 * b, &R[(b & 3) * 8]); };
  auto wait_block = [&](int b) { tmem_ld_wait_dep(&R[(b & 3) * 8]); };
  // SiLU(scale * acc + bias) of block b, in place: h = scale/2 * acc + bias/2, a = h + h tanh(h)
  auto activate = [&](int b) {
    float2* r = &R[(b & 3) * 8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 hs = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
      if constexpr (MODE != CONV_UV) hs = __ldg(reinterpret_cast<const float4*>(hrs_s + 16 * b + 4 * i));
      const float2 h01 = fma2(r[2 * i], make_float2(hs.x, hs.y), hb2);
      const float2 h23 = fma2(r[2 * i + 1], make_float2(hs.z, hs.w), hb2);
#ifndef TDZ_ABL_NOACT
      float2 a01 = fma2(h01, make_float2(tanh_approx(h01.x), tanh_approx(h01.y)), h01);
      float2 a23 = fma2(h23, make_float2(tanh_approx(h23.x), tanh_approx(h23.y)), h23);
#else
      float2 a01 = h01, a23 = h23;
#endif
      if constexpr (MASKED) {
        // frames outside [0,S) are SELECTED to zero (= the convolution's padding; their accumulators may hold
        // anything, also NaN: rows of the N operand beyond S are not written by every producer)
        const unsigned t = static_cast<unsigned>(tbase + 16 * b + 4 * i);  // negative frames wrap to huge values
        a01.x = (t + 0u < Su) ? a01.x : 0.f;
        a01.y = (t + 1u < Su) ? a01.y : 0.f;
        a23.x = (t + 2u < Su) ? a23.x : 0.f;
        a23.y = (t + 3u < Su) ? a23.y : 0.f;
      }
      r[2 * i] = a01;
      r[2 * i + 1] = a23;
    }
  };
  ld_block(0);
  ld_block(1);
  wait_block(0);
  wait_block(1);
  ld_block(2);
  activate(0);
  activate(1);
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k + 2 <= 5) wait_block(k + 2);
    if (k + 3 <= 5) ld_block(k + 3);
    if (k == 3) release();
    const int tt0 = tbase + 8 + 16 * k;    // frame of output 0 of this step
    const size_t grow0 = static_cast<size_t>(ti.srow) + tt0;
    [[maybe_unused]] float rin[16];
    if constexpr (MODE == CONV_RESX) {  // residual stream of these 16 frames: in flight during the FMAs
      const float* src = cv.x_in + grow0 * 512 + c;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if constexpr (MASKED) rin[j] = (tt0 + j < P.S) ? src[static_cast<size_t>(j) * 512] : 0.f;
        else rin[j] = src[static_cast<size_t>(j) * 512];
      }
    }
    if (k + 2 <= 5) activate(k + 2);
    // out[j] = sum_t w[t] a[16 k + j + t], j = 0..15 (tap 8 carries the +1 of y + conv(y)); window pair i of the step
    // = frames (16 k + 2 i, + 1) = pair (i & 7) of block k + (i >> 3)
    auto Pw = [&](int i) -> const float2& { return R[((k + (i >> 3)) & 3) * 8 + (i & 7)]; };
    float2 acc[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m] = mul2(make_float2(w[0], w[0]), Pw(m));
#ifndef TDZ_ABL_NOCONV
#pragma unroll
    for (int qq = 0; qq < 9; ++qq) {
      if (qq > 0) {
#pragma unroll
        for (int m = 0; m < 8; ++m) acc[m] = fma2(make_float2(w[2 * qq], w[2 * qq]), Pw(m + qq), acc[m]);
      }
      if (qq < 8) {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          acc[m].x = fmaf(w[2 * qq + 1], Pw(m + qq).y, acc[m].x);      // output 2m   <- frame 2m + 2qq + 1
          acc[m].y = fmaf(w[2 * qq + 1], Pw(m + qq + 1).x, acc[m].y);  // output 2m+1 <- frame 2m + 2qq + 2
        }
      }
    }
#else
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m] = fma2(make_float2(w[16], w[16]), Pw(m + 8), acc[m]);
#endif
    auto out = [&](int j) { return (j & 1) ? acc[j >> 1].y : acc[j >> 1].x; };
    const int nrow = MASKED ? min(P.S - tt0, 16) : 16;  // rows j < nrow are inside the sample (may be <= 0)
    auto store_rows = [&](auto* dst, size_t ld, auto conv) {
#ifdef TDZ_ABL_NOSTORE
      if (w[3] != 12345.678f) return;
#endif
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if constexpr (MASKED) {
          if (j < nrow) dst[static_cast<size_t>(j) * ld] = conv(j);
        } else {
          dst[static_cast<size_t>(j) * ld] = conv(j);
        }
      }
    };
    if constexpr (MODE == CONV_VUQK) {
      if constexpr (!IS_QK) {
        store_rows(cv.vu + grow0 * 2048 + c, 2048, [&](int j) { return __float2bfloat16(out(j)); });
      } else {
        // to_qk channels: fp32, OffsetScale + rotary + bf16 split happen in qk_heads_kernel (the 32 rotary
        // channels all sit in one TMEM lane quarter, i.e. on one scheduler: doing that work here serialises it)
        store_rows(cv.qkf + grow0 * 128 + (c - 2048), 128, [&](int j) { return out(j); });
      }
    }
    if constexpr (MODE == CONV_RESX) {
      store_rows(cv.x_out + grow0 * 512 + c, 512, [&](int j) { return rin[j] + out(j); });
    }
    if constexpr (MODE == CONV_UV) {
      store_rows(cv.xuv + grow0 * 512 + c, 512, [&](int j) { return out(j); });
      if (c < 256) store_rows(cv.xubf + grow0 * 256 + c, 256, [&](int j) { return __float2bfloat16(out(j)); });
    }
  }
}

// CG2 = true: the CTA pair of a cluster owns two neighbouring channel tiles of the same time tile and executes one
// tcgen05.mma.cta_group::2 (M = 256 channels) per k-step; each CTA stages its own weight tile and only half of the
// X rows (32 KB per k-block instead of 48).  Used where the operand traffic matters (to_out, K = 1024: 2/3 of the
// tile's L2 -> SM bytes were operands) and the channel-tile count is even; the epilogue is the same.
constexpr int CT2_STAGES = 6;
constexpr int CT2_STAGE_BYTES = GEMM_STAGE_A_BYTES + 128 * 128;  // W tile 16 KB + half an X tile 16 KB
constexpr int CT2_SMEM_BYTES = CT2_STAGES * CT2_STAGE_BYTES + 256 + 1024;

template <int MODE, bool CG2>
__device__ __forceinline__ void gemm_convt_body(const LinearParams& P) {
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
  __shared__ uint32_t s_tmem_base;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches become uniform
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
  // Weight k-blocks of this CTA's first tile (one ring of them) are requested BEFORE the wait for the producer kernel:
  // under a programmatic dependent launch they stream in while that kernel is still running.
  int pre_kb = 0;
  if (warp == 0 && lane == 0 && first < ntiles) {
    ConvTTile ti;
    convt_tile(P, nct, first, ti);
    pre_kb = nkb < CT_STAGES ? nkb : CT_STAGES;
    [[maybe_unused]] const uint32_t leader_bars = CG2 ? mapa_shared(bar_base, 0) : 0u;
    for (int kb = 0; kb < pre_kb; ++kb) {
      const uint32_t sa = smem_base + kb * CT_STAGE_BYTES;
      if constexpr (CG2) {
        if (rank == 0) mbar_arrive_expect_tx(full_bar(kb), 2 * CT_STAGE_BYTES);
        tma_load_2d_cg2(sa, &P.tmB, leader_bars + 8u * kb, kb * 64, chan_tile(ti) * 128);
      } else {
        mbar_arrive_expect_tx(full_bar(kb), CT_STAGE_BYTES);
        tma_load_2d(sa, &P.tmB, full_bar(kb), kb * 64, ti.ct * 128);
      }
    }
  }
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      ConvTTile ti;
      convt_tile(P, nct, first, ti);
      [[maybe_unused]] const uint32_t leader_bars = CG2 ? mapa_shared(bar_base, 0) : 0u;
      for (int tile = first; tile < ntiles; tile += stride, convt_next(P, tstep, ti)) {
        if (tile + stride >= ntiles) pdl_trigger();  // last tile of this CTA: the next kernel may be launched
        for (int kb = 0; kb < nkb; ++kb) {
          const bool armed = tile == first && kb < pre_kb;  // barrier armed, weight tile requested before the wait
          if (!armed) mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * CT_STAGE_BYTES;
          const int trow = ti.t0 - (kb < P.shift_kblocks ? 1 : 0);  // token shift (mossformer_block.py:204-207)
          if constexpr (CG2) {
            // both CTAs' bytes are counted on the leader's barrier; each CTA: its weight tile + its half of X
            if (rank == 0 && !armed) mbar_arrive_expect_tx(full_bar(stage), 2 * CT_STAGE_BYTES);
            const uint32_t fb = leader_bars + 8u * stage;
            if (!armed) tma_load_2d_cg2(sa, &P.tmB, fb, kb * 64, chan_tile(ti) * 128);
            tma_load_3d_cg2(sa + GEMM_STAGE_A_BYTES, &P.tmAh, fb, kb * 64, trow + 128 * static_cast<int>(rank), ti.b);
          } else {
            if (!armed) {
              mbar_arrive_expect_tx(full_bar(stage), CT_STAGE_BYTES);
              tma_load_2d(sa, &P.tmB, full_bar(stage), kb * 64, ti.ct * 128);               // weights: M operand
            }
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
    const EpiGeneric& e = P.e;
    const EpiConv& cv = P.cv;
    // No shared scratch and no barrier among the epilogue warps: the three warps of a scheduler run the SAME
    // instruction stream, and a barrier per tile kept them in lock step - all in their FMA phase together (contending
    // for the pipe), all in their store / TMEM-wait phase together (pipe idle): tile time = sum of the phases, not the
    // maximum.  Free-running warps, started one third of a step apart, keep the FMA pipe fed from one warp while
    // another stores.  The per-channel constants (80 B) and per-frame scales come straight from global memory (L1 /
    // L2 hits: the tables are a few hundred KB and shared by every CTA).
    if (tq > 0) __nanosleep(400u * tq);
    int it = 0;
    ConvTTile ti, tn;
    [[maybe_unused]] const uint32_t leader_tempty = CG2 ? mapa_shared(tempty_bar(0), 0) : 0u;
    convt_tile(P, nct, first, ti);
    for (int tile = first; tile < ntiles; tile += stride, ++it, ti = tn) {
      if (tile + stride >= ntiles) pdl_trigger();
      tn = ti;
      convt_next(P, tstep, tn);
      const int c = chan_tile(ti) * 128 + q * 32 + lane;  // output channel of this thread
      float w[17];
      float hb;
      {
        const float4* cs = reinterpret_cast<const float4*>(cv.dw_t + static_cast<size_t>(c) * CT_CPC);
        const float4 c0 = __ldg(cs), c1 = __ldg(cs + 1), c2 = __ldg(cs + 2), c3 = __ldg(cs + 3), c4 = __ldg(cs + 4);
        w[0] = c0.x; w[1] = c0.y; w[2] = c0.z; w[3] = c0.w;
        w[4] = c1.x; w[5] = c1.y; w[6] = c1.z; w[7] = c1.w;
        w[8] = c2.x; w[9] = c2.y; w[10] = c2.z; w[11] = c2.w;
        w[12] = c3.x; w[13] = c3.y; w[14] = c3.z; w[15] = c3.w;
        w[16] = c4.x;
        hb = c4.y;
      }
      // scales of this warp's 96 frames (the table is padded: rows 8 before and 256 after the samples are readable)
      const float* hrs_s = (MODE != CONV_UV) ? e.ss_in + static_cast<ptrdiff_t>(ti.srow) + (ti.t0 + col0) : nullptr;
      const int tbase = ti.t0 + col0;  // frame of accumulator column col0
      const bool all_valid = tbase >= 0 && tbase + 96 <= P.S;

      const int as = it & 1;
      mbar_wait(tfull_bar(as), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * 256 + (static_cast<uint32_t>(q * 32) << 16) + col0;
      auto release = [&]() {  // the last TMEM read of this tile is done
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG2) mbar_arrive_cluster(leader_tempty + 8u * as);
          else mbar_arrive(tempty_bar(as));
        }
      };
      // interior tiles (warp-uniform) carry no masks and no store predicates
      const bool is_qk = MODE == CONV_VUQK && c >= 2048;  // warp-uniform: the to_qk channels are one whole tile
      if (is_qk) {
        if constexpr (MODE == CONV_VUQK) {
          if (all_valid) convt_epilogue_tile<MODE, false, true>(P, ti, c, w, hb, hrs_s, tacc, tbase, release);
          else convt_epilogue_tile<MODE, true, true>(P, ti, c, w, hb, hrs_s, tacc, tbase, release);
        }
      } else if (all_valid) {
        convt_epilogue_tile<MODE, false, false>(P, ti, c, w, hb, hrs_s, tacc, tbase, release);
      } else {
        convt_epilogue_tile<MODE, true, false>(P, ti, c, w, hb, hrs_s, tacc, tbase, release);
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

// 14 warps = 4 + 4 + 3 + 3 per scheduler, and a scheduler owns 16 K registers: 128 registers per thread is the limit
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
  pdl(gemm_convt_kernel<MODE>, grid, CT_THREADS, CT_SMEM_BYTES, st)(P);
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
  pdl(gemm_convt_cg2_kernel<MODE>, grid, CT_THREADS, smem, st)(P);
  return cudaGetLastError();
}

}  // namespace tdz
