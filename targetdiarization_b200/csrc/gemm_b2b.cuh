// Back-to-back GEMM: out = W2 act(W1 x + b1) (+ b2) (+ resid) for K1 = 256 inputs, H hidden units (a multiple of 128)
// and 256 outputs, with the [rows][H] hidden activations kept ON CHIP - TMEM accumulator -> epilogue warps -> bf16 in
// shared memory as the A operand of the second MMA - instead of a round trip through HBM between two kernels:
//   * UniDeepFsmn linear -> ReLU -> project (look2hear/models/fsmn.py:131-139), H = 256;
//   * the Apollo restorer's ConvActNorm1d 1x1 pair (look2hear/models/apollo.py:150-158: Conv1d(256, 1024) -> SiLU ->
//     Conv1d(1024, 256) + residual), H = 1024: 4 KB per token of hidden traffic removed per pair.
//
//   warp 0 (lane 0)  TMA producer: the X tile (128 rows x 256, resident for the whole tile) and a ring of 32 KB weight
//                    stages in the order the MMA warp consumes them
//   warp 1 (lane 0)  tcgen05.mma issuer, software pipelined over hidden chunks of 128:
//                    G1(j): acc1[j & 1] = X W1[j]^T (N = 128, K = 256);   G2(j): acc2 += h[j & 1] W2[:, j]^T (N = 256,
//                    K = 128), issued as G1(0) G1(1) G2(0) G1(2) G2(1) ... so that the tensor core works on G1(j + 1)
//                    while the epilogue warps turn acc1[j] into h[j]
//   warps 2..9       epilogue: (1) per chunk: tcgen05.ld of acc1 -> + b1 -> activation -> bf16 -> shared memory in the
//                    128 B-swizzled K-major layout of an A operand (fence.proxy.async + mbarrier, like the gather warps
//                    of gemm_conv3.cuh); (2) per tile: acc2 -> + b2 + resid -> global
// TMEM: acc2 = columns [0, 256), acc1 double buffer = [256, 384) and [384, 512).
// Shared memory: X 64 KB | h HBUF x 32 KB | weight ring 3 x 32 KB | b1 (H floats).  HBUF = 2 where it fits (H = 256:
// measured 0.286 ms against 0.320 ms single buffered at the benchmark shape), 1 for H = 1024 (the 4 KB of b1 would not
// fit next to two buffers): the epilogue warps then wait for G2(j) - 8 MMAs - between their arithmetic on chunk j + 1
// and the stores of its result.
// (ncu on the first version, Apollo shape: the epilogue warps were the serial resource - 43 % of the stall samples on
// the first use of b1 values loaded from global memory per chunk, and two MUFU operations per SiLU (ex2 + rcp) = as
// many XU cycles per chunk as the chunk's MMAs take.  Hence b1 in shared memory and the one-MUFU form
// silu(x) = h + h tanh(h), h = x / 2, whose 2^-11 error is below the bf16 rounding of the stored activation.)
// The arithmetic is the same as in the two-kernel form (same k order, same bf16 rounding of the hidden activations), so
// the results are bit-identical to it.
//
// CG2 = true is the cta_group::2 form (opt-in with TDZ_B2B_CG2, correct - same bits - but measured 6 % SLOWER on the
// Apollo shape and equal on the FSMN shape, so the premise below is not what bounds the kernel; kept as the tested
// starting point for a 256-row form): a CTA pair owns 256 rows and the leader issues ONE M = 256 MMA per k-step; each
// CTA stages its own X / h rows but only HALF of every weight tile (G1: 64 of the chunk's 128 W1 rows, G2: 128 of the
// 256 W2 rows), so a weight stage is 16 KB instead of 32 KB and the same shared memory holds a ring twice as deep in
// MMA work: with 128-row tiles the kernel re-streams all of W1 and W2 for every tile and its rate is (bytes in flight)
// / (TMA latency), not a pipe.  Barrier protocol as in gemm_cg2_kernel: operand-full barriers live on the leader and
// count the bytes of both CTAs; releases (w_empty, x_empty, acc1_full, h_empty, acc2_full) are tcgen05.commit multicast
// to both; the epilogue warps of both CTAs arrive on the leader's acc1_empty / h_full / acc2_empty (h_full with release
// semantics after fence.proxy.async: the leader's MMA reads the peer's h through the async proxy).
#pragma once
#include "gemm_cfgs.cuh"

namespace tdz {

struct B2bParams {
  CUtensorMap tmX;    // 3-D {256, Sp, B} bf16, box {64, 128, 1}
  CUtensorMap tmW1;   // 2-D {256, H} bf16, box {64, 128}   (CG2: box {64, 64})
  CUtensorMap tmW2;   // 2-D {H, 256} bf16, box {64, 256}   (CG2: box {64, 128})
  int B, Sp, S, H;
  const float* bias1;  // [H]
  EpiGeneric e;        // bias (b2), resid / resid_ld, out_f32 / out_ld, out_bf16 / out_bf_ld, ss_out / ss_out_ld
};

constexpr int B2B_THREADS = 64 + 256;
constexpr int B2B_RING_BYTES = 3 * 32768;   // 3 stages of 32 KB, or 6 of 16 KB in the cta_group::2 form
constexpr int B2B_X_BYTES = 4 * 16384;
constexpr int b2b_smem_bytes(int hbuf, int max_h, bool has_b2) {
  return B2B_X_BYTES + hbuf * 32768 + B2B_RING_BYTES + 1024 + 256 + max_h * 4 + (has_b2 ? 256 * 4 : 0);
}

template <int ACT1, unsigned EF2, int HBUF, bool CG2>
__global__ void __launch_bounds__(B2B_THREADS, 1) gemm_b2b_kernel(const __grid_constant__ B2bParams P) {
  constexpr int B2B_H_BYTES = HBUF * 32768;
  constexpr int STAGE_BYTES = CG2 ? 16384 : 32768;
  constexpr int B2B_WSTAGES = B2B_RING_BYTES / STAGE_BYTES;
  constexpr int NCTA = CG2 ? 2 : 1;
  constexpr int W1_KK = STAGE_BYTES / 2;        // bytes of one W1 k-block inside a G1 stage (two k-blocks per stage)
  const uint32_t rank = CG2 ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sX = smem_base;
  const uint32_t sH = sX + B2B_X_BYTES;
  const uint32_t sW = sH + B2B_H_BYTES;
  const uint32_t bar_base = sW + B2B_RING_BYTES;
  auto w_full = [&](int s) { return bar_base + 8u * s; };
  auto w_empty = [&](int s) { return bar_base + 8u * (B2B_WSTAGES + s); };
  const uint32_t b0 = bar_base + 8u * 2 * B2B_WSTAGES;
  const uint32_t x_full = b0, x_empty = b0 + 8, acc2_full = b0 + 16, acc2_empty = b0 + 24;
  auto acc1_full = [&](int b) { return b0 + 32u + 8u * b; };
  auto acc1_empty = [&](int b) { return b0 + 48u + 8u * b; };
  auto h_full = [&](int b) { return b0 + 64u + 8u * b; };
  auto h_empty = [&](int b) { return b0 + 80u + 8u * b; };
  float* sbias = reinterpret_cast<float*>(smem_al + B2B_X_BYTES + B2B_H_BYTES + B2B_RING_BYTES + 256);  // [H]
  [[maybe_unused]] float* sbias2 = sbias + P.H;   // [256] b2 (exists only in instances with EF_BIAS)
  // barrier `bar` of the leader CTA as a shared::cluster address (the local barrier itself without a cluster)
  auto leader = [&](uint32_t bar) { return CG2 ? mapa_shared(bar, 0) : bar; };
  __shared__ uint32_t s_tmem_base;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches become uniform
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmX);
    tma_prefetch_desc(&P.tmW1);
    tma_prefetch_desc(&P.tmW2);
    for (int s = 0; s < B2B_WSTAGES; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 8 * NCTA);
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc1_full(b), 1);
      mbar_init(acc1_empty(b), 8 * NCTA);
    }
    for (int b = 0; b < HBUF; ++b) {
      mbar_init(h_full(b), 8 * NCTA);
      mbar_init(h_empty(b), 1);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < P.H; i += B2B_THREADS) sbias[i] = P.bias1[i];
  if constexpr ((EF2 & EF_BIAS) != 0) {
    if (threadIdx.x < 256) sbias2[threadIdx.x] = P.e.bias[threadIdx.x];
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
  if constexpr (CG2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  pdl_wait();  // (the biases staged above are weights, not another kernel's output)
  const int ntiles = P.B * P.Sp / GEMM_BLOCK_M;
  const int NC = P.H / 128;
  // work items: row tiles, or pairs of row tiles (tile = 2 w + rank; a pair's second tile may lie past the end: its
  // X rows are zero-filled by TMA and its stores are masked)
  const int nwork = (ntiles + NCTA - 1) / NCTA;
  const int w_first = static_cast<int>(blockIdx.x) / NCTA, w_step = static_cast<int>(gridDim.x) / NCTA;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto next_stage = [&]() {
        if (++stage == B2B_WSTAGES) {
          stage = 0;
          phase ^= 1u;
        }
      };
      int it = 0;
      for (int w = w_first; w < nwork; w += w_step, ++it) {
        if (w + w_step >= nwork) pdl_trigger();  // last row tile of this CTA: the next kernel may be launched
        const int m0 = (w * NCTA + static_cast<int>(rank)) * GEMM_BLOCK_M;
        const int b = m0 / P.Sp, t0 = m0 - b * P.Sp;
        mbar_wait(x_empty, (it & 1) ^ 1u);
        if (rank == 0) mbar_arrive_expect_tx(x_full, NCTA * B2B_X_BYTES);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if constexpr (CG2) tma_load_3d_cg2(sX + kb * 16384, &P.tmX, leader(x_full), kb * 64, t0, b);
          else tma_load_3d(sX + kb * 16384, &P.tmX, x_full, kb * 64, t0, b);
        }
        for (int step = 0; step <= NC; ++step) {
          if (step < NC) {  // G1(step): two stages of two W1 k-blocks (this CTA's share of the chunk's 128 rows)
            for (int s = 0; s < 2; ++s) {
              mbar_wait(w_empty(stage), phase ^ 1u);
              if (rank == 0) mbar_arrive_expect_tx(w_full(stage), NCTA * STAGE_BYTES);
              const uint32_t dst = sW + stage * STAGE_BYTES;
              if constexpr (CG2) {
                const int r0 = step * 128 + static_cast<int>(rank) * 64;
                tma_load_2d_cg2(dst, &P.tmW1, leader(w_full(stage)), (2 * s) * 64, r0);
                tma_load_2d_cg2(dst + W1_KK, &P.tmW1, leader(w_full(stage)), (2 * s + 1) * 64, r0);
              } else {
                tma_load_2d(dst, &P.tmW1, w_full(stage), (2 * s) * 64, step * 128);
                tma_load_2d(dst + W1_KK, &P.tmW1, w_full(stage), (2 * s + 1) * 64, step * 128);
              }
              next_stage();
            }
          }
          if (step >= 1) {  // G2(step - 1): two stages of one W2 k-block (this CTA's share of the 256 rows)
            for (int s = 0; s < 2; ++s) {
              mbar_wait(w_empty(stage), phase ^ 1u);
              if (rank == 0) mbar_arrive_expect_tx(w_full(stage), NCTA * STAGE_BYTES);
              const uint32_t dst = sW + stage * STAGE_BYTES;
              if constexpr (CG2)
                tma_load_2d_cg2(dst, &P.tmW2, leader(w_full(stage)), (step - 1) * 128 + s * 64, static_cast<int>(rank) * 128);
              else
                tma_load_2d(dst, &P.tmW2, w_full(stage), (step - 1) * 128 + s * 64, 0);
              next_stage();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t IDESC1 = umma_idesc(1, NCTA * GEMM_BLOCK_M, 128, 0, 0);
      constexpr uint32_t IDESC2 = umma_idesc(1, NCTA * GEMM_BLOCK_M, 256, 0, 0);
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
        if constexpr (CG2) umma_f16_cg2(d, da, db, idesc, acc);
        else umma_f16(d, da, db, idesc, acc);
      };
      auto commit = [&](uint32_t bar) {   // arrives on `bar` (of both CTAs) once the MMAs issued so far have completed
        if constexpr (CG2) umma_commit_cg2_mc(bar, 0x3);
        else umma_commit(bar);
      };
      int stage = 0;
      uint32_t phase = 0;
      auto next_stage = [&]() {
        if (++stage == B2B_WSTAGES) {
          stage = 0;
          phase ^= 1u;
        }
      };
      int it = 0;
      int c1 = 0, c2 = 0;  // running chunk counters of G1 / G2 (buffer = counter & 1, use = counter >> 1)
      for (int w = w_first; w < nwork; w += w_step, ++it) {
        mbar_wait(x_full, it & 1);
        tc_fence_after();
        for (int step = 0; step <= NC; ++step) {
          if (step < NC) {
            const int buf = c1 & 1;
            mbar_wait(acc1_empty(buf), ((c1 >> 1) & 1) ^ 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + 256 + buf * 128;
            for (int s = 0; s < 2; ++s) {
              mbar_wait(w_full(stage), phase);
              tc_fence_after();
              const uint32_t sb = sW + stage * STAGE_BYTES;
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t da = umma_smem_desc(sX + (2 * s + kk) * 16384 + k * 32u, 16u, 1024u);
                  const uint64_t db = umma_smem_desc(sb + kk * W1_KK + k * 32u, 16u, 1024u);
                  mma(tacc, da, db, IDESC1, (s | kk | k) != 0);
                }
              }
              commit(w_empty(stage));
              next_stage();
            }
            commit(acc1_full(buf));
            if (step == NC - 1) commit(x_empty);  // the X tile may be overwritten by the next one
            ++c1;
          }
          if (step >= 1) {
            const int j = step - 1;
            const int hb = c2 % HBUF;
            if constexpr (CG2) mbar_wait_cluster(h_full(hb), (c2 / HBUF) & 1);   // the peer's h rows are published too
            else mbar_wait(h_full(hb), (c2 / HBUF) & 1);
            if (j == 0) mbar_wait(acc2_empty, (it & 1) ^ 1u);
            tc_fence_after();
            for (int s = 0; s < 2; ++s) {
              mbar_wait(w_full(stage), phase);
              tc_fence_after();
              const uint32_t sa = sH + hb * 32768 + s * 16384;
              const uint32_t sb = sW + stage * STAGE_BYTES;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_smem_desc(sa + k * 32u, 16u, 1024u);
                const uint64_t db = umma_smem_desc(sb + k * 32u, 16u, 1024u);
                mma(tmem_base, da, db, IDESC2, (j | s | k) != 0);
              }
              commit(w_empty(stage));
              next_stage();
            }
            commit(h_empty(hb));
            if (j == NC - 1) commit(acc2_full);
            ++c2;
          }
        }
      }
    }
  } else {
    const int ew = warp - 2;          // 0..7
    const int q = warp & 3;           // TMEM lane quarter this warp may read
    const int half = ew >> 2;         // which half of the columns
    const int row = q * 32 + lane;
    const uint32_t lane_t = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* hbase = smem_al + B2B_X_BYTES;
    const uint32_t rsw = (row >> 3) * 1024 + (row & 7) * 128;   // row offset inside a k-block atom set
    const EpiGeneric& e = P.e;
    int it = 0, c = 0;
    // arrivals for the MMA issuer go to the leader CTA's barriers (cluster scope in the pair form)
    auto arrive = [&](uint32_t bar, bool release) {
      if constexpr (CG2) {
        if (release) mbar_arrive_cluster_release(mapa_shared(bar, 0));
        else mbar_arrive_cluster(mapa_shared(bar, 0));
      } else {
        mbar_arrive(bar);
      }
    };
    for (int w = w_first; w < nwork; w += w_step, ++it) {
      if (w + w_step >= nwork) pdl_trigger();
      const int m0 = (w * NCTA + static_cast<int>(rank)) * GEMM_BLOCK_M;
      const int b = m0 / P.Sp, t0 = m0 - b * P.Sp;
      const bool in_range = m0 < P.B * P.Sp;   // false only for the second tile of an odd last pair
      // ---- (1) hidden chunks: acc1 -> h
      for (int j = 0; j < NC; ++j, ++c) {
        const int buf = c & 1;
        mbar_wait(acc1_full(buf), (c >> 1) & 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + 256 + buf * 128 + lane_t + half * 64;
        const float* bs = sbias + j * 128 + half * 64;
        uint4 packed[8];
#pragma unroll
        for (int cc = 0; cc < 64; cc += 32) {
          float v[32];
          tmem_ld16(tacc + cc, v);
          tmem_ld16(tacc + cc + 16, v + 16);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bs + cc + i);   // same address in every lane: broadcast
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float x = v[i + k] + bb[k];
              if constexpr (ACT1 == ACT_SILU) {
                const float hx = 0.5f * x;
                x = fmaf(hx, tanh_approx(hx), hx);
              } else {
                x = act_apply<ACT1>(x);
              }
              v[i + k] = x;
            }
          }
#pragma unroll
          for (int p = 0; p < 4; ++p)
            packed[cc / 8 + p] = make_uint4(pack_bf16(v[8 * p], v[8 * p + 1]), pack_bf16(v[8 * p + 2], v[8 * p + 3]),
                                            pack_bf16(v[8 * p + 4], v[8 * p + 5]), pack_bf16(v[8 * p + 6], v[8 * p + 7]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(acc1_empty(buf), false);  // acc1[buf] may be overwritten by G1(j + 2)
        const int hbuf = c % HBUF;
        mbar_wait(h_empty(hbuf), ((c / HBUF) & 1) ^ 1u);  // G2(j - HBUF) has finished reading this h buffer
        uint8_t* hb = hbase + hbuf * 32768 + half * 16384 + rsw;   // this warp's 64 columns = k-block atom `half`
#pragma unroll
        for (int piece = 0; piece < 8; ++piece)          // 16 B pieces (8 columns) of the 128 B row, swizzled
          *reinterpret_cast<uint4*>(hb + ((piece ^ (row & 7)) << 4)) = packed[piece];
        fence_proxy_async();   // generic-proxy writes of h -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) arrive(h_full(hbuf), true);
      }
      // ---- (2) output tile: acc2 -> + b2 + resid -> global; this warp owns columns [128 half, 128 half + 128).
      // The residual values of a 16-column step are requested one step ahead (the first one before the wait for the
      // accumulator): ncu put 25 % of the kernel's stall samples on their first use when they were loaded in place.
      const bool valid = in_range && b < P.B && t0 + row < P.S;
      const size_t grow = static_cast<size_t>(m0) + row;
      float ssq0 = 0.f, ssq1 = 0.f;
      float rs_next[16];
      if constexpr ((EF2 & EF_RESID) != 0) {
        if (valid) ld_f32x16_rw(e.resid + grow * e.resid_ld + half * 128, rs_next, true);
      }
      mbar_wait(acc2_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 128; cc += 16) {
        const int col0 = half * 128 + cc;
        float v[16], rs[16];
        tmem_ld16(tmem_base + lane_t + col0, v);
        if constexpr ((EF2 & EF_RESID) != 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) rs[i] = rs_next[i];
          if (valid && cc + 16 < 128) ld_f32x16_rw(e.resid + grow * e.resid_ld + col0 + 16, rs_next, true);
        }
        tmem_ld_wait();
        if (!valid) continue;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr ((EF2 & EF_BIAS) != 0) b4 = *reinterpret_cast<const float4*>(sbias2 + col0 + i);  // broadcast
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float x = v[i + k] + bb[k];
            if constexpr ((EF2 & EF_RESID) != 0) x += rs[i + k];
            s = fmaf(x, x, s);
            v[i + k] = x;
          }
        }
        if (cc < 64) ssq0 += s;
        else ssq1 += s;
        if constexpr ((EF2 & EF_OUT_F32) != 0) {
          float4* o = reinterpret_cast<float4*>(e.out_f32 + grow * e.out_ld + col0);
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
        if constexpr ((EF2 & EF_OUT_BF16) != 0) {
          uint4* o = reinterpret_cast<uint4*>(e.out_bf16 + grow * e.out_bf_ld + col0);
          o[0] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          o[1] = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]),
                            pack_bf16(v[14], v[15]));
        }
      }
      if ((EF2 & EF_SS_OUT) != 0 && in_range) {   // one partial per 64 output columns, like LinearPanel
        e.ss_out[grow * e.ss_out_ld + half * 2] = valid ? ssq0 : 0.f;
        e.ss_out[grow * e.ss_out_ld + half * 2 + 1] = valid ? ssq1 : 0.f;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive(acc2_empty, false);
    }
  }

  tc_fence_before();
  if constexpr (CG2) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    if constexpr (CG2) tmem_dealloc_cg2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// MAX_H: the largest hidden width the instance is launched with (sizes the b1 copy): 256 with HBUF = 2, 1024 with 1.
// CG2: the cta_group::2 form (tensor maps with the half-size weight boxes), launched as clusters of two CTAs.
template <int ACT1, unsigned EF2, int HBUF, int MAX_H, bool CG2>
cudaError_t launch_gemm_b2b(const B2bParams& P, int num_sms, cudaStream_t st) {
  constexpr int smem = b2b_smem_bytes(HBUF, MAX_H, (EF2 & EF_BIAS) != 0);
  static_assert(smem <= 232448, "back-to-back GEMM: shared memory budget");
  if (P.H > MAX_H || P.H % 128 != 0) return cudaErrorInvalidValue;
  auto kernel = gemm_b2b_kernel<ACT1, EF2, HBUF, CG2>;
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t err = set_max_smem_once(reinterpret_cast<const void*>(kernel), smem, configured); err != cudaSuccess)
    return err;
  const int ntiles = P.B * P.Sp / GEMM_BLOCK_M;
  if (ntiles <= 0) return cudaSuccess;
  if constexpr (!CG2) {
    pdl(kernel, ntiles < num_sms ? ntiles : num_sms, B2B_THREADS, smem, st)(P);
    return cudaGetLastError();
  } else {
    const int npairs = (ntiles + 1) / 2;
    const int grid = 2 * npairs < num_sms ? 2 * npairs : (num_sms & ~1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(B2B_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, P);
  }
}

}  // namespace tdz
