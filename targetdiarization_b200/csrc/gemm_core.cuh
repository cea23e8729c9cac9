// Warp-specialised persistent tcgen05 GEMM core for sm_100a.
//
//   warp 0 (lane 0) : TMA producer   - fills a ring of {A,B} shared-memory stages (128 B swizzle)
//   warp 1 (lane 0) : MMA issuer     - tcgen05.mma cta_group::1, M=128, N=BLOCK_N, fp32 accumulators in TMEM
//   warps 2..5(9)   : epilogue       - tcgen05.ld of the finished accumulator, fused elementwise work, stores
//                                      (Cfg::EPI_SPLIT = 2: two warps per TMEM lane quarter, each half the columns)
//
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.  What is loaded
// for a k-block and what the epilogue does are supplied by a Cfg class, so every dense op of the
// separator is an instance of this one pipeline.
#pragma once
#include <atomic>

#include "ptx.cuh"

namespace tdz {

constexpr int GEMM_BLOCK_M = 128;
constexpr int gemm_threads(int epi_split) { return 64 + 128 * epi_split; }
constexpr int GEMM_STAGE_A_BYTES = GEMM_BLOCK_M * 128;  // 128 rows x 128 B (K-major) or 64 k-rows x 2 atoms (MN-major)

struct TileInfo {
  int m0;   // first row of the tile in the padded token space [B*Sp]
  int n0;   // first output column
  int b;    // sample index
  int t0;   // first frame of the tile inside the sample
  int nkb;  // k-blocks to accumulate
  int aux;  // Cfg specific (group index, split index ...)
};

// The opt-in dynamic shared-memory size is a per-DEVICE attribute of a kernel: set it once per device (bit mask), so
// that handles on several GPUs of one process - and concurrent first calls from several threads - both work.
inline cudaError_t set_max_smem_once(const void* kernel, int smem_bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

// Shared memory of one CTA: operand stages | barriers (256 B) | optional fp32 epilogue panel (Cfg::PANEL_BYTES).
template <int BLOCK_N, int STAGES, int PANEL_BYTES = 0>
constexpr int gemm_smem_bytes() {
  return STAGES * (GEMM_STAGE_A_BYTES + BLOCK_N * 128) + 1024 /*align slack*/ + 256 /*barriers*/ + PANEL_BYTES;
}

// What the epilogue warps get besides the accumulator address: their index among the epilogue threads and the
// shared-memory panel (nullptr when the Cfg asks for none).
struct EpiCtx {
  float* panel;
  int tid;  // 0 .. 128*EPI_SPLIT-1
};
// Barrier among the epilogue warps only (named barrier 1; producer / MMA warps never join it).
template <int NTHREADS>
__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
}

template <class Cfg>
__global__ void __launch_bounds__(gemm_threads(Cfg::EPI_SPLIT), 1) gemm_kernel(const __grid_constant__ typename Cfg::Params P) {
  constexpr int BLOCK_N = Cfg::BLOCK_N;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int STAGE_B_BYTES = BLOCK_N * 128;
  constexpr int STAGE_BYTES = GEMM_STAGE_A_BYTES + STAGE_B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                 : (2 * BLOCK_N <= 256) ? 256 : 512;
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N constraint for M=128");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  __shared__ uint32_t s_tmem_base;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches become uniform
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    Cfg::prefetch(P);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4 * Cfg::EPI_SPLIT);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&s_tmem_base), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const int ntiles = Cfg::num_tiles(P);
  // Weight k-blocks of this CTA's first tile (as many as the ring holds) are requested BEFORE the wait for the
  // producer kernel: with a programmatic dependent launch they stream in while that kernel is still running.
  int pre_kb = 0;
  if constexpr (Cfg::PREFETCH_B) {
    if (warp == 0 && lane == 0 && static_cast<int>(blockIdx.x) < ntiles) {
      TileInfo ti;
      Cfg::tile_info(P, blockIdx.x, ti);
      pre_kb = ti.nkb < STAGES ? ti.nkb : STAGES;
      for (int kb = 0; kb < pre_kb; ++kb) {
        mbar_arrive_expect_tx(full_bar(kb), STAGE_BYTES);   // (the A half arrives after the wait)
        Cfg::load_b(P, ti, kb, smem_base + kb * STAGE_BYTES + GEMM_STAGE_A_BYTES, full_bar(kb));
      }
    }
  }
  pdl_wait();  // barriers, TMEM and descriptors are ready; from here on the producer kernel's data is visible

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // last tile of this CTA: the next kernel of the stream may be launched (its CTAs take the SMs that go idle in
        // this grid's tail and run their prologue there).  Not earlier: a dependent grid of small CTAs that sits next
        // to this one for its whole duration costs the epilogue warps issue slots (measured: +4 % per forward).
        if (tile + static_cast<int>(gridDim.x) >= ntiles) pdl_trigger();
        TileInfo ti;
        Cfg::tile_info(P, tile, ti);
        for (int kb = 0; kb < ti.nkb; ++kb) {
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          bool armed = false;  // first ring of the first tile: barrier armed and weights requested before the wait
          if constexpr (Cfg::PREFETCH_B) armed = tile == static_cast<int>(blockIdx.x) && kb < pre_kb;
          if (armed) {
            if constexpr (Cfg::PREFETCH_B) Cfg::load_a(P, ti, kb, sa, full_bar(stage));
          } else {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
            Cfg::load(P, ti, kb, sa, sa + GEMM_STAGE_A_BYTES, full_bar(stage));
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t IDESC = umma_idesc(Cfg::FMT, GEMM_BLOCK_M, BLOCK_N, Cfg::A_MN, Cfg::B_MN);
      // K-major: 32 B per UMMA_K step inside the 128 B swizzle row; 8-row groups 1024 B apart.
      // MN-major: 16 k-rows (2 KB) per step; next 64-wide MN atom one 64-row box (8 KB) further.
      constexpr uint32_t A_STEP = Cfg::A_MN ? 2048u : 32u;
      constexpr uint32_t B_STEP = Cfg::B_MN ? 2048u : 32u;
      constexpr uint32_t A_LBO = Cfg::A_MN ? 8192u : 16u;
      constexpr uint32_t B_LBO = Cfg::B_MN ? 8192u : 16u;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        TileInfo ti;
        Cfg::tile_info(P, tile, ti);
        const int as = it & 1;
        mbar_wait(tempty_bar(as), ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < ti.nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + GEMM_STAGE_A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * A_STEP, A_LBO, 1024u);
            const uint64_t db = umma_smem_desc(sb + k * B_STEP, B_LBO, 1024u);
            if (Cfg::FMT == 2)
              umma_tf32(tacc, da, db, IDESC, (kb | k) != 0);
            else
              umma_f16(tacc, da, db, IDESC, (kb | k) != 0);
          }
          umma_commit(empty_bar(stage));
          if (kb == ti.nkb - 1) umma_commit(tfull_bar(as));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // epilogue warps: TMEM lane quarter is fixed by warp id % 4
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int half = (warp - 2) >> 2;  // which column half this warp owns (always 0 when EPI_SPLIT == 1)
    EpiCtx ectx;
    ectx.tid = threadIdx.x - 64;
    ectx.panel = Cfg::PANEL_BYTES > 0
                     ? reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + STAGES * STAGE_BYTES + 256)
                     : nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      if (tile + static_cast<int>(gridDim.x) >= ntiles) pdl_trigger();
      TileInfo ti;
      Cfg::tile_info(P, tile, ti);
      const int as = it & 1;
      mbar_wait(tfull_bar(as), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
      Cfg::epilogue(P, ti, tacc, row, half, ectx);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// cta_group::2 variant: the CTA pair executes ONE 256 x 256 MMA per k-step.  Each CTA stages its own 128 rows of A
// and only its own 128 of the 256 B columns (32 KB per k-block instead of 48), and its tensor core receives the
// other half of B from the partner, so the shared-memory port of an SM carries 2/3 of the bytes per MMA that the
// cta_group::1 form needs (operand reads + TMA fills exceed 128 B/clk there, which is what held the 128 x 256
// single-CTA mainloop at ~2/3 of the MMA rate).  Only the leader (cluster rank 0) issues MMAs:
//   full[s]    (leader)  <- the leader's expect_tx of both CTAs' bytes + the TMA completions of both CTAs
//   empty[s]   (each)    <- tcgen05.commit multicast by the leader
//   tfull[a]   (each)    <- tcgen05.commit multicast by the leader after the last k-block of a tile
//   tempty[a]  (leader)  <- every epilogue warp of both CTAs (remote mbarrier arrive)
template <class Cfg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(gemm_threads(Cfg::CG2_EPI_SPLIT), 1)
    gemm_cg2_kernel(const __grid_constant__ typename Cfg::Params P) {
  constexpr int BLOCK_N = Cfg::BLOCK_N;
  constexpr int STAGES = Cfg::CG2_STAGES;
  constexpr int STAGE_B_BYTES = (BLOCK_N / 2) * 128;
  constexpr int STAGE_BYTES = GEMM_STAGE_A_BYTES + STAGE_B_BYTES;
  constexpr uint32_t TMEM_COLS = 512;
  static_assert(BLOCK_N == 256, "cta_group::2 kernel is written for 256-column tiles");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  __shared__ uint32_t s_tmem_base;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches become uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    Cfg::prefetch(P);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * 4 * Cfg::CG2_EPI_SPLIT);  // epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_cg2(smem_u32(&s_tmem_base), TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  pdl_wait();

  const int nwork = Cfg::num_pair_tiles(P);
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t leader_bars = mapa_shared(bar_base, 0);  // full[s] of the leader = leader_bars + 8 s
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < nwork; w += nclusters) {
        if (w + nclusters >= nwork) pdl_trigger();  // last work item of this CTA pair (see gemm_kernel)
        TileInfo ti;
        Cfg::pair_tile_info(P, w, rank, ti);
        for (int kb = 0; kb < ti.nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          Cfg::load_cg2(P, ti, kb, sa, sa + GEMM_STAGE_A_BYTES, leader_bars + 8u * stage, rank);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t IDESC = umma_idesc(Cfg::FMT, 2 * GEMM_BLOCK_M, BLOCK_N, Cfg::A_MN, Cfg::B_MN);
      constexpr uint32_t IDESC_F16 = umma_idesc(0, 2 * GEMM_BLOCK_M, BLOCK_N, Cfg::A_MN, Cfg::B_MN);
      constexpr uint32_t A_STEP = Cfg::A_MN ? 2048u : 32u;
      constexpr uint32_t B_STEP = Cfg::B_MN ? 2048u : 32u;
      constexpr uint32_t A_LBO = Cfg::A_MN ? 8192u : 16u;
      constexpr uint32_t B_LBO = Cfg::B_MN ? 8192u : 16u;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = cluster_id; w < nwork; w += nclusters, ++it) {
        TileInfo ti;
        Cfg::pair_tile_info(P, w, 0, ti);
        const int as = it & 1;
        mbar_wait(tempty_bar(as), ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < ti.nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + GEMM_STAGE_A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * A_STEP, A_LBO, 1024u);
            const uint64_t db = umma_smem_desc(sb + k * B_STEP, B_LBO, 1024u);
            umma_f16_cg2(tacc, da, db, kb >= Cfg::F16_FROM_KB ? IDESC_F16 : IDESC, (kb | k) != 0);
          }
          umma_commit_cg2_mc(empty_bar(stage), 0x3);
          if (kb == ti.nkb - 1) umma_commit_cg2_mc(tfull_bar(as), 0x3);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int part = (warp - 2) >> 2;
    const uint32_t leader_tempty = mapa_shared(tempty_bar(0), 0);
    // The epilogue's global operands (v, u of the thread's row) are requested one tile ahead: each half of the
    // register set is re-loaded for the next tile right after its last use in the current one.
    typename Cfg::EpiPre4 pre;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    TileInfo ti, nx;
    int w = cluster_id;
    if (w < nwork) {
      Cfg::pair_tile_info(P, w, rank, ti);
      Cfg::epi_prefetch4(P, ti, row, part, 0, pre);
      Cfg::epi_prefetch4(P, ti, row, part, 1, pre);
    }
    for (int it = 0; w < nwork; w += nclusters, ++it) {
      const bool has_next = w + nclusters < nwork;
      if (!has_next) pdl_trigger();
      if (has_next) Cfg::pair_tile_info(P, w + nclusters, rank, nx);
      const int as = it & 1;
      mbar_wait(tfull_bar(as), (it >> 1) & 1);
      tc_fence_after();
      Cfg::epilogue4(P, ti, nx, has_next, lane_taddr + as * BLOCK_N, row, part, pre);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tempty + 8u * as);
      ti = nx;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_cg2(tmem_base, TMEM_COLS);
  }
}

template <class Cfg>
cudaError_t launch_gemm_cg2(const typename Cfg::Params& P, int nwork, int num_sms, cudaStream_t st) {
  constexpr int smem = Cfg::CG2_STAGES * (GEMM_STAGE_A_BYTES + (Cfg::BLOCK_N / 2) * 128) + 1024 + 256;
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(gemm_cg2_kernel<Cfg>), smem, configured); e != cudaSuccess)
    return e;
  if (nwork <= 0) return cudaSuccess;
  int grid = 2 * nwork < num_sms ? 2 * nwork : (num_sms & ~1);
  pdl(gemm_cg2_kernel<Cfg>, grid, gemm_threads(Cfg::CG2_EPI_SPLIT), smem, st)(P);
  return cudaGetLastError();
}

template <class Cfg>
cudaError_t launch_gemm(const typename Cfg::Params& P, int ntiles, int num_sms, cudaStream_t st) {
  constexpr int smem = gemm_smem_bytes<Cfg::BLOCK_N, Cfg::STAGES, Cfg::PANEL_BYTES>();
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(gemm_kernel<Cfg>), smem, configured); e != cudaSuccess)
    return e;
  if (ntiles <= 0) return cudaSuccess;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  pdl(gemm_kernel<Cfg>, grid, gemm_threads(Cfg::EPI_SPLIT), smem, st)(P);
  return cudaGetLastError();
}

}  // namespace tdz
