// 3x3 convolution (pad 1, stride 1) over NHWC pixel-major bf16 maps as an IMPLICIT GEMM: the im2col matrix is never
// written.  Same persistent tcgen05 pipeline as gemm_kernel, but the A operand of a k-block (128 pixels x 64
// im2col columns) is assembled in shared memory by four gather warps instead of a TMA box:
//
//   warp 0            TMA producer of the weight k-blocks (B operand, [N][9*C] bf16, K-major)
//   warp 1            tcgen05.mma issuer
//   warps 2..         epilogue: folded-BN bias + Hardtanh(0,20) -> bf16; optionally also out + x_next, the input of
//                     the next 3x3 convolution of the Res2Net chain (so that every conv gathers ONE map)
//   last 8 warps      gather: 16 B pieces (8 channels of one tap of one pixel) are copied from the input map with
//                     cp.async (zero-fill outside the image) into the 128 B-swizzled K-major layout that the UMMA
//                     descriptor expects, four k-blocks in flight per thread; fence.proxy.async + mbarrier arrive
//
// The explicit im2col path wrote 9x the input to HBM and read it back (50 GB per 128 x 4 s embedder batch: 19 ms of
// im2col kernels + the GEMMs' operand reads); here the input map is read nine times out of L2 instead.
#pragma once
#include "gemm_cfgs.cuh"
#include "gemm_core.cuh"

namespace tdz {

struct Conv3Params {
  LinearParams L;            // tmB = weights, S = pixels, Sp = padded pixels, N, K = 9*C, n_tiles, e.bias
  const __nv_bfloat16* a;    // input map, pixel-major, leading dimension lda, channel offset offa
  int lda, offa;
  int H, W, C;               // image rows / columns, input channels (multiple of 8)
  __nv_bfloat16* out;        // Hardtanh(0,20)(conv + bias) -> out[p][out_col0 + c], leading dimension out_ld
  int out_ld, out_col0;
  // optional second output, the input of the NEXT 3x3 convolution of a Res2Net chain: out2[p][c] =
  // bf16(out[p][c] + nx[p][nx_off + c])  (sum of two bf16 values formed in fp32, as the im2col path did)
  const __nv_bfloat16* nx;
  int nx_ld, nx_off;
  __nv_bfloat16* out2;
  int out2_ld;
};

constexpr int C3_GATHER_WARPS = 8;
constexpr int C3_MAX_KCHUNKS = 1728 / 8;  // 9 * 192 channels

template <class Cfg>
constexpr int conv3_smem_bytes() {
  return Cfg::STAGES * (GEMM_STAGE_A_BYTES + Cfg::BLOCK_N * 128) + 1024 + 256 + 128 * 8 + C3_MAX_KCHUNKS * 4;
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 16-byte asynchronous global -> shared copy; src_bytes = 0 writes zeros (padding) without touching global memory
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 add_bf16x8(const uint4& x, const uint4& y) {
  const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&x);
  const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&y);
  uint4 o;
  uint32_t* r = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    r[i] = pack_bf16(__bfloat162float(a[i].x) + __bfloat162float(b[i].x),
                     __bfloat162float(a[i].y) + __bfloat162float(b[i].y));
  return o;
}
constexpr int C3_DEPTH = 4;  // k-blocks a gather thread keeps in flight (cp.async groups)

template <class Cfg>
__global__ void __launch_bounds__(gemm_threads(Cfg::EPI_SPLIT) + 32 * C3_GATHER_WARPS, 1)
    gemm_conv3_kernel(const __grid_constant__ Conv3Params CP) {
  const LinearParams& P = CP.L;
  constexpr int BLOCK_N = Cfg::BLOCK_N;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int STAGE_B_BYTES = BLOCK_N * 128;
  constexpr int STAGE_BYTES = GEMM_STAGE_A_BYTES + STAGE_B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                 : (2 * BLOCK_N <= 256) ? 256 : 512;
  constexpr int EPI_WARPS = 4 * Cfg::EPI_SPLIT;
  static_assert(Cfg::PANEL_BYTES == 0 && Cfg::FMT == 1, "direct bf16 epilogues only");
  static_assert(C3_DEPTH - 1 < Cfg::STAGES, "a stage must be published before its ring slot comes round again");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  int2* rowinfo = reinterpret_cast<int2*>(smem_al + STAGES * STAGE_BYTES + 256);          // [128] (h, w) of a tile row
  int* tapc = reinterpret_cast<int*>(smem_al + STAGES * STAGE_BYTES + 256 + 128 * 8);     // [K/8] (dh, dw, channel)
  __shared__ uint32_t s_tmem_base;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // provably warp-uniform: role branches become uniform
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1 + C3_GATHER_WARPS);  // the weight TMA's expect_tx arrive + one arrive per gather warp
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), EPI_WARPS);
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
  // Warp 0 streams WEIGHT k-blocks only (the activation half of a stage is assembled by the gather warps): it does not
  // wait for the producer kernel - under a programmatic dependent launch the first ring of weights is in shared memory
  // when the gather warps are released.  Every other warp waits (and keeps the CTA alive until the producer is done).
  if (warp != 0) pdl_wait();
  const int ntiles = Cfg::num_tiles(P);
  const int nkb = (P.K + 63) / 64;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (tile + static_cast<int>(gridDim.x) >= ntiles) pdl_trigger();  // last tile: the next kernel may be launched
        TileInfo ti;
        Cfg::tile_info(P, tile, ti);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), STAGE_B_BYTES);
          tma_load_2d(smem_base + stage * STAGE_BYTES + GEMM_STAGE_A_BYTES, &P.tmB, full_bar(stage), kb * 64, ti.n0);
          if (++stage == STAGES) {
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
        const int as = it & 1;
        mbar_wait(tempty_bar(as), ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + GEMM_STAGE_A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * 32u, 16u, 1024u);
            const uint64_t db = umma_smem_desc(sb + k * 32u, 16u, 1024u);
            umma_f16(tacc, da, db, IDESC, (kb | k) != 0);
          }
          umma_commit(empty_bar(stage));
          if (kb == nkb - 1) umma_commit(tfull_bar(as));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ---------------- epilogue: thread = pixel; folded-BN bias + Hardtanh(0,20) -> bf16 (+ the chained sum)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int half = (warp - 2) >> 2;
    constexpr int COLS = BLOCK_N / Cfg::EPI_SPLIT;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      if (tile + static_cast<int>(gridDim.x) >= ntiles) pdl_trigger();
      TileInfo ti;
      Cfg::tile_info(P, tile, ti);
      const int as = it & 1;
      mbar_wait(tfull_bar(as), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
      const size_t p = static_cast<size_t>(ti.m0) + row;
      const bool valid = p < static_cast<size_t>(P.S);
#pragma unroll 1
      for (int cc = 0; cc < COLS; cc += 16) {
        const int col0 = ti.n0 + half * COLS + cc;
        if (col0 >= P.N) break;                // warp-uniform
        const bool full = col0 + 16 <= P.N;    // else exactly 8 columns exist
        float v[16], bias[16];
        tmem_ld16(tacc + half * COLS + cc, v);
        ld_f32x16(P.e.bias + col0, bias);
        uint4 nx0 = make_uint4(0u, 0u, 0u, 0u), nx1 = nx0;
        if (valid && CP.out2 != nullptr) {
          const uint4* np = reinterpret_cast<const uint4*>(CP.nx + p * CP.nx_ld + CP.nx_off + col0);
          nx0 = np[0];
          if (full) nx1 = np[1];
        }
        tmem_ld_wait();
        if (!valid) continue;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fminf(fmaxf(v[j] + bias[j], 0.f), 20.f);
        const uint4 o0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]),
                                    pack_bf16(v[6], v[7]));
        const uint4 o1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]),
                                    pack_bf16(v[14], v[15]));
        uint4* o = reinterpret_cast<uint4*>(CP.out + p * CP.out_ld + CP.out_col0 + col0);
        o[0] = o0;
        if (full) o[1] = o1;
        if (CP.out2 != nullptr) {
          uint4* o2 = reinterpret_cast<uint4*>(CP.out2 + p * CP.out2_ld + col0);
          o2[0] = add_bf16x8(o0, nx0);
          if (full) o2[1] = add_bf16x8(o1, nx1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
  } else {
    // ---------------- gather warps: build the im2col k-blocks of the A operand in shared memory
    // Thread g owns the 16 B piece j = g & 7 of rows r_i = 32 i + (g >> 3), i = 0..3, of every k-block: the tap /
    // channel of the piece is looked up once per k-block, the rows' image coordinates once per tile.
    constexpr int GT = 32 * C3_GATHER_WARPS;           // 256 gather threads
    constexpr int RPT = 1024 / GT;                     // rows (pieces) per thread and k-block
    const int g = threadIdx.x - 32 * (2 + EPI_WARPS);
    const int nchunks = P.K >> 3;                      // 16 B pieces per im2col row
    for (int kc = g; kc < nchunks; kc += GT) {
      const int k = kc * 8;
      const int tap = k / CP.C;
      const int dh = tap / 3 - 1, dw = tap - 3 * (tap / 3) - 1;
      // (dh + 1) | (dw + 1) << 2 | channel << 4
      tapc[kc] = (dh + 1) | ((dw + 1) << 2) | ((k - tap * CP.C) << 4);
    }
    const int H = CP.H, W = CP.W, HW = H * W;
    const int j = g & 7, r0 = g >> 3;
    const uint32_t dst0 = (r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4);  // + i * 4096 (32 rows)
    int stage = 0;
    uint32_t phase = 0;
    int last_m0 = -1;
    // k-blocks are copied with cp.async groups, C3_DEPTH of them in flight per thread; a stage is published
    // (fence.proxy.async + arrive on its full barrier) once its group has landed, C3_DEPTH - 1 issues later
    int arrive_stage = 0, issued = 0, arrived = 0;
    auto publish = [&]() {
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(arrive_stage));
      if (++arrive_stage == STAGES) arrive_stage = 0;
      ++arrived;
    };
    int rh[RPT], rw[RPT];
    const __nv_bfloat16* rbase[RPT];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      if (tile + static_cast<int>(gridDim.x) >= ntiles) pdl_trigger();
      TileInfo ti;
      Cfg::tile_info(P, tile, ti);
      if (ti.m0 != last_m0) {  // (n-tiles of the same pixel tile follow each other)
        // everybody has finished reading the previous tile's rowinfo (and, the first time, tapc is complete)
        asm volatile("bar.sync 2, %0;" ::"n"(GT) : "memory");
        if (g < 128) {
          const int p = ti.m0 + g;
          int h = -4, w = -4;  // rows past the last pixel: every tap is out of the image -> zeros
          if (p < P.S) {
            const int rem = p % HW;
            h = rem / W;
            w = rem - h * W;
          }
          rowinfo[g] = make_int2(h, w);
        }
        asm volatile("bar.sync 2, %0;" ::"n"(GT) : "memory");
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int2 hw = rowinfo[r0 + 32 * i];
          rh[i] = hw.x;
          rw[i] = hw.y;
          rbase[i] = CP.a + static_cast<size_t>(ti.m0 + r0 + 32 * i) * CP.lda + CP.offa;
        }
        last_m0 = ti.m0;
      }
      for (int kb = 0; kb < nkb; ++kb) {
        const int kc = kb * 8 + j;
        int dh = 0, dw = 0, c = 0;
        const bool in_k = kc < nchunks;  // pieces past K = 9 C are zeros
        if (in_k) {
          const int tc = tapc[kc];
          dh = (tc & 3) - 1;
          dw = ((tc >> 2) & 3) - 1;
          c = tc >> 4;
        }
        const int soff = (dh * W + dw) * CP.lda + c;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t dst = smem_base + stage * STAGE_BYTES + dst0;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int hh = rh[i] + dh, ww = rw[i] + dw;
          const bool in = in_k && hh >= 0 && hh < H && ww >= 0 && ww < W;
          cp_async16_zfill(dst + i * 4096, in ? rbase[i] + soff : CP.a, in ? 16 : 0);
        }
        cp_async_commit();
        ++issued;
        if (issued - arrived >= C3_DEPTH) {
          cp_async_wait_group<C3_DEPTH - 1>();
          publish();
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    cp_async_wait_group<0>();
    while (arrived < issued) publish();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <class Cfg>
cudaError_t launch_gemm_conv3(const Conv3Params& CP, int ntiles, int num_sms, cudaStream_t st) {
  constexpr int smem = conv3_smem_bytes<Cfg>();
  static std::atomic<unsigned long long> configured{0};
  if (cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(gemm_conv3_kernel<Cfg>), smem, configured); e != cudaSuccess)
    return e;
  if (ntiles <= 0) return cudaSuccess;
  const int grid = ntiles < num_sms ? ntiles : num_sms;
  pdl(gemm_conv3_kernel<Cfg>, grid, gemm_threads(Cfg::EPI_SPLIT) + 32 * C3_GATHER_WARPS, smem, st)(CP);
  return cudaGetLastError();
}

}  // namespace tdz
