// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld) and the UMMA shared-memory / instruction descriptors.
// Hand-written for this project; semantics follow the PTX ISA (tcgen05 chapter) as summarised in
// /opt/skills/guides/blackwell_cuda_programming.md.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace tdz {

#ifndef TDZ_HANG_GUARD
#define TDZ_HANG_GUARD 1
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- programmatic dependent launch
// Small calls (the streaming shapes: one tile per CTA on a fraction of the SMs, ~650 kernels of 5-25 us) launch every
// kernel with the programmatic-stream-serialization attribute (pdl() below, switched per API call by PdlScope): the
// grid may START while the previous kernel of the stream is still running - its CTAs are placed on free SMs, run
// their prologue (barrier init, TMEM allocation, descriptor prefetch) and block in pdl_wait() until the previous grid
// has completed and its writes are visible.  Rule: no access to memory another kernel writes (or reads and this one
// overwrites) before pdl_wait(); weights, tables and kernel parameters are fair game.  Without the attribute both
// instructions are no-ops.
//   Who triggers: only the persistent GEMM kernels, when a CTA starts its LAST tile (pdl_trigger()).  Measured on
//   B200 (C2 batch, one forward = 191 ms without the attribute): a trigger at the top of EVERY kernel 197 ms - the
//   dependent grid is released while a multi-wave grid still has CTAs to place and its waiting CTAs sit on the SMs;
//   GEMM kernels only: 190.3 ms, embedder 35.7 instead of 34.8 ms; batch-1 streaming step 5.52 -> 5.3 ms.  Hence on
//   for small calls only.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Entry of the non-persistent (SIMT) kernels: wait for the producer grid.  A grid small enough to be resident as a
// whole (at most four CTAs per SM) releases its dependent grid right away - every CTA is already placed; a larger
// grid releases it by exiting (see above: an early trigger in a multi-wave grid cost 3 %).
__device__ __forceinline__ void pdl_enter() {
  if (gridDim.x * gridDim.y * gridDim.z <= 4u * 148u) pdl_trigger();
  pdl_wait();
}

// TDZ_NO_PDL=1 (read once): never set the attribute;  TDZ_PDL_ALL=1: set it for calls of any size (A/B timing).
inline int pdl_mode() {
  static const int mode = [] {
    const char* off = getenv("TDZ_NO_PDL");
    if (off && off[0] && off[0] != '0') return 0;
    const char* all = getenv("TDZ_PDL_ALL");
    return (all && all[0] && all[0] != '0') ? 2 : 1;
  }();
  return mode;
}
inline bool& pdl_call_flag() {
  static thread_local bool on = false;
  return on;
}
// Set by an API entry point for the launches it makes (they happen on the calling thread, under the handle's mutex).
struct PdlScope {
  bool prev;
  explicit PdlScope(bool small_call) : prev(pdl_call_flag()) {
    pdl_call_flag() = pdl_mode() == 2 || (pdl_mode() == 1 && small_call);
  }
  ~PdlScope() { pdl_call_flag() = prev; }
  PdlScope(const PdlScope&) = delete;
  PdlScope& operator=(const PdlScope&) = delete;
};
template <class... KArgs>
struct PdlLaunch {
  void (*kernel)(KArgs...);
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
  template <class... Args>
  cudaError_t operator()(Args&&... args) const {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_call_flag() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
  }
};
// pdl(kernel, grid, block, smem, stream)(args...) stands where a triple-chevron launch with the same configuration would
template <class... KArgs>
PdlLaunch<KArgs...> pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream) {
  return PdlLaunch<KArgs...>{kernel, grid, block, smem, stream};
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One probe of a phase parity.  The suspend-time hint lets the hardware park the thread until the phase completes
// (or the hint expires) instead of returning at once, so a waiting warp does not burn issue slots that the
// epilogue warps of the same SM need.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return done != 0;
}
// Wait on a phase parity.  With TDZ_HANG_GUARD the kernel traps instead of hanging the GPU if a barrier is never
// satisfied (a protocol bug), so a bad launch surfaces as a CUDA error.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
#if TDZ_HANG_GUARD
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s: far beyond any legitimate wait
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}

// Acquire at CLUSTER scope: for barriers whose arrivals come from the peer CTA of a pair and publish shared-memory data.
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    if (done) return;
#if TDZ_HANG_GUARD
    if (clock64() - t0 > 4000000000ll) __trap();
#endif
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Multicast variant: the box is written to the same shared-memory offset of every CTA in `cta_mask` and completes
// bytes on the mbarrier at the same offset in each of them (thread-block cluster, NVSwitch-free on-chip broadcast).
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4, %5}], [%2], %6;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; bf16/fp16 operands, fp32 accumulate.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with tf32 operands (fp32 storage, 10-bit mantissa used).
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all MMAs previously issued by this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Same, arriving on the mbarrier at this offset in every CTA of `cta_mask` (stage release to all multicast sources).
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
// 256-bit global accesses (sm_100: LDG/STG.256): one 32 B sector per thread and instruction, i.e. half the LSU
// requests of the 128-bit form for row-per-thread epilogues.
struct __align__(32) U8 {
  uint32_t r[8];
};
__device__ __forceinline__ U8 ld_global_256(const void* p) {
  U8 v;
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]),
                 "=r"(v.r[7])
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_global_256(void* p, const U8& v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v.r[0]), "r"(v.r[1]),
               "r"(v.r[2]), "r"(v.r[3]), "r"(v.r[4]), "r"(v.r[5]), "r"(v.r[6]), "r"(v.r[7])
               : "memory");
}

// ---------------------------------------------------------------- cta_group::2 (one MMA across a CTA pair)
// A kernel that uses these must use them for every tcgen05 alloc / mma / commit it issues.
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256: each CTA of the pair supplies its own 128 rows of A and its own
// half of the N columns of B from the same shared-memory offsets, and receives its 128 rows of D in its own TMEM.
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg2_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// Relaxed: the only accesses this arrive has to order are the tcgen05.ld reads of the accumulator, which the
// preceding tcgen05.fence::before_thread_sync covers; a release at cluster scope would also wait for every
// outstanding global store of the warp (MEMBAR + ERRBAR, 12 % of the epilogue's stall samples).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// Release form: orders this thread's earlier shared-memory writes (made visible to the async proxy by a preceding
// fence.proxy.async) before the arrival - used where generic-proxy writes feed a tcgen05.mma issued by the peer CTA.
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* m, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
// TMA load into this CTA's shared memory whose completion bytes are counted on an mbarrier that may live in the
// peer CTA of the pair (`cluster_bar` is a shared::cluster address).
__device__ __forceinline__ void tma_load_3d_cg2(uint32_t dst, const CUtensorMap* m, uint32_t cluster_bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i <-> lane base+i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor, 128-byte swizzle (layout type 2), descriptor version 1 (sm_100).
//   K-major operand : rows of 128 B (64 bf16 / 32 tf32), 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major operand: 64-element (128 B) MN atoms, K rows 128 B apart, 8-row K groups SBO apart,
//                     next MN atom LBO apart.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // version = 1
  d |= static_cast<uint64_t>(2) << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor (upper 32 bits of the idesc operand): fp32 accumulate, dense.
//   fmt: 0 = fp16, 1 = bf16 (kind::f16), 2 = tf32 (kind::tf32).  a_mn / b_mn: 1 = MN-major operand.
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_mn,
                                                  uint32_t b_mn) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---------------------------------------------------------------- misc math
// exp / reciprocal on the special-function unit (MUFU.EX2 / MUFU.RCP, <= 2 ulp), flush-to-zero, no slow paths:
// in the GEMM epilogues the issue slots of the few epilogue warps are the critical resource, and IEEE division
// (a subroutine with a fix-up branch per element) costs more than the rest of the epilogue together.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
__device__ __forceinline__ float silu_fast(float x) { return silu_f(x); }
// Packed fp32 pair arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 - two fp32 operations per issue slot).  The time
// convolutions and the epilogue activations are issue-bound, and their data is laid out as channel pairs anyway.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(r)
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
        "l"(reinterpret_cast<unsigned long long&>(c)));
  return reinterpret_cast<float2&>(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(r)
      : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
  return reinterpret_cast<float2&>(r);
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// fp32 -> nearest tf32 (ties away), kept in fp32 storage; used where a buffer is only a tf32 MMA operand,
// because the tensor core itself truncates the low 13 mantissa bits (a biased rounding).
__device__ __forceinline__ float round_tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// two fp32 -> packed fp16, clamped to the finite range (an overflow must not turn into inf / NaN downstream)
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  a = fminf(fmaxf(a, -65504.f), 65504.f);
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace tdz
