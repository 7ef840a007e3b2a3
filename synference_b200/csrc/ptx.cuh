// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA, TMEM).
// Written against the PTX ISA 8.7 shipped with CUDA 12.9; no CUTLASS/CuTe dependency.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace sb2 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The suspend-time hint lets the hardware park a waiting warp for up to ~kWaitHintNs instead of returning early: waiting
// warps (writers, producer, idle epilogue groups) otherwise re-issue the try_wait loop hundreds of millions of times per
// launch and take issue slots from the MMA warp on their scheduler (ncu, round 2: 45 % of all executed instructions).
// A completed phase wakes the warp at once, so the hint adds no latency.
constexpr uint32_t kWaitHintNs = 20000u;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kWaitHintNs)
      : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug traps instead of hanging the GPU box.  Before trapping, the waiter
// records who it is in a host-mapped buffer (g_wait_dbg, set by the host; readable after the context died).
__device__ unsigned int* g_wait_dbg = nullptr;
__device__ __noinline__ void mbar_timeout(uint32_t id, uint32_t parity) {
  if (g_wait_dbg && (threadIdx.x & 31u) == 0u) {
    const unsigned slot = atomicAdd(g_wait_dbg, 1u);
    if (slot < 63u) {
      unsigned int* e = g_wait_dbg + 4 + slot * 4;
      e[0] = id; e[1] = blockIdx.x; e[2] = threadIdx.x; e[3] = parity;
    }
    __threadfence_system();
  }
  __trap();
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// try_wait suspends the warp in hardware for a while (it does not burn the issue slots the epilogue warps of the
// same scheduler need).  The watchdog turns a protocol bug into a trap instead of a hung GPU; it is wall-clock based
// (%globaltimer keeps running while a context is time-sliced or stopped in a debugger), so the bound is generous -- no
// legitimate wait of these kernels lasts a millisecond -- and -DSB2_NO_WATCHDOG removes it altogether.
#ifndef SB2_WATCHDOG_NS
#define SB2_WATCHDOG_NS 30000000000ull
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t id = 0) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3Fu) == 0u) {
      const uint64_t now = global_ns();
#ifndef SB2_NO_WATCHDOG
      if (t0 == 0) t0 = now;
      else if (now - t0 > SB2_WATCHDOG_NS) mbar_timeout(id, parity);
#endif
    }
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier (bytes), with an L2 cache hint.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "l"(cache_hint)
      : "memory");
}
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

// ---------------------------------------------------------------- tcgen05 / TMEM
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 inputs, FP32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets columns [c, c+32) of lane (base + t).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait::ld for a load issued EARLIER (prefetch): the registers only hold the data after the wait, so every later
// use must be ordered behind it -- the empty asm makes the compiler treat them as rewritten here.
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  asm volatile(""
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}

// K-major operand tile in shared memory, 128-byte rows, SWIZZLE_128B (what TMA wrote):
// 8-row groups are 1024 B apart (SBO); LBO is unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);        // start address      [0,14)
  d |= static_cast<uint64_t>(1) << 16;                           // leading byte off   [16,30)  (ignored)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // stride byte offset [32,46)
  d |= static_cast<uint64_t>(1) << 46;                           // descriptor version [46,48) = 1 on sm_100
  d |= static_cast<uint64_t>(2) << 61;                           // layout type        [61,64) = SWIZZLE_128B
  return d;
}
// kind::tf32 instruction descriptor: D=F32, A=B=TF32, both K-major, dense.
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t m, uint32_t n) {
  return (1u << 4)            // c_format  = F32
         | (2u << 7)          // a_format  = TF32
         | (2u << 10)         // b_format  = TF32
         | (0u << 15)         // a_major   = K
         | (0u << 16)         // b_major   = K
         | ((n >> 3) << 17)   // n_dim
         | ((m >> 4) << 24);  // m_dim
}

// ---------------------------------------------------------------- CTA pairs (cluster of 2, cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are counted on an mbarrier that may live in the peer CTA (leader's barrier).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, uint64_t cache_hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1),
        "l"(cache_hint)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 split over the CTA pair; issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-uniform issue: every lane of the MMA warp executes these with identical operands; `elected` is 1 in
// exactly one lane (elect_one()), which is the lane that issues.  Keeping the issue loop free of divergent
// branches lets ptxas hold descriptors in uniform registers instead of broadcasting them for every MMA.
__device__ __forceinline__ void umma_tf32_e(uint32_t elected, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(elected)
      : "memory");
}
// both operands in shared memory, bfloat16 (K = 16 per instruction: the same 32 bytes per row as a K = 8 TF32 slice)
__device__ __forceinline__ void umma_bf16_e(uint32_t elected, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm_e(uint32_t elected, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_commit_e(uint32_t elected, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"(elected)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_e(uint32_t elected, uint64_t* bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %2, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
          smem_u32(bar)),
      "h"(mask), "r"(elected)
      : "memory");
}
// Elected-lane forms of the producer's operations (same warp-uniform scheme as the MMA issue).
__device__ __forceinline__ void mbar_expect_tx_e(uint32_t elected, uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "r"(bytes), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_e(uint32_t elected, uint32_t cluster_addr) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n\t}" ::"r"(cluster_addr),
      "r"(elected)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_e(uint32_t elected, uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_addr,
                                              int c0, int c1, uint64_t cache_hint) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %6, 0;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;\n\t}"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1), "l"(cache_hint), "r"(elected)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm_e(uint32_t elected, uint32_t smem_dst, const CUtensorMap* m,
                                                  uint32_t bar_cluster_addr, int c0, int c1, uint64_t cache_hint) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %6, 0;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;\n\t}"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "l"(cache_hint),
        "r"(elected)
      : "memory");
}
// ---------------------------------------------------------------- TS-mode MMA (A operand in TMEM), TMEM stores, bulk copies
// D[tmem] (+)= A[tmem] * B[smem]: A is M=128 lanes x K=8 consecutive 32-bit columns at tmem_a (element (m, k) in lane m,
// column tmem_a + k), B a K-major SWIZZLE_128B shared-memory tile.  Warp-uniform issue, one elected lane (see umma_tf32_e).
__device__ __forceinline__ void umma_tf32_ts_e(uint32_t elected, uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(elected)
      : "memory");
}
// Same, with the B descriptor given as its LOW word only (start address >> 4 | LBO field); the high word of a K-major
// SWIZZLE_128B descriptor is a constant (SBO = 1024 B, version 1, layout 2), so the issue loop does 32-bit arithmetic.
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t kDescLoBase = 1u << 16;
__device__ __forceinline__ void umma_tf32_ts_lo_e(uint32_t elected, uint32_t tmem_d, uint32_t tmem_a, uint32_t desc_b_lo,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "r"(desc_b_lo), "r"(idesc), "r"(accumulate), "r"(elected), "r"(kDescHiSw128)
      : "memory");
}
// kind::f16 instruction descriptor with BF16 operands: D=F32, A=B=BF16, both K-major, dense (K = 16 per instruction).
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4)            // c_format  = F32
         | (1u << 7)          // a_format  = BF16
         | (1u << 10)         // b_format  = BF16
         | (0u << 15)         // a_major   = K
         | (0u << 16)         // b_major   = K
         | ((n >> 3) << 17)   // n_dim
         | ((m >> 4) << 24);  // m_dim
}
// TS-mode MMA of 16-bit operands: A is M=128 lanes x K=16 elements packed two per 32-bit column (element 2j in the low half
// of column tmem_a + j), B a K-major SWIZZLE_128B tile of the same 32 bytes per row as a K=8 tile of 32-bit elements.
__device__ __forceinline__ void umma_bf16_ts_lo_e(uint32_t elected, uint32_t tmem_d, uint32_t tmem_a, uint32_t desc_b_lo,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "r"(desc_b_lo), "r"(idesc), "r"(accumulate), "r"(elected), "r"(kDescHiSw128)
      : "memory");
}
// two floats -> one 32-bit word of two bfloat16 (round to nearest even): `even` in the low half, `odd` in the high half
__device__ __forceinline__ uint32_t pack_bf16x2(float even, float odd) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd), "f"(even));
  return r;
}
// registers -> TMEM: thread t of the warp writes columns [c, c+N) of lane (base + t)  (mirror of tmem_ld_32x32b_x32)
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 1-D bulk copy global -> shared (bytes and both addresses multiples of 16), completion counted on an mbarrier
__device__ __forceinline__ void bulk_load_e(uint32_t elected, uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar_addr) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %4, 0;\n\t"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar_addr), "r"(elected)
      : "memory");
}
// Register budget hand-over between warpgroups (all four warps of an aligned warpgroup execute the same one)
template <int kRegs> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }

// Tell the compiler a value is the same in every lane of the (converged) warp.
__device__ __forceinline__ int warp_uniform(int v) { return __shfl_sync(0xffffffffu, v, 0); }

// Arrive (once all previously issued MMAs completed) on the mbarrier at this smem offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// named barriers (ids 1..15; 0 is __syncthreads): `count` threads take part, arriving or waiting
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

}  // namespace sb2
