// Blackwell (sm_100a) primitives used by the fused EndoSurf kernels: mbarrier, bulk TMA copies,
// tcgen05 (TMEM alloc / MMA / commit / ld) and the shared-memory matrix descriptors.
// Everything is inline PTX; no CUTLASS dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace es {

// ------------------------------------------------------------------ error plumbing
// Device-side watchdog: a barrier wait that spins longer than this many polls records an error
// code and the kernel drains (every later wait returns immediately) instead of hanging the GPU.
#ifndef ES_WATCHDOG_SPINS
#define ES_WATCHDOG_SPINS (1u << 26)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Slow path of mbar_wait, kept out of line so that the hot loops only carry the try_wait + branch.  Returns false if
// the watchdog tripped (the caller keeps going; results are garbage and the error word tells the host).  `err` points
// to a global int; `code` identifies the wait site.
#ifndef ES_WAIT_BACKOFF_NS
#define ES_WAIT_BACKOFF_NS 128
#endif
static __device__ __noinline__ bool mbar_wait_slow(uint32_t bar_sa, uint32_t parity, int* err, int code) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_sa), "r"(parity)
        : "memory");
    if (ok) return true;
    if (++spins > ES_WATCHDOG_SPINS) {
      atomicCAS(err, 0, code);
      return false;
    }
    if ((spins & 0xFFFF) == 0 && *reinterpret_cast<volatile int*>(err) != 0) return false;
#if ES_WAIT_BACKOFF_NS > 0
    // Waits that are long by construction (the epilogue waiting for an accumulator or a free ring slot, the weight
    // producer and the record warp waiting for the MMAs) back off between polls: a polling warp costs ~6 issue slots per
    // iteration on a sub-partition it shares with the MMA issuer and the working epilogue warps (source-level ncu:
    // 12-13 % of the executed instructions of a chain were barrier polls).  The MMA issuer's own waits (codes 3xx) and
    // the pair relay (41x, 42x) stay tight.
    if (code < 300 || code == 400 || code == 401 || code == 500) asm volatile("nanosleep.u32 %0;" ::"n"(ES_WAIT_BACKOFF_NS));
#endif
  }
}
__device__ __forceinline__ bool mbar_wait_sa(uint32_t bar_sa, uint32_t parity, int* err, int code) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar_sa), "r"(parity)
      : "memory");
  if (ok) return true;
  return mbar_wait_slow(bar_sa, parity, err, code);
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
  return mbar_wait_sa(smem_u32(bar), parity, err, code);
}
__device__ __forceinline__ void mbar_arrive_sa(uint32_t bar_sa) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_sa) : "memory");
}

// ------------------------------------------------------------------ shared-memory stores by 32-bit shared address
template <int OFF>
__device__ __forceinline__ void sts32(uint32_t sa, uint32_t v) {
  asm volatile("st.shared.b32 [%0+%1], %2;" ::"r"(sa), "n"(OFF), "r"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts64(uint32_t sa, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0+%1], {%2, %3};" ::"r"(sa), "n"(OFF), "r"(a), "r"(b) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts128(uint32_t sa, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0+%1], {%2, %3, %4, %5};" ::"r"(sa), "n"(OFF), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ float2 lds64f(uint32_t sa) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(sa));
  return r;
}
__device__ __forceinline__ float4 lds128f(uint32_t sa) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(sa));
  return r;
}

// ------------------------------------------------------------------ proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------ bulk TMA copy global -> shared (non-tensor form)
// size must be a multiple of 16; src/dst 16-byte aligned.  Completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// the same with an L2 cache-policy operand (createpolicy): the packed weights are re-read by every CTA for every tile
// and should outlive the streaming traffic of the training launches (evict_last)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx_sa(uint32_t bar_sa, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s_sa(uint32_t smem_dst_sa, const void* gsrc, uint32_t bytes, uint32_t bar_sa) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst_sa),
      "l"(gsrc), "r"(bytes), "r"(bar_sa)
      : "memory");
}

// ------------------------------------------------------------------ bulk TMA copy shared -> global (non-tensor form)
// Completion is tracked per issuing thread with bulk async-groups.  The source must have been made visible to the
// async proxy (fence.proxy.async after the generic-proxy stores, before the barrier that hands the buffer over).
__device__ __forceinline__ void tma_bulk_s2g(void* gdst, uint32_t smem_src_sa, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src_sa), "r"(bytes)
               : "memory");
}
// the same with an L2 eviction-priority hint (plane records are streamed once: evict_first keeps the packed weights and
// the lines other kernels will re-read resident)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_bulk_s2g_hint(void* gdst, uint32_t smem_src_sa, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(smem_src_sa), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source (the buffer may be overwritten)
__device__ __forceinline__ void tma_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all bulk groups of this thread are complete (writes performed)
__device__ __forceinline__ void tma_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ TMEM
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives lane (base_lane+i), columns [c, c+32).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]) {
  static_assert(N == 16 || N == 32, "tmem_ld: 16 or 32 columns");
  if constexpr (N == 16) tmem_ld16(taddr, v);
  else tmem_ld32(taddr, v);
}
// 16 lanes x 256 bit (x2 = 16 columns): thread t of the warp receives, for TMEM lanes L = base_lane + t/4 and L + 8
// and column pairs c = col + 2*(t%4) + {0,1} and c + 8:   r0,r1 = (L, c..c+1)   r2,r3 = (L+8, c..c+1)
//                                                          r4,r5 = (L, c+8..c+9) r6,r7 = (L+8, c+8..c+9)
// (the m16n8 accumulator-fragment layout; cute SM100_TMEM_LOAD_16dp256b2x).  base_lane = 32*(warp%4) or that + 16.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float& r0, float& r1, float& r2, float& r3,
                                                   float& r4, float& r5, float& r6, float& r7) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r0), "=f"(r1), "=f"(r2), "=f"(r3), "=f"(r4), "=f"(r5), "=f"(r6), "=f"(r7)
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ UMMA descriptors (see cute/arch/mma_sm100_desc.hpp
// for the bit layout; restated here).  K-major, no swizzle ("interleave"): a core matrix is 8 rows x 16 bytes stored
// as 128 contiguous bytes; SBO = byte stride between 8-row groups, LBO = byte stride between the two K core matrices.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version for sm_100
  return d;                             // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// kind::f16, A/B = fp16 (format code 0) K-major, D = fp32 (c_format 1), M x N tile.
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, descriptors given as their low words: the high words (stride fields + version, make_smem_desc >> 32) are
// immediates, so only two 32-bit values per MMA have to reach the uniform registers.
template <uint32_t A_HI, uint32_t B_HI>
__device__ __forceinline__ void umma_f16_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(A_HI), "n"(B_HI)
      : "memory");
}
__host__ __device__ constexpr uint32_t smem_desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14);  // bits 32..45 SBO, bit 46 descriptor version
}
__host__ __device__ constexpr uint32_t smem_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
// mbarrier arrives (count 1) when all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void umma_commit_sa(uint32_t bar_sa) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_sa) : "memory");
}
// ------------------------------------------------------------------ CTA pairs (cta_group::2): two CTAs of a cluster
// on the two SMs of a TPC run ONE M = 256 MMA per instruction - each supplies its own 128 rows of A and half of the
// N rows of B from the same shared-memory offsets, each receives its own 128 rows of D in its own TMEM - so the
// shared-memory traffic of the B operand (reads and the weight stream's writes) is halved per SM.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctaid_x() {  // clusters in the grid
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t sa, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(sa), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_sa) {
  // default (cta-scope) semantics as CUTLASS' ClusterBarrier::arrive: a cluster-scope release costs a memory barrier and
  // an L1 invalidation per arrive (measured: the relay thread then limits the whole pipeline); the data this arrive
  // announces was published to the async proxy by its writers (fence.proxy.async) before their own local arrives
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_sa) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result) {  // same warp index in both CTAs of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
template <uint32_t A_HI, uint32_t B_HI>
__device__ __forceinline__ void umma2_f16_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(A_HI), "n"(B_HI)
      : "memory");
}
// arrives (count 1) on the barrier at this shared-memory offset in BOTH CTAs of the pair when all previously issued
// tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma2_commit_sa(uint32_t bar_sa) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_sa),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// bulk copy global -> the SAME shared-memory offset of every CTA in `mask` of the cluster; each destination CTA's
// mbarrier at the same offset receives the complete_tx
__device__ __forceinline__ void tma_bulk_g2s_mcast(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                                   uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// single-CTA MMAs, completion signalled to the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ fp16 hi/lo split
// x ~= hi + lo with hi = fp16(x), lo = fp16(x - hi) (round-to-nearest-even): 22 mantissa bits, i.e. a relative
// representation error of 2^-22, 16x tighter than a bf16 pair for the same three MMAs (hi*hi + hi*lo + lo*hi).
// Products of two fp16 values are exact in the fp32 accumulator.  Range: |x| must stay below 65504 (activations,
// tangents and weights of these weight-normalised 256-wide MLPs are O(1..100)); tiny values lose nothing that
// matters because the error is absolute (fp16 subnormal spacing 6e-8).
// The conversion saturates (F2FP.SATFINITE): an adjoint that outgrows the fp16 range inside a reverse chain is
// clipped to +-65504 instead of turning every later gradient into NaN.
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {  // a -> low half, b -> high half
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// The residuals a - hi.x, b - hi.y are mixed-precision adds (add.f32.f16, SASS FHADD with the half selected and
// negated by operand modifiers): 4 instructions per pair instead of 6 (unpack, unpack, FADD, FADD).
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16x2_sat(a, b);
  float ra, rb;
  asm("{\n\t.reg .b16 h0, h1, n0, n1;\n\t"
      "mov.b32 {h0, h1}, %2;\n\t"
      "neg.f16 n0, h0;\n\t"
      "neg.f16 n1, h1;\n\t"
      "add.rn.f32.f16 %0, n0, %3;\n\t"
      "add.rn.f32.f16 %1, n1, %4;\n\t}"
      : "=f"(ra), "=f"(rb)
      : "r"(hi), "f"(a), "f"(b));
  lo = pack_f16x2_sat(ra, rb);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace es
