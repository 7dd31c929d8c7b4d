// PTX wrappers shared by the tcgen05 kernels of this library (tc_gemm.cu, ru_fused.cu): mbarriers, TMA loads,
// TMEM allocation / loads, tcgen05.mma in cta_group::1 and ::2 forms, 256-bit global accesses, shared-memory
// matrix descriptors.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug traps (an error the host sees) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    if (mbar_try_wait(b, parity)) return;
    const long long t0 = clock64();
    // watchdog: a broken pipeline becomes a launch failure, not a hung GPU.  No printf here: a call in this
    // loop makes the compiler spill whatever is live across the wait -- in the epilogue that was the residual
    // prefetch, whose spill store then waited out the full DRAM latency of every load (54 % of all stall samples).
    while (!mbar_try_wait(b, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ uint32_t opaque_u32(uint32_t x) {
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
// Waits of the MMA-issuing warp, on raw shared-memory addresses: a lean spin of three instructions, no watchdog
// (a stuck pipeline still trips the watchdog of the producer and epilogue warps, which wait on the same
// hand-offs).  The issuing thread's scalar work per stage must stay below the stage's MMA time.
__device__ __forceinline__ void mbar_spin_a(uint32_t addr, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "SPIN_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SPIN_DONE;\n\t"
        "bra SPIN_WAIT;\n\t"
        "SPIN_DONE:\n\t}"
        :: "r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_commit_a(uint32_t addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mma2_commit_both_a(uint32_t addr) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(addr), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// descriptors are passed as their low words; the high word (SBO, version, swizzle mode) is a
// compile-time constant of the kernel
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint32_t desc_a_lo, uint32_t desc_b_lo, uint32_t desc_hi,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(desc_a_lo), "r"(desc_b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// same, accumulate always on (no runtime predicate: keeps the issue loop free of vector->uniform moves)
__device__ __forceinline__ void mma_f16_ss_acc(uint32_t tmem_d, uint32_t desc_a_lo, uint32_t desc_b_lo, uint32_t desc_hi,
                                               uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.eq.u32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(desc_a_lo), "r"(desc_b_lo), "r"(desc_hi), "r"(idesc) : "memory");
}
// ---- cta_group::2 (a pair of CTAs on one TPC computes a 256-row tile; B is split between them) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// Relaxed: the hand-back of a TMEM buffer orders TMEM reads (tcgen05.wait::ld + tcgen05.fence), not memory.
// A release at cluster scope compiles to MEMBAR + ERRBAR and waits for the warp's outstanding global stores
// (ncu: 21 % of the epilogue warps' samples in pair mode).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose bytes are credited to a barrier of the pair's leader CTA (cluster address)
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_f16_ss(uint32_t tmem_d, uint32_t desc_a_lo, uint32_t desc_b_lo, uint32_t desc_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(desc_a_lo), "r"(desc_b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma2_f16_ss_acc(uint32_t tmem_d, uint32_t desc_a_lo, uint32_t desc_b_lo, uint32_t desc_hi,
                                                uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.eq.u32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(desc_a_lo), "r"(desc_b_lo), "r"(desc_hi), "r"(idesc) : "memory");
}
// the same with separate descriptor high words for A and B (operands staged with different swizzle modes)
__device__ __forceinline__ void mma2_f16_ss_acc_hh(uint32_t tmem_d, uint32_t desc_a_lo, uint32_t desc_a_hi, uint32_t desc_b_lo,
                                                   uint32_t desc_b_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.eq.u32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(desc_a_lo), "r"(desc_a_hi), "r"(desc_b_lo), "r"(desc_b_hi), "r"(idesc) : "memory");
}
// completion of the pair's MMAs arrives on the barrier at the same offset in both CTAs
// 256-bit global accesses (sm_100): one full 32-byte sector per lane per instruction
__device__ __forceinline__ void ldg256(const void* p, float (&r)[8]) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
                 : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(void* p, const float (&r)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]) : "memory");
}
__device__ __forceinline__ void stg256u(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address and byte
// offsets in 16-byte units, version 1 (Blackwell), LBO unused for swizzled K-major layouts,
// SBO = one 8-row swizzle group, layout 2 = SWIZZLE_128B / 4 = SWIZZLE_64B.
// Base offset (bits 49-51) stays 0 even for starts that are not aligned to the swizzle period.
// Low word: start address >> 4 | LBO (= 1) << 16; high word: SBO | version << 14 | layout << 29.
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t addr) { return (addr >> 4) | (1u << 16); }
template <int BK>
__device__ __forceinline__ constexpr uint32_t smem_desc_hi() {
    return ((8u * BK * 2u) >> 4) | (1u << 14) | ((BK == 64 ? 2u : 4u) << 29);
}



}  // namespace
