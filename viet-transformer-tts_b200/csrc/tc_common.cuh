// tc_common.cuh -- sm_100a primitives: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 / TMEM.
// Hand-written inline PTX; no CUTLASS dependency.
#pragma once

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime)
#include "common.cuh"

namespace vtts {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_addr(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    return ok != 0;
}
// out of line: the diagnostic must not bloat every wait site (instruction-cache footprint of the big kernels)
static __device__ __noinline__ void mbar_timeout() {
    printf("vtts: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
    __trap();
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
    if (mbar_try_wait_addr(bar_addr, parity)) return;
    const long long t0 = clock64();
    for (uint32_t spins = 1; !mbar_try_wait_addr(bar_addr, parity); ++spins)   // clock checked every 64 probes: short loop
        if ((spins & 63u) == 0 && clock64() - t0 > 4000000000LL) mbar_timeout();
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
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
// Bounded wait: a protocol bug traps (CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    for (uint32_t spins = 1; !mbar_try_wait(bar, parity); ++spins)
        if ((spins & 63u) == 0 && clock64() - t0 > 4000000000LL) mbar_timeout();  // ~2 s
}

// Relaxed wait for warps that are not on the critical path (epilogue): the hardware suspends the
// warp for up to `hint_ns` per probe instead of hot-polling the barrier unit, which the TMA /
// tcgen05.commit arrivals and the single MMA-issuing thread also depend on.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t *bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    for (uint32_t spins = 1; !mbar_try_wait_hint(bar, parity, 2000000u); ++spins)
        if ((spins & 63u) == 0 && clock64() - t0 > 4000000000LL) mbar_timeout();
}

// bulk L2 prefetch of a contiguous global range (16-byte aligned, size a multiple of 16): one instruction by one thread
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Producers run several stages ahead of the tensor pipe, so their waits for a free stage are long and not latency
// critical: suspend (VTTS_PRODUCER_SPIN=1 restores the hot poll) instead of burning issue slots and power next to the
// epilogue warps.
#ifndef VTTS_PRODUCER_SPIN
__device__ __forceinline__ void mbar_wait_producer(uint64_t *bar, uint32_t parity) { mbar_wait_relaxed(bar, parity); }
#else
__device__ __forceinline__ void mbar_wait_producer(uint64_t *bar, uint32_t parity) { mbar_wait(bar, parity); }
#endif

// ---- TMA ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// L2 eviction policy for streams that are dead after this read (activations, residual): evict-first keeps the lines
// this kernel WRITES in L2 instead, where the next kernel (which walks the tiles in the opposite direction) finds them.
__device__ __forceinline__ uint64_t l2_policy(bool evict_first) {
    uint64_t pol;
    if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
// (plain ld.global, not .nc: these streams were written by the preceding kernel, whose lifetime can overlap this
// kernel's under programmatic dependent launch - .nc is only defined for data that is read-only for the whole kernel)
__device__ __forceinline__ float4 ldg_f4_hint(const float *p, uint64_t policy) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(policy));
    return v;
}

// multicast: the box lands at the same shared-memory offset of every CTA in `mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_3d_mc(void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
          "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05 / TMEM ------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// same, arriving on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// One pipeline step fused into a single asm block so the instruction order is ours:
//   1. non-blocking probe of the NEXT stage's full-barrier (its ~100-300 cycle latency then overlaps
//      the MMA issue instead of idling the shallow MMA queue),
//   2. the K-step MMAs of the CURRENT stage,
//   3. tcgen05.commit -> this stage's empty-barrier,
//   4. materialise the probe result (consumed at the top of the next step).
__device__ __forceinline__ uint32_t umma_step4(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t idesc,
                                               uint32_t acc_first, uint32_t next_full_addr, uint32_t next_parity,
                                               uint32_t this_empty_addr) {
    uint32_t ready;
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
        "add.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, 4;\n\tadd.u64 b, %3, 4;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, 6;\n\tadd.u64 b, %3, 6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(ready)
        : "r"(tmem_d), "l"(a0), "l"(b0), "r"(idesc), "r"(acc_first), "r"(next_full_addr), "r"(next_parity),
          "r"(this_empty_addr)
        : "memory");
    return ready;
}
__device__ __forceinline__ uint32_t umma_step2(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t idesc,
                                               uint32_t acc_first, uint32_t next_full_addr, uint32_t next_parity,
                                               uint32_t this_empty_addr) {
    uint32_t ready;
    asm volatile(
        "{\n\t.reg .pred p, q, t;\n\t.reg .b64 a, b;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
        "add.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(ready)
        : "r"(tmem_d), "l"(a0), "l"(b0), "r"(idesc), "r"(acc_first), "r"(next_full_addr), "r"(next_parity),
          "r"(this_empty_addr)
        : "memory");
    return ready;
}
// Whole-warp variant: every lane executes it converged with warp-uniform operands (so the compiler keeps the
// descriptors in uniform registers, no per-MMA R2UR waterfall); elect.sync picks the lane that issues.
__device__ __forceinline__ uint32_t umma_step4_warp(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t idesc,
                                                    uint32_t acc_first, uint32_t next_full_addr, uint32_t next_parity,
                                                    uint32_t this_empty_addr) {
    uint32_t ready;
    asm volatile(
        "{\n\t.reg .pred p, q, t, e;\n\t.reg .b64 a, b;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
        "add.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, 4;\n\tadd.u64 b, %3, 4;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, 6;\n\tadd.u64 b, %3, 6;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(ready)
        : "r"(tmem_d), "l"(a0), "l"(b0), "r"(idesc), "r"(acc_first), "r"(next_full_addr), "r"(next_parity),
          "r"(this_empty_addr)
        : "memory");
    return ready;
}
// Whole-warp variant: every lane executes it converged with warp-uniform operands (so the compiler keeps the
// descriptors in uniform registers, no per-MMA R2UR waterfall); elect.sync picks the lane that issues.
__device__ __forceinline__ uint32_t umma_step2_warp(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t idesc,
                                                    uint32_t acc_first, uint32_t next_full_addr, uint32_t next_parity,
                                                    uint32_t this_empty_addr) {
    uint32_t ready;
    asm volatile(
        "{\n\t.reg .pred p, q, t, e;\n\t.reg .b64 a, b;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
        "add.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(ready)
        : "r"(tmem_d), "l"(a0), "l"(b0), "r"(idesc), "r"(acc_first), "r"(next_full_addr), "r"(next_parity),
          "r"(this_empty_addr)
        : "memory");
    return ready;
}
// Two taps per step (narrow layers): the weight stage holds two consecutive tap tiles (tap stride a_tap_step),
// b1 is the activation descriptor of the second tap.  Halves the per-MMA barrier/commit overhead.
__device__ __forceinline__ uint32_t umma_step2x2_warp(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t idesc,
                                                    uint32_t acc_first, uint32_t next_full_addr, uint32_t next_parity,
                                                    uint32_t this_empty_addr, uint64_t b1, uint64_t a_tap_step) {
    uint32_t ready;
    asm volatile(
        "{\n\t.reg .pred p, q, t, e;\n\t.reg .b64 a, b;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
        "add.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, %10;\n\tadd.u64 a, a, 0;\n\tadd.u64 b, %9, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, %10;\n\tadd.u64 a, a, 2;\n\tadd.u64 b, %9, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(ready)
        : "r"(tmem_d), "l"(a0), "l"(b0), "r"(idesc), "r"(acc_first), "r"(next_full_addr), "r"(next_parity),
          "r"(this_empty_addr), "l"(b1), "l"(a_tap_step)
        : "memory");
    return ready;
}
// Two taps per step (narrow layers): the weight stage holds two consecutive tap tiles (tap stride a_tap_step),
// b1 is the activation descriptor of the second tap.  Halves the per-MMA barrier/commit overhead.
__device__ __forceinline__ uint32_t umma_step4x2_warp(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t idesc,
                                                    uint32_t acc_first, uint32_t next_full_addr, uint32_t next_parity,
                                                    uint32_t this_empty_addr, uint64_t b1, uint64_t a_tap_step) {
    uint32_t ready;
    asm volatile(
        "{\n\t.reg .pred p, q, t, e;\n\t.reg .b64 a, b;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %4, p;\n\t"
        "add.u64 a, %2, 2;\n\tadd.u64 b, %3, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, 4;\n\tadd.u64 b, %3, 4;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, 6;\n\tadd.u64 b, %3, 6;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, %10;\n\tadd.u64 a, a, 0;\n\tadd.u64 b, %9, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, %10;\n\tadd.u64 a, a, 2;\n\tadd.u64 b, %9, 2;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, %10;\n\tadd.u64 a, a, 4;\n\tadd.u64 b, %9, 4;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "add.u64 a, %2, %10;\n\tadd.u64 a, a, 6;\n\tadd.u64 b, %9, 6;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %4, t;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(ready)
        : "r"(tmem_d), "l"(a0), "l"(b0), "r"(idesc), "r"(acc_first), "r"(next_full_addr), "r"(next_parity),
          "r"(this_empty_addr), "l"(b1), "l"(a_tap_step)
        : "memory");
    return ready;
}
// One K-slice of KSTEPS (2 or 4) 16-element steps, whole converged warp with warp-uniform operands (descriptors stay in
// uniform registers); elect.sync picks the issuing lane.  No barrier traffic: the caller commits once per accumulator.
template <int KSTEPS>
__device__ __forceinline__ void umma_slice_warp(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t idesc, uint32_t acc_first) {
    static_assert(KSTEPS == 2 || KSTEPS == 4, "KSTEPS");
    if (KSTEPS == 2) {
        asm volatile(
            "{\n\t.reg .pred p, t, e;\n\t.reg .b64 a, b;\n\t"
            "elect.sync _|e, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "setp.eq.b32 t, 0, 0;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "add.u64 a, %1, 2;\n\tadd.u64 b, %2, 2;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n\t}"
            ::"r"(tmem_d), "l"(a0), "l"(b0), "r"(idesc), "r"(acc_first)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p, t, e;\n\t.reg .b64 a, b;\n\t"
            "elect.sync _|e, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "setp.eq.b32 t, 0, 0;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "add.u64 a, %1, 2;\n\tadd.u64 b, %2, 2;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n\t"
            "add.u64 a, %1, 4;\n\tadd.u64 b, %2, 4;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n\t"
            "add.u64 a, %1, 6;\n\tadd.u64 b, %2, 6;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %3, t;\n\t}"
            ::"r"(tmem_d), "l"(a0), "l"(b0), "r"(idesc), "r"(acc_first)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit_elect(uint64_t *bar) {   // whole converged warp; one lane commits
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// 16 lanes x 256 bits x2: the mma-style accumulator fragment.  Thread T of the warp gets, for the 16 TMEM lanes
// starting at taddr's lane and the 16 columns starting at its column:
//   r[0], r[1] = lane T/4,     columns 2*(T%4), 2*(T%4)+1        r[4..7]: the same for columns +8
//   r[2], r[3] = lane T/4 + 8, columns 2*(T%4), 2*(T%4)+1
__device__ __forceinline__ void tmem_ld_16x256_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// 16 lanes x 256 bits x1: the same fragment for 8 columns (r[0], r[1] = lane T/4; r[2], r[3] = lane T/4 + 8;
// columns 2*(T%4), 2*(T%4)+1)
__device__ __forceinline__ void tmem_ld_16x256_x1(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
// registers -> TMEM: 32 lanes x 32 consecutive fp32 columns (thread = TMEM lane)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Four 8x8 16-bit matrices, stored transposed: thread T passes the shared-memory address of row T%8 of matrix T/8
// (16 bytes per row); register i is the thread's fragment of matrix i (element pair (T/4, 2*(T%4)), (T/4, 2*(T%4)+1)),
// which lands in rows 2*(T%4), 2*(T%4)+1 at 16-bit column T/4.
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b),
                 "r"(c), "r"(d)
                 : "memory");
}

// ---- descriptors ---------------------------------------------------------------------------------
// Instruction descriptor, kind::f16: 16-bit x 16-bit -> fp32, both operands K-major.
// fmt: 0 = bf16 operands, 1 = fp16 operands (same tensor-core rate; fp16 has 3 more mantissa bits).
__host__ __device__ constexpr uint32_t make_idesc_16(int M, int N, int fmt) {
    return (1u << 4)                                  // c_format = F32
           | ((fmt == 0 ? 1u : 0u) << 7)              // a_format: 1 = BF16, 0 = F16
           | ((fmt == 0 ? 1u : 0u) << 10)             // b_format
           | ((uint32_t)(N >> 3) << 17)               // n_dim
           | ((uint32_t)(M >> 4) << 24);              // m_dim
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) { return make_idesc_16(M, N, 0); }
// Shared-memory matrix descriptor, K-major, swizzled rows of `row_bytes` (128 -> SW128, 64 -> SW64).
// Rows are stored at a pitch of row_bytes; 8-row swizzle atoms are contiguous (SBO = 8 * row_bytes).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, int row_bytes, uint32_t base_offset) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);            // start address
    d |= (uint64_t)1 << 16;                                // LBO (unused for swizzled K-major)
    d |= (uint64_t)(((8 * row_bytes) >> 4) & 0x3FFF) << 32;  // SBO
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= layout << 61;
    return d;
}

// ---- host: tensor-map encoding through the runtime's driver entry point ---------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// 16-bit element tensor (bf16 or fp16: moved as opaque 16-bit words), `rank` dims
// (dim 0 innermost/contiguous), zero OOB fill.
int make_tmap_bf16(CUtensorMap *out, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                   const uint32_t *box, int swizzle_bytes);

}  // namespace tc
}  // namespace vtts
