// tma.cuh — mbarrier + bulk-copy (TMA engine) primitives for sm_100a, raw PTX.
//   cp.async.bulk            -> SASS UBLKCP   (contiguous 1-D bulk copy global -> shared on the TMA engine, completes on an mbarrier)
// Tensor-map copies (cp.async.bulk.tensor, SASS UTMALDG) are deliberately not used: a box cannot wrap, zero fill only covers
// Remove(0), box sides are capped at 256 elements (a 1 KiB + halo Float32 row needs two boxes) and the dense box layout would
// drop the "global and shared addresses agree mod 128" placement every kernel relies on; whole-row bulk copies with
// producer-side addressing cover every boundary with one code path and reach 0.92-0.97 of the HBM roofline.
#pragma once
#include <cstdint>

namespace sb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// Make barrier initialisation visible to the async proxy before any bulk copy may signal it.
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// Producer-side wait for a free stage. A failed try_wait returns at once, so a producer whose consumers are the bottleneck
// spins at one probe per ~10 ns and competes with them for issue slots (ncu r02e, box3d: SYNCS + YIELD + BRA of the two
// producer warps were 17 % of all issued instructions); sleeping between probes costs nothing — the ring holds several
// stages of slack. SB200_PRODUCER_BACKOFF_NS = 0 keeps the plain loop.
#ifndef SB200_PRODUCER_BACKOFF_NS
#define SB200_PRODUCER_BACKOFF_NS 0
#endif
// try_wait with a suspend-time hint: the thread may sleep inside the instruction for up to `ns` nanoseconds before it reports
// "not yet" (the plain form returns at once; __nanosleep(200) between probes did not slow the loop measurably: r02f, one probe
// every ~12 ns, 39.6 M NANOSLEEP instructions in a 0.8 ms kernel).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_producer(uint64_t* bar, uint32_t parity, unsigned backoff_ns = SB200_PRODUCER_BACKOFF_NS) {
    if (backoff_ns == 0) {
        while (!mbar_try_wait(bar, parity)) {
        }
        return;
    }
    while (!mbar_try_wait_hint(bar, parity, backoff_ns)) __nanosleep(backoff_ns);
}

// Predicated shared-memory load as ONE instruction (`@p LDS`): the end lanes of a warp fetch the cells next to its span
// without a divergent branch (ptxas turns `if (lane == 0) x = *p;` into BSSY / BRA / BSYNC sequences).
__device__ __forceinline__ void lds_if(float& v, const void* p, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.shared.f32 %0, [%1];\n\t}" : "+f"(v) : "r"(smem_u32(p)), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void lds_if(double& v, const void* p, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q ld.shared.f64 %0, [%1];\n\t}" : "+d"(v) : "r"(smem_u32(p)), "r"((int)pred) : "memory");
}

// Contiguous bulk copy global -> shared; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Element-granular asynchronous copy (LDGSTS) of one 4- or 8-byte element, for layouts the bulk-copy engine cannot
// address (rows that are not 16-byte aligned, Halo padding), and the arrive-on-completion that ties all prior cp.async
// of the executing thread to an mbarrier (.noinc: the arrival is part of the barrier's expected count).
template <typename T> __device__ __forceinline__ void cp_async_elem(T* dst_smem, const T* src_gmem) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "cp.async sizes");
    if constexpr (sizeof(T) == 4)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Order generic-proxy accesses to shared memory before subsequent async-proxy (TMA) accesses.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace sb
