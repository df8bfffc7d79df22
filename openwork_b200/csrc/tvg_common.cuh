// Shared device helpers for the two-view geometry kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TVG_SENTINEL_KEY 0xFFFFFFFFu
#define TVG_KEY_SHIFT 23                       // key = dist << 23 | trainIdx  (dist <= 256, idx < 2^23)
#define TVG_KEY_IDX_MASK ((1u << TVG_KEY_SHIFT) - 1u)

namespace tvg {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + TMA (bulk async copy) ---------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- top-2 of packed (dist, idx) keys ---------------------------------------------------
// min on keys == (lowest distance, then lowest train index): OpenCV's strict-< insertion order.
__device__ __forceinline__ void top2_insert(uint32_t& m1, uint32_t& m2, uint32_t k) {
    m2 = min(m2, max(m1, k));
    m1 = min(m1, k);
}
__device__ __forceinline__ void top2_merge(uint32_t& a1, uint32_t& a2, uint32_t b1, uint32_t b2) {
    uint32_t lo = min(a1, b1);
    uint32_t hi = max(a1, b1);
    a2 = min(hi, min(a2, b2));
    a1 = lo;
}

// ---- 64-bit warp shuffles ---------------------------------------------------------------
__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_f64(v, m);
    return v;
}

}  // namespace tvg
