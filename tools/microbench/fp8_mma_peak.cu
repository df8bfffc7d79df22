// Measured e4m3 tensor-core peak of a B200: back-to-back tcgen05.mma kind::f8f6f4 (SASS: UTCQMMA) with both operands
// resident in shared memory and the accumulator in tensor memory -- no loads, no epilogue.  One CTA per SM, one
// elected thread issues.  Shapes: the matcher's (M=128, N=256, K=32, f16 accumulate) and the same with fp32
// accumulate; cta_group::1.  This is the denominator bench.py uses for the tensor-core matcher's roofline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp8_mma_peak fp8_mma_peak.cu && ./fp8_mma_peak
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {   // K-major, no swizzle: LBO 128 B, SBO 2048 B
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(2048u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}

constexpr int M = 128, N = 256, KB = 256;          // one MMA consumes K = 32 bytes; 8 MMAs walk a 256-byte row
constexpr int SMEM = M * KB + N * KB + 64;

template <int F32ACC>
__global__ void __launch_bounds__(64, 1) peak_kernel(int iters, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + M * KB + N * KB);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < (M + N) * KB / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;  // +1.0 e4m3
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    // idesc: D format bits 4-5 (0 = f16, 1 = f32), A = B = e4m3 (0), K-major, N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = ((uint32_t)F32ACC << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (threadIdx.x < 32) {
        const uint64_t a = smem_desc(smem_u32(smem)), b = smem_desc(smem_u32(smem + M * KB));
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < KB / 32; ++k)
                    mma_f8(tmem + (uint32_t)((it & 1) * N), a + (uint64_t)(k * 16), b + (uint64_t)(k * 16), idesc, k > 0);
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        __syncwarp();
        uint32_t ok = 0, spins = 0;
        while (!ok) {
            if (++spins > (1u << 28)) __trap();   // bounded: a protocol error must not hang the device
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
    if (threadIdx.x == 0 && sink) sink[blockIdx.x] = (unsigned long long)tmem;
}

template <int F32ACC>
void run(const char* name, int sms, int iters) {
    cudaFuncSetAttribute(peak_kernel<F32ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    peak_kernel<F32ACC><<<sms, 64, SMEM>>>(iters / 8, nullptr);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        peak_kernel<F32ACC><<<sms, 64, SMEM>>>(iters, nullptr);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double flop = (double)sms * iters * (KB / 32) * 2.0 * M * N * 32;
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"tflops\": %.1f, \"frac_of_nominal_4500\": %.3f, \"error\": \"%s\"}\n", name, best,
           flop / (best * 1e-3) / 1e12, flop / (best * 1e-3) / 1e12 / 4500.0, cudaGetErrorString(err));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    run<0>("tcgen05.mma kind::f8f6f4 e4m3 x e4m3, M=128 N=256 K=32, f16 accumulate, cta_group::1", p.multiProcessorCount, 20000);
    run<1>("tcgen05.mma kind::f8f6f4 e4m3 x e4m3, M=128 N=256 K=32, f32 accumulate, cta_group::1", p.multiProcessorCount, 20000);
    return 0;
}
