// Micro-benchmark of the integer pipes the matcher lives on ("confirm by micro-benchmark"):
// POPC (XU pipe), LOP3 (ALU pipe), VIMNMX (ALU pipe), IMAD (FMA pipe) throughput in lanes/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu && ./int_pipes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITERS = 4096, UNROLL = 8;

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t seed, uint32_t mul) {
    uint32_t a[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) a[i] = seed + threadIdx.x * 2654435761u + i;
    uint32_t b = seed ^ 0x9e3779b9u, c = seed * 3u + 1u;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
            if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == 2) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(mul), "r"(c));
            if (OP == 4) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        }
        b += 0x01010101u;  // keeps the min chain from being hoisted
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
double run(const char* name, int sms, double mhz) {
    uint32_t* out;
    const int blocks = sms * 2, threads = 1024;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 1u, 3u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<OP><<<blocks, threads>>>(out, r, 3u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 5.0 * blocks * threads * (double)ITERS * UNROLL;
    const double per_clk_sm = ops / (ms * 1e-3) / (mhz * 1e6) / sms;
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"Gops_s\": %.1f, \"lanes_per_clk_per_sm_at_max_clock\": %.2f}\n", name, ms,
           ops / ms / 1e6, per_clk_sm);
    cudaFree(out);
    return per_clk_sm;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1e3;
    printf("{\"device\": \"%s\", \"sms\": %d, \"max_sm_mhz\": %.0f}\n", p.name, p.multiProcessorCount, mhz);
    run<0>("POPC (XU pipe)", p.multiProcessorCount, mhz);
    run<1>("LOP3 (ALU pipe)", p.multiProcessorCount, mhz);
    run<2>("VIMNMX.U32 (ALU pipe)", p.multiProcessorCount, mhz);
    run<3>("IMAD (FMA pipe)", p.multiProcessorCount, mhz);
    run<4>("VIMNMX.U16x2 (ALU pipe)", p.multiProcessorCount, mhz);
    return 0;
}
