// FP32 FMA throughput: scalar FFMA (3 register sources) vs packed fma.rn.f32x2 (sm_100+), lanes/clk/SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr int ITERS = 4096, UNROLL = 8;

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, float seed) {
    float a[UNROLL], b2[UNROLL];
    unsigned long long p[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { a[i] = seed + threadIdx.x + i; b2[i] = seed * 0.5f + i; p[i] = (unsigned long long)__float_as_uint(a[i]) << 32 | __float_as_uint(b2[i]); }
    float x = seed * 1.0001f, y = seed * 0.999f;
    unsigned long long xx = (unsigned long long)__float_as_uint(x) << 32 | __float_as_uint(y), yy = xx ^ 0x0000000100000001ull;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(x), "f"(y));
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(xx), "l"(yy));
            if (OP == 2) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b2[i]), "f"(x));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) s += a[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, int sms, double mhz, int flops_per_op) {
    float* out; const int blocks = sms * 2, threads = 1024;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<OP><<<blocks, threads>>>(out, 1.0f + r);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 5.0 * blocks * threads * (double)ITERS * UNROLL;
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"instr_lanes_per_clk_per_sm\": %.2f, \"TFLOP_s\": %.1f}\n", name, ms,
           ops / (ms * 1e-3) / (mhz * 1e6) / sms, ops * flops_per_op / (ms * 1e-3) / 1e12);
    cudaFree(out);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    run<0>("FFMA d=d*x+y (2 distinct sources + acc)", p.multiProcessorCount, khz / 1e3, 2);
    run<2>("FFMA d=b*x+d (3 distinct sources)", p.multiProcessorCount, khz / 1e3, 2);
    run<1>("FFMA2 fma.rn.f32x2", p.multiProcessorCount, khz / 1e3, 4);
    return 0;
}
