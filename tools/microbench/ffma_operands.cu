// FFMA issue rate as a function of how many FRESH register operands an instruction reads (B200, sm_100a).
// Hypothesis under test: an SM sub-partition's register file feeds about two fresh 32-bit operands per clock;
// a third has to come from the operand-reuse cache, a uniform register, a constant or an immediate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_operands ffma_operands.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int ITERS = 2048, N = 12;

template <int OP>
__global__ void __launch_bounds__(256) k(float* out, const float* in, float seed) {
    float a[N], b[N], c[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 32 + i]; c[i] = in[threadIdx.x + 64 + i]; }
    const float x = in[threadIdx.x + 100], y = in[threadIdx.x + 101];
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(c[i]));            // 3 fresh (acc, b_i, c_i)
            if (OP == 1) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(x));               // 2 fresh + x reusable
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(x), "f"(y));                  // 1 fresh + x, y reusable
            if (OP == 3) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[i]) : "f"(b[i]), "f"(c[i]), "f"(seed)); // 2 fresh + kernel-argument constant
            if (OP == 4) asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(a[i]) : "f"(b[i]));                       // square-accumulate: 2 distinct
            if (OP == 5) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[i]) : "f"(b[i]), "f"(c[i]), "f"(b[(i + 1) % N]));   // 3 fresh, no accumulator dependence
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, int sms, double mhz, const float* in) {
    float* out; const int blocks = sms * 8, threads = 256;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, in, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<OP><<<blocks, threads>>>(out, in, 1.0f + r);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 5.0 * blocks * threads * (double)ITERS * N;
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"ffma_lanes_per_clk_per_sm\": %.2f, \"frac_of_128\": %.3f}\n", name, ms,
           ops / (ms * 1e-3) / (mhz * 1e6) / sms, ops / (ms * 1e-3) / (mhz * 1e6) / sms / 128.0);
    cudaFree(out);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* in; cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    const int sms = p.multiProcessorCount; const double mhz = khz / 1e3;
    run<0>("3 fresh register operands (a_i = b_i * c_i + a_i)", sms, mhz, in);
    run<1>("2 fresh + 1 reusable (a_i = b_i * x + a_i)", sms, mhz, in);
    run<2>("1 fresh + 2 reusable (a_i = a_i * x + y)", sms, mhz, in);
    run<3>("2 fresh + constant-bank operand (a_i = b_i * c_i + param)", sms, mhz, in);
    run<4>("2 distinct (a_i = b_i * b_i + a_i)", sms, mhz, in);
    run<5>("3 fresh, independent destination (a_i = b_i * c_i + b_{i+1})", sms, mhz, in);
    return 0;
}
