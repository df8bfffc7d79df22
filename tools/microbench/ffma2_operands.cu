// Packed fp32 FMA (fma.rn.f32x2 -> FFMA2, new with sm_100): FMA lanes per clock per SM as a function of how many
// FRESH 64-bit register-pair operands an instruction reads -- companion of ffma_operands.cu, where a scalar FFMA with
// three fresh registers issues at half rate.  Question: does the pair form deliver 128 FMA/clk/SM with three fresh
// operands (one issue slot per two FMAs, register pairs read as one 64-bit access)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_operands ffma2_operands.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int ITERS = 2048, N = 10;

template <int OP>
__global__ void __launch_bounds__(256) k(float* out, const float* in, float seed) {
    unsigned long long a[N], b[N], c[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        a[i] = reinterpret_cast<const unsigned long long*>(in)[threadIdx.x + i];
        b[i] = reinterpret_cast<const unsigned long long*>(in)[threadIdx.x + 32 + i];
        c[i] = reinterpret_cast<const unsigned long long*>(in)[threadIdx.x + 64 + i];
    }
    const unsigned long long x = reinterpret_cast<const unsigned long long*>(in)[threadIdx.x + 100];
    const unsigned long long y = reinterpret_cast<const unsigned long long*>(in)[threadIdx.x + 101];
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(c[i]));   // 3 fresh pairs
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(x));      // 2 fresh + reusable
            if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x), "l"(y));         // 1 fresh + 2 reusable
            if (OP == 3) asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(a[i]) : "l"(b[i]));              // square-accumulate
            if (OP == 4) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a[i]) : "l"(b[i]), "l"(c[i]), "l"(b[(i + 1) % N]));  // 3 fresh, independent dest
            if (OP == 5) {   // mixed stream: FFMA2 (3 fresh) + one integer-pipe instruction per two, as in the scoring loop
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b[i]), "l"(c[i]));
                if (i & 1) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(reinterpret_cast<unsigned*>(&c[i])[0]) : "r"(0u));
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s ^= a[i] ^ c[i];
    reinterpret_cast<unsigned long long*>(out)[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, int sms, double mhz, const float* in) {
    float* out; const int blocks = sms * 8, threads = 256;
    cudaMalloc(&out, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, in, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<OP><<<blocks, threads>>>(out, in, 1.0f + r);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = 2.0 * 5.0 * blocks * threads * (double)ITERS * N;
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"fma_lanes_per_clk_per_sm\": %.2f, \"frac_of_128\": %.3f}\n", name, ms,
           fmas / (ms * 1e-3) / (mhz * 1e6) / sms, fmas / (ms * 1e-3) / (mhz * 1e6) / sms / 128.0);
    cudaFree(out);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float* in; cudaMalloc(&in, 8192); cudaMemset(in, 0, 8192);
    const int sms = p.multiProcessorCount; const double mhz = khz / 1e3;
    run<0>("FFMA2, 3 fresh pair operands (a_i = b_i * c_i + a_i)", sms, mhz, in);
    run<1>("FFMA2, 2 fresh + 1 reusable (a_i = b_i * x + a_i)", sms, mhz, in);
    run<2>("FFMA2, 1 fresh + 2 reusable (a_i = a_i * x + y)", sms, mhz, in);
    run<3>("FFMA2, 2 distinct (a_i = b_i * b_i + a_i)", sms, mhz, in);
    run<4>("FFMA2, 3 fresh, independent destination", sms, mhz, in);
    run<5>("FFMA2 (3 fresh) + 1 PRMT per two", sms, mhz, in);
    return 0;
}
