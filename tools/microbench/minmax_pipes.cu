// Which pipe do the packed 16-bit min / max instructions of the matcher's epilogue run on, and how fast?  (B200)
//   VIMNMX.U16x2 (__vminu2 / __vmaxu2), VIMNMX3.U16x2 (__vimin3_u16x2), HMNMX2 (__hmin2 / __hmax2), and the two kinds
//   interleaved -- if they run on different pipes the mix costs less than the sum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o minmax_pipes minmax_pipes.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
constexpr int ITERS = 4096, N = 8;

template <int OP>
__global__ void __launch_bounds__(256) k(unsigned* out, const unsigned* in) {
    unsigned a[N], b[N];
    __half2 h[N], g[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 64 + i];
        h[i] = *reinterpret_cast<const __half2*>(&in[threadIdx.x + 128 + i]); g[i] = *reinterpret_cast<const __half2*>(&in[threadIdx.x + 192 + i]);
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (OP == 0) { a[i] = __vminu2(a[i], b[i]); b[i] = __vmaxu2(b[i], a[(i + 1) % N]); }
            if (OP == 1) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) % N]); b[i] = __vimax3_u16x2(b[i], a[i], b[(i + 3) % N]); }
            if (OP == 2) { h[i] = __hmin2(h[i], g[i]); g[i] = __hmax2(g[i], h[(i + 1) % N]); }
            if (OP == 3) { a[i] = __vminu2(a[i], b[i]); h[i] = __hmin2(h[i], g[i]); b[i] = __vmaxu2(b[i], a[(i + 1) % N]); g[i] = __hmax2(g[i], h[(i + 1) % N]); }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s ^= a[i] ^ b[i] ^ *reinterpret_cast<unsigned*>(&h[i]) ^ *reinterpret_cast<unsigned*>(&g[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, int sms, double mhz, const unsigned* in, int ops_per_iter) {
    unsigned* out; const int blocks = sms * 8, threads = 256;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, in); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<OP><<<blocks, threads>>>(out, in);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 5.0 * blocks * threads * (double)ITERS * N * ops_per_iter;
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"instr_lanes_per_clk_per_sm\": %.2f}\n", name, ms, ops / (ms * 1e-3) / (mhz * 1e6) / sms);
    cudaFree(out);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    unsigned* in; cudaMalloc(&in, 4096); cudaMemset(in, 0x11, 4096);
    const int sms = p.multiProcessorCount; const double mhz = khz / 1e3;
    run<0>("VIMNMX.U16x2 (vminu2 + vmaxu2)", sms, mhz, in, 2);
    run<1>("VIMNMX3.U16x2 (vimin3 + vimax3)", sms, mhz, in, 2);
    run<2>("HMNMX2 (hmin2 + hmax2)", sms, mhz, in, 2);
    run<3>("mixed: 2 VIMNMX.U16x2 + 2 HMNMX2 per step", sms, mhz, in, 4);
    return 0;
}
