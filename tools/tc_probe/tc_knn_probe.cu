// Probe: 256-bit Hamming kNN(k=2) on the 5th-gen tensor cores.  bits -> e4m3 (+1/-1), D = Q.T^T via tcgen05.mma
// kind::f8f6f4 (fp32 accumulate, exact), hamming = (256 - dot) / 2, top-2 in the TMEM epilogue.
// Standalone (own main) so the layout/descriptor assumptions can be checked before the kernel goes into the library.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

namespace tc {
constexpr int QB = 128;            // query rows per MMA (M)
constexpr int QSB = 256;           // query rows per CTA item (two accumulators)
constexpr int TB = 256;            // train rows per tile (N)
constexpr int KBYTES = 256;        // bytes per unpacked descriptor
constexpr int A_BYTES = QSB * KBYTES;   // 64 KB
constexpr int B_BYTES = TB * KBYTES;    // 64 KB
constexpr int STAGES = 2;
constexpr int THREADS = 32 * 10;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TB >> 3) << 17) | ((uint32_t)(QB >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
    long long spins = 0;
    while (!mbar_try(bar, parity)) {
        if (++spins > (1ll << 24)) { printf("mbar timeout tag %d block %d thread %d\n", tag, blockIdx.x, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    // K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; LBO (K-adjacent core matrices) = 128 B,
    // SBO (next 8-row group) = 2048 B; version 1 (sm_100)
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(2048u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_f8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// bits -> e4m3 +1 (0x38) / -1 (0xB8), written in the MMA's no-swizzle K-major core-matrix order:
// out[(row / 8)][chunk 0..15][row % 8][16 B]; one thread = one 16-byte piece (consecutive threads, consecutive pieces).
__global__ void unpack_kernel(const uint8_t* __restrict__ des, uint8_t* __restrict__ out, int n, int n_pad, int frames) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_frame = (long long)n_pad * 16;
    if (t >= per_frame * frames) return;
    const int f = (int)(t / per_frame);
    const int p = (int)(t % per_frame);
    const int r = p & 7, c = (p >> 3) & 15, g = p >> 7;
    const int row = g * 8 + r;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (row < n) {
        const uint8_t* d = des + ((size_t)f * n + row) * 32 + 2 * c;
        const uint32_t bits = (uint32_t)d[0] | ((uint32_t)d[1] << 8);
        auto ex = [](uint32_t b4) { return 0xB8B8B8B8u ^ ((((b4 & 0xFu) * 0x00204081u) & 0x01010101u) << 7); };
        o = make_uint4(ex(bits), ex(bits >> 4), ex(bits >> 8), ex(bits >> 12));
    }
    reinterpret_cast<uint4*>(out)[t] = o;
}

struct Params {
    const uint8_t* uq;   // unpacked query frames  [frames][n_pad][256]
    const uint8_t* ut;   // unpacked train frames
    int n_pad, nq, nt, batch;
    int32_t* idx;        // [batch][nq][2]
    int32_t* dist;       // [batch][nq][2]
};

__global__ void __launch_bounds__(THREADS, 1) knn_tc_kernel(Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_smem = smem;                         // 64 KB: two 128-row halves
    uint8_t* b_smem = smem + A_BYTES;               // STAGES x 64 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + STAGES * B_BYTES);
    uint64_t* full = bars;              // [STAGES]
    uint64_t* empty = bars + STAGES;    // [STAGES]
    uint64_t* a_full = bars + 2 * STAGES;
    uint64_t* a_empty = a_full + 1;
    uint64_t* acc_full = a_empty + 1;   // [2]
    uint64_t* acc_empty = acc_full + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qsbs = (p.nq + QSB - 1) / QSB;
    const int items = p.batch * qsbs;
    const int ttiles = (p.nt + TB - 1) / TB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int h = 0; h < 2; ++h) { mbar_init(acc_full + h, 1); mbar_init(acc_empty + h, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t step = 0, it = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
                const int b = item / qsbs, sb = item % qsbs;
                mbar_wait(a_empty, (it & 1) ^ 1, 1);
                mbar_expect_tx(a_full, A_BYTES);
                tma_load_1d(a_smem, p.uq + ((size_t)b * p.n_pad + (size_t)sb * QSB) * KBYTES, A_BYTES, a_full);
                for (int j = 0; j < ttiles; ++j, ++step) {
                    const uint32_t s = step % STAGES, ph = (step / STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1, 2);
                    mbar_expect_tx(full + s, B_BYTES);
                    tma_load_1d(b_smem + s * B_BYTES, p.ut + ((size_t)b * p.n_pad + (size_t)j * TB) * KBYTES, B_BYTES, full + s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t step = 0, it = 0, use = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
                mbar_wait(a_full, it & 1, 3);
                for (int j = 0; j < ttiles; ++j, ++step, ++use) {
                    const uint32_t s = step % STAGES, ph = (step / STAGES) & 1;
                    mbar_wait(full + s, ph, 4);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(b_smem + s * B_BYTES);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(acc_empty + h, (use & 1) ^ 1, 5);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(a_smem + h * QB * KBYTES);
#pragma unroll
                        for (int k = 0; k < KBYTES / 32; ++k)
                            mma_f8(tmem_base + h * TB, smem_desc(a_addr + k * 256), smem_desc(b_addr + k * 256), IDESC, k > 0);
                        mma_commit(acc_full + h);
                    }
                    mma_commit(empty + s);
                }
                mma_commit(a_empty);
            }
        }
    } else {
        const int e = warp - 2, h = e >> 2, quarter = warp & 3;
        const int row_in_sb = h * QB + quarter * 32 + lane;
        uint32_t use = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int b = item / qsbs, sb = item % qsbs;
            uint32_t g1 = 0xFFFFFFFFu, g2 = 0xFFFFFFFFu;       // dist << 23 | train idx
            for (int j = 0; j < ttiles; ++j, ++use) {
                mbar_wait(acc_full + h, use & 1, 6);
                tc_fence_after();
                const int cols = min(TB, p.nt - j * TB);
                for (int c0 = 0; c0 < cols; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * TB + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (c0 + i < cols) {
                            const int dot = (int)__uint_as_float(v[i]);
                            const uint32_t key = ((uint32_t)((256 - dot) >> 1) << 23) | (uint32_t)(j * TB + c0 + i);
                            g2 = min(g2, max(g1, key));
                            g1 = min(g1, key);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + h);
            }
            const int q = sb * QSB + row_in_sb;
            if (q < p.nq) {
                const size_t o = ((size_t)b * p.nq + q) * 2;
                p.idx[o] = g1 == 0xFFFFFFFFu ? -1 : (int)(g1 & 0x7FFFFFu);
                p.dist[o] = g1 == 0xFFFFFFFFu ? -1 : (int)(g1 >> 23);
                p.idx[o + 1] = g2 == 0xFFFFFFFFu ? -1 : (int)(g2 & 0x7FFFFFu);
                p.dist[o + 1] = g2 == 0xFFFFFFFFu ? -1 : (int)(g2 >> 23);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}
constexpr int SMEM_BYTES = A_BYTES + STAGES * B_BYTES + 256;
}  // namespace tc

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 2048;
    const int batch = argc > 2 ? atoi(argv[2]) : 4;
    const int iters = argc > 3 ? atoi(argv[3]) : 5;
    const int n_pad = (n + 255) / 256 * 256;
    std::vector<uint8_t> q((size_t)batch * n * 32), t((size_t)batch * n * 32);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (auto& x : t) x = (uint8_t)rnd();
    for (size_t i = 0; i < q.size(); ++i) q[i] = (uint8_t)rnd();
    // make most queries noisy copies of a train row so that small distances and ties exist
    for (int b = 0; b < batch; ++b)
        for (int i = 0; i < n; i += 2) {
            const int src = (int)(rnd() % n);
            memcpy(&q[((size_t)b * n + i) * 32], &t[((size_t)b * n + src) * 32], 32);
            for (int f = 0; f < (int)(rnd() % 40); ++f) q[((size_t)b * n + i) * 32 + rnd() % 32] ^= (uint8_t)(1u << (rnd() % 8));
        }
    for (int b = 0; b < batch; ++b)   // duplicate train rows: forced ties
        for (int i = 0; i < 16 && i + 100 < n; ++i) memcpy(&t[((size_t)b * n + i + 100) * 32], &t[((size_t)b * n + i) * 32], 32);
    uint8_t *dq, *dt, *uq, *ut; int32_t *didx, *ddist;
    CK(cudaMalloc(&dq, q.size())); CK(cudaMalloc(&dt, t.size()));
    CK(cudaMalloc(&uq, (size_t)batch * n_pad * 256)); CK(cudaMalloc(&ut, (size_t)batch * n_pad * 256));
    CK(cudaMalloc(&didx, (size_t)batch * n * 2 * 4)); CK(cudaMalloc(&ddist, (size_t)batch * n * 2 * 4));
    CK(cudaMemcpy(dq, q.data(), q.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, t.data(), t.size(), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(tc::knn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    tc::Params p{uq, ut, n_pad, n, n, batch, didx, ddist};
    const int items = batch * ((n + 255) / 256);
    const int grid = items < sms ? items : sms;
    const long long pieces = (long long)batch * n_pad * 16;
    auto run = [&]() {
        tc::unpack_kernel<<<(unsigned)((pieces + 255) / 256), 256>>>(dq, uq, n, n_pad, batch);
        tc::unpack_kernel<<<(unsigned)((pieces + 255) / 256), 256>>>(dt, ut, n, n_pad, batch);
        tc::knn_tc_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES>>>(p);
    };
    run();
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> idx((size_t)batch * n * 2), dist((size_t)batch * n * 2);
    CK(cudaMemcpy(idx.data(), didx, idx.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dist.data(), ddist, dist.size() * 4, cudaMemcpyDeviceToHost));
    // CPU check on the first and last pair
    long long bad = 0;
    for (int b : {0, batch - 1}) {
        for (int i = 0; i < n; ++i) {
            uint32_t m1 = 0xFFFFFFFFu, m2 = 0xFFFFFFFFu;
            const uint64_t* a = reinterpret_cast<const uint64_t*>(&q[((size_t)b * n + i) * 32]);
            for (int j = 0; j < n; ++j) {
                const uint64_t* c = reinterpret_cast<const uint64_t*>(&t[((size_t)b * n + j) * 32]);
                const uint32_t d = __builtin_popcountll(a[0] ^ c[0]) + __builtin_popcountll(a[1] ^ c[1]) +
                                   __builtin_popcountll(a[2] ^ c[2]) + __builtin_popcountll(a[3] ^ c[3]);
                const uint32_t key = (d << 23) | (uint32_t)j;
                if (key < m1) { m2 = m1; m1 = key; } else if (key < m2) m2 = key;
            }
            const size_t o = ((size_t)b * n + i) * 2;
            const bool ok = idx[o] == (int)(m1 & 0x7FFFFF) && dist[o] == (int)(m1 >> 23) &&
                            idx[o + 1] == (int)(m2 & 0x7FFFFF) && dist[o + 1] == (int)(m2 >> 23);
            if (!ok && bad++ < 8)
                printf("MISMATCH b %d q %d: gpu (%d,%d)(%d,%d) cpu (%d,%d)(%d,%d)\n", b, i, idx[o], dist[o], idx[o + 1], dist[o + 1],
                       (int)(m1 & 0x7FFFFF), (int)(m1 >> 23), (int)(m2 & 0x7FFFFF), (int)(m2 >> 23));
        }
    }
    printf("checked 2 pairs of %d x %d: %lld mismatching rows\n", n, n, bad);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    float ms_all = 0, ms_mm = 0;
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) run();
    cudaEventRecord(e1);
    for (int i = 0; i < iters; ++i) tc::knn_tc_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES>>>(p);
    cudaEventRecord(e2);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms_all, e0, e1); cudaEventElapsedTime(&ms_mm, e1, e2);
    const double cmp = (double)batch * n * n;
    printf("n %d batch %d: unpack+match %.3f ms (%.1f Gcmp/s), match only %.3f ms (%.1f Gcmp/s)\n", n, batch, ms_all / iters,
           cmp / (ms_all / iters) * 1e-6, ms_mm / iters, cmp / (ms_mm / iters) * 1e-6);
    return bad ? 1 : 0;
}
