// 256-bit Hamming kNN (k=2) + Lowe ratio flag on the 5th-generation tensor cores -- the batched form of
// cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(curr_des, last_des, k=2)   (MonocularSlam/slam.py:126-131).
//
// A descriptor bit b becomes the e4m3 value +1 (b = 1) or -1 (b = 0); the dot product of two such rows is
// 256 - 2 * hamming, an integer of magnitude <= 256, so  D = Q . T^T  (tcgen05.mma kind::f8f6f4, f16 accumulators in
// tensor memory) is EXACT and  hamming = (256 - dot) / 2.  Results are bit-identical to the integer-pipe kernel
// (hamming_knn2.cu) and to OpenCV, including the lowest-train-index tie-break.
//
//   unpack_train_kernel   train descriptors -> e4m3 rows in the MMA's K-major, no-swizzle core-matrix order
//                         [row / 8][16-byte K chunk][row % 8], so a 256-row tile is one contiguous 64 KB TMA copy.
//   knn2_tc_kernel        persistent, warp-specialised, one CTA per SM.  Item = (pair, 256 query rows):
//                           warp 0      TMA producer: train tiles into a 2-stage shared-memory ring
//                           warp 1      MMA issuer: per tile 2 x 8 tcgen05.mma (M=128, N=256, K=32) into the two
//                                       accumulators (2 x 256 TMEM columns), tcgen05.commit -> mbarriers
//                           warps 2..9  epilogue, 4 warps per accumulator (thread = query row = TMEM lane):
//                                       expand their query rows into the A operand in shared memory (no unpacked
//                                       query buffer), then per tile tcgen05.ld (16-bit packed) -> hand the
//                                       accumulator back -> top-2 on packed 16-bit keys in registers.
//   Epilogue arithmetic per TWO train columns (one register holds their two f16 dots):
//       HFMA2  1152 - dot/2 = 1024 + d          -> f16 bit pattern 0x6400 | d   (d = Hamming distance 0..256)
//       IMAD   bits * 128 + C_p                  -> per 16-bit lane (d << 7) | p, p = column pair 0..127 of the tile
//                                                   (the 0x6400 terms overflow into constants that C_p cancels)
//       VIMNMX.U16x2 / VIMNMX3.U16x2              -> tile-local top-2 of both lanes, four independent lists
//   and once per tile the four 16-bit candidates are folded into the 32-bit keys  dist << 23 | train index.
#include <cuda_fp16.h>

#include <type_traits>

#include "tvg_common.cuh"
#include "tvg_kernels.h"

namespace tvg {
namespace {

constexpr int TC_QB = 128;                    // query rows per MMA (M)
constexpr int TC_NQB = 2;                     // query blocks (accumulators) per item
constexpr int TC_QI = TC_NQB * TC_QB;         // query rows per item
constexpr int TC_TN = 256;                    // train rows per tile (MMA N)
constexpr int TC_KB = 256;                    // bytes per e4m3 row
constexpr int TC_A_BLK = TC_QB * TC_KB;       // 32 KB
constexpr int TC_B_TILE = TC_TN * TC_KB;      // 64 KB
constexpr int TC_NST = 2;
#ifndef TC_LISTS
#define TC_LISTS 4   // independent top-2 lists per thread in the epilogue (instruction-level parallelism)
#endif
constexpr int TC_THREADS = 32 * (2 + 4 * TC_NQB);
constexpr int TC_SMEM = TC_NQB * TC_A_BLK + TC_NST * TC_B_TILE + 256;
// instruction descriptor: D = f16 (bits 4-5 = 0), A = B = e4m3 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t TC_IDESC = ((uint32_t)(TC_TN >> 3) << 17) | ((uint32_t)(TC_QB >> 4) << 24);

__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool tc_mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a protocol error traps (the host sees a launch failure) instead of hanging the device
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    unsigned int spins = 0;
    while (!tc_mbar_try(bar, parity))
        if (++spins > (1u << 28)) __trap();
}
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// D[tmem] (+)= A[smem] . B[smem]^T, e4m3 x e4m3
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 64 columns of f16 accumulators -> 32 registers, two adjacent columns per register
__device__ __forceinline__ void tc_ld64_pack16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 4 descriptor bits -> 4 e4m3 bytes: +1 = 0x38, -1 = 0xB8
__device__ __forceinline__ uint32_t tc_expand4(uint32_t b4) {
    return 0xB8B8B8B8u ^ ((((b4 & 0xFu) * 0x00204081u) & 0x01010101u) << 7);
}
// shared-memory operand descriptor: K-major, no swizzle; a core matrix is 8 rows x 16 B contiguous, the K-adjacent
// core matrix 128 B further (leading byte offset), the next 8-row group 2048 B further (stride byte offset)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(2048u >> 4) << 32) | (1ull << 46);
}

struct TcArgs {
    KnnParams k;
    uint8_t* unpacked;     // [batch][n_pad][256] e4m3 train rows in core-matrix order
    int n_pad, batch, qis, items;
    int splits, tiles_per_split;   // train tiles of a pair are cut into `splits` ranges (one item each) when the
                                   // (pair, query block) items alone cannot fill the SMs; partial top-2s are merged
                                   // by knn2_tc_merge_kernel
};

__device__ __forceinline__ void tc_counts(const TcArgs& a, int b, int& fq, int& ft, int& nq, int& nt) {
    fq = a.k.pair_q ? a.k.pair_q[b] : b;
    ft = a.k.pair_t ? a.k.pair_t[b] : b;
    nq = a.k.q_count ? min(a.k.q_count[fq], a.k.n_max) : a.k.nq;
    nt = a.k.t_count ? min(a.k.t_count[ft], a.k.n_max) : a.k.nt;
}

// one thread = four 16-byte pieces, 256 pieces apart (consecutive threads write consecutive pieces; four independent
// load -> expand -> store chains per thread keep more bytes in flight)
constexpr int UNPACK_PER_THREAD = 4;

__global__ void __launch_bounds__(256) unpack_train_kernel(TcArgs a) {
    const int b = blockIdx.y;
    int fq, ft, nq, nt;
    tc_counts(a, b, fq, ft, nq, nt);
    const int rows = (nt + TC_TN - 1) / TC_TN * TC_TN;       // only the tiles the matcher will load
    const int piece0 = blockIdx.x * (256 * UNPACK_PER_THREAD) + threadIdx.x;
    uint32_t bits[UNPACK_PER_THREAD];
#pragma unroll
    for (int u = 0; u < UNPACK_PER_THREAD; ++u) {
        const int piece = piece0 + u * 256;
        const int r = piece & 7, c = (piece >> 3) & 15, row = (piece >> 7) * 8 + r;
        bits[u] = 0x10000u;   // marker: zero row
        if (piece < rows * 16 && row < nt)
            bits[u] = (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(a.k.t_base + ((size_t)ft * a.k.n_max + row) * 32 + 2 * c));
    }
    uint4* out = reinterpret_cast<uint4*>(a.unpacked + (size_t)b * a.n_pad * TC_KB);
#pragma unroll
    for (int u = 0; u < UNPACK_PER_THREAD; ++u) {
        const int piece = piece0 + u * 256;
        if (piece >= rows * 16) continue;
        const uint32_t v = bits[u];
        out[piece] = (v & 0x10000u) ? make_uint4(0, 0, 0, 0)
                                    : make_uint4(tc_expand4(v), tc_expand4(v >> 4), tc_expand4(v >> 8), tc_expand4(v >> 12));
    }
}

__device__ __forceinline__ void tc_emit(const KnnParams& p, size_t o, uint32_t k1, uint32_t k2) {
    int2 id, di;
    id.x = (k1 == TVG_SENTINEL_KEY) ? -1 : (int)(k1 & TVG_KEY_IDX_MASK);
    di.x = (k1 == TVG_SENTINEL_KEY) ? -1 : (int)(k1 >> TVG_KEY_SHIFT);
    id.y = (k2 == TVG_SENTINEL_KEY) ? -1 : (int)(k2 & TVG_KEY_IDX_MASK);
    di.y = (k2 == TVG_SENTINEL_KEY) ? -1 : (int)(k2 >> TVG_KEY_SHIFT);
    reinterpret_cast<int2*>(p.idx)[o] = id;
    reinterpret_cast<int2*>(p.dist)[o] = di;
    if (p.keep) p.keep[o] = (k2 != TVG_SENTINEL_KEY) && ((double)di.x < p.ratio * (double)di.y);
}

__global__ void __launch_bounds__(TC_THREADS, 1) knn2_tc_kernel(TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + TC_NQB * TC_A_BLK;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_NQB * TC_A_BLK + TC_NST * TC_B_TILE);
    uint64_t* full = bars;                       // [NST]  TMA tile landed
    uint64_t* empty = full + TC_NST;             // [NST]  MMAs that read the tile are complete
    uint64_t* a_full = empty + TC_NST;           // [NQB]  query block written (4 epilogue warps)
    uint64_t* acc_full = a_full + TC_NQB;        // [NQB]  accumulator complete
    uint64_t* acc_empty = acc_full + TC_NQB;     // [NQB]  accumulator read back (4 epilogue warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + TC_NQB);

    const KnnParams& p = a.k;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_NST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int h = 0; h < TC_NQB; ++h) { mbar_init(a_full + h, 4); mbar_init(acc_full + h, 1); mbar_init(acc_empty + h, 4); }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // An item is live when it has query rows and the pair has train rows; every role walks the same item sequence.
    auto item_info = [&](int item, int& b, int& qi, int& fq, int& nq, int& nt, int& j0, int& j1) -> bool {
        const int sp = item % a.splits;
        const int rest = item / a.splits;
        b = rest / a.qis;
        qi = rest % a.qis;
        int ft;
        tc_counts(a, b, fq, ft, nq, nt);
        const int ttiles = (nt + TC_TN - 1) / TC_TN;
        j0 = sp * a.tiles_per_split;
        j1 = min(ttiles, j0 + a.tiles_per_split);
        return qi * TC_QI < nq && j0 < j1;
    };

    if (warp == 0) {
        uint32_t step = 0;
        for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
            int b, qi, fq, nq, nt, j0, j1;
            if (!item_info(item, b, qi, fq, nq, nt, j0, j1)) continue;
            for (int j = j0; j < j1; ++j, ++step) {
                const uint32_t s = step % TC_NST, ph = (step / TC_NST) & 1;
                tc_mbar_wait(empty + s, ph ^ 1);
                if (tc_elect_one()) {
                    mbar_arrive_expect_tx(full + s, TC_B_TILE);
                    tma_load_1d(b_smem + s * TC_B_TILE, a.unpacked + ((size_t)b * a.n_pad + (size_t)j * TC_TN) * TC_KB, TC_B_TILE,
                                full + s);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        uint32_t step = 0, it = 0;
        for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
            int b, qi, fq, nq, nt, j0, j1;
            if (!item_info(item, b, qi, fq, nq, nt, j0, j1)) continue;
            for (int j = j0; j < j1; ++j, ++step) {
                const uint32_t s = step % TC_NST, ph = (step / TC_NST) & 1;
                tc_mbar_wait(full + s, ph);
                tc_fence_after();
                const uint64_t b_desc = tc_smem_desc(smem_u32(b_smem + s * TC_B_TILE));
#pragma unroll
                for (int h = 0; h < TC_NQB; ++h) {
                    if (j == j0) tc_mbar_wait(a_full + h, it & 1);
                    tc_mbar_wait(acc_empty + h, (step & 1) ^ 1);
                    tc_fence_after();
                    const uint64_t a_desc = tc_smem_desc(smem_u32(a_smem + h * TC_A_BLK));
                    if (tc_elect_one()) {
#pragma unroll
                        for (int k = 0; k < TC_KB / 32; ++k)   // 32 K-bytes = two core matrices = 256 B = 16 address units
                            tc_mma_f8(tmem_base + h * TC_TN, a_desc + (uint64_t)(k * 16), b_desc + (uint64_t)(k * 16), TC_IDESC, k > 0);
                        tc_commit(acc_full + h);
                        if (h == TC_NQB - 1) tc_commit(empty + s);
                    }
                    __syncwarp();
                }
            }
            ++it;
        }
    } else {
        const int e = warp - 2, h = e >> 2, quarter = warp & 3;   // a warp reads the TMEM lanes 32 * (warp % 4) ...
        const int row_in_blk = quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * TC_TN);
        // query row -> e4m3, into this block's shared-memory operand: [row / 8][16-byte K chunk][row % 8]
        auto write_a = [&](int fq, int nq, int qi) {
            const int q = qi * TC_QI + h * TC_QB + row_in_blk;
            uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
            const bool live = q < nq;
            if (live) {
                const uint4* src = reinterpret_cast<const uint4*>(p.q_base + ((size_t)fq * p.n_max + q) * 32);
                d0 = __ldg(src);
                d1 = __ldg(src + 1);
            }
            const uint32_t w[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            uint8_t* dst = a_smem + h * TC_A_BLK + (row_in_blk >> 3) * 2048 + (row_in_blk & 7) * 16;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const uint32_t bits = w[c >> 1] >> ((c & 1) * 16);
                uint4 o = make_uint4(0, 0, 0, 0);
                if (live) o = make_uint4(tc_expand4(bits), tc_expand4(bits >> 4), tc_expand4(bits >> 8), tc_expand4(bits >> 12));
                *reinterpret_cast<uint4*>(dst + c * 128) = o;
            }
            fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(a_full + h);
        };
        auto next_live = [&](int item, int& fq, int& nq, int& qi) -> bool {
            for (item += gridDim.x; item < a.items; item += gridDim.x) {
                int b, nt, j0, j1;
                if (item_info(item, b, qi, fq, nq, nt, j0, j1)) return true;
            }
            return false;
        };
        uint32_t use = 0;
        {
            int fq, nq, qi;
            if (next_live((int)blockIdx.x - (int)gridDim.x, fq, nq, qi)) write_a(fq, nq, qi);
        }
        for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
            int b, qi, fq, nq, nt, j0, j1;
            const bool live = item_info(item, b, qi, fq, nq, nt, j0, j1);
            uint32_t g1 = TVG_SENTINEL_KEY, g2 = TVG_SENTINEL_KEY;       // dist << 23 | train idx
            if (!live) j1 = j0;
            for (int j = j0; j < j1; ++j, ++use) {
                tc_mbar_wait(acc_full + h, use & 1);
                tc_fence_after();
                uint32_t v[4][32];
#pragma unroll
                for (int c = 0; c < 4; ++c) tc_ld64_pack16(taddr + c * 64, v[c]);
                tc_ld_wait();
                // the accumulator is in registers now: hand it back to the MMA warp before the arithmetic
                tc_fence_before();
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(acc_empty + h);
                // last tile of the item: every MMA that read this A block is complete -> stage the next item's rows
                if (j == j1 - 1) {
                    int fq2, nq2, qi2;
                    if (next_live(item, fq2, nq2, qi2)) write_a(fq2, nq2, qi2);
                }
                uint32_t m1[TC_LISTS], m2[TC_LISTS];
#pragma unroll
                for (int l = 0; l < TC_LISTS; ++l) m1[l] = m2[l] = 0xFFFFFFFFu;
                const int cols = min(TC_TN, nt - j * TC_TN);
                auto chunk = [&](auto masked) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            uint32_t key[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const uint32_t pl = (uint32_t)(c * 32 + i + u);
                                const __half2 y = __hfma2(*reinterpret_cast<const __half2*>(&v[c][i + u]), __float2half2_rn(-0.5f),
                                                          __float2half2_rn(1152.0f));
                                key[u] = *reinterpret_cast<const uint32_t*>(&y) * p.w128 + (pl | (((pl - 0x32u) & 0xFFFFu) << 16));
                                if (decltype(masked)::value) {
                                    if (2 * (int)pl >= cols) key[u] |= 0x0000FFFFu;
                                    if (2 * (int)pl + 1 >= cols) key[u] |= 0xFFFF0000u;
                                }
                            }
                            // two keys per update: 5 min/max (one 3-input) instead of 6
                            const int l = (i >> 1) & (TC_LISTS - 1);
                            const uint32_t lo = __vminu2(key[0], key[1]), hi = __vmaxu2(key[0], key[1]);
                            m2[l] = __vimin3_u16x2(m2[l], hi, __vmaxu2(m1[l], lo));
                            m1[l] = __vminu2(m1[l], lo);
                        }
                };
                if (cols >= TC_TN) chunk(std::false_type{}); else chunk(std::true_type{});
                // fold the four lists, then the four 16-bit candidates into the 32-bit global keys
#pragma unroll
                for (int l = 1; l < TC_LISTS; ++l) {
                    const uint32_t hi = __vmaxu2(m1[0], m1[l]);
                    m1[0] = __vminu2(m1[0], m1[l]);
                    m2[0] = __vminu2(hi, __vminu2(m2[0], m2[l]));
                }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const uint32_t k16 = (((w & 2) ? m2[0] : m1[0]) >> ((w & 1) * 16)) & 0xFFFFu;
                    if (k16 != 0xFFFFu) {
                        const uint32_t key = ((k16 >> 7) << TVG_KEY_SHIFT) | (uint32_t)(j * TC_TN + 2 * (k16 & 127u) + (w & 1));
                        top2_insert(g1, g2, key);
                    }
                }
            }
            const int q = qi * TC_QI + h * TC_QB + row_in_blk;
            if (q < nq) {
                if (a.splits == 1) tc_emit(p, (size_t)b * p.n_max + q, g1, g2);
                else p.partial[((size_t)b * p.n_max + q) * a.splits + item % a.splits] = make_uint2(g1, g2);
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

// split-Nt: fold the per-split partial top-2s of one query (thread = query) and emit
__global__ void __launch_bounds__(256) knn2_tc_merge_kernel(TcArgs a) {
    const int b = blockIdx.y, q = blockIdx.x * blockDim.x + threadIdx.x;
    int fq, ft, nq, nt;
    tc_counts(a, b, fq, ft, nq, nt);
    if (q >= nq) return;
    const int ttiles = (nt + TC_TN - 1) / TC_TN;
    const int used = (ttiles + a.tiles_per_split - 1) / a.tiles_per_split;   // splits that held tiles for this pair
    uint32_t g1 = TVG_SENTINEL_KEY, g2 = TVG_SENTINEL_KEY;
    const uint2* part = a.k.partial + ((size_t)b * a.k.n_max + q) * a.splits;
    for (int sp = 0; sp < used; ++sp) {
        const uint2 v = part[sp];
        top2_merge(g1, g2, v.x, v.y);
    }
    tc_emit(a.k, (size_t)b * a.k.n_max + q, g1, g2);
}

}  // namespace

// Optional L2-resident mode (l2_budget > 0): the e4m3 train rows are produced and consumed in sub-batches of pairs
// small enough to stay in L2 (one buffer, reused: the next sub-batch overwrites lines that are still dirty in L2), so
// they never travel to HBM and back.  Measured at 1024 pairs of 2048 x 2048 (ncu --cache-control none): 40 MB budget ->
// 134 MB read + 39 MB written per call (= the algorithmic bytes) instead of 671 + 572 MB, for +8 % time (14 sub-batch
// launches: prologue / drain of the persistent kernel per launch); 80 MB no longer stays resident (827 MB).

static int tc_pairs_per_launch(int batch, int n_max, int items_per_pair, int sm_count, size_t l2_budget) {
    if (l2_budget == 0) return batch;
    const size_t per_pair = ((size_t)n_max + TC_TN - 1) / TC_TN * TC_TN * TC_KB;
    long long cap = (long long)(l2_budget / per_pair);
    if (cap < 1) cap = 1;
    if (cap >= batch) return batch;
    // largest sub-batch <= cap whose item count fills whole waves of the persistent grid best
    int best = (int)cap;
    double best_eff = 0.0;
    for (long long s = cap; s >= 1 && s > cap - 2LL * sm_count; --s) {
        const long long items = s * items_per_pair, waves = (items + sm_count - 1) / sm_count;
        const double eff = (double)items / (double)(waves * sm_count);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = (int)s; }
    }
    return best;
}

size_t knn2_tc_workspace_bytes(int batch, int n_max, size_t l2_budget) {
    const size_t per_pair = ((size_t)n_max + TC_TN - 1) / TC_TN * TC_TN * TC_KB;
    if (l2_budget == 0) return (size_t)batch * per_pair;
    long long cap = (long long)(l2_budget / per_pair);
    if (cap < 1) cap = 1;
    if (cap > batch) cap = batch;
    return (size_t)cap * per_pair;
}

int knn2_tc_splits(int batch, int q_rows, int n_max, int sm_count) {
    const int qis = (q_rows + TC_QI - 1) / TC_QI > 0 ? (q_rows + TC_QI - 1) / TC_QI : 1;
    const long long base = (long long)batch * qis;
    const int ttiles = (n_max + TC_TN - 1) / TC_TN;
    long long splits = base >= sm_count ? 1 : sm_count / base;
    if (splits > ttiles) splits = ttiles;
    if (splits < 1) splits = 1;
    const int tps = (ttiles + (int)splits - 1) / (int)splits;
    return tps > 0 ? (ttiles + tps - 1) / tps : 1;
}

// p.partial must hold batch * n_max * splits uint2 when splits = knn2_tc_splits(...) > 1
cudaError_t launch_knn2_tc(const KnnParams& p, int batch, int q_rows, uint8_t* workspace, int sm_count, size_t l2_budget,
                           int* launches, cudaStream_t stream) {
    TcArgs a{};
    a.unpacked = workspace;
    a.n_pad = (p.n_max + TC_TN - 1) / TC_TN * TC_TN;
    a.qis = (q_rows + TC_QI - 1) / TC_QI;
    if (a.qis < 1) a.qis = 1;
    a.splits = knn2_tc_splits(batch, q_rows, p.n_max, sm_count);
    const int ttiles = (p.n_max + TC_TN - 1) / TC_TN;
    a.tiles_per_split = (ttiles + a.splits - 1) / a.splits;
    if (a.tiles_per_split < 1) a.tiles_per_split = 1;
    *launches = 0;
    const int t_rows = p.t_count ? p.n_max : p.nt;
    cudaError_t e = cudaFuncSetAttribute(knn2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return e;
    const int per = tc_pairs_per_launch(batch, p.n_max, a.qis * a.splits, sm_count, l2_budget);
    for (int b0 = 0; b0 < batch; b0 += per) {
        const int nb = batch - b0 < per ? batch - b0 : per;
        // the kernels index pairs from 0: shift every per-pair array to pair b0
        KnnParams k = p;
        if (p.pair_q) k.pair_q = p.pair_q + b0;
        else {   // default pairing: pair b takes frame b
            k.q_base = p.q_base + (size_t)b0 * p.n_max * 32;
            if (p.q_count) k.q_count = p.q_count + b0;
        }
        if (p.pair_t) k.pair_t = p.pair_t + b0;
        else {
            k.t_base = p.t_base + (size_t)b0 * p.n_max * 32;
            if (p.t_count) k.t_count = p.t_count + b0;
        }
        k.idx = p.idx + (size_t)b0 * p.n_max * 2;
        k.dist = p.dist + (size_t)b0 * p.n_max * 2;
        if (p.keep) k.keep = p.keep + (size_t)b0 * p.n_max;
        if (p.partial) k.partial = p.partial + (size_t)b0 * p.n_max * a.splits;
        a.k = k;
        a.batch = nb;
        a.items = nb * a.qis * a.splits;
        if (t_rows > 0) {
            dim3 ug((unsigned)(((size_t)a.n_pad * 16 + 256 * UNPACK_PER_THREAD - 1) / (256 * UNPACK_PER_THREAD)), (unsigned)nb);
            unpack_train_kernel<<<ug, 256, 0, stream>>>(a);
            ++*launches;
        }
        const int grid = a.items < sm_count ? a.items : sm_count;
        knn2_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(a);
        ++*launches;
        if (a.splits > 1) {
            dim3 mg((unsigned)((q_rows + 255) / 256 > 0 ? (q_rows + 255) / 256 : 1), (unsigned)nb);
            knn2_tc_merge_kernel<<<mg, 256, 0, stream>>>(a);
            ++*launches;
        }
    }
    return cudaGetLastError();
}

}  // namespace tvg
