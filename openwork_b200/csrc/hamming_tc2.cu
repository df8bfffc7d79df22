// 256-bit Hamming kNN (k=2) + Lowe ratio flag on the 5th-generation tensor cores, CTA-pair form  --  the batched
// cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(curr_des, last_des, k=2)   (MonocularSlam/slam.py:126-131).
//
// Same arithmetic as hamming_tc.cu (bit -> e4m3 +-1, dot = 256 - 2 * hamming exactly in f16 accumulators, top-2 on
// packed keys in the tensor-memory epilogue: bit-identical to OpenCV, lowest-train-index tie-break), different
// machine mapping:
//   * two CTAs on the two SMs of a TPC form a cluster and issue ONE tcgen05.mma.cta_group::2 (M = 256: 128 query rows
//     from each CTA; N = 256 train rows, 128 held by each CTA).  Every train row is expanded once per pair of SMs and
//     each SM reads half of the B operand from its shared memory: the shared-memory bandwidth that bounds the
//     single-CTA form (A + B operand reads + tile writes = 121 of 128 B/clk) drops to ~75 B/clk;
//   * the train descriptors are expanded INSIDE the kernel: three warps per CTA turn the raw 32-byte rows (read
//     straight from the frame pool, prefetched one tile ahead in registers) into the MMA's K-major core-matrix
//     layout in shared memory.  No unpack kernel, no 512 KB-per-pair scratch: DRAM traffic = the algorithmic bytes.
//
//   warp 0        leader CTA: MMA issuer (elected lane): per tile 2 x 8 tcgen05.mma (M=256, N=256, K=32) into the two
//                 accumulators (2 x 256 TMEM columns on both SMs); tcgen05.commit multicast to both CTAs' mbarriers
//   warps 1..3    expanders: 128 train rows of the tile -> 32 KB operand half in a 4-stage ring
//   warps 4..11   epilogue, 4 warps per accumulator (thread = query row = TMEM lane), as in hamming_tc.cu; the read-back
//                 of a tile is split: the first 64 columns are waited for alone and their top-2 arithmetic runs while
//                 the other 192 load, then the accumulator goes back to the MMA warp and the rest follows (the
//                 tensor-memory load used to sit in front of every tile's arithmetic: 0.768 -> 0.714 ms per call)
// Barriers the MMA issuer waits on live in the leader CTA; the peer arrives remotely (mapa + mbarrier.arrive).
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

#include "tvg_common.cuh"
#include "tvg_kernels.h"

namespace tvg {
namespace {

constexpr int T2_QB = 128;                      // query rows per accumulator per CTA (= TMEM lanes)
constexpr int T2_NQB = 2;                       // accumulators per CTA
constexpr int T2_QC = T2_NQB * T2_QB;           // query rows per CTA per item
constexpr int T2_QI = 2 * T2_QC;                // query rows per cluster item
constexpr int T2_TN = 256;                      // train rows per tile (MMA N)
constexpr int T2_TH = T2_TN / 2;                // train rows expanded by one CTA
constexpr int T2_KB = 256;                      // bytes per e4m3 row
constexpr int T2_A_BLK = T2_QB * T2_KB;         // 32 KB
constexpr int T2_B_HALF = T2_TH * T2_KB;        // 32 KB
#ifndef T2_NST_N
#define T2_NST_N 4
#endif
#ifndef T2_EXP_WARPS_N
#define T2_EXP_WARPS_N 3
#endif
constexpr int T2_NST = T2_NST_N;                // train-tile ring stages (tools/sweep_tc2_knobs.sh times the alternatives)
constexpr int T2_EXP_WARPS = T2_EXP_WARPS_N;    // expander warps per CTA
#ifndef T2_CSPLIT_N
#define T2_CSPLIT_N 1
#endif
constexpr int T2_CSPLIT = T2_CSPLIT_N;          // epilogue warps per (accumulator, TMEM lane quarter): each takes TN / CSPLIT columns
constexpr int T2_EPI_WARPS = 4 * T2_NQB * T2_CSPLIT;   // (measured: 16 warps with half the columns each are slower than 8:
                                                // 0.93 vs 0.88 ms -- the epilogue is bound by the integer pipe, not by latency)
constexpr int T2_THREADS = 32 * (1 + T2_EXP_WARPS + T2_EPI_WARPS);
constexpr int T2_CW = T2_TN / T2_CSPLIT;        // columns per epilogue warp and tile (128)
// Column halves: the accumulator of a query block is filled and drained as two independent 128-column halves (MMA
// N = 128, four full / empty barrier pairs instead of two), so that the tensor-memory load of one half is in flight
// while the epilogue's arithmetic runs on the other and the MMAs of a half start as soon as that half has been read.
#ifndef T2_HALVES_N
#define T2_HALVES_N 0
#endif
#ifndef T2_SPLIT_LD_N
#define T2_SPLIT_LD_N 1                         // 64-column loads of a tile whose arithmetic precedes the accumulator's release
#endif
#ifndef T2_FIRST_LD_N
#define T2_FIRST_LD_N 1                         // 64-column loads waited for before that arithmetic starts (>= T2_SPLIT_LD_N)
#endif
#ifndef T2_EARLY_LD_N
#define T2_EARLY_LD_N 0                         // next tile's first load behind the last quarter of the arithmetic (measured: 0.83 ms against 0.72 -- slower)
#endif
constexpr int T2_NC = T2_HALVES_N ? 2 : 1;      // column halves per accumulator
constexpr int T2_CN = T2_TN / T2_NC;            // MMA N
static_assert(T2_NC == 1 || T2_CSPLIT == 1, "column halves replace the column split of the epilogue warps");
constexpr int T2_XCH = T2_NQB * T2_QB * 8;      // top-2 hand-off between the column halves of a row (uint2 per row)
constexpr int T2_CONST = 2 * 4096;              // the two constant operand blocks of the offset MMA (see below)
constexpr int T2_SMEM = T2_NQB * T2_A_BLK + T2_NST * T2_B_HALF + T2_XCH + T2_CONST + 256;
constexpr int T2_UNITS = T2_TH * 4;             // 8-byte input units of a half tile: unit = quarter * 128 + row, so the
                                                // lanes of a warp take consecutive rows (conflict-free 16-byte stores)
constexpr int T2_UPT = (T2_UNITS + 32 * T2_EXP_WARPS - 1) / (32 * T2_EXP_WARPS);   // units per expander thread (6)
#ifndef T2_LISTS
#define T2_LISTS 4
#endif
// instruction descriptor: D = f16, A = B = e4m3, both K-major, N >> 3 at bit 17, M >> 4 at bit 24 (M = 256 over the pair)
constexpr uint32_t T2_IDESC = ((uint32_t)(T2_CN >> 3) << 17) | ((uint32_t)((2 * T2_QB) >> 4) << 24);
constexpr uint64_t T2_B_HALF_STEP = (uint64_t)(((T2_TH / T2_NC) / 8) * 2048) >> 4;   // B rows of the next column half, in descriptor units

__device__ __forceinline__ uint32_t t2_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void t2_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.  Default (.release.cta) semantics
// as in CUTLASS's ClusterBarrier::arrive: what the barrier orders here is either tensor-memory traffic (tcgen05
// fences) or generic-proxy shared-memory stores already made visible to the async proxy by fence.proxy.async; the
// .release.cluster / .acquire.cluster forms cost a MEMBAR.GPU per arrive and an L1 invalidate (CCTL.IVALL) per wait
// (measured: 1.09 ms per call against 0.xx with the plain forms).
__device__ __forceinline__ void t2_arrive_remote(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ bool t2_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a protocol error traps (the host sees a launch failure) instead of hanging the device
__device__ __forceinline__ void t2_wait(uint64_t* bar, uint32_t parity) {
    unsigned int spins = 0;
    while (!t2_try(bar, parity))
        if (++spins > (1u << 27)) __trap();
}
__device__ __forceinline__ bool t2_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void t2_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
// completion of every MMA issued so far -> one arrival on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void t2_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void t2_ld64_pack16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void t2_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void t2_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void t2_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 descriptor bits -> 16 e4m3 bytes (+1 = 0x38 for a set bit, -1 = 0xB8).  Per nibble: a carry-free multiply spreads
// bit j to the top bit of byte j (n * 0x10204080 for the low nibble in place, n * 0x01020408 for the high nibble in
// place: no shift needed), one LOP3 masks and flips: (prod & m80) ^ xb8.  9 integer-pipe + 4 multiply instructions per
// 16 bytes.  m80 / xb8 are 0x80808080 / 0xB8B8B8B8 for a live row and 0 / 0 for a padding row (all-zero operand row),
// so padding costs no branch and no extra instruction.
__device__ __forceinline__ uint4 t2_expand16(uint32_t bits, uint32_t m80, uint32_t xb8) {
    const uint32_t hi = bits >> 8;
    uint4 o;
    o.x = (((bits & 0x0Fu) * 0x10204080u) & m80) ^ xb8;
    o.y = (((bits & 0xF0u) * 0x01020408u) & m80) ^ xb8;
    o.z = (((hi & 0x0Fu) * 0x10204080u) & m80) ^ xb8;
    o.w = (((hi & 0xF0u) * 0x01020408u) & m80) ^ xb8;
    return o;
}
// shared-memory operand descriptor: K-major, no swizzle; a core matrix is 8 rows x 16 B contiguous, the K-adjacent
// core matrix 128 B further (leading byte offset), the next 8-row group 2048 B further (stride byte offset)
__device__ __forceinline__ uint64_t t2_smem_desc(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(2048u >> 4) << 32) | (1ull << 46);
}

// The epilogue wants the accumulator to hold 1024 + hamming, whose f16 bit pattern is 0x6400 | hamming (its ulp is 1
// there).  With A = +-1 and B = -+0.5 the eight data MMAs leave -dot / 2 = hamming - 128 (every partial sum a
// multiple of 0.5 below 129 in magnitude: exact in f16), and a NINTH MMA over 32 constant K-bytes (A = 1.0, B = 36.0:
// 32 * 36 = 1152) adds the offset inside the tensor core.  That replaces one HFMA2 per two comparisons in the epilogue
// -- the integer / half pipe is what bounds it -- by 12.5 % more work on the tensor pipe, which has the headroom.
// Both constant blocks hold one byte value everywhere, so their layout is immaterial: 128 rows x 32 B, 8-row groups
// 256 B apart.
__device__ __forceinline__ uint64_t t2_smem_desc_const(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
}

struct Tc2Args {
    KnnParams k;
    int batch, qis, items;          // qis = 512-row query super-blocks per pair; items = cluster items
    int splits, tiles_per_split;    // train tiles of a pair cut into `splits` ranges when items alone cannot fill the GPU
    int ablate;                     // timing experiments only (env TVG_TC2_ABLATE; wrong results): 1 = no epilogue
                                    // arithmetic, 2 = no operand expansion, 3 = both
};

__device__ __forceinline__ void t2_counts(const Tc2Args& a, int b, int& fq, int& ft, int& nq, int& nt) {
    fq = a.k.pair_q ? a.k.pair_q[b] : b;
    ft = a.k.pair_t ? a.k.pair_t[b] : b;
    nq = a.k.q_count ? min(a.k.q_count[fq], a.k.n_max) : a.k.nq;
    nt = a.k.t_count ? min(a.k.t_count[ft], a.k.n_max) : a.k.nt;
}

__device__ __forceinline__ void t2_emit(const KnnParams& p, size_t o, uint32_t k1, uint32_t k2) {
    int2 id, di;
    id.x = (k1 == TVG_SENTINEL_KEY) ? -1 : (int)(k1 & TVG_KEY_IDX_MASK);
    di.x = (k1 == TVG_SENTINEL_KEY) ? -1 : (int)(k1 >> TVG_KEY_SHIFT);
    id.y = (k2 == TVG_SENTINEL_KEY) ? -1 : (int)(k2 & TVG_KEY_IDX_MASK);
    di.y = (k2 == TVG_SENTINEL_KEY) ? -1 : (int)(k2 >> TVG_KEY_SHIFT);
    reinterpret_cast<int2*>(p.idx)[o] = id;
    reinterpret_cast<int2*>(p.dist)[o] = di;
    if (p.keep) p.keep[o] = (k2 != TVG_SENTINEL_KEY) && ((double)di.x < p.ratio * (double)di.y);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1) knn2_tc2_kernel(Tc2Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_smem = smem;
    uint8_t* b_smem = smem + T2_NQB * T2_A_BLK;
    uint2* xch = reinterpret_cast<uint2*>(smem + T2_NQB * T2_A_BLK + T2_NST * T2_B_HALF);
    uint8_t* c_smem = smem + T2_NQB * T2_A_BLK + T2_NST * T2_B_HALF + T2_XCH;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T2_NQB * T2_A_BLK + T2_NST * T2_B_HALF + T2_XCH + T2_CONST);
    uint64_t* full = bars;                       // [NST]  operand halves of both CTAs written      (waited on in the leader)
    uint64_t* empty = full + T2_NST;             // [NST]  MMAs that read the stage are complete    (commit, both CTAs)
    uint64_t* a_full = empty + T2_NST;           // [NQB]  query blocks of both CTAs written        (leader)
    uint64_t* acc_full = a_full + T2_NQB;        // [NQB][NC]  accumulator (half) complete          (commit, both CTAs)
    uint64_t* acc_empty = acc_full + T2_NQB * T2_NC;   // [NQB][NC]  read back by both CTAs             (leader)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + T2_NQB * T2_NC);

    const KnnParams& p = a.k;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = t2_cta_rank();
    const int cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T2_NST; ++s) { mbar_init(full + s, 2 * T2_EXP_WARPS); mbar_init(empty + s, 1); }
        for (int h = 0; h < T2_NQB; ++h) mbar_init(a_full + h, 8);
        for (int h = 0; h < T2_NQB * T2_NC; ++h) { mbar_init(acc_full + h, 1); mbar_init(acc_empty + h, 8 * T2_CSPLIT); }
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < T2_CONST / 4; i += T2_THREADS)   // A: 1.0 (0x38), B: 36.0 (0x61)
        reinterpret_cast<uint32_t*>(c_smem)[i] = i < T2_CONST / 8 ? 0x38383838u : 0x61616161u;
    fence_proxy_async();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    t2_fence_before();
    __syncthreads();
    t2_cluster_sync();      // both CTAs' barriers are initialised before anyone arrives remotely
    t2_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // A cluster item is live when its 512-row query super-block has rows and the pair has train tiles in the item's
    // range; every role of both CTAs walks the same item sequence.
    auto item_info = [&](int item, int& b, int& qi, int& fq, int& ft, int& nq, int& nt, int& j0, int& j1) -> bool {
        const int sp = item % a.splits;
        const int rest = item / a.splits;
        b = rest / a.qis;
        qi = rest % a.qis;
        t2_counts(a, b, fq, ft, nq, nt);
        const int ttiles = (nt + T2_TN - 1) / T2_TN;
        j0 = sp * a.tiles_per_split;
        j1 = min(ttiles, j0 + a.tiles_per_split);
        return qi * T2_QI < nq && j0 < j1;
    };

    if (warp == 0) {
        // ---------------------------------------------------------------- MMA issuer (leader CTA only)
        if (rank == 0) {
            uint32_t step = 0, it = 0;
            const uint64_t c_a_desc = t2_smem_desc_const(smem_u32(c_smem)), c_b_desc = t2_smem_desc_const(smem_u32(c_smem + 4096));
            for (int item = cluster; item < a.items; item += n_clusters) {
                int b, qi, fq, ft, nq, nt, j0, j1;
                if (!item_info(item, b, qi, fq, ft, nq, nt, j0, j1)) continue;
                for (int j = j0; j < j1; ++j, ++step) {
                    const uint32_t s = step % T2_NST, ph = (step / T2_NST) & 1;
                    t2_wait(full + s, ph);
                    t2_fence_after();
                    const uint64_t b_desc = t2_smem_desc(smem_u32(b_smem + s * T2_B_HALF));
#pragma unroll
                    for (int h = 0; h < T2_NQB; ++h) {
                        if (j == j0) t2_wait(a_full + h, it & 1);
                        const uint64_t a_desc = t2_smem_desc(smem_u32(a_smem + h * T2_A_BLK));
#pragma unroll
                        for (int c = 0; c < T2_NC; ++c) {
                            t2_wait(acc_empty + h * T2_NC + c, (step & 1) ^ 1);
                            t2_fence_after();
                            if (t2_elect_one()) {
                                const uint32_t d = tmem_base + h * T2_TN + c * T2_CN;
                                const uint64_t bd = b_desc + (uint64_t)c * T2_B_HALF_STEP;
#pragma unroll
                                for (int k = 0; k < T2_KB / 32; ++k)   // 32 K-bytes = two core matrices = 256 B = 16 address units
                                    t2_mma(d, a_desc + (uint64_t)(k * 16), bd + (uint64_t)(k * 16), T2_IDESC, k > 0);
                                t2_mma(d, c_a_desc, c_b_desc, T2_IDESC, 1);   // + 1152, last
                                t2_commit(acc_full + h * T2_NC + c);
                                if (h == T2_NQB - 1 && c == T2_NC - 1) t2_commit(empty + s);
                            }
                            __syncwarp();
                        }
                    }
                }
                ++it;
            }
        }
    } else if (warp <= T2_EXP_WARPS) {
        // ---------------------------------------------------------------- expanders (both CTAs)
        const int et = threadIdx.x - 32;                        // 0 .. 95
        // walk the (item, tile) sequence with a one-tile look-ahead for the raw loads
        int item = cluster - n_clusters, j = 0, j1 = 0, ft = 0, nt = 0;
        auto advance = [&]() -> bool {                          // next live (item, tile); false at the end
            if (++j < j1) return true;
            for (item += n_clusters; item < a.items; item += n_clusters) {
                int b, qi, fq, nq, j0;
                if (item_info(item, b, qi, fq, ft, nq, nt, j0, j1)) { j = j0; return true; }
            }
            return false;
        };
        auto load = [&](uint2 (&raw)[T2_UPT]) {
            const int r0 = j * T2_TN + (int)rank * T2_TH;
#pragma unroll
            for (int u = 0; u < T2_UPT; ++u) {
                const int unit = et + u * 32 * T2_EXP_WARPS;
                const int row = r0 + (unit & (T2_TH - 1));
                raw[u] = make_uint2(0u, 0u);
                if (unit < T2_UNITS && row < nt)
                    raw[u] = __ldg(reinterpret_cast<const uint2*>(p.t_base + ((size_t)ft * p.n_max + row) * 32 + (unit >> 7) * 8));
            }
        };
        uint2 cur[T2_UPT], nxt[T2_UPT];
        bool more = advance();
        int cur_rows = 0;                                        // valid rows of this CTA's half of the current tile
        if (more) { load(cur); cur_rows = nt - (j * T2_TN + (int)rank * T2_TH); }
        uint32_t step = 0;
        while (more) {
            const int rows_here = cur_rows;
            more = advance();
            if (more) { load(nxt); cur_rows = nt - (j * T2_TN + (int)rank * T2_TH); }
            const uint32_t s = step % T2_NST, ph = (step / T2_NST) & 1;
            t2_wait(empty + s, ph ^ 1);
            uint8_t* dst_stage = b_smem + s * T2_B_HALF;
#pragma unroll
            for (int u = 0; u < T2_UPT; ++u) {
                const int unit = et + u * 32 * T2_EXP_WARPS;
                if (unit >= T2_UNITS || (a.ablate & 2)) continue;
                const int row = unit & (T2_TH - 1), quarter = unit >> 7;
                uint8_t* dst = dst_stage + (row >> 3) * 2048 + (row & 7) * 16 + quarter * 4 * 128;
                const bool live = row < rows_here;               // rows beyond the frame: zero operand rows
                const uint32_t m80 = live ? 0x80808080u : 0u, xb8 = live ? 0x30303030u : 0u;   // bit 1 -> -0.5, bit 0 -> +0.5
                const uint32_t w0 = cur[u].x, w1 = cur[u].y;
                *reinterpret_cast<uint4*>(dst) = t2_expand16(w0, m80, xb8);
                *reinterpret_cast<uint4*>(dst + 128) = t2_expand16(w0 >> 16, m80, xb8);
                *reinterpret_cast<uint4*>(dst + 256) = t2_expand16(w1, m80, xb8);
                *reinterpret_cast<uint4*>(dst + 384) = t2_expand16(w1 >> 16, m80, xb8);
            }
            fence_proxy_async();   // generic-proxy stores -> visible to the tensor cores' async-proxy reads
            __syncwarp();
            if (lane == 0) t2_arrive_remote(full + s, 0);
#pragma unroll
            for (int u = 0; u < T2_UPT; ++u) cur[u] = nxt[u];
            ++step;
        }
    } else {
        // ---------------------------------------------------------------- epilogue (both CTAs)
        // a warp reads TMEM lanes 32 * (warp % 4) ...; warps e and e + 4 of an accumulator share the rows and split the columns
        const int e = warp - (1 + T2_EXP_WARPS), h = e / (4 * T2_CSPLIT), sub = (e >> 2) % T2_CSPLIT, quarter = warp & 3;
        const int row_in_blk = quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(h * T2_TN + sub * T2_CW);
        const int pair_bar = 1 + h * 4 + quarter;               // named barrier of the two warps that share these rows
        // query row -> e4m3, into this block's shared-memory operand: [row / 8][16-byte K chunk][row % 8]
        auto write_a = [&](int fq, int nq, int qi) {
            const int q = qi * T2_QI + (int)rank * T2_QC + h * T2_QB + row_in_blk;
            uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
            const bool live = q < nq;
            if (live) {
                const uint4* src = reinterpret_cast<const uint4*>(p.q_base + ((size_t)fq * p.n_max + q) * 32);
                d0 = __ldg(src);
                d1 = __ldg(src + 1);
            }
            const uint32_t w[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            uint8_t* dst = a_smem + h * T2_A_BLK + (row_in_blk >> 3) * 2048 + (row_in_blk & 7) * 16;
#pragma unroll
            const uint32_t m80 = live ? 0x80808080u : 0u, xb8 = live ? 0xB8B8B8B8u : 0u;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const uint32_t bits = w[c >> 1] >> ((c & 1) * 16);
                *reinterpret_cast<uint4*>(dst + c * 128) = t2_expand16(bits, m80, xb8);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) t2_arrive_remote(a_full + h, 0);
        };
        auto next_live = [&](int item, int& fq, int& nq, int& qi) -> bool {
            for (item += n_clusters; item < a.items; item += n_clusters) {
                int b, ft, nt, j0, j1;
                if (item_info(item, b, qi, fq, ft, nq, nt, j0, j1)) return true;
            }
            return false;
        };
        uint32_t use = 0;
        uint32_t va[2][32], vb[2][32];   // column-half form: the two register halves of the accumulator read-back
        bool early0 = false;             // half 0 of the tile about to be processed is already on its way into va
        (void)va; (void)vb; (void)early0;
        {
            int fq, nq, qi;
            if (sub == 0 && next_live(cluster - n_clusters, fq, nq, qi)) write_a(fq, nq, qi);
        }
        for (int item = cluster; item < a.items; item += n_clusters) {
            int b, qi, fq, ft, nq, nt, j0, j1;
            const bool live = item_info(item, b, qi, fq, ft, nq, nt, j0, j1);
            uint32_t g1 = TVG_SENTINEL_KEY, g2 = TVG_SENTINEL_KEY;       // dist << 23 | train idx
            if (!live) j1 = j0;
            if constexpr (T2_NC == 2) {
                // ---- column halves: while the arithmetic runs on one half (64 registers), the tensor-memory load of the
                // next half is in flight into the other 64; a half goes back to the MMA warp as soon as it has been read
                auto ld_half = [&](int c, uint32_t (&v)[2][32]) {
                    t2_ld64_pack16(taddr + c * T2_CN, v[0]);
                    t2_ld64_pack16(taddr + c * T2_CN + 64, v[1]);
                };
                for (int j = j0; j < j1; ++j, ++use) {
                    const int cols = min(T2_TN, nt - j * T2_TN);
                    uint32_t m1[T2_LISTS], m2[T2_LISTS];
#pragma unroll
                    for (int l = 0; l < T2_LISTS; ++l) m1[l] = m2[l] = 0xFFFFFFFFu;
                    // columns [0, 64) of a half come from the leader CTA's train rows, [64, 128) from the peer's: register
                    // i of load g of half c holds tile rows 2 * pl and 2 * pl + 1 with pl = 64 g + 32 c + i
                    auto arith = [&](const uint32_t (&v)[2][32], auto cc, auto masked) {
                        constexpr int c = decltype(cc)::value;
#pragma unroll
                        for (int g = 0; g < 2; ++g)
#pragma unroll
                            for (int i = 0; i < 32; i += 2) {
                                uint32_t key[2];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const uint32_t pl = (uint32_t)(64 * g + 32 * c + i + u);
                                    key[u] = v[g][i + u] * p.w128 + (pl | (((pl - 0x32u) & 0xFFFFu) << 16));
                                    if (decltype(masked)::value) {
                                        if (2 * (int)pl >= cols) key[u] |= 0x0000FFFFu;
                                        if (2 * (int)pl + 1 >= cols) key[u] |= 0xFFFF0000u;
                                    }
                                }
                                const int l = (i >> 1) & (T2_LISTS - 1);
                                const uint32_t lo = __vminu2(key[0], key[1]), hi = __vmaxu2(key[0], key[1]);
                                m2[l] = __vimin3_u16x2(m2[l], hi, __vmaxu2(m1[l], lo));
                                m1[l] = __vminu2(m1[l], lo);
                            }
                    };
                    // ---- half 0
                    if (!early0) {
                        t2_wait(acc_full + h * 2, use & 1);
                        t2_fence_after();
                        ld_half(0, va);
                    }
                    t2_ld_wait();
                    t2_fence_before();
                    __syncwarp();
                    if (lane == 0) t2_arrive_remote(acc_empty + h * 2, 0);
                    // half 1 of the same tile is usually complete by now (its MMAs follow half 0's)
                    const bool early1 = __all_sync(0xffffffffu, t2_try(acc_full + h * 2 + 1, use & 1));
                    if (early1) {
                        t2_fence_after();
                        ld_half(1, vb);
                    }
                    if (!(a.ablate & 1)) {
                        if (cols >= T2_TN) arith(va, std::integral_constant<int, 0>{}, std::false_type{});
                        else arith(va, std::integral_constant<int, 0>{}, std::true_type{});
                    }
                    // ---- half 1
                    if (!early1) {
                        t2_wait(acc_full + h * 2 + 1, use & 1);
                        t2_fence_after();
                        ld_half(1, vb);
                    }
                    t2_ld_wait();
                    t2_fence_before();
                    __syncwarp();
                    if (lane == 0) t2_arrive_remote(acc_empty + h * 2 + 1, 0);
                    // last tile of the item: every MMA that read this A block is complete -> stage the next item's rows
                    if (j == j1 - 1) {
                        int fq2, nq2, qi2;
                        if (next_live(item, fq2, nq2, qi2)) write_a(fq2, nq2, qi2);
                    }
                    // half 0 of the item's next tile, if it is complete already
                    early0 = false;
                    if (j + 1 < j1) {
                        early0 = __all_sync(0xffffffffu, t2_try(acc_full + h * 2, (use + 1) & 1));
                        if (early0) {
                            t2_fence_after();
                            ld_half(0, va);
                        }
                    }
                    if (a.ablate & 1) { g1 = min(g1, va[0][0] ^ vb[1][1]); continue; }
                    if (cols >= T2_TN) arith(vb, std::integral_constant<int, 1>{}, std::false_type{});
                    else arith(vb, std::integral_constant<int, 1>{}, std::true_type{});
#pragma unroll
                    for (int l = 1; l < T2_LISTS; ++l) {
                        const uint32_t hi = __vmaxu2(m1[0], m1[l]);
                        m1[0] = __vminu2(m1[0], m1[l]);
                        m2[0] = __vminu2(hi, __vminu2(m2[0], m2[l]));
                    }
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const uint32_t k16 = (((w & 2) ? m2[0] : m1[0]) >> ((w & 1) * 16)) & 0xFFFFu;
                        if (k16 != 0xFFFFu) {
                            const uint32_t key = ((k16 >> 7) << TVG_KEY_SHIFT) | (uint32_t)(j * T2_TN + 2 * (k16 & 127u) + (w & 1));
                            top2_insert(g1, g2, key);
                        }
                    }
                }
            } else {
                uint32_t v[T2_CW / 64][32];
                bool early = false;      // the first 64 columns of the tile about to be processed are already on their way into v[0]
                for (int j = j0; j < j1; ++j, ++use) {
                    if (!early) {
                        t2_wait(acc_full + h, use & 1);
                        t2_fence_after();
                    }
                    uint32_t m1[T2_LISTS], m2[T2_LISTS];
#pragma unroll
                    for (int l = 0; l < T2_LISTS; ++l) m1[l] = m2[l] = 0xFFFFFFFFu;
                    const int cols = min(T2_TN, nt - j * T2_TN);
                    auto chunk = [&](auto masked, auto c_lo, auto c_hi) {
#pragma unroll
                        for (int c = decltype(c_lo)::value; c < decltype(c_hi)::value; ++c)
#pragma unroll
                            for (int i = 0; i < 32; i += 2) {
                                uint32_t key[2];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const uint32_t pl = (uint32_t)(sub * (T2_CW / 2) + c * 32 + i + u);
                                    // accumulator = 1024 + d in f16 = bit pattern 0x6400 | d per 16-bit lane; * 128 + C_p ->
                                    // (d << 7) | p per lane (the 0x6400 terms overflow into constants that C_p cancels)
                                    key[u] = v[c][i + u] * p.w128 + (pl | (((pl - 0x32u) & 0xFFFFu) << 16));
                                    if (decltype(masked)::value) {
                                        if (2 * (int)pl >= cols) key[u] |= 0x0000FFFFu;
                                        if (2 * (int)pl + 1 >= cols) key[u] |= 0xFFFF0000u;
                                    }
                                }
                                // two keys per update: 5 min/max (one 3-input) instead of 6
                                const int l = (i >> 1) & (T2_LISTS - 1);
                                const uint32_t lo = __vminu2(key[0], key[1]), hi = __vmaxu2(key[0], key[1]);
                                m2[l] = __vimin3_u16x2(m2[l], hi, __vmaxu2(m1[l], lo));
                                m1[l] = __vminu2(m1[l], lo);
                            }
                    };
                    constexpr int NLD = T2_CW / 64;              // tensor-memory loads per tile and warp (4)
                    constexpr int SPL = (T2_SPLIT_LD_N && NLD >= 2) ? T2_SPLIT_LD_N : 0;   // loads whose arithmetic runs before the release
                    using C0 = std::integral_constant<int, 0>;
                    using CS = std::integral_constant<int, SPL>;
                    using CN = std::integral_constant<int, NLD>;
                    if (SPL > 0) {
                        // first half of the columns, then the second half's load runs behind the first arithmetic
                        if (!early) {
#pragma unroll
                            for (int c = 0; c < T2_FIRST_LD_N; ++c) t2_ld64_pack16(taddr + c * 64, v[c]);
                        }
                        t2_ld_wait();
#pragma unroll
                        for (int c = T2_FIRST_LD_N; c < NLD; ++c) t2_ld64_pack16(taddr + c * 64, v[c]);
                        if (!(a.ablate & 1)) {
                            if (cols >= T2_TN) chunk(std::false_type{}, C0{}, CS{}); else chunk(std::true_type{}, C0{}, CS{});
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < NLD; ++c) t2_ld64_pack16(taddr + c * 64, v[c]);
                    }
                    t2_ld_wait();
                    // the accumulator is in registers now: hand it back to the MMA warp before the (rest of the) arithmetic
                    t2_fence_before();
                    __syncwarp();
                    if (lane == 0) t2_arrive_remote(acc_empty + h, 0);
                    // last tile of the item: every MMA that read this A block is complete -> stage the next item's rows
                    if (sub == 0 && j == j1 - 1) {
                        int fq2, nq2, qi2;
                        if (next_live(item, fq2, nq2, qi2)) write_a(fq2, nq2, qi2);
                    }
                    if (a.ablate & 1) { g1 = min(g1, v[0][0] ^ v[1][1]); continue; }
                    constexpr bool EARLY = T2_EARLY_LD_N && SPL == 1 && T2_FIRST_LD_N == 1 && NLD >= 3;
                    using CL = std::integral_constant<int, EARLY ? NLD - 1 : NLD>;
                    if (cols >= T2_TN) chunk(std::false_type{}, CS{}, CL{}); else chunk(std::true_type{}, CS{}, CL{});
                    early = false;
                    if (EARLY) {
                        // the item's next tile: if its MMAs are complete already, its first 64 columns load (into v[0],
                        // consumed above) behind the last quarter of this tile's arithmetic
                        if (j + 1 < j1) {
                            early = __all_sync(0xffffffffu, t2_try(acc_full + h, (use + 1) & 1));
                            if (early) {
                                t2_fence_after();
                                t2_ld64_pack16(taddr, v[0]);
                            }
                        }
                        if (cols >= T2_TN) chunk(std::false_type{}, CL{}, CN{}); else chunk(std::true_type{}, CL{}, CN{});
                    }
    #pragma unroll
                    for (int l = 1; l < T2_LISTS; ++l) {
                        const uint32_t hi = __vmaxu2(m1[0], m1[l]);
                        m1[0] = __vminu2(m1[0], m1[l]);
                        m2[0] = __vminu2(hi, __vminu2(m2[0], m2[l]));
                    }
    #pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        const uint32_t k16 = (((w & 2) ? m2[0] : m1[0]) >> ((w & 1) * 16)) & 0xFFFFu;
                        if (k16 != 0xFFFFu) {
                            const uint32_t key = ((k16 >> 7) << TVG_KEY_SHIFT) | (uint32_t)(j * T2_TN + 2 * (k16 & 127u) + (w & 1));
                            top2_insert(g1, g2, key);
                        }
                    }
                }
            }
            // the two column halves of a row meet: the upper half hands its top-2 over through shared memory
            if (T2_CSPLIT > 1) {
                if (sub == 1) xch[h * T2_QB + row_in_blk] = make_uint2(g1, g2);
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            }
            if (sub == 0) {
                if (T2_CSPLIT > 1) {
                    const uint2 o = xch[h * T2_QB + row_in_blk];
                    top2_merge(g1, g2, o.x, o.y);
                }
                const int q = qi * T2_QI + (int)rank * T2_QC + h * T2_QB + row_in_blk;
                if (live && q < nq) {
                    if (a.splits == 1) t2_emit(p, (size_t)b * p.n_max + q, g1, g2);
                    else p.partial[((size_t)b * p.n_max + q) * a.splits + item % a.splits] = make_uint2(g1, g2);
                }
            }
            if (T2_CSPLIT > 1) asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // the slot is free again
        }
    }
    t2_fence_before();
    __syncthreads();
    t2_cluster_sync();      // the peer may still read this CTA's shared memory / signal its barriers until here
    if (warp == 0) {
        t2_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// split-Nt: fold the per-split partial top-2s of one query (thread = query) and emit
__global__ void __launch_bounds__(256) knn2_tc2_merge_kernel(Tc2Args a) {
    const int b = blockIdx.y, q = blockIdx.x * blockDim.x + threadIdx.x;
    int fq, ft, nq, nt;
    t2_counts(a, b, fq, ft, nq, nt);
    if (q >= nq) return;
    const int ttiles = (nt + T2_TN - 1) / T2_TN;
    const int used = (ttiles + a.tiles_per_split - 1) / a.tiles_per_split;   // splits that held tiles for this pair
    uint32_t g1 = TVG_SENTINEL_KEY, g2 = TVG_SENTINEL_KEY;
    const uint2* part = a.k.partial + ((size_t)b * a.k.n_max + q) * a.splits;
    for (int sp = 0; sp < used; ++sp) {
        const uint2 v = part[sp];
        top2_merge(g1, g2, v.x, v.y);
    }
    t2_emit(a.k, (size_t)b * a.k.n_max + q, g1, g2);
}

}  // namespace

int knn2_tc2_splits(int batch, int q_rows, int n_max, int sm_count) {
    const int qis = (q_rows + T2_QI - 1) / T2_QI > 0 ? (q_rows + T2_QI - 1) / T2_QI : 1;
    const long long base = (long long)batch * qis;
    const int clusters = sm_count / 2;
    const int ttiles = (n_max + T2_TN - 1) / T2_TN;
    long long splits = base >= clusters ? 1 : clusters / base;
    if (splits > ttiles) splits = ttiles;
    if (splits < 1) splits = 1;
    const int tps = (ttiles + (int)splits - 1) / (int)splits;
    return tps > 0 ? (ttiles + tps - 1) / tps : 1;
}

// p.partial must hold batch * n_max * splits uint2 when splits = knn2_tc2_splits(...) > 1
cudaError_t launch_knn2_tc2(const KnnParams& p, int batch, int q_rows, int sm_count, int* launches, cudaStream_t stream) {
    Tc2Args a{};
    a.k = p;
    a.batch = batch;
    a.qis = (q_rows + T2_QI - 1) / T2_QI;
    if (a.qis < 1) a.qis = 1;
    a.splits = knn2_tc2_splits(batch, q_rows, p.n_max, sm_count);
    const int ttiles = (p.n_max + T2_TN - 1) / T2_TN;
    a.tiles_per_split = (ttiles + a.splits - 1) / a.splits;
    if (a.tiles_per_split < 1) a.tiles_per_split = 1;
    a.items = batch * a.qis * a.splits;
    {
        const char* env = getenv("TVG_TC2_ABLATE");
        a.ablate = env ? atoi(env) : 0;
    }
    *launches = 0;
    cudaError_t e = cudaFuncSetAttribute(knn2_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM);
    if (e != cudaSuccess) return e;
    const int clusters = a.items < sm_count / 2 ? a.items : sm_count / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(T2_THREADS);
    cfg.dynamicSmemBytes = T2_SMEM;
    cfg.stream = stream;
    e = cudaLaunchKernelEx(&cfg, knn2_tc2_kernel, a);   // the cluster shape is the kernel's __cluster_dims__(2, 1, 1)
    if (e != cudaSuccess) return e;
    ++*launches;
    if (a.splits > 1) {
        dim3 mg((unsigned)((q_rows + 255) / 256 > 0 ? (q_rows + 255) / 256 : 1), (unsigned)batch);
        knn2_tc2_merge_kernel<<<mg, 256, 0, stream>>>(a);
        ++*launches;
    }
    return cudaGetLastError();
}

}  // namespace tvg
