// Brute-force 256-bit Hamming kNN (k=2) + Lowe ratio flag  --  replaces
// cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(curr_des, last_des, k=2) and the ratio test of
// /root/reference/MonocularSlam/slam.py:126-137.
//
// Design (sm_100a):
//   * one thread owns R query descriptors in registers (8 x u32 each, fetched with two 128-bit loads);
//   * the CTA streams its slice of train rows through a STAGES-deep shared-memory ring filled by
//     TMA bulk copies (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP); every lane of a warp reads
//     the SAME train row -> shared-memory broadcast, conflict-free, no swizzle needed;
//   * distance = popcount(q ^ t) on the integer pipes: LOP3 (xor / carry-save adders) + POPC;
//   * top-2 kept as packed keys  dist<<23 | trainIdx , so unsigned min() IS "lowest distance, then
//     lowest train index" (OpenCV's strict-< insertion order) and results are bit-exact;
//   * when Nq alone cannot fill 148 SMs the train rows are split over gridDim.y; the last CTA to
//     finish a query tile merges the per-split partial top-2 lists with a warp-shuffle butterfly and
//     applies the ratio test ( d1 < ratio * d2 , strict, evaluated in fp64 like the Python float compare).
#include "tvg_common.cuh"
#include "tvg_kernels.h"

namespace tvg {

__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// popcount of the 256-bit XOR.  VARIANT 0: 8 LOP3 + 8 POPC.  VARIANT 1: Harley-Seal carry-save
// adders fold the eight words into one weight-1 pair, one weight-2 and one weight-4 word:
// 16 LOP3 + 4 POPC (trades POPC-pipe work, 16 lanes/clk/SM, for ALU-pipe work, 64 lanes/clk/SM).
// VARIANT 2: three CSAs, 14 LOP3 + 5 POPC.
template <int VARIANT>
__device__ __forceinline__ uint32_t hamming256(const uint32_t (&q)[8], const uint4& t0, const uint4& t1) {
    const uint32_t x0 = q[0] ^ t0.x, x1 = q[1] ^ t0.y, x2 = q[2] ^ t0.z, x3 = q[3] ^ t0.w;
    const uint32_t x4 = q[4] ^ t1.x, x5 = q[5] ^ t1.y, x6 = q[6] ^ t1.z, x7 = q[7] ^ t1.w;
    if (VARIANT == 0) {
        return (__popc(x0) + __popc(x1)) + (__popc(x2) + __popc(x3)) + (__popc(x4) + __popc(x5)) +
               (__popc(x6) + __popc(x7));
    } else if (VARIANT == 1) {
        const uint32_t sa = lop3_xor3(x0, x1, x2), ca = lop3_maj(x0, x1, x2);
        const uint32_t sb = lop3_xor3(x3, x4, x5), cb = lop3_maj(x3, x4, x5);
        const uint32_t sc = lop3_xor3(x6, x7, sa), cc = lop3_maj(x6, x7, sa);
        const uint32_t sd = lop3_xor3(ca, cb, cc), cd = lop3_maj(ca, cb, cc);
        // weight 1: sb, sc ; weight 2: sd ; weight 4: cd
        return (__popc(sb) + __popc(sc)) + 2u * __popc(sd) + 4u * __popc(cd);
    } else {
        const uint32_t sa = lop3_xor3(x0, x1, x2), ca = lop3_maj(x0, x1, x2);
        const uint32_t sb = lop3_xor3(x3, x4, x5), cb = lop3_maj(x3, x4, x5);
        const uint32_t sc = lop3_xor3(x6, x7, sa), cc = lop3_maj(x6, x7, sa);
        return (__popc(sb) + __popc(sc)) + 2u * ((__popc(ca) + __popc(cb)) + __popc(cc));
    }
}

// Final (idx, dist) pair of one query + the Lowe ratio flag (slam.py:131: m.distance < 0.75*n.distance).
__device__ __forceinline__ void knn2_emit(const KnnParams& p, size_t o, uint32_t k1, uint32_t k2) {
    int2 id, di;
    id.x = (k1 == TVG_SENTINEL_KEY) ? -1 : (int)(k1 & TVG_KEY_IDX_MASK);
    di.x = (k1 == TVG_SENTINEL_KEY) ? -1 : (int)(k1 >> TVG_KEY_SHIFT);
    id.y = (k2 == TVG_SENTINEL_KEY) ? -1 : (int)(k2 & TVG_KEY_IDX_MASK);
    di.y = (k2 == TVG_SENTINEL_KEY) ? -1 : (int)(k2 >> TVG_KEY_SHIFT);
    reinterpret_cast<int2*>(p.idx)[o] = id;
    reinterpret_cast<int2*>(p.dist)[o] = di;
    if (p.keep) p.keep[o] = (k2 != TVG_SENTINEL_KEY) && ((double)di.x < p.ratio * (double)di.y);
}

// mad.lo with a multiplier the compiler cannot see (kernel parameter): stays an IMAD on the FMA pipe
// instead of being strength-reduced to LEA / SHF on the (saturated) ALU pipe.
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// VARIANT 3: Harley-Seal with the first-level carry-save adders sharing XOR work between the query and
// the train row.  With x_i = q_i ^ t_i :  sum(x0,x1,x2) = Q012 ^ T012  and
// maj(x0,x1,x2) = x0 ^ ((x0^x1) & (x0^x2)) = (q0^t0) ^ ((Q01^T01) & (Q02^T02)), where Q01 = q0^q1 etc. are
// computed once per query and T01 = t0^t1 etc. once per train row (by the CTA when the TMA tile lands).
// 14 LOP3 + 4 POPC per comparison instead of 16 + 4.
// Expanded train row (12 words): t0 T01 T02 T012 | t3 T34 T35 T345 | t6 T67 - -
// Query combos (10 words):       q0 Q01 Q02 Q012   q3 Q34 Q35 Q345   q6 Q67
__device__ __forceinline__ uint32_t lop3_and_xor(uint32_t a, uint32_t b, uint32_t c) {  // a & (b ^ c)
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x60;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lop3_maj_via(uint32_t a, uint32_t b, uint32_t c) {  // a ^ (b & (a ^ c))
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xB8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ void expand_query(const uint32_t (&q)[8], uint32_t (&qc)[10]) {
    qc[0] = q[0]; qc[1] = q[0] ^ q[1]; qc[2] = q[0] ^ q[2]; qc[3] = q[0] ^ q[1] ^ q[2];
    qc[4] = q[3]; qc[5] = q[3] ^ q[4]; qc[6] = q[3] ^ q[5]; qc[7] = q[3] ^ q[4] ^ q[5];
    qc[8] = q[6]; qc[9] = q[6] ^ q[7];
}
__device__ __forceinline__ uint32_t hamming256_x(const uint32_t (&qc)[10], const uint4& e0, const uint4& e1,
                                                 const uint2& e2, uint32_t w2) {
    // CSA over (x0,x1,x2)
    const uint32_t ua = qc[1] ^ e0.y;
    const uint32_t va = lop3_and_xor(ua, qc[2], e0.z);
    const uint32_t ca = lop3_xor3(qc[0], e0.x, va);
    const uint32_t sa = qc[3] ^ e0.w;
    // CSA over (x3,x4,x5)
    const uint32_t ub = qc[5] ^ e1.y;
    const uint32_t vb = lop3_and_xor(ub, qc[6], e1.z);
    const uint32_t cb = lop3_xor3(qc[4], e1.x, vb);
    const uint32_t sb = qc[7] ^ e1.w;
    // CSA over (x6,x7,sa)
    const uint32_t uc = qc[9] ^ e2.y;          // x6 ^ x7
    const uint32_t x6 = qc[8] ^ e2.x;
    const uint32_t sc = uc ^ sa;
    const uint32_t cc = lop3_maj_via(x6, uc, sa);
    // CSA over the three carries
    const uint32_t sd = lop3_xor3(ca, cb, cc), cd = lop3_maj(ca, cb, cc);
    // weight 1: sb, sc ; weight 2: sd ; weight 4: cd
    return imad(imad(__popc(cd), w2, __popc(sd)), w2, __popc(sb) + __popc(sc));
}

// Publish this CTA's top-2 (directly when the train rows are not split, else as partials that the last
// CTA of the query tile merges with a warp-shuffle butterfly) and apply the ratio test.
template <int R, int THREADS>
__device__ __forceinline__ void knn2_finish(const KnnParams& p, int b, int qt, int split, int q0, int nq,
                                            const uint32_t (&m1)[R], const uint32_t (&m2)[R], int* s_is_last) {
    constexpr int BQ = R * THREADS;
    const int tid = threadIdx.x;
    const int32_t ob = b * p.n_max;
    if (p.splits == 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int qi = q0 + r * THREADS + tid;
            if (qi < nq) knn2_emit(p, (size_t)(ob + qi), m1[r], m2[r]);
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int qi = q0 + r * THREADS + tid;
        if (qi < nq) p.partial[(size_t)(ob + qi) * p.splits + split] = make_uint2(m1[r], m2[r]);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int* ctr = p.tile_counters + (size_t)b * gridDim.x + qt;
        const unsigned int old = atomicAdd(ctr, 1u);
        *s_is_last = (old == (unsigned)p.splits - 1u);
        if (*s_is_last) *ctr = 0u;  // self-resetting: the workspace stays zeroed between launches
    }
    __syncthreads();
    if (!*s_is_last) return;
    __threadfence();

    // Each warp merges MB queries per iteration so that MB independent partial loads (and butterflies)
    // are in flight: the merge is latency-bound, not throughput-bound.
    constexpr int MB = 8;
    const int warp = tid >> 5, lane = tid & 31;
    const int qend = min(nq, q0 + BQ);
    for (int qb = q0 + warp * MB; qb < qend; qb += (THREADS / 32) * MB) {
        uint32_t a1[MB], a2[MB];
#pragma unroll
        for (int u = 0; u < MB; ++u) {
            a1[u] = TVG_SENTINEL_KEY;
            a2[u] = TVG_SENTINEL_KEY;
        }
        for (int s = lane; s < p.splits; s += 32) {
            uint2 v[MB];
#pragma unroll
            for (int u = 0; u < MB; ++u) {
                const int qi = min(qb + u, qend - 1);
                v[u] = __ldcg(p.partial + (size_t)(ob + qi) * p.splits + s);
            }
#pragma unroll
            for (int u = 0; u < MB; ++u) top2_merge(a1[u], a2[u], v[u].x, v[u].y);
        }
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {  // warp-shuffle top-2 butterfly, MB queries interleaved
#pragma unroll
            for (int u = 0; u < MB; ++u) {
                const uint32_t b1 = __shfl_xor_sync(0xffffffffu, a1[u], m);
                const uint32_t b2 = __shfl_xor_sync(0xffffffffu, a2[u], m);
                top2_merge(a1[u], a2[u], b1, b2);
            }
        }
#pragma unroll
        for (int u = 0; u < MB; ++u)
            if (lane == u && qb + u < qend) knn2_emit(p, (size_t)(ob + qb + u), a1[u], a2[u]);
    }
}

// Top-2 update for four keys of one query.  LAZY: most groups cannot change the top-2, so test the group
// minimum against the current second-best first (3 ALU ops per 4 comparisons instead of 12).
template <bool LAZY>
__device__ __forceinline__ void top2_insert4(uint32_t& m1, uint32_t& m2, uint32_t k0, uint32_t k1, uint32_t k2,
                                             uint32_t k3) {
    if (LAZY) {
        const uint32_t kmin = min(min(k0, k1), min(k2, k3));
        if (kmin >= m2) return;
    }
    top2_insert(m1, m2, k0);
    top2_insert(m1, m2, k1);
    top2_insert(m1, m2, k2);
    top2_insert(m1, m2, k3);
}

template <int R, int THREADS, int CH, int STAGES, int VARIANT, bool LAZY, bool PACK, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB) knn2_kernel(const KnnParams p) {
    static_assert(!PACK || (R % 2 == 0 && CH == 128), "packed top-2 pairs queries and needs 7-bit row indices");
    constexpr int BQ = R * THREADS;
    constexpr bool X = (VARIANT == 3);
    constexpr int QW = X ? 10 : 8;
    __shared__ alignas(128) uint4 tile[STAGES][CH * 2];
    __shared__ alignas(16) uint4 xtile[X ? CH * 3 : 1];   // expanded rows (VARIANT 3)
    __shared__ alignas(8) uint64_t full_bar[STAGES];
    __shared__ int s_is_last;

    const int tid = threadIdx.x;
    const int b = blockIdx.z, split = blockIdx.y, qt = blockIdx.x;
    const int fq = p.pair_q ? p.pair_q[b] : b;
    const int ft = p.pair_t ? p.pair_t[b] : b;
    const int nq = p.q_count ? min(p.q_count[fq], p.n_max) : p.nq;
    const int nt = p.t_count ? min(p.t_count[ft], p.n_max) : p.nt;
    const int q0 = qt * BQ;
    if (q0 >= nq) return;  // whole tile empty for every split: nobody touches the counter

    const uint4* __restrict__ qv = reinterpret_cast<const uint4*>(p.q_base + (size_t)fq * p.n_max * 32);
    const uint8_t* __restrict__ tb = p.t_base + (size_t)ft * p.n_max * 32;
    const int row0 = min(nt, split * p.rows_per_split);
    const int row1 = min(nt, row0 + p.rows_per_split);
    const int nrows = row1 - row0;
    const int nchunks = (nrows + CH - 1) / CH;
    const uint32_t key_mul = p.key_mul, w2 = p.w2;

    uint32_t qd[R][QW];
    uint32_t m1[R], m2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int qi = q0 + r * THREADS + tid;
        uint4 a = make_uint4(0, 0, 0, 0), c = a;
        if (qi < nq) {
            a = __ldg(qv + 2 * (size_t)qi);
            c = __ldg(qv + 2 * (size_t)qi + 1);
        }
        const uint32_t raw[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        if (X) {
            uint32_t qc[10];
            expand_query(raw, qc);
#pragma unroll
            for (int i = 0; i < QW; ++i) qd[r][i] = qc[i];
        } else {
#pragma unroll
            for (int i = 0; i < QW; ++i) qd[r][i] = raw[i];
        }
        m1[r] = TVG_SENTINEL_KEY;
        m2[r] = TVG_SENTINEL_KEY;
    }

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int c) {
        const int st = c % STAGES;
        const int rows = min(CH, nrows - c * CH);
        const uint32_t bytes = (uint32_t)rows * 32u;
        mbar_arrive_expect_tx(&full_bar[st], bytes);
        tma_load_1d(&tile[st][0], tb + (size_t)(row0 + c * CH) * 32, bytes, &full_bar[st]);
    };
    if (tid == 0) {
        for (int c = 0; c < STAGES - 1 && c < nchunks; ++c) issue(c);
    }

    // distance of query r against row j of the current chunk
    auto dist_of = [&](const uint4* __restrict__ tp, int j, int r) -> uint32_t {
        if (X) {
            const uint4 e0 = xtile[3 * j], e1 = xtile[3 * j + 1];
            const uint4 e2 = xtile[3 * j + 2];
            uint32_t qc[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) qc[i] = qd[r][i < QW ? i : 0];
            return hamming256_x(qc, e0, e1, make_uint2(e2.x, e2.y), w2);
        } else {
            uint32_t q8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) q8[i] = qd[r][i];
            return hamming256<VARIANT>(q8, tp[2 * j], tp[2 * j + 1]);
        }
    };
    auto key_of = [&](const uint4* __restrict__ tp, int j, int r, uint32_t jg) -> uint32_t {
        return imad(dist_of(tp, j, r), key_mul, jg);
    };
    // PACK: chunk-local 16-bit keys  dist<<7 | row-in-chunk  of two queries share one register, so one
    // VIMNMX.U16x2 updates two top-2 lists; flushed into the 32-bit global keys at the end of every chunk.
    uint32_t mp1[PACK ? R / 2 : 1], mp2[PACK ? R / 2 : 1];
#pragma unroll
    for (int i = 0; i < (PACK ? R / 2 : 1); ++i) { mp1[i] = 0xFFFFFFFFu; mp2[i] = 0xFFFFFFFFu; }
    auto pack_insert = [&](const uint4* __restrict__ tp, int j, int rp) {
        const uint32_t da = dist_of(tp, j, 2 * rp), db = dist_of(tp, j, 2 * rp + 1);
        const uint32_t kp = imad(imad(db, p.w65536, da), p.w128, (uint32_t)j * 0x10001u);
        mp2[rp] = __vminu2(mp2[rp], __vmaxu2(mp1[rp], kp));
        mp1[rp] = __vminu2(mp1[rp], kp);
    };
    auto pack_flush = [&](uint32_t jbase) {
#pragma unroll
        for (int rp = 0; rp < (PACK ? R / 2 : 0); ++rp) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t k1 = h ? (mp1[rp] >> 16) : (mp1[rp] & 0xFFFFu);
                const uint32_t k2 = h ? (mp2[rp] >> 16) : (mp2[rp] & 0xFFFFu);
                const uint32_t g1 = (k1 == 0xFFFFu) ? TVG_SENTINEL_KEY : (((k1 >> 7) << TVG_KEY_SHIFT) | (jbase + (k1 & 127u)));
                const uint32_t g2 = (k2 == 0xFFFFu) ? TVG_SENTINEL_KEY : (((k2 >> 7) << TVG_KEY_SHIFT) | (jbase + (k2 & 127u)));
                top2_merge(m1[2 * rp + h], m2[2 * rp + h], g1, g2);
            }
            mp1[rp] = 0xFFFFFFFFu;
            mp2[rp] = 0xFFFFFFFFu;
        }
    };

    for (int c = 0; c < nchunks; ++c) {
        const int st = c % STAGES;
        if (tid == 0 && c + STAGES - 1 < nchunks) issue(c + STAGES - 1);
        mbar_wait(&full_bar[st], (uint32_t)((c / STAGES) & 1));
        const int rows = min(CH, nrows - c * CH);
        const uint4* __restrict__ tp = &tile[st][0];
        if (X) {
            // cooperative expansion of the landed rows: one row per thread
            for (int j = tid; j < rows; j += THREADS) {
                const uint4 a = tp[2 * j], d = tp[2 * j + 1];
                xtile[3 * j] = make_uint4(a.x, a.x ^ a.y, a.x ^ a.z, a.x ^ a.y ^ a.z);
                xtile[3 * j + 1] = make_uint4(a.w, a.w ^ d.x, a.w ^ d.y, a.w ^ d.x ^ d.y);
                xtile[3 * j + 2] = make_uint4(d.z, d.z ^ d.w, 0u, 0u);
            }
            __syncthreads();
        }
        const uint32_t jbase = (uint32_t)(row0 + c * CH);
        int j = 0;
        if (PACK) {
            for (; j + 4 <= rows; j += 4) {
#pragma unroll
                for (int rp = 0; rp < R / 2; ++rp) {
                    pack_insert(tp, j, rp);
                    pack_insert(tp, j + 1, rp);
                    pack_insert(tp, j + 2, rp);
                    pack_insert(tp, j + 3, rp);
                }
            }
            for (; j < rows; ++j) {
#pragma unroll
                for (int rp = 0; rp < R / 2; ++rp) pack_insert(tp, j, rp);
            }
            pack_flush(jbase);
        } else {
            for (; j + 4 <= rows; j += 4) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t k0 = key_of(tp, j, r, jbase + j);
                    const uint32_t k1 = key_of(tp, j + 1, r, jbase + j + 1);
                    const uint32_t k2 = key_of(tp, j + 2, r, jbase + j + 2);
                    const uint32_t k3 = key_of(tp, j + 3, r, jbase + j + 3);
                    top2_insert4<LAZY>(m1[r], m2[r], k0, k1, k2, k3);
                }
            }
            for (; j < rows; ++j) {
#pragma unroll
                for (int r = 0; r < R; ++r) top2_insert(m1[r], m2[r], key_of(tp, j, r, jbase + j));
            }
        }
        __syncthreads();  // every warp is done with stage st (and xtile) before they are refilled
    }
    knn2_finish<R, THREADS>(p, b, qt, split, q0, nq, m1, m2, &s_is_last);
}

template <int R, int THREADS, int CH, int STAGES, int VARIANT, bool LAZY, bool PACK, int MINB = 1>
static cudaError_t launch_cfg(const KnnParams& p, int batch, int qtiles, cudaStream_t stream) {
    dim3 grid(qtiles, p.splits, batch);
    knn2_kernel<R, THREADS, CH, STAGES, VARIANT, LAZY, PACK, MINB><<<grid, THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

int knn2_block_queries(int r) { return r * 128; }

// variant: 0..3 popcount form; +4 = lazy top-2 insertion (1, 3 only); +8 = packed 16-bit top-2 (1, 3 only)
cudaError_t launch_knn2(const KnnParams& p, int batch, int qtiles, int r, int variant, cudaStream_t stream) {
#define TVG_KNN_CASE(RR, VV, LZ, PK) \
    if (r == RR && variant == VV + (LZ ? 4 : 0) + (PK ? 8 : 0))  \
        return launch_cfg<RR, 128, 128, 3, VV, LZ, PK>(p, batch, qtiles, stream);
#define TVG_KNN_R(RR)                                                                                       \
    TVG_KNN_CASE(RR, 0, false, false) TVG_KNN_CASE(RR, 1, false, false) TVG_KNN_CASE(RR, 2, false, false)   \
    TVG_KNN_CASE(RR, 3, false, false) TVG_KNN_CASE(RR, 1, true, false) TVG_KNN_CASE(RR, 3, true, false)     \
    TVG_KNN_CASE(RR, 1, false, true) TVG_KNN_CASE(RR, 3, false, true)
    TVG_KNN_R(2) TVG_KNN_R(4) TVG_KNN_R(8)
#undef TVG_KNN_R
#undef TVG_KNN_CASE
    return cudaErrorInvalidValue;
}

}  // namespace tvg
