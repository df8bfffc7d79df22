// Ratio-test compaction + keypoint gather + K-normalisation.
//
// Replaces the Python loop of /root/reference/MonocularSlam/slam.py:130-141 (collect idx1 / idx2 in
// ASCENDING QUERY ORDER, gather curr_kps[idx1] / last_kps[idx2]) and the two normalize(K, .) calls of
// fit_essential_matrix (slam.py:151-152, normalize at :87-92).  A block-wide prefix sum keeps the
// reference's list order; no atomics.  One CTA per pair.
#include "tvg_common.cuh"
#include "tvg_kernels.h"

namespace tvg {

constexpr int COMPACT_THREADS = 1024;

__global__ void __launch_bounds__(COMPACT_THREADS) compact_kernel(const CompactParams p) {
    __shared__ int warp_tot[COMPACT_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fq = p.pair_q ? p.pair_q[b] : b;
    const int ft = p.pair_t ? p.pair_t[b] : b;
    const int nq = p.q_count ? min(p.q_count[fq], p.n_max) : p.nq;
    const size_t ob = (size_t)b * p.n_max;
    const uint8_t* keep = p.keep + ob;
    const int per = (nq + COMPACT_THREADS - 1) / COMPACT_THREADS;
    const int lo = min(nq, tid * per), hi = min(nq, lo + per);

    int cnt = 0;
    for (int i = lo; i < hi; ++i) cnt += keep[i] ? 1 : 0;
    // exclusive scan of cnt over the block
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = warp_tot[lane];
        int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += u;
        }
        warp_tot[lane] = s - v;  // exclusive warp offsets
        if (lane == 31) p.m_count[b] = s;
    }
    __syncthreads();
    int pos = warp_tot[warp] + incl - cnt;

    const double* kq = p.kps_q ? p.kps_q + (size_t)fq * p.n_max * 2 : nullptr;
    const double* kt = p.kps_t ? p.kps_t + (size_t)ft * p.n_max * 2 : nullptr;
    for (int i = lo; i < hi; ++i) {
        if (!keep[i]) continue;
        const int ti = p.idx[(ob + i) * 2];
        p.qidx[ob + pos] = i;
        p.tidx[ob + pos] = ti;
        if (kq) {
            const double cx = kq[2 * i], cy = kq[2 * i + 1];
            const double lx = kt[2 * ti], ly = kt[2 * ti + 1];
            // (Kinv @ [x, y, 1])[:2]
            const double sx = p.kinv[0] * lx + p.kinv[1] * ly + p.kinv[2];
            const double sy = p.kinv[3] * lx + p.kinv[4] * ly + p.kinv[5];
            const double dx = p.kinv[0] * cx + p.kinv[1] * cy + p.kinv[2];
            const double dy = p.kinv[3] * cx + p.kinv[4] * cy + p.kinv[5];
            p.src[(ob + pos) * 2] = sx;
            p.src[(ob + pos) * 2 + 1] = sy;
            p.dst[(ob + pos) * 2] = dx;
            p.dst[(ob + pos) * 2 + 1] = dy;
            p.corr32[ob + pos] = make_float4((float)sx, (float)sy, (float)dx, (float)dy);
        }
        ++pos;
    }
}

// ---- large-n form (n_max > COMPACT_BIG_N): a block owns COMPACT_BIG_ROWS consecutive queries; per-block keep counts,
// then every block scatters its survivors behind the counts of the blocks before it (same ascending order).
constexpr int COMPACT_BIG_N = 4096;
constexpr int COMPACT_BIG_ROWS = 2048;
constexpr int COMPACT_BIG_THREADS = 256;
int compact_big_blocks(int n_max) { return n_max > COMPACT_BIG_N ? (n_max + COMPACT_BIG_ROWS - 1) / COMPACT_BIG_ROWS : 0; }

__global__ void __launch_bounds__(COMPACT_BIG_THREADS) compact_count_kernel(const CompactParams p, int nblk) {
    __shared__ int s_w[COMPACT_BIG_THREADS / 32];
    const int blk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int fq = p.pair_q ? p.pair_q[b] : b;
    const int nq = p.q_count ? min(p.q_count[fq], p.n_max) : p.nq;
    const uint8_t* keep = p.keep + (size_t)b * p.n_max;
    const int lo = blk * COMPACT_BIG_ROWS, hi = min(nq, lo + COMPACT_BIG_ROWS);
    int cnt = 0;
    for (int i = lo + tid; i < hi; i += COMPACT_BIG_THREADS) cnt += keep[i] ? 1 : 0;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, m);
    if ((tid & 31) == 0) s_w[tid >> 5] = cnt;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < COMPACT_BIG_THREADS / 32; ++w) t += s_w[w];
        p.blk_cnt[(size_t)b * nblk + blk] = t;
    }
}

__global__ void __launch_bounds__(COMPACT_BIG_THREADS) compact_scatter_kernel(const CompactParams p, int nblk) {
    __shared__ int s_w[COMPACT_BIG_THREADS / 32];
    const int blk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fq = p.pair_q ? p.pair_q[b] : b;
    const int ft = p.pair_t ? p.pair_t[b] : b;
    const int nq = p.q_count ? min(p.q_count[fq], p.n_max) : p.nq;
    const size_t ob = (size_t)b * p.n_max;
    const uint8_t* keep = p.keep + ob;
    int base = 0, total = 0;
    for (int k = 0; k < nblk; ++k) {
        const int c = p.blk_cnt[(size_t)b * nblk + k];
        if (k < blk) base += c;
        total += c;
    }
    if (blk == 0 && tid == 0) p.m_count[b] = total;
    const double* kq = p.kps_q ? p.kps_q + (size_t)fq * p.n_max * 2 : nullptr;
    const double* kt = p.kps_t ? p.kps_t + (size_t)ft * p.n_max * 2 : nullptr;
    const int lo = blk * COMPACT_BIG_ROWS, hi = min(nq, lo + COMPACT_BIG_ROWS);
    int run = base;
    for (int i0 = lo; i0 < hi; i0 += COMPACT_BIG_THREADS) {
        const int i = i0 + tid;
        const bool in = i < hi && keep[i];
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_w[warp] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < COMPACT_BIG_THREADS / 32; ++w) {
            if (w < warp) off += s_w[w];
            tot += s_w[w];
        }
        if (in) {
            const int pos = run + off + __popc(bal & ((1u << lane) - 1u));
            const int ti = p.idx[(ob + i) * 2];
            p.qidx[ob + pos] = i;
            p.tidx[ob + pos] = ti;
            if (kq) {
                const double cx = kq[2 * i], cy = kq[2 * i + 1];
                const double lx = kt[2 * ti], ly = kt[2 * ti + 1];
                const double sx = p.kinv[0] * lx + p.kinv[1] * ly + p.kinv[2];
                const double sy = p.kinv[3] * lx + p.kinv[4] * ly + p.kinv[5];
                const double dx = p.kinv[0] * cx + p.kinv[1] * cy + p.kinv[2];
                const double dy = p.kinv[3] * cx + p.kinv[4] * cy + p.kinv[5];
                p.src[(ob + pos) * 2] = sx;
                p.src[(ob + pos) * 2 + 1] = sy;
                p.dst[(ob + pos) * 2] = dx;
                p.dst[(ob + pos) * 2 + 1] = dy;
                p.corr32[ob + pos] = make_float4((float)sx, (float)sy, (float)dx, (float)dy);
            }
        }
        run += tot;
        __syncthreads();
    }
}

cudaError_t launch_compact(const CompactParams& p, int batch, cudaStream_t stream) {
    const int nblk = p.blk_cnt ? compact_big_blocks(p.n_max) : 0;
    if (nblk > 0) {
        dim3 grid(nblk, batch);
        compact_count_kernel<<<grid, COMPACT_BIG_THREADS, 0, stream>>>(p, nblk);
        compact_scatter_kernel<<<grid, COMPACT_BIG_THREADS, 0, stream>>>(p, nblk);
    } else {
        compact_kernel<<<batch, COMPACT_THREADS, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace tvg
