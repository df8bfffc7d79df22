// Seeded RANSAC over 8-point essential-matrix hypotheses  --  replaces
// skimage.measure.ransac((norm_last, norm_curr), EssentialMatrixTransform, min_samples=8,
//                        residual_threshold=0.005, max_trials=H)
// at /root/reference/MonocularSlam/slam.py:155-159 (algorithm: SURVEY.md Appendix A.2 / A.3), with the
// per-trial rng.choice replaced by a host-supplied sample table so hypotheses equal the oracle's.
//
// Kernels
//   ransac_fit      one thread = one hypothesis, register-resident fp64: Hartley normalisation,
//                   8x9 DLT null vector (Householder), 3x3 one-sided Jacobi SVD, (s,s,0) projection,
//                   de-normalisation.  Emits E (fp64) and a Frobenius-normalised fp32 copy (SoA).
//   ransac_score    fp32 Sampson scoring: thread = hypothesis (E in registers), correspondence chunk
//                   resident in shared memory (TMA bulk copy), broadcast LDS.128 per correspondence.
//                   Per-thread count / uncertainty-band count / sum r^2 -> deterministic partials.
//   ransac_reduce   sums the partials, then lists every hypothesis whose count upper bound reaches the
//                   best lower bound (fp32 cannot be trusted closer than the band).
//   ransac_rescore  exact fp64 residuals for the listed candidates (count, sum r^2 over ALL points).
//   ransac_refit    picks best = (max count, min sum r^2, earliest), writes the PRE-refit inlier mask,
//                   compacts the inliers and accumulates the refit system A^T A (fp64, Hartley-normalised).
//   ransac_solve    one warp per pair: warp-parallel Jacobi eigen-solve of A^T A, (s,s,0) projection,
//                   de-normalisation, extract_Rt on the refit E.
#include "tvg_linalg.cuh"
#include "tvg_kernels.h"

namespace tvg {

// ------------------------------------------------------------------------------------ fit
constexpr int FIT_THREADS = 64;
#ifndef FIT_MIN_CTAS_N
#define FIT_MIN_CTAS_N 1
#endif
// resident CTAs per SM the register allocation aims at: 1 = no cap (154 registers, 12 warps per SM: 0.133 ms per 1 Mi
// hypotheses); 8 (128 registers, 16 warps, 46 doubles spilled) measured 0.152 ms
constexpr int FIT_MIN_CTAS = FIT_MIN_CTAS_N;

__device__ __forceinline__ bool hartley8(const double (&x)[8], const double (&y)[8], double& s, double& cx,
                                         double& cy) {
    double sx = 0.0, sy = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sx += x[k]; sy += y[k]; }
    cx = sx * 0.125;
    cy = sy * 0.125;
    double q = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double dx = x[k] - cx, dy = y[k] - cy;
        q = fma(dx, dx, q);
        q = fma(dy, dy, q);
    }
    const double ms = q * 0.125;                  // rms^2; s = sqrt(2) / rms
    if (!(ms > 0.0)) return false;
    s = 1.4142135623730951 * rsqrt(ms);
    return isfinite(s);
}

__global__ void __launch_bounds__(FIT_THREADS, FIT_MIN_CTAS) ransac_fit_kernel(const RansacParams p) {
    const int b = blockIdx.y;
    const int h = blockIdx.x * FIT_THREADS + threadIdx.x;
    if (h >= p.H) return;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const size_t ob = (size_t)b * p.n_max;
    const size_t oh = (size_t)b * p.H + h;

    int idx[8];
    bool ok = (M > 8);  // skimage: 0 < min_samples < num_samples
    if (ok) {
        if (p.samples) {
            const int32_t* s = p.samples + (size_t)b * p.sample_stride + (size_t)h * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                idx[k] = s[k];
                ok = ok && (idx[k] >= 0) && (idx[k] < M);
            }
        } else {
            const uint32_t* w = p.words + (size_t)b * p.sample_stride + (size_t)h * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int v = (int)(((unsigned long long)w[k] * (unsigned long long)M) >> 32);
                // bump until distinct from every earlier slot (slot order, +1 mod M)
                bool clash = true;
                while (clash) {
                    clash = false;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < k && idx[j] == v) clash = true;
                    if (clash) v = (v + 1 == M) ? 0 : v + 1;
                }
                idx[k] = v;
            }
        }
    }
    if (p.samples_out) {
#pragma unroll
        for (int k = 0; k < 8; ++k) p.samples_out[oh * 8 + k] = ok ? idx[k] : -1;
    }
    if (h == 0) {   // per-pair state of the grid-wide reduce (few pairs, many hypotheses)
        p.n_cand[b] = 0;
        if (p.lb) p.lb[b] = -1;
    }

    double E[9];
    if (ok) {
        double sx[8], sy[8], dx[8], dy[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double2 s = reinterpret_cast<const double2*>(p.src)[ob + idx[k]];
            const double2 d = reinterpret_cast<const double2*>(p.dst)[ob + idx[k]];
            sx[k] = s.x; sy[k] = s.y; dx[k] = d.x; dy[k] = d.y;
        }
        double ss, csx, csy, sd, cdx, cdy;
        ok = hartley8(sx, sy, ss, csx, csy);
        ok = hartley8(dx, dy, sd, cdx, cdy) && ok;
        if (ok) {
            double a[8][9];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const double xs = ss * sx[k] - ss * csx, ys = ss * sy[k] - ss * csy;
                const double xd = sd * dx[k] - sd * cdx, yd = sd * dy[k] - sd * cdy;
                a[k][0] = xd * xs; a[k][1] = xd * ys; a[k][2] = xd;
                a[k][3] = yd * xs; a[k][4] = yd * ys; a[k][5] = yd;
                a[k][6] = xs;      a[k][7] = ys;      a[k][8] = 1.0;
            }
            double n[9], En[9];
            null_vector_8x9(a, n);
            if (!essential_project_fast(n, En)) essential_project(n, En);
            essential_denormalise(En, ss, csx, csy, sd, cdx, cdy, E);
#pragma unroll
            for (int i = 0; i < 9; ++i) ok = ok && isfinite(E[i]);
        }
    }
    double fro = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        if (!ok) E[i] = __longlong_as_double(0x7ff8000000000000LL);
        fro = fma(E[i], E[i], fro);
    }
    const double inv = ok ? rsqrt(fro) : 0.0;
    p.valid[oh] = ok ? 1 : 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        p.E64[oh * 9 + i] = E[i];
        p.E32[((size_t)b * 9 + i) * p.H + h] = ok ? (float)(E[i] * inv) : 0.0f;
    }
}

cudaError_t launch_ransac_fit(const RansacParams& p, int batch, cudaStream_t stream) {
    dim3 grid((p.H + FIT_THREADS - 1) / FIT_THREADS, batch);
    ransac_fit_kernel<<<grid, FIT_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

// (fp32 scoring: ransac_score.cu)

// ------------------------------------------------------------------------------------ reduce + candidates
constexpr int REDUCE_THREADS = 1024;

__global__ void __launch_bounds__(REDUCE_THREADS) ransac_reduce_kernel(const RansacParams p) {
    __shared__ int s_red[32];
    __shared__ int s_lb;
    __shared__ int s_ncand;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const int nchunks = min(p.chunks, (M + p.chunk - 1) / p.chunk);
    const size_t oh = (size_t)b * p.H;
    int lb = -1;
    for (int h = tid; h < p.H; h += REDUCE_THREADS) {
        int c = 0, u = 0;
        float s = 0.0f;
        for (int k = 0; k < nchunks; ++k) {
            const size_t o = ((size_t)b * p.chunks + k) * p.H + h;
            const int2 v = p.part_cnt[o];
            c += v.x;
            u += v.y;
            s += p.part_sse[o];
        }
        const bool ok = p.valid[oh + h] != 0;
        p.counts[oh + h] = ok ? c : -1;
        p.band[oh + h] = ok ? u : 0;
        p.sse32[oh + h] = ok ? s : __int_as_float(0x7f800000);
        if (ok) lb = max(lb, c - u);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) lb = max(lb, __shfl_xor_sync(0xffffffffu, lb, m));
    if ((tid & 31) == 0) s_red[tid >> 5] = lb;
    if (tid == 0) s_ncand = 0;
    __syncthreads();
    if (tid < 32) {
        int v = s_red[tid];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, m));
        if (tid == 0) s_lb = v;
    }
    __syncthreads();
    lb = s_lb;
    for (int h = tid; h < p.H; h += REDUCE_THREADS) {
        const int c = p.counts[oh + h];
        if (c >= 0 && c + p.band[oh + h] >= lb) {
            const int slot = atomicAdd(&s_ncand, 1);
            p.cand[oh + slot] = h;
        }
    }
    __syncthreads();
    if (tid == 0) {
        p.n_cand[b] = s_ncand;
        if (b == 0) *p.score_queue = 0u;   // the scoring kernel's work queue, ready for the next launch
    }
}

// Grid-wide form for few pairs with many hypotheses (one huge pair: H = 65536): the per-hypothesis sums and the
// best lower bound (atomicMax), then the candidate list (atomicAdd slots: the list order is arbitrary, the selection
// rule downstream is a total order).  p.lb / p.n_cand are initialised by the fit kernel.
constexpr int REDUCE_BIG_THREADS = 256;

__global__ void __launch_bounds__(REDUCE_BIG_THREADS) ransac_reduce_big_sum_kernel(const RansacParams p) {
    __shared__ int s_red[REDUCE_BIG_THREADS / 32];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int h = blockIdx.x * REDUCE_BIG_THREADS + tid;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const int nchunks = min(p.chunks, (M + p.chunk - 1) / p.chunk);
    const size_t oh = (size_t)b * p.H;
    int lb = -1;
    if (h < p.H) {
        int c = 0, u = 0;
        float s = 0.0f;
        for (int k = 0; k < nchunks; ++k) {
            const size_t o = ((size_t)b * p.chunks + k) * p.H + h;
            const int2 v = p.part_cnt[o];
            c += v.x;
            u += v.y;
            s += p.part_sse[o];
        }
        const bool ok = p.valid[oh + h] != 0;
        p.counts[oh + h] = ok ? c : -1;
        p.band[oh + h] = ok ? u : 0;
        p.sse32[oh + h] = ok ? s : __int_as_float(0x7f800000);
        if (ok) lb = c - u;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) lb = max(lb, __shfl_xor_sync(0xffffffffu, lb, m));
    if ((tid & 31) == 0) s_red[tid >> 5] = lb;
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int w = 1; w < REDUCE_BIG_THREADS / 32; ++w) lb = max(lb, s_red[w]);
        if (lb >= 0) atomicMax(&p.lb[b], lb);
    }
}

__global__ void __launch_bounds__(REDUCE_BIG_THREADS) ransac_reduce_big_list_kernel(const RansacParams p) {
    const int b = blockIdx.y;
    const int h = blockIdx.x * REDUCE_BIG_THREADS + threadIdx.x;
    const size_t oh = (size_t)b * p.H;
    if (h < p.H) {
        const int c = p.counts[oh + h];
        if (c >= 0 && c + p.band[oh + h] >= p.lb[b]) p.cand[oh + atomicAdd(&p.n_cand[b], 1)] = h;
    }
    if (blockIdx.x == 0 && b == 0 && threadIdx.x == 0) *p.score_queue = 0u;
}

cudaError_t launch_ransac_reduce(const RansacParams& p, int batch, cudaStream_t stream) {
    if (p.lb) {
        dim3 grid((p.H + REDUCE_BIG_THREADS - 1) / REDUCE_BIG_THREADS, batch);
        ransac_reduce_big_sum_kernel<<<grid, REDUCE_BIG_THREADS, 0, stream>>>(p);
        ransac_reduce_big_list_kernel<<<grid, REDUCE_BIG_THREADS, 0, stream>>>(p);
        return cudaGetLastError();
    }
    ransac_reduce_kernel<<<batch, REDUCE_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}
int ransac_reduce_launches(const RansacParams& p) { return p.lb ? 2 : 1; }

// ------------------------------------------------------------------------------------ rescore (fp64)
constexpr int RESCORE_THREADS = 256;

template <int THREADS>
__device__ __forceinline__ double block_sum_f64(double v, double* s_buf) {
    v = warp_sum_f64(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) t += s_buf[w];
    return t;
}

__global__ void __launch_bounds__(RESCORE_THREADS) ransac_rescore_kernel(const RansacParams p) {
    __shared__ double s_buf[RESCORE_THREADS / 32];
    const int b = blockIdx.y;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const int ncand = p.n_cand[b];
    const size_t ob = (size_t)b * p.n_max, oh = (size_t)b * p.H;
    const double2* src = reinterpret_cast<const double2*>(p.src) + ob;
    const double2* dst = reinterpret_cast<const double2*>(p.dst) + ob;
    for (int ci = blockIdx.x; ci < ncand; ci += gridDim.x) {
        const int h = p.cand[oh + ci];
        double E[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) E[i] = p.E64[(oh + h) * 9 + i];
        double cnt = 0.0, sse = 0.0;
        for (int m = threadIdx.x; m < M; m += RESCORE_THREADS) {
            const double2 s = src[m], d = dst[m];
            const double r = sampson_residual(E, s.x, s.y, d.x, d.y);
            cnt += (r < p.thr) ? 1.0 : 0.0;
            sse = fma(r, r, sse);
        }
        cnt = block_sum_f64<RESCORE_THREADS>(cnt, s_buf);
        sse = block_sum_f64<RESCORE_THREADS>(sse, s_buf);
        if (threadIdx.x == 0) {
            p.cand_cnt[oh + h] = (int)cnt;
            p.cand_sse[oh + h] = sse;
            p.counts[oh + h] = (int)cnt;  // candidates report their exact count
        }
    }
}

// Large-M form (p.cand_part != NULL): a block re-scores one candidate over one range of REFIT_BIG_ROWS correspondences;
// the per-range (count, sum r^2) partials are summed in range order by ransac_select_big_kernel, so the exact sums do
// not depend on the launch geometry (nor on how many ranks share the hypotheses in tvg_huge_pair).
constexpr int RESCORE_BIG_ROWS = 2048;   // == REFIT_BIG_ROWS
constexpr int RESCORE_BIG_SLOTS = 64;

__global__ void __launch_bounds__(RESCORE_THREADS) ransac_rescore_big_kernel(const RansacParams p, int nblk) {
    __shared__ double s_buf[RESCORE_THREADS / 32];
    const int b = blockIdx.z, blk = blockIdx.y;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const int ncand = p.n_cand[b];
    const size_t ob = (size_t)b * p.n_max, oh = (size_t)b * p.H;
    const double2* src = reinterpret_cast<const double2*>(p.src) + ob;
    const double2* dst = reinterpret_cast<const double2*>(p.dst) + ob;
    const int lo = blk * RESCORE_BIG_ROWS, hi = min(M, lo + RESCORE_BIG_ROWS);
    for (int ci = blockIdx.x; ci < ncand; ci += gridDim.x) {
        const int h = p.cand[oh + ci];
        double E[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) E[i] = p.E64[(oh + h) * 9 + i];
        double cnt = 0.0, sse = 0.0;
        for (int m = lo + threadIdx.x; m < hi; m += RESCORE_THREADS) {
            const double2 s = src[m], d = dst[m];
            const double r = sampson_residual(E, s.x, s.y, d.x, d.y);
            cnt += (r < p.thr) ? 1.0 : 0.0;
            sse = fma(r, r, sse);
        }
        cnt = block_sum_f64<RESCORE_THREADS>(cnt, s_buf);
        sse = block_sum_f64<RESCORE_THREADS>(sse, s_buf);
        if (threadIdx.x == 0) p.cand_part[(oh + ci) * nblk + blk] = make_double2(cnt, sse);
    }
}

cudaError_t launch_ransac_rescore(const RansacParams& p, int batch, cudaStream_t stream) {
    if (p.cand_part) {
        const int nblk = (p.n_max + RESCORE_BIG_ROWS - 1) / RESCORE_BIG_ROWS > 0 ? (p.n_max + RESCORE_BIG_ROWS - 1) / RESCORE_BIG_ROWS : 1;
        dim3 grid(p.H < RESCORE_BIG_SLOTS ? p.H : RESCORE_BIG_SLOTS, nblk, batch);
        ransac_rescore_big_kernel<<<grid, RESCORE_THREADS, 0, stream>>>(p, nblk);
        return cudaGetLastError();
    }
    const int gx = batch >= 148 ? 4 : (p.H < 592 ? p.H : 592);
    dim3 grid(gx, batch);
    ransac_rescore_kernel<<<grid, RESCORE_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ select + mask + refit
constexpr int REFIT_THREADS = 256;

// skimage's selection rule as a TOTAL order, so that the winner does not depend on the order
// in which candidates are visited: most inliers, then smallest sum of squared residuals (NaN counts as +inf), then
// earliest trial.  The scan starts from skimage's initial state (0 inliers, rss = inf, no model): a hypothesis
// with 0 inliers is a model only when its rss is finite (`rss < inf`).
struct BestSel {
    int c, h;
    double s;
};
__device__ __forceinline__ BestSel best_init() { return BestSel{0, 0x7fffffff, __longlong_as_double(0x7ff0000000000000LL)}; }
__device__ __forceinline__ double sse_key(double s) { return (s == s) ? s : __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ void best_update(BestSel& b, int c, double s, int h) {
    s = sse_key(s);
    if ((c > b.c) || (c == b.c && (s < b.s || (s == b.s && h < b.h)))) { b.c = c; b.s = s; b.h = h; }
}
__device__ __forceinline__ bool best_is_model(const BestSel& b) {
    return b.h != 0x7fffffff && (b.c > 0 || b.s < __longlong_as_double(0x7ff0000000000000LL));
}
// block-wide reduction of per-thread BestSel (REFIT_THREADS threads); the result is valid on thread 0
__device__ __forceinline__ BestSel best_block_reduce(BestSel v, int* s_bc, int* s_bh, double* s_bs) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const int oc = __shfl_xor_sync(0xffffffffu, v.c, m);
        const int ohh = __shfl_xor_sync(0xffffffffu, v.h, m);
        const double os = shfl_xor_f64(v.s, m);
        best_update(v, oc, os, ohh);
    }
    if (lane == 0) { s_bc[warp] = v.c; s_bs[warp] = v.s; s_bh[warp] = v.h; }
    __syncthreads();
    if (tid == 0)
        for (int w = 1; w < REFIT_THREADS / 32; ++w) best_update(v, s_bc[w], s_bs[w], s_bh[w]);
    return v;
}
constexpr int REFIT_WS = 64;   // doubles of hand-off per pair between the accumulate and solve kernels
int ransac_refit_ws_doubles() { return REFIT_WS; }

// Smallest-eigenvalue eigenvector of a symmetric 9x9 matrix by parallel-order cyclic Jacobi on one warp.
// G and V live in shared memory; each round applies four disjoint rotations (round-robin tournament).
__device__ void jacobi_eig9_warp(double (*G)[9], double (*V)[9], double* cs /*[8]*/, double (&out)[9]) {
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < 81; i += 32) V[i / 9][i % 9] = (i / 9 == i % 9) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, dg = 0.0;
        for (int i = lane; i < 81; i += 32) {
            const int r = i / 9, c = i % 9;
            const double g = G[r][c];
            if (r == c) dg = fma(g, g, dg); else off = fma(g, g, off);
        }
        off = warp_sum_f64(off);
        dg = warp_sum_f64(dg);
        if (!(off > 1e-30 * dg)) break;
        for (int round = 0; round < 9; ++round) {
            if (lane < 4) {
                const int k = lane + 1;
                const int pp = (round + k) % 9, qq = (round + 9 - k) % 9;
                const double app = G[pp][pp], aqq = G[qq][qq], apq = G[pp][qq];
                double c = 1.0, s = 0.0;
                if (apq != 0.0) {
                    const double theta = (aqq - app) / (2.0 * apq);
                    const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                    c = rsqrt(fma(t, t, 1.0));
                    s = t * c;
                }
                cs[2 * lane] = c;
                cs[2 * lane + 1] = s;
            }
            __syncwarp();
            // columns: G <- G J, V <- V J      (72 element-pair tasks)
            for (int task = lane; task < 72; task += 32) {
                const int which = task / 36, rem = task % 36;
                const int k = rem / 9, i = rem % 9;
                const int pp = (round + k + 1) % 9, qq = (round + 9 - (k + 1)) % 9;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                double(*Mx)[9] = which ? V : G;
                const double a = Mx[i][pp], bq = Mx[i][qq];
                Mx[i][pp] = c * a - s * bq;
                Mx[i][qq] = s * a + c * bq;
            }
            __syncwarp();
            // rows: G <- J^T G
            for (int task = lane; task < 36; task += 32) {
                const int k = task / 9, j = task % 9;
                const int pp = (round + k + 1) % 9, qq = (round + 9 - (k + 1)) % 9;
                const double c = cs[2 * k], s = cs[2 * k + 1];
                const double a = G[pp][j], bq = G[qq][j];
                G[pp][j] = c * a - s * bq;
                G[qq][j] = s * a + c * bq;
            }
            __syncwarp();
        }
    }
    int best = 0;
    double lam = G[0][0];
    for (int i = 1; i < 9; ++i)
        if (G[i][i] < lam) { lam = G[i][i]; best = i; }
    for (int i = 0; i < 9; ++i) out[i] = V[i][best];
}

// The same eigenvector by inverse iteration on the Cholesky factor of G + delta I (delta = 1e-13 tr G keeps the factor
// positive when the smallest eigenvalue is numerically zero, i.e. on noise-free data; the shift does not move the
// eigenvectors).  One thread, everything in registers: g = the upper triangle of G in row order (45 values, the layout
// of refit_ws), L packed the same way by rows of the lower triangle, every index static.  The well-separated case --
// an inlier set explained by one essential matrix -- converges in 3-10 solves of 72 multiply-adds; returns false (the
// caller falls back to the Jacobi sweep) when the two smallest eigenvalues are too close to separate in 40 solves or
// the factorisation breaks down.
__device__ __forceinline__ constexpr int tri_lo(int i, int j) { return i * (i + 1) / 2 + j; }          // j <= i
__device__ __forceinline__ constexpr int tri_up(int i, int j) { return i * 9 - i * (i - 1) / 2 + (j - i); }  // j >= i

__device__ __forceinline__ bool eig9_smallest_inverse_iteration(const double (&g)[45], double (&x)[9]) {
    double tr = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) tr += g[tri_up(i, i)];
    if (!(tr > 0.0) || !isfinite(tr)) return false;
    const double delta = 1e-13 * tr;
    double L[45], idiag[9];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        double d = g[tri_up(j, j)] + delta;
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-L[tri_lo(j, k)], L[tri_lo(j, k)], d);
        ok = ok && (d > 0.0);
        const double inv = rsqrt(d);
        idiag[j] = inv;
        L[tri_lo(j, j)] = d * inv;
#pragma unroll
        for (int i = j + 1; i < 9; ++i) {
            double v = g[tri_up(j, i)];
#pragma unroll
            for (int k = 0; k < j; ++k) v = fma(-L[tri_lo(i, k)], L[tri_lo(j, k)], v);
            L[tri_lo(i, j)] = v * inv;
        }
    }
    if (!ok) return false;
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = 1.0 / 3.0;
    for (int it = 0; it < 40; ++it) {
        double y[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {            // L y = x
            double v = x[i];
#pragma unroll
            for (int k = 0; k < i; ++k) v = fma(-L[tri_lo(i, k)], y[k], v);
            y[i] = v * idiag[i];
        }
#pragma unroll
        for (int i = 8; i >= 0; --i) {           // L^T z = y   (z overwrites y)
            double v = y[i];
#pragma unroll
            for (int k = i + 1; k < 9; ++k) v = fma(-L[tri_lo(k, i)], y[k], v);
            y[i] = v * idiag[i];
        }
        double nn = 0.0, dot = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) { nn = fma(y[i], y[i], nn); dot = fma(y[i], x[i], dot); }
        if (!(nn > 0.0) || !isfinite(nn)) return false;
        const double sc = (dot < 0.0 ? -1.0 : 1.0) * rsqrt(nn);
        double diff = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const double z = y[i] * sc;
            diff = fmax(diff, fabs(z - x[i]));
            x[i] = z;
        }
        if (diff < 2e-15) return true;
    }
    return false;
}

__global__ void __launch_bounds__(REFIT_THREADS, 2) ransac_refit_kernel(const RansacParams p) {
    __shared__ double s_buf[REFIT_THREADS / 32];
    __shared__ double s_acc[45][REFIT_THREADS / 32];
    __shared__ int s_cnt[REFIT_THREADS / 32], s_bc[REFIT_THREADS / 32], s_bh[REFIT_THREADS / 32];
    __shared__ double s_bs[REFIT_THREADS / 32];
    __shared__ int s_best;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const size_t ob = (size_t)b * p.n_max, oh = (size_t)b * p.H;
    const int ncand = p.n_cand[b];
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    // ---- best = (max count, min sum r^2, earliest) over the exactly re-scored candidates
    BestSel sel = best_init();
    for (int ci = tid; ci < ncand; ci += REFIT_THREADS) {
        const int h = p.cand[oh + ci];
        best_update(sel, p.cand_cnt[oh + h], p.cand_sse[oh + h], h);
    }
    sel = best_block_reduce(sel, s_bc, s_bh, s_bs);
    if (tid == 0) {
        const bool have = best_is_model(sel);
        s_best = have ? sel.h : -1;
        p.best[b] = s_best;
        p.best_sse[b] = have ? p.cand_sse[oh + sel.h] : qnan;
    }
    __syncthreads();
    const int best = s_best;
    if (best < 0) {  // no valid hypothesis (fewer than 9 correspondences or every fit degenerate)
        for (int m = tid; m < M; m += REFIT_THREADS) p.mask[ob + m] = 0;
        if (tid == 0) {
            p.n_inl[b] = 0;
            p.status[b] = (M <= 8) ? 1 : 2;
            for (int i = 0; i < 9; ++i) { p.E[b * 9 + i] = qnan; p.E_best[b * 9 + i] = qnan; }
            for (int i = 0; i < 16; ++i) p.Rt[b * 16 + i] = qnan;
            p.refit_ws[(size_t)b * REFIT_WS + 51] = 0.0;
        }
        return;
    }
    double Eb[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Eb[i] = p.E64[(oh + best) * 9 + i];

    // ---- pre-refit inlier mask (fp64), ordered compaction, centroid sums
    const double2* src = reinterpret_cast<const double2*>(p.src) + ob;
    const double2* dst = reinterpret_cast<const double2*>(p.dst) + ob;
    const int per = (M + REFIT_THREADS - 1) / REFIT_THREADS;
    const int lo = min(M, tid * per), hi = min(M, lo + per);
    // (a thread owns at most REFIT_BIG_N / REFIT_THREADS = 16 consecutive rows: its inlier flags live in one word;
    // rows are loaded four at a time so that the loads of a group are in flight together)
    int cnt = 0;
    unsigned bits = 0u;
    double sum[4] = {0.0, 0.0, 0.0, 0.0};
    for (int m0 = lo; m0 < hi; m0 += 4) {
        double2 s4[4], d4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (m0 + k < hi) { s4[k] = __ldg(src + m0 + k); d4[k] = __ldg(dst + m0 + k); }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (m0 + k >= hi) break;
            const double2 s = s4[k], d = d4[k];
            const bool in = sampson_residual(Eb, s.x, s.y, d.x, d.y) < p.thr;
            p.mask[ob + m0 + k] = in ? 1 : 0;
            if (in) {
                bits |= 1u << (m0 + k - lo);
                ++cnt; sum[0] += s.x; sum[1] += s.y; sum[2] += d.x; sum[3] += d.y;
            }
        }
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_cnt[warp] = incl;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < REFIT_THREADS / 32; ++w) {
        if (w < warp) base += s_cnt[w];
        total += s_cnt[w];
    }
    {
        int pos = base + incl - cnt;
        const int fq = p.pair_q ? p.pair_q[b] : b;
        const int ft = p.pair_t ? p.pair_t[b] : b;
        const double2* kq = p.kps_q ? reinterpret_cast<const double2*>(p.kps_q) + (size_t)fq * p.n_max : nullptr;
        const double2* kt = p.kps_t ? reinterpret_cast<const double2*>(p.kps_t) + (size_t)ft * p.n_max : nullptr;
        for (unsigned rem = bits; rem;) {       // four inliers at a time: index loads, then gathers, then stores
            int mm[4], qi[4], ti[4];
            double2 a4[4], b4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mm[k] = rem ? lo + __ffs(rem) - 1 : -1;
                rem &= rem - 1u;
            }
            if (kq) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (mm[k] >= 0) { qi[k] = __ldg(p.qidx + ob + mm[k]); ti[k] = __ldg(p.tidx + ob + mm[k]); }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (mm[k] >= 0) { a4[k] = __ldg(kq + qi[k]); b4[k] = __ldg(kt + ti[k]); }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mm[k] < 0) break;
                p.inl_idx[ob + pos] = mm[k];
                if (kq) {
                    reinterpret_cast<double2*>(p.pix_curr)[ob + pos] = a4[k];
                    reinterpret_cast<double2*>(p.pix_last)[ob + pos] = b4[k];
                }
                ++pos;
            }
        }
    }
    if (tid == 0) {
        p.n_inl[b] = total;
        p.status[b] = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) p.E_best[b * 9 + i] = Eb[i];
    }
    double* ws = p.refit_ws + (size_t)b * REFIT_WS;
    if (total < 8) {  // skimage's refit would raise; report the un-refit model
        if (tid == 0) {
            ws[51] = 1.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) ws[52 + i] = Eb[i];
        }
        return;
    }
    // ---- Hartley normalisation of the inlier set
    double c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = block_sum_f64<REFIT_THREADS>(sum[k], s_buf) / (double)total;
    double qs = 0.0, qd = 0.0;
    for (unsigned rem = bits; rem;) {
        int mm[4];
        double2 s4[4], d4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mm[k] = rem ? lo + __ffs(rem) - 1 : -1;
            rem &= rem - 1u;
            if (mm[k] >= 0) { s4[k] = __ldg(src + mm[k]); d4[k] = __ldg(dst + mm[k]); }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (mm[k] < 0) break;
            const double2 s = s4[k], d = d4[k];
            qs = fma(s.x - c[0], s.x - c[0], qs); qs = fma(s.y - c[1], s.y - c[1], qs);
            qd = fma(d.x - c[2], d.x - c[2], qd); qd = fma(d.y - c[3], d.y - c[3], qd);
        }
    }
    qs = block_sum_f64<REFIT_THREADS>(qs, s_buf);
    qd = block_sum_f64<REFIT_THREADS>(qd, s_buf);
    const double ss = 1.4142135623730951 / sqrt(qs / (double)total);
    const double sd = 1.4142135623730951 / sqrt(qd / (double)total);
    // ---- A^T A (upper triangle, 45 sums) in fp64
    double acc[45];
#pragma unroll
    for (int i = 0; i < 45; ++i) acc[i] = 0.0;
    double2 s_nx = make_double2(0.0, 0.0), d_nx = s_nx;
    unsigned rem = bits;
    if (rem) { const int m = lo + __ffs(rem) - 1; s_nx = __ldg(src + m); d_nx = __ldg(dst + m); }
    while (rem) {
        const double2 s = s_nx, d = d_nx;
        rem &= rem - 1u;
        if (rem) { const int m = lo + __ffs(rem) - 1; s_nx = __ldg(src + m); d_nx = __ldg(dst + m); }   // next row in flight
        const double xs = ss * s.x - ss * c[0], ys = ss * s.y - ss * c[1];
        const double xd = sd * d.x - sd * c[2], yd = sd * d.y - sd * c[3];
        const double a[9] = {xd * xs, xd * ys, xd, yd * xs, yd * ys, yd, xs, ys, 1.0};
        int t = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = i; j < 9; ++j) { acc[t] = fma(a[i], a[j], acc[t]); ++t; }
    }
#pragma unroll
    for (int i = 0; i < 45; ++i) {
        const double v = warp_sum_f64(acc[i]);
        if (lane == 0) s_acc[i][warp] = v;
    }
    __syncthreads();
    if (tid < 45) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < REFIT_THREADS / 32; ++w) v += s_acc[tid][w];
        ws[tid] = v;
    }
    if (tid == 0) {
        ws[45] = ss; ws[46] = c[0]; ws[47] = c[1]; ws[48] = sd; ws[49] = c[2]; ws[50] = c[3];
        ws[51] = 2.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) ws[52 + i] = Eb[i];
    }
}

// ---- large-M form of the refit accumulation (n_max > REFIT_BIG_N): the same steps as ransac_refit_kernel cut into
// grid-wide phases so that a single huge pair (C3: M ~ 6.5 k, C5: M ~ 52 k) is not walked by one CTA.  A block owns
// REFIT_BIG_ROWS consecutive correspondences; every cross-block quantity is a fixed-order sum of per-block partials,
// so the result does not depend on the launch (and not on the number of ranks in dist.huge_pair).
//   select  (1 CTA / pair)      best hypothesis, "no model" outputs
//   mask    (blocks x pairs)    pre-refit inlier mask, per-block inlier count and coordinate sums
//   scatter (blocks x pairs)    ordered inlier list + pixel gather, per-block squared distances to the centroid
//   ata     (blocks x pairs)    per-block A^T A (45 sums)
//   combine (1 CTA / pair)      fixed-order sums -> refit_ws, the hand-off ransac_solve_kernel reads
constexpr int REFIT_BIG_N = 4096;
constexpr int REFIT_BIG_ROWS = 2048;
constexpr int REFIT_BIG_STRIDE = 64;    // doubles per (pair, block): [0] count, [1..4] sums, [5..6] q sums, [8..52] A^T A
static int refit_nblk(int n_max) { return n_max > REFIT_BIG_ROWS ? (n_max + REFIT_BIG_ROWS - 1) / REFIT_BIG_ROWS : 1; }
static_assert(REFIT_BIG_N <= 32 * REFIT_THREADS, "ransac_refit_kernel keeps a thread's inlier flags in one 32-bit word");
int ransac_refit_big_blocks(int n_max, bool force) { return (n_max > REFIT_BIG_N || force) ? refit_nblk(n_max) : 0; }
int ransac_refit_big_stride() { return REFIT_BIG_STRIDE; }

// "no model" outputs of a pair (fewer than 9 correspondences, or every hypothesis degenerate)
__device__ void write_no_model(const RansacParams& p, int b, int M) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    p.best[b] = -1;
    p.best_sse[b] = qnan;
    p.n_inl[b] = 0;
    p.status[b] = (M <= 8) ? 1 : 2;
    for (int i = 0; i < 9; ++i) { p.E[b * 9 + i] = qnan; p.E_best[b * 9 + i] = qnan; }
    for (int i = 0; i < 16; ++i) p.Rt[b * 16 + i] = qnan;
    p.refit_ws[(size_t)b * REFIT_WS + 51] = 0.0;
}

// Sums the per-range re-score partials of every candidate in range order, applies the selection rule and publishes
// the winner: p.best / p.best_sse / p.E_best, or -- when the hypotheses of this pair are shared between ranks
// (p.rec != NULL, tvg_huge_pair) -- this rank's record {count, global hypothesis index, sum r^2, E} for the exchange.
__global__ void __launch_bounds__(REFIT_THREADS) ransac_select_big_kernel(const RansacParams p, int nblk) {
    __shared__ int s_bc[REFIT_THREADS / 32], s_bh[REFIT_THREADS / 32];
    __shared__ double s_bs[REFIT_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const size_t oh = (size_t)b * p.H;
    const int ncand = p.n_cand[b];
    BestSel sel = best_init();
    for (int ci = tid; ci < ncand; ci += REFIT_THREADS) {
        const int h = p.cand[oh + ci];
        double cnt = 0.0, sse = 0.0;
        for (int k = 0; k < nblk; ++k) {
            const double2 v = p.cand_part[(oh + ci) * nblk + k];
            cnt += v.x;
            sse += v.y;
        }
        p.cand_cnt[oh + h] = (int)cnt;
        p.cand_sse[oh + h] = sse;
        p.counts[oh + h] = (int)cnt;   // candidates report their exact count
        best_update(sel, (int)cnt, sse, h);
    }
    sel = best_block_reduce(sel, s_bc, s_bh, s_bs);
    if (tid == 0) {
        const bool have = best_is_model(sel);
        if (p.rec) {
            BestRec r;
            r.count = have ? sel.c : 0;
            r.hyp = have ? sel.h + p.h_off : -1;
            r.sse = have ? p.cand_sse[oh + sel.h] : 0.0;
            for (int i = 0; i < 9; ++i) r.E[i] = have ? p.E64[(oh + sel.h) * 9 + i] : 0.0;
            r.pad = 0.0;
            p.rec[b] = r;
        } else if (have) {
            p.best[b] = sel.h;
            p.best_sse[b] = p.cand_sse[oh + sel.h];
            for (int i = 0; i < 9; ++i) p.E_best[b * 9 + i] = p.E64[(oh + sel.h) * 9 + i];
        } else {
            write_no_model(p, b, M);
        }
    }
}

// After the exchange: the global winner among the ranks' records (same total order) becomes p.best / p.E_best on
// every rank.  One pair per launch (tvg_huge_pair).
__global__ void ransac_merge_records_kernel(const RansacParams p, const BestRec* recs, int world) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int M = p.m_count ? min(p.m_count[0], p.n_max) : p.m;
    BestSel sel = best_init();
    int who = -1;
    for (int r = 0; r < world; ++r) {
        if (recs[r].hyp < 0) continue;
        const int before = sel.h;
        best_update(sel, recs[r].count, recs[r].sse, recs[r].hyp);
        if (sel.h != before) who = r;
    }
    if (who < 0) {
        write_no_model(p, 0, M);
        return;
    }
    p.best[0] = recs[who].hyp;
    p.best_sse[0] = recs[who].sse;
    for (int i = 0; i < 9; ++i) p.E_best[i] = recs[who].E[i];
}

__global__ void __launch_bounds__(REFIT_THREADS) ransac_refit_big_mask_kernel(const RansacParams p, int nblk) {
    __shared__ double s_buf[REFIT_THREADS / 32];
    const int blk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const size_t ob = (size_t)b * p.n_max;
    const int best = p.best[b];
    const int lo = blk * REFIT_BIG_ROWS, hi = min(M, lo + REFIT_BIG_ROWS);
    double* out = p.refit_big + ((size_t)b * nblk + blk) * REFIT_BIG_STRIDE;
    if (best < 0) {
        for (int m = lo + tid; m < hi; m += REFIT_THREADS) p.mask[ob + m] = 0;
        return;
    }
    double Eb[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Eb[i] = p.E_best[b * 9 + i];
    const double2* src = reinterpret_cast<const double2*>(p.src) + ob;
    const double2* dst = reinterpret_cast<const double2*>(p.dst) + ob;
    double cnt = 0.0, sum[4] = {0.0, 0.0, 0.0, 0.0};
    for (int m = lo + tid; m < hi; m += REFIT_THREADS) {
        const double2 s = src[m], d = dst[m];
        const bool in = sampson_residual(Eb, s.x, s.y, d.x, d.y) < p.thr;
        p.mask[ob + m] = in ? 1 : 0;
        if (in) { cnt += 1.0; sum[0] += s.x; sum[1] += s.y; sum[2] += d.x; sum[3] += d.y; }
    }
    cnt = block_sum_f64<REFIT_THREADS>(cnt, s_buf);
#pragma unroll
    for (int k = 0; k < 4; ++k) sum[k] = block_sum_f64<REFIT_THREADS>(sum[k], s_buf);
    if (tid == 0) {
        out[0] = cnt;
#pragma unroll
        for (int k = 0; k < 4; ++k) out[1 + k] = sum[k];
    }
}

// inlier total and centroid of a pair from the per-block partials (same order in every kernel)
__device__ __forceinline__ void refit_big_stats(const double* part, int nblk, int blk, int& base, int& total, double (&c)[4]) {
    base = 0; total = 0;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < nblk; ++k) {
        const double* q = part + (size_t)k * REFIT_BIG_STRIDE;
        const int n = (int)q[0];
        if (k < blk) base += n;
        total += n;
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] += q[1 + i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = s[i] / (double)total;
}

__global__ void __launch_bounds__(REFIT_THREADS) ransac_refit_big_scatter_kernel(const RansacParams p, int nblk) {
    __shared__ double s_buf[REFIT_THREADS / 32];
    __shared__ int s_wcnt[REFIT_THREADS / 32];
    const int blk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const size_t ob = (size_t)b * p.n_max;
    const int best = p.best[b];
    if (best < 0) return;
    double* part = p.refit_big + (size_t)b * nblk * REFIT_BIG_STRIDE;
    int base, total;
    double c[4];
    refit_big_stats(part, nblk, blk, base, total, c);
    if (blk == 0 && tid == 0) {
        p.n_inl[b] = total;
        p.status[b] = 0;
    }
    const double2* src = reinterpret_cast<const double2*>(p.src) + ob;
    const double2* dst = reinterpret_cast<const double2*>(p.dst) + ob;
    const int fq = p.pair_q ? p.pair_q[b] : b;
    const int ft = p.pair_t ? p.pair_t[b] : b;
    const double2* kq = p.kps_q ? reinterpret_cast<const double2*>(p.kps_q) + (size_t)fq * p.n_max : nullptr;
    const double2* kt = p.kps_t ? reinterpret_cast<const double2*>(p.kps_t) + (size_t)ft * p.n_max : nullptr;
    const int lo = blk * REFIT_BIG_ROWS, hi = min(M, lo + REFIT_BIG_ROWS);
    double qs = 0.0, qd = 0.0;
    int run = base;
    for (int m0 = lo; m0 < hi; m0 += REFIT_THREADS) {   // ordered compaction, 256 consecutive rows per round
        const int m = m0 + tid;
        const bool in = m < hi && p.mask[ob + m];
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < REFIT_THREADS / 32; ++w) {
            if (w < warp) off += s_wcnt[w];
            tot += s_wcnt[w];
        }
        if (in) {
            const int pos = run + off + __popc(bal & ((1u << lane) - 1u));
            p.inl_idx[ob + pos] = m;
            if (kq) {
                reinterpret_cast<double2*>(p.pix_curr)[ob + pos] = kq[p.qidx[ob + m]];
                reinterpret_cast<double2*>(p.pix_last)[ob + pos] = kt[p.tidx[ob + m]];
            }
            const double2 s = src[m], d = dst[m];
            qs = fma(s.x - c[0], s.x - c[0], qs); qs = fma(s.y - c[1], s.y - c[1], qs);
            qd = fma(d.x - c[2], d.x - c[2], qd); qd = fma(d.y - c[3], d.y - c[3], qd);
        }
        run += tot;
        __syncthreads();
    }
    qs = block_sum_f64<REFIT_THREADS>(qs, s_buf);
    qd = block_sum_f64<REFIT_THREADS>(qd, s_buf);
    if (tid == 0) {
        double* out = part + (size_t)blk * REFIT_BIG_STRIDE;
        out[5] = qs;
        out[6] = qd;
    }
}

__global__ void __launch_bounds__(REFIT_THREADS) ransac_refit_big_ata_kernel(const RansacParams p, int nblk) {
    __shared__ double s_acc[45][REFIT_THREADS / 32];
    const int blk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const size_t ob = (size_t)b * p.n_max;
    if (p.best[b] < 0) return;
    double* part = p.refit_big + (size_t)b * nblk * REFIT_BIG_STRIDE;
    int base, total;
    double c[4];
    refit_big_stats(part, nblk, blk, base, total, c);
    if (total < 8) return;
    double qs = 0.0, qd = 0.0;
    for (int k = 0; k < nblk; ++k) { qs += part[(size_t)k * REFIT_BIG_STRIDE + 5]; qd += part[(size_t)k * REFIT_BIG_STRIDE + 6]; }
    const double ss = 1.4142135623730951 / sqrt(qs / (double)total);
    const double sd = 1.4142135623730951 / sqrt(qd / (double)total);
    const double2* src = reinterpret_cast<const double2*>(p.src) + ob;
    const double2* dst = reinterpret_cast<const double2*>(p.dst) + ob;
    const int lo = blk * REFIT_BIG_ROWS, hi = min(M, lo + REFIT_BIG_ROWS);
    double acc[45];
#pragma unroll
    for (int i = 0; i < 45; ++i) acc[i] = 0.0;
    for (int m = lo + tid; m < hi; m += REFIT_THREADS) {
        if (!p.mask[ob + m]) continue;
        const double2 s = src[m], d = dst[m];
        const double xs = ss * s.x - ss * c[0], ys = ss * s.y - ss * c[1];
        const double xd = sd * d.x - sd * c[2], yd = sd * d.y - sd * c[3];
        const double a[9] = {xd * xs, xd * ys, xd, yd * xs, yd * ys, yd, xs, ys, 1.0};
        int t = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = i; j < 9; ++j) { acc[t] = fma(a[i], a[j], acc[t]); ++t; }
    }
#pragma unroll
    for (int i = 0; i < 45; ++i) {
        const double v = warp_sum_f64(acc[i]);
        if (lane == 0) s_acc[i][warp] = v;
    }
    __syncthreads();
    if (tid < 45) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < REFIT_THREADS / 32; ++w) v += s_acc[tid][w];
        part[(size_t)blk * REFIT_BIG_STRIDE + 8 + tid] = v;
    }
}

__global__ void __launch_bounds__(64) ransac_refit_big_combine_kernel(const RansacParams p, int nblk) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int best = p.best[b];
    if (best < 0) return;
    const double* part = p.refit_big + (size_t)b * nblk * REFIT_BIG_STRIDE;
    double* ws = p.refit_ws + (size_t)b * REFIT_WS;
    int base, total;
    double c[4];
    refit_big_stats(part, nblk, 0, base, total, c);
    if (total < 8) {
        if (tid == 0) {
            ws[51] = 1.0;
            for (int i = 0; i < 9; ++i) ws[52 + i] = p.E_best[b * 9 + i];
        }
        return;
    }
    if (tid < 45) {
        double v = 0.0;
        for (int k = 0; k < nblk; ++k) v += part[(size_t)k * REFIT_BIG_STRIDE + 8 + tid];
        ws[tid] = v;
    }
    if (tid == 0) {
        double qs = 0.0, qd = 0.0;
        for (int k = 0; k < nblk; ++k) { qs += part[(size_t)k * REFIT_BIG_STRIDE + 5]; qd += part[(size_t)k * REFIT_BIG_STRIDE + 6]; }
        ws[45] = 1.4142135623730951 / sqrt(qs / (double)total); ws[46] = c[0]; ws[47] = c[1];
        ws[48] = 1.4142135623730951 / sqrt(qd / (double)total); ws[49] = c[2]; ws[50] = c[3];
        ws[51] = 2.0;
        for (int i = 0; i < 9; ++i) ws[52 + i] = p.E_best[b * 9 + i];
    }
}

// One warp per pair: smallest eigenvector of A^T A, (s,s,0) projection, de-normalisation, extract_Rt.
// Split from the accumulation kernel so that all pairs' (latency-bound) solves run concurrently.
constexpr int SOLVE_WARPS = 4;

__global__ void __launch_bounds__(SOLVE_WARPS * 32) ransac_solve_kernel(const RansacParams p, int batch) {
    __shared__ double s_G[SOLVE_WARPS][9][9], s_V[SOLVE_WARPS][9][9], s_cs[SOLVE_WARPS][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * SOLVE_WARPS + warp;
    if (b >= batch) return;
    const double* ws = p.refit_ws + (size_t)b * REFIT_WS;
    const int flag = (int)ws[51];
    if (flag == 0) return;   // no model: NaNs already written
    double E[9];
    if (flag == 2) {
        double n[9];
        int conv = 0;
        if (lane == 0) {
            double g[45];
#pragma unroll
            for (int t = 0; t < 45; ++t) g[t] = ws[t];
            conv = eig9_smallest_inverse_iteration(g, n) ? 1 : 0;
        }
        conv = __shfl_sync(0xffffffffu, conv, 0);
        if (!conv) {   // near-degenerate inlier sets: the warp-wide Jacobi sweep
            for (int t = lane; t < 45; t += 32) {
                int i = 0, rem = t;
                while (rem >= 9 - i) { rem -= 9 - i; ++i; }
                const int j = i + rem;
                const double v = ws[t];
                s_G[warp][i][j] = v;
                s_G[warp][j][i] = v;
            }
            __syncwarp();
            jacobi_eig9_warp(s_G[warp], s_V[warp], s_cs[warp], n);
        }
        if (lane == 0) {
            double En[9];
            if (!essential_project_fast(n, En)) essential_project(n, En);
            essential_denormalise(En, ws[45], ws[46], ws[47], ws[48], ws[49], ws[50], E);
            bool fin = true;
#pragma unroll
            for (int i = 0; i < 9; ++i) fin = fin && isfinite(E[i]);
            if (!fin) {
#pragma unroll
                for (int i = 0; i < 9; ++i) E[i] = ws[52 + i];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) E[i] = ws[52 + i];
    }
    if (lane == 0) {
        double Rt[16];
        extract_rt(E, Rt);
#pragma unroll
        for (int i = 0; i < 9; ++i) p.E[b * 9 + i] = E[i];
#pragma unroll
        for (int i = 0; i < 16; ++i) p.Rt[b * 16 + i] = Rt[i];
    }
}

// big path only: candidate sums + selection (publishes the winner, or this rank's record when p.rec is set)
cudaError_t launch_ransac_select_big(const RansacParams& p, int batch, cudaStream_t stream) {
    ransac_select_big_kernel<<<batch, REFIT_THREADS, 0, stream>>>(p, refit_nblk(p.n_max));
    return cudaGetLastError();
}

cudaError_t launch_ransac_merge_records(const RansacParams& p, const BestRec* recs, int world, cudaStream_t stream) {
    ransac_merge_records_kernel<<<1, 32, 0, stream>>>(p, recs, world);
    return cudaGetLastError();
}

// mask + inlier compaction + refit accumulation + solve.  Big path (p.refit_big): p.best / p.E_best are already
// published (launch_ransac_select_big, or launch_ransac_merge_records after an exchange).
cudaError_t launch_ransac_refit(const RansacParams& p, int batch, cudaStream_t stream) {
    if (p.refit_big) {
        const int nblk = refit_nblk(p.n_max);
        dim3 grid(nblk, batch);
        ransac_refit_big_mask_kernel<<<grid, REFIT_THREADS, 0, stream>>>(p, nblk);
        ransac_refit_big_scatter_kernel<<<grid, REFIT_THREADS, 0, stream>>>(p, nblk);
        ransac_refit_big_ata_kernel<<<grid, REFIT_THREADS, 0, stream>>>(p, nblk);
        ransac_refit_big_combine_kernel<<<batch, 64, 0, stream>>>(p, nblk);
    } else {
        ransac_refit_kernel<<<batch, REFIT_THREADS, 0, stream>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ransac_solve_kernel<<<(batch + SOLVE_WARPS - 1) / SOLVE_WARPS, SOLVE_WARPS * 32, 0, stream>>>(p, batch);
    return cudaGetLastError();
}

}  // namespace tvg
