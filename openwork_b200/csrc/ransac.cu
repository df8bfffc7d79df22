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

__device__ __forceinline__ bool hartley8(const double (&x)[8], const double (&y)[8], double& s, double& cx,
                                         double& cy) {
    double sx = 0.0, sy = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sx += x[k]; sy += y[k]; }
    cx = sx / 8.0;
    cy = sy / 8.0;
    double q = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const double dx = x[k] - cx, dy = y[k] - cy;
        q = fma(dx, dx, q);
        q = fma(dy, dy, q);
    }
    const double rms = sqrt(q / 8.0);
    if (!(rms > 0.0)) return false;
    s = 1.4142135623730951 / rms;
    return isfinite(s);
}

__global__ void __launch_bounds__(FIT_THREADS) ransac_fit_kernel(const RansacParams p) {
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
            essential_project(n, En);
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

// ------------------------------------------------------------------------------------ score (fp32)
constexpr int SCORE_THREADS = 128;
constexpr int SCORE_CHUNK = 256;          // correspondences per shared-memory stage (4 KB)
constexpr int SCORE_MAX_ROWS = 32768;     // rows per work item: the packed 16-bit lane counters must not wrap

// One Sampson evaluation in fp32: 19 FFMA (+ 1 MUFU.RCP + 1 FFMA when sum r^2 is wanted) + 2 integer-pipe
// instructions for BOTH counters: the sign bits of t (inlier) and bt (inside the fp32 uncertainty band) are
// replicated into the two 16-bit lanes of one word by a single PRMT (selector 0xFFBB: bytes 0,1 = sign of t, bytes
// 2,3 = sign of bt), and one IADD subtracts that word from the packed accumulator:
//     acc - 0x0000FFFF  ==  acc + 1 - 0x10000   -> low lane += 1, high lane -= 1
//     acc - 0xFFFF0000  ==  acc + 0x10000       -> high lane += 1
// so after <= 32768 rows  low = count, high = (band - count) mod 2^16.
// (Measured alternatives: SHF + IADD per counter = 4 integer instructions per evaluation, round 1; packed
// fma.rn.f32x2 over two hypotheses: FFMA2 issues at half the FFMA rate on B200, tools/microbench/fp32_pipes.cu.)
template <bool SSE>
__device__ __forceinline__ void score_eval(const float (&e)[9], const float4 c, float thr2, float band,
                                           float c1, float k1, unsigned& acc, float& sse) {
    // c = (x, y, u, v): x = src (last), u = dst (curr)
    const float f0 = fmaf(e[0], c.x, fmaf(e[1], c.y, e[2]));
    const float f1 = fmaf(e[3], c.x, fmaf(e[4], c.y, e[5]));
    const float f2 = fmaf(e[6], c.x, fmaf(e[7], c.y, e[8]));
    const float g0 = fmaf(e[0], c.z, fmaf(e[3], c.w, e[6]));
    const float g1 = fmaf(e[1], c.z, fmaf(e[4], c.w, e[7]));
    const float num = fmaf(c.z, f0, fmaf(c.w, f1, f2));
    // too close for fp32 to call:  |t| < band*den + floor.  The relative part covers rounding of num^2 and
    // thr^2*den; the absolute floor covers cancellation in num for points near the epipole (den -> 0):
    // 2*thr*sqrt(den)*delta <= (band_rel/2)*thr^2*den + delta^2/(band_rel/2)  with  delta ~ 3e-7 (|E|_F = 1,
    // |x| <~ 1); the other half of band_rel is the rounding allowance.  The floor costs no instruction: with
    // c1 = floor/band added to den and thr^2*c1 added to num^2 (two FMULs become FFMAs), t is unchanged and
    // band*den' = band*den + floor.
    const float den = fmaf(f0, f0, fmaf(f1, f1, fmaf(g0, g0, fmaf(g1, g1, c1))));
    const float n2 = fmaf(num, num, k1);
    const float t = fmaf(-thr2, den, n2);              // r < thr  <=>  num^2 - thr^2 den < 0
    const float bt = fmaf(-band, den, fabsf(t));
    unsigned w;
    asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(w) : "r"(__float_as_uint(t)), "r"(__float_as_uint(bt)));
    acc -= w;
    if (SSE) {
        float rinv;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(den - c1));
        sse = fmaf(n2 - k1, rinv, sse);
    }
}

// Persistent CTAs pull work items from a device-side queue (atomic counter): item = (pair, row slot, block of
// HPT * 128 hypotheses).  A thread owns HPT hypotheses (E in registers); the item's correspondences stream through a
// two-stage shared-memory ring filled by TMA bulk copies; one broadcast LDS.128 feeds HPT evaluations.  Items whose
// row slot lies beyond the pair's match count (known only on the device) cost one queue pop.
template <int HPT, bool SSE>
__global__ void __launch_bounds__(SCORE_THREADS) ransac_score_kernel(const RansacParams p, int items, int hblocks) {
    extern __shared__ __align__(128) unsigned char score_smem[];
    float4(*corr)[SCORE_CHUNK] = reinterpret_cast<float4(*)[SCORE_CHUNK]>(score_smem);
    __shared__ alignas(8) uint64_t bar[2];
    __shared__ int s_item;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    const float thr2 = (float)(p.thr * p.thr);
    const float band = p.band_rel * thr2;
    const float c1 = p.band_floor / band;   // den offset that realises the absolute floor
    const float k1 = thr2 * c1;             // matching num^2 offset: t stays num^2 - thr^2 den
    uint32_t g = 0;                         // chunks consumed so far by this CTA: stage = g & 1, parity = (g >> 1) & 1
    for (;;) {
        __syncthreads();                    // barrier init visible / previous item's s_item and stages released
        if (threadIdx.x == 0) s_item = (int)atomicAdd(p.score_queue, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= items) break;
        const int hb = item % hblocks;
        const int slot = (item / hblocks) % p.chunks;
        const int b = item / (hblocks * p.chunks);
        const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
        const int r_begin = slot * p.chunk;
        const int r_end = min(M, r_begin + p.chunk);
        if (r_end <= r_begin && slot > 0) continue;     // slot 0 always writes (zero counts when M == 0)
        const int n = (r_end > r_begin) ? (r_end - r_begin + SCORE_CHUNK - 1) / SCORE_CHUNK : 0;
        const size_t oc = ((size_t)b * p.chunks + slot) * p.H;
        const float4* __restrict__ src = p.corr32 + (size_t)b * p.n_max + r_begin;
        auto issue = [&](int i) {
            const int rows = min(SCORE_CHUNK, r_end - r_begin - i * SCORE_CHUNK);
            const uint32_t st = (g + (uint32_t)i) & 1u;
            mbar_arrive_expect_tx(&bar[st], (uint32_t)rows * 16u);
            tma_load_1d(corr[st], src + (size_t)i * SCORE_CHUNK, (uint32_t)rows * 16u, &bar[st]);
        };
        if (threadIdx.x == 0) {
            if (n > 0) issue(0);
            if (n > 1) issue(1);
        }
        float e[HPT][9];
        int h[HPT];
        unsigned acc[HPT];
        float sse[HPT];
#pragma unroll
        for (int k = 0; k < HPT; ++k) {
            h[k] = (hb * HPT + k) * SCORE_THREADS + threadIdx.x;
#pragma unroll
            for (int i = 0; i < 9; ++i) e[k][i] = (h[k] < p.H) ? p.E32[((size_t)b * 9 + i) * p.H + h[k]] : 0.0f;
            acc[k] = 0u; sse[k] = 0.0f;
        }
        for (int i = 0; i < n; ++i) {
            const uint32_t gi = g + (uint32_t)i;
            mbar_wait(&bar[gi & 1u], (gi >> 1) & 1u);
            const int rows = min(SCORE_CHUNK, r_end - r_begin - i * SCORE_CHUNK);
            const float4* __restrict__ cp = corr[gi & 1u];
#pragma unroll(16 / HPT)
            for (int j = 0; j < rows; ++j) {
                const float4 c = cp[j];
#pragma unroll
                for (int k = 0; k < HPT; ++k) score_eval<SSE>(e[k], c, thr2, band, c1, k1, acc[k], sse[k]);
            }
            __syncthreads();   // everyone is done with this stage before TMA refills it
            if (threadIdx.x == 0 && i + 2 < n) issue(i + 2);
        }
        g += (uint32_t)n;
#pragma unroll
        for (int k = 0; k < HPT; ++k)
            if (h[k] < p.H) {
                const unsigned cnt = acc[k] & 0xFFFFu;
                const unsigned unc = ((acc[k] >> 16) + cnt) & 0xFFFFu;
                p.part_cnt[oc + h[k]] = make_int2((int)cnt, (int)unc);
                p.part_sse[oc + h[k]] = sse[k];
            }
    }
}

// Work decomposition of the scoring launch.  hpt hypotheses per thread; `slots` row ranges of `rows_per_slot`
// correspondences each (partial sums merged by the reduce kernel); persistent grid of `ctas_per_sm` CTAs per SM
// (dynamic shared memory is padded so that no more fit), chosen so that the item count fills whole rounds.
static int score_max_ctas(int hpt) {
    static int cached[5] = {0, 0, 0, 0, 0};
    if (!cached[hpt]) {
        int n = 0;
        cudaError_t e = (hpt == 4)
            ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ransac_score_kernel<4, false>, SCORE_THREADS, 2 * SCORE_CHUNK * 16)
            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ransac_score_kernel<2, false>, SCORE_THREADS, 2 * SCORE_CHUNK * 16);
        if (e != cudaSuccess || n < 1) { cudaGetLastError(); n = hpt == 4 ? 5 : 8; }
        cached[hpt] = n;
    }
    return cached[hpt];
}

void ransac_score_plan(int batch, int n_max, int H, int sm_count, int cfg_hpt, int cfg_ctas, ScorePlan* out) {
    ScorePlan pl{};
    pl.hpt = (cfg_hpt == 4) ? 4 : 2;
    const int per_block = pl.hpt * SCORE_THREADS;
    pl.hblocks = (H + per_block - 1) / per_block;
    const int chunks = (n_max + SCORE_CHUNK - 1) / SCORE_CHUNK > 0 ? (n_max + SCORE_CHUNK - 1) / SCORE_CHUNK : 1;
    int max_c = score_max_ctas(pl.hpt);
    if (cfg_ctas > 0 && cfg_ctas < max_c) max_c = cfg_ctas;
    // row slots: one when (pair, hypothesis block) items alone give >= 4 rounds of the persistent grid, else split
    // the rows (a single huge pair); never more than SCORE_MAX_ROWS rows per slot
    const long long base = (long long)batch * pl.hblocks;
    long long want = (4LL * sm_count * max_c + base - 1) / base;
    const long long need = ((long long)chunks * SCORE_CHUNK + SCORE_MAX_ROWS - 1) / SCORE_MAX_ROWS;
    if (want < need) want = need;
    if (want < 1) want = 1;
    if (want > chunks) want = chunks;
    const int cps = (chunks + (int)want - 1) / (int)want;       // chunks per slot
    pl.slots = (chunks + cps - 1) / cps;
    pl.rows_per_slot = cps * SCORE_CHUNK;
    pl.items = (int)(base * pl.slots);
    // CTAs per SM: the largest count (>= 60 % of the maximum) whose rounds are filled best
    int best_c = max_c;
    double best_eff = 0.0;
    for (int c = max_c; c >= 1 && c * 10 >= max_c * 6; --c) {
        const long long per_round = (long long)sm_count * c;
        const long long rounds = (pl.items + per_round - 1) / per_round;
        const double eff = (double)pl.items / (double)(rounds * per_round);
        if (eff > best_eff + 0.01) { best_eff = eff; best_c = c; }
    }
    if (cfg_ctas > 0) best_c = cfg_ctas < score_max_ctas(pl.hpt) ? cfg_ctas : score_max_ctas(pl.hpt);
    pl.ctas_per_sm = best_c;
    const long long g = (long long)sm_count * best_c;
    pl.grid = (int)(pl.items < g ? pl.items : g);
    if (pl.grid < 1) pl.grid = 1;
    // pad the dynamic shared memory so that exactly ctas_per_sm CTAs fit on an SM (228 KB, 1 KB reserved per CTA)
    int smem = (int)((228 * 1024) / best_c) - 1024 - 256;
    smem &= ~127;
    if (smem < 2 * SCORE_CHUNK * 16) smem = 2 * SCORE_CHUNK * 16;
    if (smem > 200 * 1024) smem = 200 * 1024;
    pl.smem = smem;
    *out = pl;
}

cudaError_t launch_ransac_score(const RansacParams& p, int batch, const ScorePlan& pl, cudaStream_t stream) {
    // the fp32 sum of squared residuals is diagnostic (selection uses the exact fp64 re-score): computed
    // only when the caller asked for it
    (void)batch;
#define TVG_SCORE_LAUNCH(HPT, SSE)                                                                                    \
    do {                                                                                                              \
        cudaError_t e__ = cudaFuncSetAttribute(ransac_score_kernel<HPT, SSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem); \
        if (e__ != cudaSuccess) return e__;                                                                           \
        ransac_score_kernel<HPT, SSE><<<pl.grid, SCORE_THREADS, pl.smem, stream>>>(p, pl.items, pl.hblocks);         \
    } while (0)
    if (pl.hpt == 4) {
        if (p.want_sse) TVG_SCORE_LAUNCH(4, true); else TVG_SCORE_LAUNCH(4, false);
    } else {
        if (p.want_sse) TVG_SCORE_LAUNCH(2, true); else TVG_SCORE_LAUNCH(2, false);
    }
#undef TVG_SCORE_LAUNCH
    return cudaGetLastError();
}

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

cudaError_t launch_ransac_reduce(const RansacParams& p, int batch, cudaStream_t stream) {
    ransac_reduce_kernel<<<batch, REDUCE_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

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

cudaError_t launch_ransac_rescore(const RansacParams& p, int batch, cudaStream_t stream) {
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
    int cnt = 0;
    double sum[4] = {0.0, 0.0, 0.0, 0.0};
    for (int m = lo; m < hi; ++m) {
        const double2 s = src[m], d = dst[m];
        const bool in = sampson_residual(Eb, s.x, s.y, d.x, d.y) < p.thr;
        p.mask[ob + m] = in ? 1 : 0;
        if (in) { ++cnt; sum[0] += s.x; sum[1] += s.y; sum[2] += d.x; sum[3] += d.y; }
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
        for (int m = lo; m < hi; ++m) {
            if (!p.mask[ob + m]) continue;
            p.inl_idx[ob + pos] = m;
            if (kq) {
                reinterpret_cast<double2*>(p.pix_curr)[ob + pos] = kq[p.qidx[ob + m]];
                reinterpret_cast<double2*>(p.pix_last)[ob + pos] = kt[p.tidx[ob + m]];
            }
            ++pos;
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
    for (int m = lo; m < hi; ++m) {
        if (!p.mask[ob + m]) continue;
        const double2 s = src[m], d = dst[m];
        qs = fma(s.x - c[0], s.x - c[0], qs); qs = fma(s.y - c[1], s.y - c[1], qs);
        qd = fma(d.x - c[2], d.x - c[2], qd); qd = fma(d.y - c[3], d.y - c[3], qd);
    }
    qs = block_sum_f64<REFIT_THREADS>(qs, s_buf);
    qd = block_sum_f64<REFIT_THREADS>(qd, s_buf);
    const double ss = 1.4142135623730951 / sqrt(qs / (double)total);
    const double sd = 1.4142135623730951 / sqrt(qd / (double)total);
    // ---- A^T A (upper triangle, 45 sums) in fp64
    double acc[45];
#pragma unroll
    for (int i = 0; i < 45; ++i) acc[i] = 0.0;
    for (int m = lo; m < hi; ++m) {
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
int ransac_refit_big_blocks(int n_max) { return n_max > REFIT_BIG_N ? (n_max + REFIT_BIG_ROWS - 1) / REFIT_BIG_ROWS : 0; }
int ransac_refit_big_stride() { return REFIT_BIG_STRIDE; }

__global__ void __launch_bounds__(REFIT_THREADS) ransac_refit_big_select_kernel(const RansacParams p) {
    __shared__ int s_bc[REFIT_THREADS / 32], s_bh[REFIT_THREADS / 32];
    __shared__ double s_bs[REFIT_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    const size_t oh = (size_t)b * p.H;
    const int ncand = p.n_cand[b];
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    BestSel sel = best_init();
    for (int ci = tid; ci < ncand; ci += REFIT_THREADS) {
        const int h = p.cand[oh + ci];
        best_update(sel, p.cand_cnt[oh + h], p.cand_sse[oh + h], h);
    }
    sel = best_block_reduce(sel, s_bc, s_bh, s_bs);
    if (tid == 0) {
        const bool have = best_is_model(sel);
        const int best = have ? sel.h : -1;
        p.best[b] = best;
        p.best_sse[b] = have ? p.cand_sse[oh + sel.h] : qnan;
        if (best < 0) {
            p.n_inl[b] = 0;
            p.status[b] = (M <= 8) ? 1 : 2;
            for (int i = 0; i < 9; ++i) { p.E[b * 9 + i] = qnan; p.E_best[b * 9 + i] = qnan; }
            for (int i = 0; i < 16; ++i) p.Rt[b * 16 + i] = qnan;
            p.refit_ws[(size_t)b * REFIT_WS + 51] = 0.0;
        }
    }
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
    for (int i = 0; i < 9; ++i) Eb[i] = p.E64[((size_t)b * p.H + best) * 9 + i];
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
#pragma unroll
        for (int i = 0; i < 9; ++i) p.E_best[b * 9 + i] = p.E64[((size_t)b * p.H + best) * 9 + i];
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
            for (int i = 0; i < 9; ++i) ws[52 + i] = p.E64[((size_t)b * p.H + best) * 9 + i];
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
        for (int i = 0; i < 9; ++i) ws[52 + i] = p.E64[((size_t)b * p.H + best) * 9 + i];
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
        for (int t = lane; t < 45; t += 32) {
            int i = 0, rem = t;
            while (rem >= 9 - i) { rem -= 9 - i; ++i; }
            const int j = i + rem;
            const double v = ws[t];
            s_G[warp][i][j] = v;
            s_G[warp][j][i] = v;
        }
        __syncwarp();
        double n[9];
        jacobi_eig9_warp(s_G[warp], s_V[warp], s_cs[warp], n);
        if (lane == 0) {
            double En[9];
            essential_project(n, En);
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

cudaError_t launch_ransac_refit(const RansacParams& p, int batch, cudaStream_t stream) {
    const int nblk = p.refit_big ? ransac_refit_big_blocks(p.n_max) : 0;
    if (nblk > 0) {
        dim3 grid(nblk, batch);
        ransac_refit_big_select_kernel<<<batch, REFIT_THREADS, 0, stream>>>(p);
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
