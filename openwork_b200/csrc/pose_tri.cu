// Pose extraction, pose chaining and DLT triangulation  --  replaces extract_Rt (slam.py:167-186),
// `curr_pose = Rt @ last_pose` (slam.py:261), triangulate (slam.py:202-221) and check_points
// (slam.py:233-237) of /root/reference/MonocularSlam/slam.py.
//
// triangulate: one thread per point, everything in registers: K-normalise the two pixel observations,
// build the 4x4 DLT system, one-sided Jacobi SVD (fp64), take the right singular vector of the smallest
// singular value, divide by w.  Algorithmic traffic 64 B/point (2 x (x,y) f64 in, (x,y,z,1) f64 out).
#include "tvg_linalg.cuh"
#include "tvg_kernels.h"

namespace tvg {

__global__ void extract_rt_kernel(const ExtractRtParams p, int batch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    double E[9], Rt[16];
#pragma unroll
    for (int i = 0; i < 9; ++i) E[i] = p.E[(size_t)b * 9 + i];
    extract_rt(E, Rt);
#pragma unroll
    for (int i = 0; i < 16; ++i) p.Rt[(size_t)b * 16 + i] = Rt[i];
}

cudaError_t launch_extract_rt(const ExtractRtParams& p, int batch, cudaStream_t stream) {
    extract_rt_kernel<<<(batch + 31) / 32, 32, 0, stream>>>(p, batch);
    return cudaGetLastError();
}

// pose[f+1] = Rt[f] @ pose[f]  (slam.py:261): a sequential chain per sequence.  One warp per sequence: the lanes load
// 32 consecutive Rt matrices at once (the loads do not depend on the chain), then the chain steps run on values
// broadcast by shuffles -- every lane computes the same product, lane j keeps step j -- so a 500-pair sequence costs
// 500 short register-resident steps instead of 500 dependent global round trips.  Same fma order as a plain loop.
constexpr int POSE_WARPS = 4;

__global__ void __launch_bounds__(POSE_WARPS * 32) pose_chain_kernel(const PoseChainParams p, int nseq) {
    const int s = blockIdx.x * POSE_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= nseq) return;
    const int b0 = p.seq_off ? p.seq_off[s] : 0, b1 = p.seq_off ? p.seq_off[s + 1] : p.batch;
    double cur[16];
    const double* init = (s == 0 && p.first_init) ? p.first_init : (p.init_pose ? p.init_pose + (size_t)s * 16 : nullptr);
#pragma unroll
    for (int i = 0; i < 16; ++i) cur[i] = init ? init[i] : ((i % 5 == 0) ? 1.0 : 0.0);
    if (lane < 16) p.poses[(size_t)(b0 + s) * 16 + lane] = init ? init[lane] : ((lane % 5 == 0) ? 1.0 : 0.0);
    for (int base = b0; base < b1; base += 32) {
        const int mine_b = base + lane;
        double R[16], mine[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            R[i] = (mine_b < b1) ? p.Rt[(size_t)mine_b * 16 + i] : 0.0;
            mine[i] = 0.0;
        }
        const int steps = min(32, b1 - base);
        for (int j = 0; j < steps; ++j) {
            double Rj[16], nxt[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int lo = __shfl_sync(0xffffffffu, __double2loint(R[i]), j);
                const int hi = __shfl_sync(0xffffffffu, __double2hiint(R[i]), j);
                Rj[i] = __hiloint2double(hi, lo);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    double a = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) a = fma(Rj[r * 4 + k], cur[k * 4 + c], a);
                    nxt[r * 4 + c] = a;
                }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                cur[i] = nxt[i];
                if (lane == j) mine[i] = nxt[i];
            }
        }
        if (mine_b < b1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) p.poses[(size_t)(mine_b + s + 1) * 16 + i] = mine[i];
        }
    }
}

// np.linalg.inv of the two poses of every pair (slam.py:204-205), one thread per pair, after the chain.
__global__ void pose_pair_inverse_kernel(const PoseChainParams p, int nseq) {
    const int total = p.seq_off ? p.seq_off[nseq] : p.batch;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < total; b += gridDim.x * blockDim.x) {
        int s = 0;
        if (p.seq_off) {   // last sequence whose first pair is <= b
            int lo = 0, hi = nseq - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (p.seq_off[mid] <= b) lo = mid; else hi = mid - 1;
            }
            s = lo;
        }
        inverse4(p.poses + (size_t)(b + s) * 16, p.pair_pinv + (size_t)b * 32);
        inverse4(p.poses + (size_t)(b + s + 1) * 16, p.pair_pinv + (size_t)b * 32 + 16);
    }
}

// batch_hint = total pairs when the host knows it (sizes the second grid), else 0
cudaError_t launch_pose_chain(const PoseChainParams& p, int nseq, int batch_hint, cudaStream_t stream) {
    pose_chain_kernel<<<(nseq + POSE_WARPS - 1) / POSE_WARPS, POSE_WARPS * 32, 0, stream>>>(p, nseq);
    if (p.pair_pinv) {
        const int blocks = batch_hint > 0 ? (batch_hint + 63) / 64 : 592;
        pose_pair_inverse_kernel<<<blocks, 64, 0, stream>>>(p, nseq);
    }
    return cudaGetLastError();
}

// np.linalg.inv(pose) for a list of 4x4 matrices (slam.py:204-205), one thread each
__global__ void invert4_kernel(const double* __restrict__ in, double* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inverse4(in + (size_t)i * 16, out + (size_t)i * 16);
}

cudaError_t launch_invert4(const double* in, double* out, int n, cudaStream_t stream) {
    invert4_kernel<<<(n + 31) / 32, 32, 0, stream>>>(in, out, n);
    return cudaGetLastError();
}

constexpr int TRI_THREADS = 128;

__device__ __forceinline__ void dlt_point(const double* __restrict__ P1, const double* __restrict__ P2,
                                          double x1, double y1, double x2, double y2, double (&X)[4]) {
    // rows: x1*P1[2]-P1[0], y1*P1[2]-P1[1], x2*P2[2]-P2[0], y2*P2[2]-P2[1]   (slam.py:212-216)
    double G[4][4], V[4][4];  // G[c][r]: column c of A
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        G[c][0] = x1 * P1[8 + c] - P1[c];
        G[c][1] = y1 * P1[8 + c] - P1[4 + c];
        G[c][2] = x2 * P2[8 + c] - P2[c];
        G[c][3] = y2 * P2[8 + c] - P2[4 + c];
    }
    jacobi_onesided<4, 4>(G, V, 30);
    double nmin = 0.0;
    int best = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double n = G[c][0] * G[c][0] + G[c][1] * G[c][1] + G[c][2] * G[c][2] + G[c][3] * G[c][3];
        if (c == 0 || n < nmin) { nmin = n; best = c; }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c == best) {
            const double w = V[c][3];
            X[0] = V[c][0] / w; X[1] = V[c][1] / w; X[2] = V[c][2] / w; X[3] = w / w;
        }
}

// ---- fast path of the DLT solve -----------------------------------------------------------------------------------
// vt[3] of svd(A) (slam.py:217-218) is the dominant eigenvector of (A^T A)^-1 = adj(A) adj(A)^T / det^2.  The adjugate
// of the 4x4 system needs no division (Laplace expansion over the 2x2 minors of rows 0,1 and rows 2,3: 72 multiply-adds)
// and stays accurate when A is nearly singular -- which is the normal case here, the exact solution being a null
// vector.  C = adj adj^T is then squared repeatedly (C, C^2, C^4, ...: 44 multiply-adds per step) until it is rank one
// to working precision: the separation after k steps is (sigma_4 / sigma_3)^(2^(k+1)), 3-5 steps on the benchmark scenes.
// tr(C^2) = tr(C)^2 exactly when C is rank one, and tr(C^2) is the trace of the next iterate, so the test is free.
// ~450 fp64 instructions per point against ~4300 for the one-sided Jacobi SVD; a point that has not separated after
// TRI_MAX_SQUARINGS steps (sigma_4 / sigma_3 > 0.995: no parallax at all) is left to the Jacobi kernel.
constexpr int TRI_MAX_SQUARINGS = 12;

__device__ __forceinline__ double rcp_newton(double x) {   // 1 / x to ~1 ulp: MUFU seed + two Newton steps
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}

// symmetric 4x4 as 10 doubles: 00 01 02 03 11 12 13 22 23 33
__device__ __forceinline__ void sym4_square(const double (&c)[10], double (&d)[10]) {
    d[0] = fma(c[0], c[0], fma(c[1], c[1], fma(c[2], c[2], c[3] * c[3])));
    d[1] = fma(c[0], c[1], fma(c[1], c[4], fma(c[2], c[5], c[3] * c[6])));
    d[2] = fma(c[0], c[2], fma(c[1], c[5], fma(c[2], c[7], c[3] * c[8])));
    d[3] = fma(c[0], c[3], fma(c[1], c[6], fma(c[2], c[8], c[3] * c[9])));
    d[4] = fma(c[1], c[1], fma(c[4], c[4], fma(c[5], c[5], c[6] * c[6])));
    d[5] = fma(c[1], c[2], fma(c[4], c[5], fma(c[5], c[7], c[6] * c[8])));
    d[6] = fma(c[1], c[3], fma(c[4], c[6], fma(c[5], c[8], c[6] * c[9])));
    d[7] = fma(c[2], c[2], fma(c[5], c[5], fma(c[7], c[7], c[8] * c[8])));
    d[8] = fma(c[2], c[3], fma(c[5], c[6], fma(c[7], c[8], c[8] * c[9])));
    d[9] = fma(c[3], c[3], fma(c[6], c[6], fma(c[8], c[8], c[9] * c[9])));
}

// Returns false (X untouched) when the iteration has not separated the smallest singular direction.
__device__ __forceinline__ bool dlt_point_fast(const double* __restrict__ P1, const double* __restrict__ P2, double x1,
                                               double y1, double x2, double y2, double (&X)[4]) {
    double a[4][4];   // rows of A, the same expressions as slam.py:212-216
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        a[0][c] = x1 * P1[8 + c] - P1[c];
        a[1][c] = y1 * P1[8 + c] - P1[4 + c];
        a[2][c] = x2 * P2[8 + c] - P2[c];
        a[3][c] = y2 * P2[8 + c] - P2[4 + c];
    }
    const double s0 = a[0][0] * a[1][1] - a[1][0] * a[0][1], s1 = a[0][0] * a[1][2] - a[1][0] * a[0][2];
    const double s2 = a[0][0] * a[1][3] - a[1][0] * a[0][3], s3 = a[0][1] * a[1][2] - a[1][1] * a[0][2];
    const double s4 = a[0][1] * a[1][3] - a[1][1] * a[0][3], s5 = a[0][2] * a[1][3] - a[1][2] * a[0][3];
    const double c5 = a[2][2] * a[3][3] - a[3][2] * a[2][3], c4 = a[2][1] * a[3][3] - a[3][1] * a[2][3];
    const double c3 = a[2][1] * a[3][2] - a[3][1] * a[2][2], c2 = a[2][0] * a[3][3] - a[3][0] * a[2][3];
    const double c1 = a[2][0] * a[3][2] - a[3][0] * a[2][2], c0 = a[2][0] * a[3][1] - a[3][0] * a[2][1];
    double b[4][4];   // adj(A)
    b[0][0] = a[1][1] * c5 - a[1][2] * c4 + a[1][3] * c3;
    b[0][1] = -a[0][1] * c5 + a[0][2] * c4 - a[0][3] * c3;
    b[0][2] = a[3][1] * s5 - a[3][2] * s4 + a[3][3] * s3;
    b[0][3] = -a[2][1] * s5 + a[2][2] * s4 - a[2][3] * s3;
    b[1][0] = -a[1][0] * c5 + a[1][2] * c2 - a[1][3] * c1;
    b[1][1] = a[0][0] * c5 - a[0][2] * c2 + a[0][3] * c1;
    b[1][2] = -a[3][0] * s5 + a[3][2] * s2 - a[3][3] * s1;
    b[1][3] = a[2][0] * s5 - a[2][2] * s2 + a[2][3] * s1;
    b[2][0] = a[1][0] * c4 - a[1][1] * c2 + a[1][3] * c0;
    b[2][1] = -a[0][0] * c4 + a[0][1] * c2 - a[0][3] * c0;
    b[2][2] = a[3][0] * s4 - a[3][1] * s2 + a[3][3] * s0;
    b[2][3] = -a[2][0] * s4 + a[2][1] * s2 - a[2][3] * s0;
    b[3][0] = -a[1][0] * c3 + a[1][1] * c1 - a[1][2] * c0;
    b[3][1] = a[0][0] * c3 - a[0][1] * c1 + a[0][2] * c0;
    b[3][2] = -a[3][0] * s3 + a[3][1] * s1 - a[3][2] * s0;
    b[3][3] = a[2][0] * s3 - a[2][1] * s1 + a[2][2] * s0;
    double C[10];
    {
        int t = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = i; j < 4; ++j)
                C[t++] = fma(b[i][0], b[j][0], fma(b[i][1], b[j][1], fma(b[i][2], b[j][2], b[i][3] * b[j][3])));
    }
    double tr = C[0] + C[4] + C[7] + C[9];
    if (!(tr > 0.0) || !isfinite(tr)) return false;
    bool conv = false;
    for (int k = 0; k < TRI_MAX_SQUARINGS; k += 2) {   // two squarings per trip, ping-pong between C and D (no copies)
        if ((k & 3) == 0) {   // keep the entries in range (any scale will do: the approximate reciprocal as is)
            double r;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(tr));
#pragma unroll
            for (int i = 0; i < 10; ++i) C[i] *= r;
            tr = C[0] + C[4] + C[7] + C[9];
        }
        double D[10];
        sym4_square(C, D);
        const double trd = D[0] + D[4] + D[7] + D[9];
        const bool conv1 = trd >= (1.0 - 2e-14) * tr * tr;
        sym4_square(D, C);                             // one more squaring costs less than a divergent exit
        tr = C[0] + C[4] + C[7] + C[9];
        conv = conv1 || tr >= (1.0 - 2e-14) * trd * trd;
        if (conv) break;
    }
    if (!conv || !(tr > 0.0)) return false;
    // the column with the largest diagonal entry is (a multiple of) the eigenvector
    double v0 = C[0], v1 = C[1], v2 = C[2], v3 = C[3], best = C[0];
    if (C[4] > best) { best = C[4]; v0 = C[1]; v1 = C[4]; v2 = C[5]; v3 = C[6]; }
    if (C[7] > best) { best = C[7]; v0 = C[2]; v1 = C[5]; v2 = C[7]; v3 = C[8]; }
    if (C[9] > best) { best = C[9]; v0 = C[3]; v1 = C[6]; v2 = C[8]; v3 = C[9]; }
    if (v3 == 0.0) { X[0] = v0 / v3; X[1] = v1 / v3; X[2] = v2 / v3; X[3] = v3 / v3; return true; }
    const double r = rcp_newton(v3);   // X /= X[3]  (slam.py:220)
    X[0] = v0 * r; X[1] = v1 * r; X[2] = v2 * r; X[3] = 1.0;
    return true;
}

// z > 0 flag and the optional point filters, shared by the two triangulation kernels
__device__ __forceinline__ void tri_store(const TriParams& p, const double* sP, size_t o, double x1, double y1, double x2,
                                          double y2, const double (&X)[4]) {
    double4* out = reinterpret_cast<double4*>(p.points4d);
    out[o] = make_double4(X[0], X[1], X[2], X[3]);
    if (p.good) p.good[o] = (X[2] > 0.0) ? 1 : 0;  // check_points: world z > 0
    if (p.reproj || p.parallax_cos) {
        // The filters slam.py:236 leaves as TODO (no reference behaviour; spec: oracle/two_view.py point_quality).
        const double* P1 = sP;
        const double* P2 = sP + 16;
        double c1[3], c2[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            c1[r] = P1[r * 4] * X[0] + P1[r * 4 + 1] * X[1] + P1[r * 4 + 2] * X[2] + P1[r * 4 + 3];
            c2[r] = P2[r * 4] * X[0] + P2[r * 4 + 1] * X[1] + P2[r * 4 + 2] * X[2] + P2[r * 4 + 3];
        }
        if (p.reproj) {
            const double e1 = hypot(c1[0] / c1[2] - x1, c1[1] / c1[2] - y1);
            const double e2 = hypot(c2[0] / c2[2] - x2, c2[1] / c2[2] - y2);
            p.reproj[o] = p.focal * fmax(e1, e2);
        }
        if (p.parallax_cos) {
            // rays from the two camera centres to X, in the world frame: R^T * (camera-frame point)
            double r1[3], r2[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                r1[k] = P1[k] * c1[0] + P1[4 + k] * c1[1] + P1[8 + k] * c1[2];
                r2[k] = P2[k] * c2[0] + P2[4 + k] * c2[1] + P2[8 + k] * c2[2];
            }
            const double d = r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2];
            const double n1 = r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2];
            const double n2 = r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2];
            p.parallax_cos[o] = d * rsqrt(n1 * n2);
        }
    }
}

// JACOBI = false: the fast solve for every point; a point it cannot separate gets w = NaN as a marker (and one is
// added to *p.n_fallback).  JACOBI = true: the one-sided Jacobi SVD for exactly those points -- a second launch with a
// small grid that walks all (pair, block-of-points) items only when the counter is not zero.
template <bool JACOBI>
__global__ void __launch_bounds__(TRI_THREADS) triangulate_kernel(const TriParams p, int batch) {
    __shared__ double sP[32];
    const int per_pair = (p.n_max + TRI_THREADS - 1) / TRI_THREADS;
    if (JACOBI && *p.n_fallback == 0) return;
    const int items = JACOBI ? batch * per_pair : 1;
    for (int item = JACOBI ? blockIdx.x : 0; item < items; item += gridDim.x) {
        const int b = JACOBI ? item / per_pair : blockIdx.y;
        const int blk = JACOBI ? item % per_pair : blockIdx.x;
        const int n = p.n_pts ? min(p.n_pts[b], p.n_max) : p.n;
        if (blk * TRI_THREADS >= n) continue;
        __syncthreads();
        if (threadIdx.x < 32) {
            const double* P = (threadIdx.x < 16) ? p.P1 + (size_t)b * p.p1_stride : p.P2 + (size_t)b * p.p2_stride;
            sP[threadIdx.x] = P[threadIdx.x & 15];
        }
        __syncthreads();
        const int i = blk * TRI_THREADS + threadIdx.x;
        if (i >= n) continue;
        const size_t o = (size_t)b * p.n_max + i;
        if (JACOBI) {
            const double w = p.points4d[o * 4 + 3];
            if (w == w) continue;                    // solved by the fast kernel
        }
        const double2 a = reinterpret_cast<const double2*>(p.pix1)[o];
        const double2 c = reinterpret_cast<const double2*>(p.pix2)[o];
        const double x1 = p.kinv[0] * a.x + p.kinv[1] * a.y + p.kinv[2];
        const double y1 = p.kinv[3] * a.x + p.kinv[4] * a.y + p.kinv[5];
        const double x2 = p.kinv[0] * c.x + p.kinv[1] * c.y + p.kinv[2];
        const double y2 = p.kinv[3] * c.x + p.kinv[4] * c.y + p.kinv[5];
        double X[4];
        if (JACOBI) {
            dlt_point(sP, sP + 16, x1, y1, x2, y2, X);
        } else if (!dlt_point_fast(sP, sP + 16, x1, y1, x2, y2, X)) {
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            reinterpret_cast<double4*>(p.points4d)[o] = make_double4(qnan, qnan, qnan, qnan);
            atomicAdd(p.n_fallback, 1);
            continue;
        }
        tri_store(p, sP, o, x1, y1, x2, y2, X);
    }
}

// resets the fallback counter for the next call (stream order: after the Jacobi kernel)
__global__ void triangulate_reset_kernel(int* n_fallback) { *n_fallback = 0; }

cudaError_t launch_triangulate(const TriParams& p, int batch, cudaStream_t stream) {
    dim3 grid((p.n_max + TRI_THREADS - 1) / TRI_THREADS, batch);
    triangulate_kernel<false><<<grid, TRI_THREADS, 0, stream>>>(p, batch);
    // points the fast solve could not separate (none on ordinary scenes): the blocks of the second launch return at
    // once while the counter is zero
    const long long items = (long long)grid.x * batch;
    triangulate_kernel<true><<<(unsigned)(items < 592 ? items : 592), TRI_THREADS, 0, stream>>>(p, batch);
    triangulate_reset_kernel<<<1, 1, 0, stream>>>(p.n_fallback);
    return cudaGetLastError();
}

// Point-cloud append of Map.add_observation (slam.py:37-40) for whole sequences on the device: per sequence,
// (0,0,0,1) for the first frame (slam.py:248) followed by every pair's points with z > 0, in frame order --
// the `points` half of the (poses, points) tuple Map.display() hands to the viewer (slam.py:81-84).
constexpr int MAP_THREADS = 256;

__global__ void __launch_bounds__(MAP_THREADS) map_accumulate_kernel(const MapParams p) {
    __shared__ int s_cnt[MAP_THREADS / 32];
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double4* out = reinterpret_cast<double4*>(p.map_points) + (size_t)s * p.cap;
    long long base = 0;
    if (p.include_origin) {
        if (tid == 0 && p.cap > 0) out[0] = make_double4(0.0, 0.0, 0.0, 1.0);
        base = 1;
    }
    for (int b = p.seq_off[s]; b < p.seq_off[s + 1]; ++b) {
        const int n = min(p.n_pts[b], p.n_max);
        const size_t ob = (size_t)b * p.n_max;
        const int per = (n + MAP_THREADS - 1) / MAP_THREADS;
        const int lo = min(n, tid * per), hi = min(n, lo + per);
        int cnt = 0;
        for (int i = lo; i < hi; ++i) cnt += p.good[ob + i] ? 1 : 0;
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        __syncthreads();   // s_cnt reuse across iterations
        if (lane == 31) s_cnt[warp] = incl;
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < MAP_THREADS / 32; ++w) {
            if (w < warp) off += s_cnt[w];
            total += s_cnt[w];
        }
        long long pos = base + off + incl - cnt;
        for (int i = lo; i < hi; ++i) {
            if (!p.good[ob + i]) continue;
            if (pos < p.cap) out[pos] = reinterpret_cast<const double4*>(p.points4d)[ob + i];
            ++pos;
        }
        if (tid == 0 && p.pair_off) p.pair_off[b] = (int32_t)base;
        base += total;
    }
    if (tid == 0) p.map_count[s] = (int32_t)base;
}

cudaError_t launch_map_accumulate(const MapParams& p, int nseq, cudaStream_t stream) {
    map_accumulate_kernel<<<nseq, MAP_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

// Cheirality vote over the four (R, +-t) decompositions of E (north star (3); NOT in the reference,
// which uses the trace / t_z heuristics).  One CTA per pair; thread-strided over correspondences.
constexpr int CHEIR_THREADS = 256;

__global__ void __launch_bounds__(CHEIR_THREADS) cheirality_kernel(const CheiralityParams p) {
    __shared__ double sRt[4][16];
    __shared__ int s_votes[4];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int M = p.m_count ? min(p.m_count[b], p.n_max) : p.m;
    if (tid == 0) {
        double E[9], u[3][3], s[3], v[3][3];
        for (int i = 0; i < 9; ++i) E[i] = p.E[(size_t)b * 9 + i];
        svd3(E, u, s, v);
        for (int cand = 0; cand < 4; ++cand) {
            const double sw = (cand < 2) ? 1.0 : -1.0;     // W or W^T
            const double st = (cand & 1) ? -1.0 : 1.0;     // +u3 or -u3
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c)
                    sRt[cand][r * 4 + c] = sw * (u[1][r] * v[0][c] - u[0][r] * v[1][c]) + u[2][r] * v[2][c];
                sRt[cand][r * 4 + 3] = st * u[2][r];
            }
            sRt[cand][12] = 0.0; sRt[cand][13] = 0.0; sRt[cand][14] = 0.0; sRt[cand][15] = 1.0;
            s_votes[cand] = 0;
        }
    }
    __syncthreads();
    const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    int votes[4] = {0, 0, 0, 0};
    const size_t ob = (size_t)b * p.n_max;
    for (int m = tid; m < M; m += CHEIR_THREADS) {
        if (p.mask && !p.mask[ob + m]) continue;
        const double2 s = reinterpret_cast<const double2*>(p.src)[ob + m];
        const double2 d = reinterpret_cast<const double2*>(p.dst)[ob + m];
        for (int cand = 0; cand < 4; ++cand) {
            double X[4];
            dlt_point(I4, sRt[cand], s.x, s.y, d.x, d.y, X);
            const double z2 = sRt[cand][8] * X[0] + sRt[cand][9] * X[1] + sRt[cand][10] * X[2] + sRt[cand][11];
            votes[cand] += (X[2] > 0.0 && z2 > 0.0) ? 1 : 0;
        }
    }
    for (int cand = 0; cand < 4; ++cand) {
        int v = votes[cand];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if ((tid & 31) == 0) atomicAdd(&s_votes[cand], v);
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int cand = 0; cand < 4; ++cand) {
            p.votes[b * 4 + cand] = s_votes[cand];
            if (s_votes[cand] > s_votes[best]) best = cand;
            for (int i = 0; i < 16; ++i) p.Rts[((size_t)b * 4 + cand) * 16 + i] = sRt[cand][i];
        }
        p.choice[b] = best;
        if (p.Rt_out)
            for (int i = 0; i < 16; ++i) p.Rt_out[(size_t)b * 16 + i] = sRt[best][i];
    }
}

cudaError_t launch_cheirality(const CheiralityParams& p, int batch, cudaStream_t stream) {
    cheirality_kernel<<<batch, CHEIR_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tvg
