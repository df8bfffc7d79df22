// Sliding-window bundle adjustment over W consecutive frames, batched over windows -- the back-end step
// MonocularSlam/slam.py:267 leaves as a TODO ("g2o bundle adjustment").  No reference behaviour exists; the
// specification is oracle/bundle_window.py:
//   minimise  sum over visible (point j, camera c) of | proj(P_c X_j) - x_cj |^2      (normalised image coordinates)
//   over the poses of the cameras c >= n_fixed (6 DoF twists, left-multiplied) and the points X_j,
//   Levenberg-Marquardt with Marquardt diagonal scaling, points eliminated by the Schur complement (3x3 blocks),
//   a fixed number of iterations, a step accepted iff the cost decreases (lambda /= 10, else *= 10).
// One CTA per window, W <= 8 cameras.  All arithmetic fp64, every cross-thread sum in a fixed order (deterministic):
//   pass 0   thread = point: V = sum_c Jp^T Jp (damped), its inverse, g_p              -> per-point scratch (12 doubles)
//   chunks   of 256 / (free cameras) points, two phases per chunk, nothing but the per-point scratch touches HBM:
//            (i)  thread = (point, free camera a): W_a = Jc^T Jp and Y_a = W_a V^-1 into shared memory (72 KB);
//                 U_a = sum Jc^T Jc, g_a = sum Jc^T r, sum Y_a g_p in registers
//            (ii) thread = (camera pair a <= b, point slice): the 6x6 block sum_j Y_a W_b^T in registers
//            then fixed-order folds: per-camera sums over the point slots, Schur blocks slice after slice
//   solve    Gaussian elimination with partial pivoting of the 6 (W - n_fixed) system in shared memory
//   update   thread = point: X - V^-1 (g_p + sum_a Jp^T (Jc d_a)) with the Jacobians recomputed, trial cost, accept / reject
#include "bundle_common.cuh"
#include "tvg_kernels.h"

namespace tvg {
namespace {
using namespace ba;

constexpr int BW_MAXW = 8;
constexpr int BW_MAXV = BW_MAXW - 1;          // free cameras
constexpr int BW_MAXS = 6 * BW_MAXV;          // reduced system size (42)
constexpr int BW_PT = 12;                     // per-point scratch doubles: V^-1 (6), g_p (3), trial X (3)
constexpr int BW_CAM = 36;                    // per (point, free camera), shared memory only: W (18), Y = W V^-1 (18)
constexpr int BW_SMEM = BA_THREADS * BW_CAM * 8;   // dynamic shared memory: one (point, camera) block per thread

// residual and Jacobian blocks of one observation: Jp (2x3, w.r.t. the point), Jc (2x6, w.r.t. the camera's twist)
__device__ __forceinline__ void obs_blocks(const Cam& c, const double (&X)[3], double x, double y, double (&r)[2],
                                           double (&Jp)[2][3], double (&Jc)[2][6]) {
    double b[3];
    to_cam(c, X, b);
    const double ib = 1.0 / b[2];
    const double ub = b[0] * ib, vb = b[1] * ib;
    r[0] = ub - x; r[1] = vb - y;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        Jp[0][k] = ib * (c.R[k] - ub * c.R[6 + k]);
        Jp[1][k] = ib * (c.R[3 + k] - vb * c.R[6 + k]);
    }
    const double j0[3] = {ib, 0.0, -ub * ib}, j1[3] = {0.0, ib, -vb * ib};
    Jc[0][0] = j0[1] * (-b[2]) + j0[2] * b[1];
    Jc[0][1] = j0[0] * b[2] + j0[2] * (-b[0]);
    Jc[0][2] = j0[0] * (-b[1]) + j0[1] * b[0];
    Jc[1][0] = j1[1] * (-b[2]) + j1[2] * b[1];
    Jc[1][1] = j1[0] * b[2] + j1[2] * (-b[0]);
    Jc[1][2] = j1[0] * (-b[1]) + j1[1] * b[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) { Jc[0][3 + k] = j0[k]; Jc[1][3 + k] = j1[k]; }
}

__device__ __forceinline__ double obs_cost(const Cam& c, const double (&X)[3], double x, double y) {
    double b[3];
    to_cam(c, X, b);
    const double r0 = b[0] / b[2] - x, r1 = b[1] / b[2] - y;
    return r0 * r0 + r1 * r1;
}

__global__ void __launch_bounds__(BA_THREADS) bundle_window_kernel(const BundleWindowParams p) {
    extern __shared__ __align__(16) double s_blk[];   // [points per chunk][free cameras][36]; later [33][BA_THREADS]
    __shared__ double s_part[33][BA_THREADS / 32];
    __shared__ double s_red[33];
    __shared__ double s_sum[BW_MAXV][33];
    __shared__ double s_S[BW_MAXS][BW_MAXS + 1];     // reduced camera system, last column = right-hand side
    __shared__ double s_d[BW_MAXS];
    __shared__ Cam s_cam[BW_MAXW], s_trial[BW_MAXW];
    __shared__ double s_lambda, s_cost;
    __shared__ int s_ok, s_accepted, s_piv;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int W = p.W, nf = p.n_fixed, nv = W - nf, ns = 6 * nv;
    const int n = p.n_pts ? min(p.n_pts[b], p.n_max) : p.n;
    const size_t ob = (size_t)b * p.n_max;
    const double2* obs = reinterpret_cast<const double2*>(p.obs) + ob * W;
    const uint8_t* vis = p.vis + ob * W;
    double* X4 = p.points4d + ob * 4;
    double* spt = p.scratch + ob * BW_PT;
    if (tid < W) cam_from_pose(p.poses + ((size_t)b * W + tid) * 16, s_cam[tid]);
    if (tid == 0) { s_lambda = p.lambda0; s_accepted = 0; }
    __syncthreads();
    auto norm = [&](double2 px, double& x, double& y) {   // K^-1 [u, v, 1]
        x = fma(p.kinv[0], px.x, fma(p.kinv[1], px.y, p.kinv[2]));
        y = fma(p.kinv[3], px.x, fma(p.kinv[4], px.y, p.kinv[5]));
    };
    auto total_cost = [&](const Cam* cams, bool trial_points) -> double {
        double acc[1] = {0.0};
        for (int j = tid; j < n; j += BA_THREADS) {
            const double* src = trial_points ? spt + (size_t)j * BW_PT + 9 : X4 + (size_t)j * 4;
            const double X[3] = {src[0], src[1], src[2]};
            for (int c = 0; c < W; ++c) {
                if (!vis[(size_t)j * W + c]) continue;
                double x, y;
                norm(obs[(size_t)j * W + c], x, y);
                acc[0] += obs_cost(cams[c], X, x, y);
            }
        }
        block_sum_vec<1>(acc, s_part, s_red);
        return s_red[0];
    };
    {
        const double c0 = total_cost(s_cam, false);
        if (tid == 0) { s_cost = c0; p.cost[b * 2] = c0; }
        __syncthreads();
    }
    for (int it = 0; it < p.iters; ++it) {
        const double lambda = s_lambda;
        // ---- pass 0: per-point V^-1 and g_p over every camera that sees the point
        for (int j = tid; j < n; j += BA_THREADS) {
            const double X[3] = {X4[(size_t)j * 4], X4[(size_t)j * 4 + 1], X4[(size_t)j * 4 + 2]};
            double V[6] = {0, 0, 0, 0, 0, 0}, gp[3] = {0, 0, 0};   // 00 01 02 11 12 22
            bool seen = false;
            for (int c = 0; c < W; ++c) {
                if (!vis[(size_t)j * W + c]) continue;
                seen = true;
                double x, y, r[2], Jp[2][3], Jc[2][6];
                norm(obs[(size_t)j * W + c], x, y);
                obs_blocks(s_cam[c], X, x, y, r, Jp, Jc);
                V[0] += Jp[0][0] * Jp[0][0] + Jp[1][0] * Jp[1][0];
                V[1] += Jp[0][0] * Jp[0][1] + Jp[1][0] * Jp[1][1];
                V[2] += Jp[0][0] * Jp[0][2] + Jp[1][0] * Jp[1][2];
                V[3] += Jp[0][1] * Jp[0][1] + Jp[1][1] * Jp[1][1];
                V[4] += Jp[0][1] * Jp[0][2] + Jp[1][1] * Jp[1][2];
                V[5] += Jp[0][2] * Jp[0][2] + Jp[1][2] * Jp[1][2];
#pragma unroll
                for (int k = 0; k < 3; ++k) gp[k] += Jp[0][k] * r[0] + Jp[1][k] * r[1];
            }
            double* o = spt + (size_t)j * BW_PT;
            if (!seen) {   // a point nobody sees: neutral block, stays where it is
                o[0] = 1.0; o[1] = 0.0; o[2] = 0.0; o[3] = 1.0; o[4] = 0.0; o[5] = 1.0;
                o[6] = 0.0; o[7] = 0.0; o[8] = 0.0;
                continue;
            }
            V[0] *= (1.0 + lambda); V[3] *= (1.0 + lambda); V[5] *= (1.0 + lambda);
            const double c00 = V[3] * V[5] - V[4] * V[4], c01 = V[4] * V[2] - V[1] * V[5], c02 = V[1] * V[4] - V[3] * V[2];
            const double id = 1.0 / (V[0] * c00 + V[1] * c01 + V[2] * c02);
            o[0] = c00 * id; o[1] = c01 * id; o[2] = c02 * id;
            o[3] = (V[0] * V[5] - V[2] * V[2]) * id; o[4] = (V[2] * V[1] - V[0] * V[4]) * id;
            o[5] = (V[0] * V[3] - V[1] * V[1]) * id;
            o[6] = gp[0]; o[7] = gp[1]; o[8] = gp[2];
        }
        for (int e = tid; e < ns * (ns + 1); e += BA_THREADS) s_S[e / (ns + 1)][e % (ns + 1)] = 0.0;
        __syncthreads();
        // ---- chunks of PCH points: (i) W / Y blocks of every (point, free camera) into shared memory + per-camera
        // sums, (ii) the Schur blocks of every camera pair from shared memory
        {
            const int PCH = BA_THREADS / nv;                 // points per chunk
            const int cp = tid / nv, ca = tid % nv;          // phase (i): this thread's point slot and free camera
            const bool worker = tid < PCH * nv;
            const int npair = nv * (nv + 1) / 2;
            const int nslice = max(BA_THREADS / npair, 1);
            const int pr = tid / nslice, sl = tid % nslice;  // phase (ii): camera pair and point slice
            int a = 0, bb = 0;
            if (pr < npair) {
                int rest = pr;
                while (rest >= nv - a) { rest -= nv - a; ++a; }
                bb = a + rest;
            }
            const Cam cam = s_cam[nf + ca];
            double acc[33];   // U_a (21), g_a (6), sum Y_a g_p (6) of camera ca over this thread's points
#pragma unroll
            for (int i = 0; i < 33; ++i) acc[i] = 0.0;
            double blk[6][6];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int c = 0; c < 6; ++c) blk[i][c] = 0.0;
            for (int j0 = 0; j0 < n; j0 += PCH) {
                if (worker) {
                    const int j = j0 + cp;
                    double* oc = s_blk + (size_t)(cp * nv + ca) * BW_CAM;
                    if (j < n && vis[(size_t)j * W + nf + ca]) {
                        const double X[3] = {X4[(size_t)j * 4], X4[(size_t)j * 4 + 1], X4[(size_t)j * 4 + 2]};
                        const double* o = spt + (size_t)j * BW_PT;
                        const double Vi[3][3] = {{o[0], o[1], o[2]}, {o[1], o[3], o[4]}, {o[2], o[4], o[5]}};
                        const double gp[3] = {o[6], o[7], o[8]};
                        double x, y, r[2], Jp[2][3], Jc[2][6];
                        norm(obs[(size_t)j * W + nf + ca], x, y);
                        obs_blocks(cam, X, x, y, r, Jp, Jc);
                        int t = 0;
#pragma unroll
                        for (int i = 0; i < 6; ++i) {
                            double Wr[3], Yr[3];
#pragma unroll
                            for (int k = 0; k < 3; ++k) Wr[k] = Jc[0][i] * Jp[0][k] + Jc[1][i] * Jp[1][k];
#pragma unroll
                            for (int k = 0; k < 3; ++k) Yr[k] = Wr[0] * Vi[0][k] + Wr[1] * Vi[1][k] + Wr[2] * Vi[2][k];
#pragma unroll
                            for (int k = 0; k < 3; ++k) { oc[i * 3 + k] = Wr[k]; oc[18 + i * 3 + k] = Yr[k]; }
#pragma unroll
                            for (int c = i; c < 6; ++c) acc[t++] += Jc[0][i] * Jc[0][c] + Jc[1][i] * Jc[1][c];
                            acc[21 + i] += Jc[0][i] * r[0] + Jc[1][i] * r[1];
                            acc[27 + i] += Yr[0] * gp[0] + Yr[1] * gp[1] + Yr[2] * gp[2];
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < BW_CAM; ++i) oc[i] = 0.0;   // unseen (or past the end): contributes exact zeros
                    }
                }
                __syncthreads();
                if (pr < npair) {
                    const int lim = min(PCH, n - j0);
                    for (int jl = sl; jl < lim; jl += nslice) {
                        const double* ya = s_blk + (size_t)(jl * nv + a) * BW_CAM + 18;
                        const double* wb = s_blk + (size_t)(jl * nv + bb) * BW_CAM;
                        double y[18], w[18];
#pragma unroll
                        for (int k = 0; k < 18; ++k) { y[k] = ya[k]; w[k] = wb[k]; }
#pragma unroll
                        for (int i = 0; i < 6; ++i)
#pragma unroll
                            for (int c = 0; c < 6; ++c)
                                blk[i][c] = fma(y[i * 3], w[c * 3],
                                                fma(y[i * 3 + 1], w[c * 3 + 1], fma(y[i * 3 + 2], w[c * 3 + 2], blk[i][c])));
                    }
                }
                __syncthreads();
            }
            // per-camera sums: every thread's 33 partials through shared memory, summed over the point slots in order
            if (worker) {
#pragma unroll
                for (int e = 0; e < 33; ++e) s_blk[(size_t)e * BA_THREADS + tid] = acc[e];
            }
            __syncthreads();
            for (int q = tid; q < nv * 33; q += BA_THREADS) {
                const int qa = q / 33, e = q % 33;
                double t = 0.0;
                for (int c2 = 0; c2 < PCH; ++c2) t += s_blk[(size_t)e * BA_THREADS + c2 * nv + qa];
                s_sum[qa][e] = t;
            }
            __syncthreads();
            if (tid < nv) {
                int t = 0;
                for (int i = 0; i < 6; ++i)
                    for (int c = i; c < 6; ++c) {
                        const double u = s_sum[tid][t++] * (i == c ? 1.0 + lambda : 1.0);
                        s_S[6 * tid + i][6 * tid + c] = u;
                        s_S[6 * tid + c][6 * tid + i] = u;
                    }
                for (int i = 0; i < 6; ++i) s_S[6 * tid + i][ns] = -(s_sum[tid][21 + i] - s_sum[tid][27 + i]);
            }
            __syncthreads();
            // Schur complement: S[(a,.)][(b,.)] -= the blocks, slice after slice (fixed order)
            for (int sfold = 0; sfold < nslice; ++sfold) {
                if (pr < npair && sl == sfold) {
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int c = 0; c < 6; ++c) {
                            if (a == bb && c < i) continue;              // diagonal block: upper triangle, mirrored
                            s_S[6 * a + i][6 * bb + c] -= blk[i][c];
                            if (!(a == bb && c == i)) s_S[6 * bb + c][6 * a + i] -= blk[i][c];
                        }
                }
                __syncthreads();
            }
        }
        // ---- solve: Gaussian elimination with partial pivoting, all threads
        if (tid == 0) s_ok = 1;
        __syncthreads();
        for (int k = 0; k < ns; ++k) {
            if (tid == 0) {
                int piv = k;
                double best = fabs(s_S[k][k]);
                for (int r = k + 1; r < ns; ++r)
                    if (fabs(s_S[r][k]) > best) { best = fabs(s_S[r][k]); piv = r; }
                if (!(best > 0.0)) s_ok = 0;
                s_piv = piv;
            }
            __syncthreads();
            if (!s_ok) break;
            const int piv = s_piv;
            if (piv != k && tid <= ns) {
                const double tmp = s_S[k][tid];
                s_S[k][tid] = s_S[piv][tid];
                s_S[piv][tid] = tmp;
            }
            __syncthreads();
            const double inv = 1.0 / s_S[k][k];
            const int cols = ns + 1 - k;                              // columns k .. ns
            for (int e = tid; e < (ns - k - 1) * cols; e += BA_THREADS) {
                const int r = k + 1 + e / cols, c = k + e % cols;
                if (c == k) continue;                                 // the column being eliminated: handled below
                s_S[r][c] -= (s_S[r][k] * inv) * s_S[k][c];
            }
            __syncthreads();
        }
        if (tid == 0) {
            bool ok = s_ok != 0;
            if (ok)
                for (int k = ns - 1; k >= 0; --k) {
                    double v = s_S[k][ns];
                    for (int c = k + 1; c < ns; ++c) v -= s_S[k][c] * s_d[c];
                    s_d[k] = v / s_S[k][k];
                }
            for (int k = 0; k < ns && ok; ++k)
                if (!isfinite(s_d[k])) ok = false;
            s_ok = ok ? 1 : 0;
        }
        __syncthreads();
        if (s_ok && tid < W) {
            if (tid < nf) s_trial[tid] = s_cam[tid];
            else {
                double d[6];
                for (int k = 0; k < 6; ++k) d[k] = s_d[6 * (tid - nf) + k];
                cam_update(s_cam[tid], d, s_trial[tid]);
            }
        }
        __syncthreads();
        // ---- back-substitution for the points, trial cost
        double trial = 0.0;
        if (s_ok) {
            for (int j = tid; j < n; j += BA_THREADS) {
                double* o = spt + (size_t)j * BW_PT;
                double e3[3] = {o[6], o[7], o[8]};
                bool seen = false;
                for (int c = 0; c < W; ++c) seen = seen || vis[(size_t)j * W + c];
                const double X[3] = {X4[(size_t)j * 4], X4[(size_t)j * 4 + 1], X4[(size_t)j * 4 + 2]};
                for (int a = 0; a < nv; ++a) {            // e3 += W_a^T d_a = Jp^T (Jc d_a), Jacobians recomputed
                    if (!vis[(size_t)j * W + nf + a]) continue;
                    double x, y, r[2], Jp[2][3], Jc[2][6];
                    norm(obs[(size_t)j * W + nf + a], x, y);
                    obs_blocks(s_cam[nf + a], X, x, y, r, Jp, Jc);
                    double s2[2] = {0.0, 0.0};
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        s2[0] += Jc[0][i] * s_d[6 * a + i];
                        s2[1] += Jc[1][i] * s_d[6 * a + i];
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) e3[k] += Jp[0][k] * s2[0] + Jp[1][k] * s2[1];
                }
                if (seen) {
                    o[9] = X[0] - (o[0] * e3[0] + o[1] * e3[1] + o[2] * e3[2]);
                    o[10] = X[1] - (o[1] * e3[0] + o[3] * e3[1] + o[4] * e3[2]);
                    o[11] = X[2] - (o[2] * e3[0] + o[4] * e3[1] + o[5] * e3[2]);
                } else {
                    o[9] = X[0]; o[10] = X[1]; o[11] = X[2];
                }
            }
            __syncthreads();
            trial = total_cost(s_trial, true);
        }
        const bool accept = s_ok && (trial < s_cost);   // NaN never accepts
        if (accept)
            for (int j = tid; j < n; j += BA_THREADS) {
                const double* o = spt + (size_t)j * BW_PT;
                X4[(size_t)j * 4] = o[9]; X4[(size_t)j * 4 + 1] = o[10]; X4[(size_t)j * 4 + 2] = o[11];
            }
        __syncthreads();
        if (accept && tid < W) s_cam[tid] = s_trial[tid];
        if (tid == 0) {
            if (accept) {
                s_cost = trial;
                s_lambda = fmax(lambda * 0.1, 1e-15);
                ++s_accepted;
            } else {
                s_lambda = fmin(lambda * 10.0, 1e12);
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        p.cost[b * 2 + 1] = s_cost;
        if (p.accepted) p.accepted[b] = s_accepted;
    }
    if (tid >= nf && tid < W) {   // pose = inv([R t; 0 1]); the fixed cameras keep their input bits
        double P[16], inv[16];
        for (int r = 0; r < 3; ++r) {
            for (int k = 0; k < 3; ++k) P[r * 4 + k] = s_cam[tid].R[r * 3 + k];
            P[r * 4 + 3] = s_cam[tid].t[r];
        }
        P[12] = 0.0; P[13] = 0.0; P[14] = 0.0; P[15] = 1.0;
        inverse4(P, inv);
        double* out = p.poses + ((size_t)b * W + tid) * 16;
        for (int i = 0; i < 16; ++i) out[i] = inv[i];
    }
}

}  // namespace

size_t bundle_window_scratch_doubles(int batch, int n_max) {
    return (size_t)batch * (n_max > 0 ? n_max : 1) * BW_PT;
}
int bundle_window_max_cameras() { return BW_MAXW; }

cudaError_t launch_bundle_window(const BundleWindowParams& p, int batch, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(bundle_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM);
    if (e != cudaSuccess) return e;
    bundle_window_kernel<<<batch, ba::BA_THREADS, BW_SMEM, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tvg
