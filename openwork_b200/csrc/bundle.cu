// Two-view bundle adjustment, batched over frame pairs -- the back-end step MonocularSlam/slam.py:267 leaves as a
// TODO ("g2o bundle adjustment").  No reference behaviour exists; the specification is oracle/bundle.py:
//   minimise  sum_j | proj(P1 X_j) - x1_j |^2 + | proj(P2 X_j) - x2_j |^2      (normalised image coordinates)
//   over the second view's pose (6 DoF twist, left-multiplied; the first view stays fixed) and the points X_j,
//   Levenberg-Marquardt with Marquardt diagonal scaling, points eliminated by the Schur complement (3x3 blocks),
//   a fixed number of iterations, a step accepted iff the cost decreases (lambda /= 10, else *= 10).
// One CTA per pair; a thread owns points j = tid, tid + 256, ...; all arithmetic fp64; the 6x6 reduced camera system
// is assembled with a fixed-order block reduction (deterministic) and solved by thread 0.
#include "bundle_common.cuh"
#include "tvg_kernels.h"

namespace tvg {
namespace {
using namespace ba;

constexpr int BA_NRED = 54;   // U (21) + g_c (6) + sum W V^-1 W^T (21) + sum W V^-1 g_p (6)

// squared reprojection error of one point in both views
__device__ __forceinline__ double point_cost(const Cam& c1, const Cam& c2, const double (&X)[3], double x1, double y1,
                                             double x2, double y2) {
    double a[3], b[3];
    to_cam(c1, X, a);
    to_cam(c2, X, b);
    const double r0 = a[0] / a[2] - x1, r1 = a[1] / a[2] - y1, r2 = b[0] / b[2] - x2, r3 = b[1] / b[2] - y2;
    return (r0 * r0 + r1 * r1) + (r2 * r2 + r3 * r3);
}

// residuals and Jacobian blocks of one point: Jp1, Jp2 (2x3 each, w.r.t. the point), Jc (2x6, w.r.t. the twist of view 2)
__device__ __forceinline__ void point_blocks(const Cam& c1, const Cam& c2, const double (&X)[3], double x1, double y1,
                                             double x2, double y2, double (&r)[4], double (&Jp1)[2][3],
                                             double (&Jp2)[2][3], double (&Jc)[2][6]) {
    double a[3], b[3];
    to_cam(c1, X, a);
    to_cam(c2, X, b);
    const double ia = 1.0 / a[2], ib = 1.0 / b[2];
    const double ua = a[0] * ia, va = a[1] * ia, ub = b[0] * ib, vb = b[1] * ib;
    r[0] = ua - x1; r[1] = va - y1; r[2] = ub - x2; r[3] = vb - y2;
    // d(u,v)/dXc = [[1/z, 0, -u/z], [0, 1/z, -v/z]]
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        Jp1[0][k] = ia * (c1.R[k] - ua * c1.R[6 + k]);
        Jp1[1][k] = ia * (c1.R[3 + k] - va * c1.R[6 + k]);
        Jp2[0][k] = ib * (c2.R[k] - ub * c2.R[6 + k]);
        Jp2[1][k] = ib * (c2.R[3 + k] - vb * c2.R[6 + k]);
    }
    // Xc' = Xc + w x Xc + v:  dXc/dw = -[Xc]x, dXc/dv = I
    const double j0[3] = {ib, 0.0, -ub * ib}, j1[3] = {0.0, ib, -vb * ib};
    // -[b]x = [[0, b2, -b1], [-b2, 0, b0], [b1, -b0, 0]]
    Jc[0][0] = j0[1] * (-b[2]) + j0[2] * b[1];
    Jc[0][1] = j0[0] * b[2] + j0[2] * (-b[0]);
    Jc[0][2] = j0[0] * (-b[1]) + j0[1] * b[0];
    Jc[1][0] = j1[1] * (-b[2]) + j1[2] * b[1];
    Jc[1][1] = j1[0] * b[2] + j1[2] * (-b[0]);
    Jc[1][2] = j1[0] * (-b[1]) + j1[1] * b[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) { Jc[0][3 + k] = j0[k]; Jc[1][3 + k] = j1[k]; }
}

// V = Jp^T Jp with Marquardt damping, its inverse (symmetric 3x3, adjugate), W = Jc^T Jp2, g_p = Jp^T r
__device__ __forceinline__ void point_system(const double (&r)[4], const double (&Jp1)[2][3], const double (&Jp2)[2][3],
                                             const double (&Jc)[2][6], double lambda, double (&Vi)[3][3],
                                             double (&W)[6][3], double (&gp)[3]) {
    double V[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            V[i][k] = (Jp1[0][i] * Jp1[0][k] + Jp1[1][i] * Jp1[1][k]) + (Jp2[0][i] * Jp2[0][k] + Jp2[1][i] * Jp2[1][k]);
        gp[i] = (Jp1[0][i] * r[0] + Jp1[1][i] * r[1]) + (Jp2[0][i] * r[2] + Jp2[1][i] * r[3]);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) V[i][i] *= (1.0 + lambda);
    const double c00 = V[1][1] * V[2][2] - V[1][2] * V[2][1];
    const double c01 = V[1][2] * V[2][0] - V[1][0] * V[2][2];
    const double c02 = V[1][0] * V[2][1] - V[1][1] * V[2][0];
    const double det = V[0][0] * c00 + V[0][1] * c01 + V[0][2] * c02;
    const double id = 1.0 / det;
    Vi[0][0] = c00 * id; Vi[0][1] = (V[0][2] * V[2][1] - V[0][1] * V[2][2]) * id; Vi[0][2] = (V[0][1] * V[1][2] - V[0][2] * V[1][1]) * id;
    Vi[1][0] = c01 * id; Vi[1][1] = (V[0][0] * V[2][2] - V[0][2] * V[2][0]) * id; Vi[1][2] = (V[0][2] * V[1][0] - V[0][0] * V[1][2]) * id;
    Vi[2][0] = c02 * id; Vi[2][1] = (V[0][1] * V[2][0] - V[0][0] * V[2][1]) * id; Vi[2][2] = (V[0][0] * V[1][1] - V[0][1] * V[1][0]) * id;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) W[a][k] = Jc[0][a] * Jp2[0][k] + Jc[1][a] * Jp2[1][k];
}

__global__ void __launch_bounds__(BA_THREADS) bundle_pairs_kernel(const BundleParams p) {
    __shared__ double s_part[BA_NRED][BA_THREADS / 32];
    __shared__ double s_red[BA_NRED];
    __shared__ double s_delta[6];
    __shared__ Cam s_c2, s_trial;
    __shared__ double s_lambda, s_cost;
    __shared__ int s_ok, s_accepted;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = p.n_pts ? min(p.n_pts[b], p.n_max) : p.n;
    const size_t ob = (size_t)b * p.n_max;
    const double2* pix1 = reinterpret_cast<const double2*>(p.pix1) + ob;
    const double2* pix2 = reinterpret_cast<const double2*>(p.pix2) + ob;
    double* X4 = p.points4d + ob * 4;
    double* Xt = p.trial + ob * 3;
    const uint8_t* use = p.use ? p.use + ob : nullptr;
    Cam c1;
    cam_from_pose(p.pose1 + (size_t)b * p.pose1_stride, c1);
    if (tid == 0) {
        cam_from_pose(p.pose2 + (size_t)b * p.pose2_stride, s_c2);
        s_lambda = p.lambda0;
        s_accepted = 0;
    }
    __syncthreads();
    auto norm = [&](double2 px, double& x, double& y) {   // K^-1 [u, v, 1]
        x = fma(p.kinv[0], px.x, fma(p.kinv[1], px.y, p.kinv[2]));
        y = fma(p.kinv[3], px.x, fma(p.kinv[4], px.y, p.kinv[5]));
    };
    // initial cost
    {
        Cam c2 = s_c2;
        double acc[1] = {0.0};
        for (int j = tid; j < n; j += BA_THREADS) {
            if (use && !use[j]) continue;
            double x1, y1, x2, y2;
            norm(pix1[j], x1, y1);
            norm(pix2[j], x2, y2);
            const double X[3] = {X4[j * 4], X4[j * 4 + 1], X4[j * 4 + 2]};
            acc[0] += point_cost(c1, c2, X, x1, y1, x2, y2);
        }
        block_sum_vec<1>(acc, s_part, s_red);
        if (tid == 0) {
            s_cost = s_red[0];
            p.cost[b * 2] = s_red[0];
        }
        __syncthreads();
    }
    for (int it = 0; it < p.iters; ++it) {
        const Cam c2 = s_c2;
        const double lambda = s_lambda;
        // ---- reduced camera system
        double acc[BA_NRED];
#pragma unroll
        for (int i = 0; i < BA_NRED; ++i) acc[i] = 0.0;
        for (int j = tid; j < n; j += BA_THREADS) {
            if (use && !use[j]) continue;
            double x1, y1, x2, y2, r[4], Jp1[2][3], Jp2[2][3], Jc[2][6], Vi[3][3], W[6][3], gp[3];
            norm(pix1[j], x1, y1);
            norm(pix2[j], x2, y2);
            const double X[3] = {X4[j * 4], X4[j * 4 + 1], X4[j * 4 + 2]};
            point_blocks(c1, c2, X, x1, y1, x2, y2, r, Jp1, Jp2, Jc);
            point_system(r, Jp1, Jp2, Jc, lambda, Vi, W, gp);
            double WV[6][3];
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int k = 0; k < 3; ++k) WV[a][k] = W[a][0] * Vi[0][k] + W[a][1] * Vi[1][k] + W[a][2] * Vi[2][k];
            int t = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = a; c < 6; ++c) {
                    acc[t] += Jc[0][a] * Jc[0][c] + Jc[1][a] * Jc[1][c];
                    acc[27 + t] += WV[a][0] * W[c][0] + WV[a][1] * W[c][1] + WV[a][2] * W[c][2];
                    ++t;
                }
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                acc[21 + a] += Jc[0][a] * r[2] + Jc[1][a] * r[3];
                acc[48 + a] += WV[a][0] * gp[0] + WV[a][1] * gp[1] + WV[a][2] * gp[2];
            }
        }
        block_sum_vec<BA_NRED>(acc, s_part, s_red);
        if (tid == 0) {
            // S = U (1 + lambda on the diagonal) - sum W V^-1 W^T,  rhs = -(g_c - sum W V^-1 g_p); Gauss with pivoting
            double S[6][7];
            int t = 0;
            for (int a = 0; a < 6; ++a)
                for (int c = a; c < 6; ++c) {
                    const double u = s_red[t] * (a == c ? 1.0 + lambda : 1.0) - s_red[27 + t];
                    S[a][c] = u; S[c][a] = u;
                    ++t;
                }
            for (int a = 0; a < 6; ++a) S[a][6] = -(s_red[21 + a] - s_red[48 + a]);
            bool ok = true;
            for (int k = 0; k < 6; ++k) {
                int piv = k;
                double best = fabs(S[k][k]);
                for (int r = k + 1; r < 6; ++r)
                    if (fabs(S[r][k]) > best) { best = fabs(S[r][k]); piv = r; }
                if (!(best > 0.0)) { ok = false; break; }
                if (piv != k)
                    for (int c = 0; c < 7; ++c) { const double tmp = S[k][c]; S[k][c] = S[piv][c]; S[piv][c] = tmp; }
                const double inv = 1.0 / S[k][k];
                for (int r = k + 1; r < 6; ++r) {
                    const double f = S[r][k] * inv;
                    for (int c = k; c < 7; ++c) S[r][c] -= f * S[k][c];
                }
            }
            double d[6] = {0, 0, 0, 0, 0, 0};
            if (ok)
                for (int k = 5; k >= 0; --k) {
                    double v = S[k][6];
                    for (int c = k + 1; c < 6; ++c) v -= S[k][c] * d[c];
                    d[k] = v / S[k][k];
                }
            for (int k = 0; k < 6; ++k) {
                if (!isfinite(d[k])) ok = false;
                s_delta[k] = d[k];
            }
            s_ok = ok ? 1 : 0;
            if (ok) cam_update(s_c2, d, s_trial);
        }
        __syncthreads();
        // ---- back-substitution for the points, trial cost
        double tc[1] = {0.0};
        if (s_ok) {
            const Cam ct = s_trial;
            double dc[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) dc[k] = s_delta[k];
            for (int j = tid; j < n; j += BA_THREADS) {
                if (use && !use[j]) continue;
                double x1, y1, x2, y2, r[4], Jp1[2][3], Jp2[2][3], Jc[2][6], Vi[3][3], W[6][3], gp[3];
                norm(pix1[j], x1, y1);
                norm(pix2[j], x2, y2);
                const double X[3] = {X4[j * 4], X4[j * 4 + 1], X4[j * 4 + 2]};
                point_blocks(c1, c2, X, x1, y1, x2, y2, r, Jp1, Jp2, Jc);
                point_system(r, Jp1, Jp2, Jc, lambda, Vi, W, gp);
                double e[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double wd = 0.0;
#pragma unroll
                    for (int a = 0; a < 6; ++a) wd += W[a][k] * dc[a];
                    e[k] = gp[k] + wd;
                }
                double Xn[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) Xn[k] = X[k] - (Vi[k][0] * e[0] + Vi[k][1] * e[1] + Vi[k][2] * e[2]);
                Xt[j * 3] = Xn[0]; Xt[j * 3 + 1] = Xn[1]; Xt[j * 3 + 2] = Xn[2];
                tc[0] += point_cost(c1, ct, Xn, x1, y1, x2, y2);
            }
        }
        block_sum_vec<1>(tc, s_part, s_red);
        const bool accept = s_ok && (s_red[0] < s_cost);   // NaN never accepts
        if (accept)
            for (int j = tid; j < n; j += BA_THREADS) {
                if (use && !use[j]) continue;
                X4[j * 4] = Xt[j * 3]; X4[j * 4 + 1] = Xt[j * 3 + 1]; X4[j * 4 + 2] = Xt[j * 3 + 2];
            }
        __syncthreads();
        if (tid == 0) {
            if (accept) {
                s_c2 = s_trial;
                s_cost = s_red[0];
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
        // pose2 = inv([R t; 0 1])
        double P[16], inv[16];
        for (int r = 0; r < 3; ++r) {
            for (int k = 0; k < 3; ++k) P[r * 4 + k] = s_c2.R[r * 3 + k];
            P[r * 4 + 3] = s_c2.t[r];
        }
        P[12] = 0.0; P[13] = 0.0; P[14] = 0.0; P[15] = 1.0;
        inverse4(P, inv);
        double* out = p.pose2 + (size_t)b * p.pose2_stride;
        for (int i = 0; i < 16; ++i) out[i] = inv[i];
    }
}

}  // namespace

cudaError_t launch_bundle_pairs(const BundleParams& p, int batch, cudaStream_t stream) {
    bundle_pairs_kernel<<<batch, BA_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tvg
