// Small register-resident fp64 linear algebra used by the RANSAC fit, refit, pose and triangulation
// kernels: one-sided (Hestenes) Jacobi SVD, Householder null vector of an 8x9 system, the
// essential-matrix (s,s,0) projection, extract_Rt, a pivoted 4x4 inverse.
#pragma once
#include "tvg_common.cuh"

namespace tvg {

// fp64 reciprocal / division / square root to ~1 ulp without the IEEE slow paths of the built-in operators (each of
// those is 25-40 instructions with branches): MUFU seed (rcp.approx.ftz.f64, >= 20 bits) + two Newton steps.  Used
// where a last-bit difference is immaterial: Jacobi rotation parameters, Householder reflectors, normalisations.
// Denormal arguments are flushed (1 / denormal = inf); every caller guards the divisor.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}
__device__ __forceinline__ double fast_div(double a, double b) { return a * fast_rcp(b); }
__device__ __forceinline__ double fast_sqrt(double x) { return (x > 0.0) ? x * rsqrt(x) : 0.0; }

// One-sided Jacobi SVD:  A V = G  with the columns of G mutually orthogonal on exit.
// G[c][r] holds column c of A (length MR), V[c][r] column c of V (starts as identity).
// Fully unrolled over column pairs so every index is static (register resident).
template <int MR, int N>
__device__ __forceinline__ void jacobi_onesided(double (&G)[N][MR], double (&V)[N][N], int max_sweeps) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < N - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
                for (int i = 0; i < MR; ++i) {
                    alpha = fma(G[p][i], G[p][i], alpha);
                    beta = fma(G[q][i], G[q][i], beta);
                    gamma = fma(G[p][i], G[q][i], gamma);
                }
                if (gamma * gamma > 1e-26 * alpha * beta && gamma != 0.0) {
                    rotated = true;
                    const double zeta = fast_div(beta - alpha, 2.0 * gamma);
                    const double t = copysign(fast_rcp(fabs(zeta) + fast_sqrt(fma(zeta, zeta, 1.0))), zeta);
                    const double c = rsqrt(fma(t, t, 1.0));
                    const double s = c * t;
#pragma unroll
                    for (int i = 0; i < MR; ++i) {
                        const double gp = G[p][i], gq = G[q][i];
                        G[p][i] = c * gp - s * gq;
                        G[q][i] = s * gp + c * gq;
                    }
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double vp = V[p][i], vq = V[q][i];
                        V[p][i] = c * vp - s * vq;
                        V[q][i] = s * vp + c * vq;
                    }
                }
            }
        }
        if (!rotated) break;
    }
}

__device__ __forceinline__ void cross3(const double (&a)[3], const double (&b)[3], double (&c)[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// SVD of a row-major 3x3 matrix by one-sided Jacobi, singular values sorted descending.
// U columns u[i][.], V columns v[i][.] ; u[2] = u[0] x u[1], v[2] = v[0] x v[1]  (det +1 both,
// which is what extract_Rt's det fixes produce from LAPACK's factors, slam.py:171-172).
__device__ __forceinline__ void svd3(const double (&M)[9], double (&u)[3][3], double (&s)[3], double (&v)[3][3]) {
    double G[3][3], V[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) G[c][r] = M[r * 3 + c];
    jacobi_onesided<3, 3>(G, V, 30);
    double n[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) n[c] = fast_sqrt(G[c][0] * G[c][0] + G[c][1] * G[c][1] + G[c][2] * G[c][2]);
    // order: i0 = largest, i1 = second (static-index selection via swaps)
#define TVG_SWAPCOL(a, b)                                                            \
    {                                                                                \
        if (n[a] < n[b]) {                                                           \
            double tmp = n[a]; n[a] = n[b]; n[b] = tmp;                              \
            _Pragma("unroll") for (int r = 0; r < 3; ++r) {                          \
                tmp = G[a][r]; G[a][r] = G[b][r]; G[b][r] = tmp;                     \
                tmp = V[a][r]; V[a][r] = V[b][r]; V[b][r] = tmp;                     \
            }                                                                        \
        }                                                                            \
    }
    TVG_SWAPCOL(0, 1) TVG_SWAPCOL(0, 2) TVG_SWAPCOL(1, 2)
#undef TVG_SWAPCOL
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double inv = fast_rcp(n[c]);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            u[c][r] = G[c][r] * inv;
            v[c][r] = V[c][r];
        }
        s[c] = n[c];
    }
    s[2] = n[2];
    cross3(u[0], u[1], u[2]);
    cross3(v[0], v[1], v[2]);
}

// EssentialMatrixTransform.estimate's internal constraint:
// S = ((s0+s1)/2, (s0+s1)/2, 0);  out = U diag(S) V^T   (row-major 3x3).
__device__ __forceinline__ void essential_project(const double (&M)[9], double (&out)[9]) {
    double u[3][3], s[3], v[3][3];
    svd3(M, u, s, v);
    const double sb = 0.5 * (s[0] + s[1]);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) out[r * 3 + c] = sb * (u[0][r] * v[0][c] + u[1][r] * v[1][c]);
}

// The same projection without a full SVD (the per-hypothesis fit runs it a million times per batch):
//   U diag(1,1,0) V^T = M [p q] B^(-1/2) [p q]^T   with  v3 = right singular vector of the smallest singular value,
//   {p, q} an orthonormal basis of its complement, B = [p q]^T M^T M [p q]  (2x2, closed-form inverse square root),
//   (s0 + s1) / 2 = tr(B^(1/2)) / 2.
// v3 is the dominant eigenvector of adj(M^T M) = V diag(l2 l3, l1 l3, l1 l2) V^T, found by repeated squaring of that
// symmetric 3x3 (18 multiply-adds a step; separation (s3 / s2)^(2^(k+1)) after k steps; tr(C^2) = tr(C)^2 when done).
// ~250 fp64 operations against ~1000 for the one-sided Jacobi SVD.  Returns false when the two smallest singular
// values are too close to separate in 10 steps (then s2 ~ s3 and the caller falls back to the SVD).
__device__ __forceinline__ bool essential_project_fast(const double (&M)[9], double (&out)[9]) {
    // S = M^T M: 00 01 02 11 12 22
    const double a = fma(M[0], M[0], fma(M[3], M[3], M[6] * M[6])), b = fma(M[0], M[1], fma(M[3], M[4], M[6] * M[7]));
    const double c = fma(M[0], M[2], fma(M[3], M[5], M[6] * M[8])), d = fma(M[1], M[1], fma(M[4], M[4], M[7] * M[7]));
    const double e = fma(M[1], M[2], fma(M[4], M[5], M[7] * M[8])), f = fma(M[2], M[2], fma(M[5], M[5], M[8] * M[8]));
    double C[6] = {d * f - e * e, c * e - b * f, b * e - c * d, a * f - c * c, b * c - a * e, a * d - b * b};   // adj(S)
    double tr = C[0] + C[3] + C[5];
    if (!(tr > 0.0) || !isfinite(tr)) return false;
    bool conv = false;
    for (int k = 0; k < 10; k += 2) {
        {   // keep the entries in range (any scale will do)
            double r;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(tr));
#pragma unroll
            for (int i = 0; i < 6; ++i) C[i] *= r;
            tr = C[0] + C[3] + C[5];
        }
        double D[6];
        D[0] = fma(C[0], C[0], fma(C[1], C[1], C[2] * C[2]));
        D[1] = fma(C[0], C[1], fma(C[1], C[3], C[2] * C[4]));
        D[2] = fma(C[0], C[2], fma(C[1], C[4], C[2] * C[5]));
        D[3] = fma(C[1], C[1], fma(C[3], C[3], C[4] * C[4]));
        D[4] = fma(C[1], C[2], fma(C[3], C[4], C[4] * C[5]));
        D[5] = fma(C[2], C[2], fma(C[4], C[4], C[5] * C[5]));
        const double trd = D[0] + D[3] + D[5];
        const bool conv1 = trd >= (1.0 - 2e-14) * tr * tr;
        C[0] = fma(D[0], D[0], fma(D[1], D[1], D[2] * D[2]));
        C[1] = fma(D[0], D[1], fma(D[1], D[3], D[2] * D[4]));
        C[2] = fma(D[0], D[2], fma(D[1], D[4], D[2] * D[5]));
        C[3] = fma(D[1], D[1], fma(D[3], D[3], D[4] * D[4]));
        C[4] = fma(D[1], D[2], fma(D[3], D[4], D[4] * D[5]));
        C[5] = fma(D[2], D[2], fma(D[4], D[4], D[5] * D[5]));
        tr = C[0] + C[3] + C[5];
        conv = conv1 || tr >= (1.0 - 2e-14) * trd * trd;
        if (conv) break;
    }
    if (!conv || !(tr > 0.0)) return false;
    // v3 = the column with the largest diagonal entry, normalised
    double v[3] = {C[0], C[1], C[2]};
    double best = C[0];
    if (C[3] > best) { best = C[3]; v[0] = C[1]; v[1] = C[3]; v[2] = C[4]; }
    if (C[5] > best) { best = C[5]; v[0] = C[2]; v[1] = C[4]; v[2] = C[5]; }
    {
        const double inv = rsqrt(fma(v[0], v[0], fma(v[1], v[1], v[2] * v[2])));
        v[0] *= inv; v[1] *= inv; v[2] *= inv;
    }
    // p = normalise(v3 x e_k) for the axis k of the smallest |component|, q = v3 x p
    double p[3], q[3];
    const double a0 = fabs(v[0]), a1 = fabs(v[1]), a2 = fabs(v[2]);
    if (a0 <= a1 && a0 <= a2) { p[0] = 0.0; p[1] = v[2]; p[2] = -v[1]; }
    else if (a1 <= a2) { p[0] = -v[2]; p[1] = 0.0; p[2] = v[0]; }
    else { p[0] = v[1]; p[1] = -v[0]; p[2] = 0.0; }
    {
        const double inv = rsqrt(fma(p[0], p[0], fma(p[1], p[1], p[2] * p[2])));
        p[0] *= inv; p[1] *= inv; p[2] *= inv;
    }
    cross3(v, p, q);
    double Mp[3], Mq[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        Mp[r] = fma(M[r * 3], p[0], fma(M[r * 3 + 1], p[1], M[r * 3 + 2] * p[2]));
        Mq[r] = fma(M[r * 3], q[0], fma(M[r * 3 + 1], q[1], M[r * 3 + 2] * q[2]));
    }
    const double b11 = fma(Mp[0], Mp[0], fma(Mp[1], Mp[1], Mp[2] * Mp[2]));
    const double b12 = fma(Mp[0], Mq[0], fma(Mp[1], Mq[1], Mp[2] * Mq[2]));
    const double b22 = fma(Mq[0], Mq[0], fma(Mq[1], Mq[1], Mq[2] * Mq[2]));
    const double det = fma(b11, b22, -b12 * b12);
    if (!(det > 0.0)) return false;
    const double dd = fast_sqrt(det);                 // s0 * s1
    const double tt = fast_sqrt(b11 + b22 + 2.0 * dd);   // s0 + s1
    // B^(1/2) = (B + dd I) / tt with determinant dd, so B^(-1/2) = adj(B + dd I) / (tt dd); times (s0 + s1) / 2 = tt / 2
    const double h = 0.5 * fast_rcp(dd);
    const double i11 = (b22 + dd) * h, i12 = -b12 * h, i22 = (b11 + dd) * h;
    double W0[3], W1[3];   // [Mp Mq] * (tt/2) B^(-1/2)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        W0[r] = fma(Mp[r], i11, Mq[r] * i12);
        W1[r] = fma(Mp[r], i12, Mq[r] * i22);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) out[r * 3 + cc] = fma(W0[r], p[cc], W1[r] * q[cc]);
    return true;
}

// params = T_d^T * E * T_s   with T = [[s,0,-s cx],[0,s,-s cy],[0,0,1]]
__device__ __forceinline__ void essential_denormalise(const double (&En)[9], double ss, double csx, double csy,
                                                     double sd, double cdx, double cdy, double (&E)[9]) {
    double X[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        X[r * 3 + 0] = En[r * 3 + 0] * ss;
        X[r * 3 + 1] = En[r * 3 + 1] * ss;
        X[r * 3 + 2] = En[r * 3 + 2] - ss * (csx * En[r * 3 + 0] + csy * En[r * 3 + 1]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        E[0 * 3 + c] = sd * X[0 * 3 + c];
        E[1 * 3 + c] = sd * X[1 * 3 + c];
        E[2 * 3 + c] = X[2 * 3 + c] - sd * (cdx * X[0 * 3 + c] + cdy * X[1 * 3 + c]);
    }
}

// Null vector of the 8x9 DLT system whose rows are a[k][0..8]: Householder QR of A^T (9x8),
// n = H_0 ... H_7 e_8.  Backward stable, no pivoting, fixed cost, all indices static.
__device__ __forceinline__ void null_vector_8x9(double (&a)[8][9], double (&n)[9]) {
    double beta[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double sig = 0.0;
#pragma unroll
        for (int i = k; i < 9; ++i) sig = fma(a[k][i], a[k][i], sig);
        const double nrm = fast_sqrt(sig);
        const double alpha = (a[k][k] >= 0.0) ? -nrm : nrm;
        const double vtv = 2.0 * (sig - alpha * a[k][k]);
        beta[k] = (vtv > 0.0) ? 2.0 * fast_rcp(vtv) : 0.0;
        a[k][k] -= alpha;
#pragma unroll
        for (int j = k + 1; j < 8; ++j) {
            double dot = 0.0;
#pragma unroll
            for (int i = k; i < 9; ++i) dot = fma(a[k][i], a[j][i], dot);
            const double f = beta[k] * dot;
#pragma unroll
            for (int i = k; i < 9; ++i) a[j][i] = fma(-f, a[k][i], a[j][i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) n[i] = (i == 8) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        double dot = 0.0;
#pragma unroll
        for (int i = k; i < 9; ++i) dot = fma(a[k][i], n[i], dot);
        const double f = beta[k] * dot;
#pragma unroll
        for (int i = k; i < 9; ++i) n[i] = fma(-f, a[k][i], n[i]);
    }
}

// skimage-style residual in fp64:
// |x'^T E x| / sqrt((Ex)_0^2 + (Ex)_1^2 + (E^T x')_0^2 + (E^T x')_1^2)
__device__ __forceinline__ double sampson_residual(const double (&E)[9], double x, double y, double u, double v) {
    const double f0 = E[0] * x + E[1] * y + E[2];
    const double f1 = E[3] * x + E[4] * y + E[5];
    const double f2 = E[6] * x + E[7] * y + E[8];
    const double g0 = E[0] * u + E[3] * v + E[6];
    const double g1 = E[1] * u + E[4] * v + E[7];
    const double num = u * f0 + v * f1 + f2;
    return fabs(num) / sqrt(f0 * f0 + f1 * f1 + g0 * g0 + g1 * g1);
}

// U and V of a numerically rank-two 3x3 matrix (an essential matrix: what extract_Rt is given) without iterating:
// v2 = the largest cross product of two rows (the null vector), {p, q} an orthonormal basis of its complement, then ONE
// exact Jacobi rotation diagonalises the 2x2 Gram matrix of M [p q]; u_i = M v_i / s_i.  Same conventions as svd3
// (s0 >= s1, u2 = u0 x u1, v2 = v0 x v1).  ~120 fp64 operations against ~1000.  Returns false -- caller falls back to
// svd3 -- unless |M v2| <= 1e-10 |M|_F and s1 / s0 > ~1e-4.
__device__ __forceinline__ bool svd3_rank2(const double (&M)[9], double (&u)[3][3], double (&v)[3][3]) {
    const double r0[3] = {M[0], M[1], M[2]}, r1[3] = {M[3], M[4], M[5]}, r2[3] = {M[6], M[7], M[8]};
    double w[3], c1[3], c2[3];
    cross3(r0, r1, w);
    cross3(r0, r2, c1);
    cross3(r1, r2, c2);
    double nw = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double n1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
    const double n2 = c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2];
    if (n1 > nw) { nw = n1; w[0] = c1[0]; w[1] = c1[1]; w[2] = c1[2]; }
    if (n2 > nw) { nw = n2; w[0] = c2[0]; w[1] = c2[1]; w[2] = c2[2]; }
    double fro2 = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) fro2 = fma(M[i], M[i], fro2);
    if (!(nw > 1e-8 * fro2 * fro2) || !isfinite(fro2)) return false;
    const double iw = rsqrt(nw);
    const double z[3] = {w[0] * iw, w[1] * iw, w[2] * iw};
    double res = 0.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double e = M[r * 3] * z[0] + M[r * 3 + 1] * z[1] + M[r * 3 + 2] * z[2];
        res = fma(e, e, res);
    }
    if (res > 1e-20 * fro2) return false;
    // p = e_k x z for the axis k where |z_k| is smallest (never parallel to z), q = z x p
    const double a0 = fabs(z[0]), a1 = fabs(z[1]), a2 = fabs(z[2]);
    double ax[3] = {0.0, 0.0, 0.0};
    if (a0 <= a1 && a0 <= a2) ax[0] = 1.0;
    else if (a1 <= a2) ax[1] = 1.0;
    else ax[2] = 1.0;
    double pv[3], qv[3];
    cross3(ax, z, pv);
    const double ip = rsqrt(pv[0] * pv[0] + pv[1] * pv[1] + pv[2] * pv[2]);
    pv[0] *= ip; pv[1] *= ip; pv[2] *= ip;
    cross3(z, pv, qv);
    double cp[3], cq[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        cp[r] = M[r * 3] * pv[0] + M[r * 3 + 1] * pv[1] + M[r * 3 + 2] * pv[2];
        cq[r] = M[r * 3] * qv[0] + M[r * 3 + 1] * qv[1] + M[r * 3 + 2] * qv[2];
    }
    const double alpha = cp[0] * cp[0] + cp[1] * cp[1] + cp[2] * cp[2];
    const double beta = cq[0] * cq[0] + cq[1] * cq[1] + cq[2] * cq[2];
    const double gamma = cp[0] * cq[0] + cp[1] * cq[1] + cp[2] * cq[2];
    double c = 1.0, sn = 0.0;
    if (gamma * gamma > 1e-32 * alpha * beta && gamma != 0.0) {
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = copysign(1.0 / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0))), zeta);
        c = 1.0 / sqrt(fma(t, t, 1.0));
        sn = c * t;
    }
    double g0[3], g1[3], v0[3], v1[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        g0[r] = c * cp[r] - sn * cq[r];
        g1[r] = sn * cp[r] + c * cq[r];
        v0[r] = c * pv[r] - sn * qv[r];
        v1[r] = sn * pv[r] + c * qv[r];
    }
    const double m0 = g0[0] * g0[0] + g0[1] * g0[1] + g0[2] * g0[2];
    const double m1 = g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2];
    const bool sw = m0 < m1;
    const double i0 = 1.0 / sqrt(sw ? m1 : m0), i1 = 1.0 / sqrt(sw ? m0 : m1);
    if (!isfinite(i0) || !isfinite(i1)) return false;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        u[0][r] = (sw ? g1[r] : g0[r]) * i0;
        u[1][r] = (sw ? g0[r] : g1[r]) * i1;
        v[0][r] = sw ? v1[r] : v0[r];
        v[1][r] = sw ? v0[r] : v1[r];
    }
    cross3(u[0], u[1], u[2]);
    cross3(v[0], v[1], v[2]);
    return true;
}

// extract_Rt (slam.py:167-186): R = U W V^T, or U W^T V^T when trace(R) < 0; t = U[:,2] with t_z >= 0.
__device__ __forceinline__ void extract_rt(const double (&E)[9], double (&Rt)[16]) {
    double u[3][3], s[3], v[3][3];
    if (!svd3_rank2(E, u, v)) svd3(E, u, s, v);
    // U W = [u1, -u0, u2] ; U W^T = [-u1, u0, u2]   (columns)
    double Ra[9], Rb[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double x = u[1][r] * v[0][c] - u[0][r] * v[1][c];
            const double y = u[2][r] * v[2][c];
            Ra[r * 3 + c] = x + y;
            Rb[r * 3 + c] = -x + y;
        }
    const bool use_b = (Ra[0] + Ra[4] + Ra[8]) < 0.0;
    const double sg = (u[2][2] < 0.0) ? -1.0 : 1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Rt[r * 4 + c] = use_b ? Rb[r * 3 + c] : Ra[r * 3 + c];
        Rt[r * 4 + 3] = sg * u[2][r];
    }
    Rt[12] = 0.0; Rt[13] = 0.0; Rt[14] = 0.0; Rt[15] = 1.0;
}

// General 4x4 inverse, Gauss-Jordan with partial pivoting (np.linalg.inv at slam.py:204-205).
__device__ inline void inverse4(const double* A, double* Ainv) {
    double m[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            m[r][c] = A[r * 4 + c];
            m[r][4 + c] = (r == c) ? 1.0 : 0.0;
        }
    for (int k = 0; k < 4; ++k) {
        int piv = k;
        double best = fabs(m[k][k]);
        for (int r = k + 1; r < 4; ++r)
            if (fabs(m[r][k]) > best) { best = fabs(m[r][k]); piv = r; }
        if (piv != k)
            for (int c = 0; c < 8; ++c) { const double t = m[k][c]; m[k][c] = m[piv][c]; m[piv][c] = t; }
        const double inv = 1.0 / m[k][k];
        for (int c = 0; c < 8; ++c) m[k][c] *= inv;
        for (int r = 0; r < 4; ++r) {
            if (r == k) continue;
            const double f = m[r][k];
            for (int c = 0; c < 8; ++c) m[r][c] = fma(-f, m[k][c], m[r][c]);
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) Ainv[r * 4 + c] = m[r][4 + c];
}

}  // namespace tvg
