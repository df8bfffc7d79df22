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

// extract_Rt (slam.py:167-186): R = U W V^T, or U W^T V^T when trace(R) < 0; t = U[:,2] with t_z >= 0.
__device__ __forceinline__ void extract_rt(const double (&E)[9], double (&Rt)[16]) {
    double u[3][3], s[3], v[3][3];
    svd3(E, u, s, v);
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
