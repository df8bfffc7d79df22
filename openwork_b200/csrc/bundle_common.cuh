// Device helpers shared by the two-view (bundle.cu) and the sliding-window (bundle_window.cu) bundle adjustment:
// world -> camera transforms, the left twist update, fixed-order block reductions.
#pragma once
#include "tvg_common.cuh"
#include "tvg_linalg.cuh"

namespace tvg {
namespace ba {

constexpr int BA_THREADS = 256;

struct Cam {
    double R[9], t[3];
};

__device__ __forceinline__ void cam_from_pose(const double* pose, Cam& c) {   // P = inv(pose), top 3x4
    double inv[16];
    inverse4(pose, inv);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < 3; ++k) c.R[r * 3 + k] = inv[r * 4 + k];
        c.t[r] = inv[r * 4 + 3];
    }
}

__device__ __forceinline__ void to_cam(const Cam& c, const double (&X)[3], double (&Xc)[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) Xc[r] = fma(c.R[r * 3], X[0], fma(c.R[r * 3 + 1], X[1], fma(c.R[r * 3 + 2], X[2], c.t[r])));
}

// exp([w]x) R, exp([w]x) t + v   (Rodrigues)
__device__ inline void cam_update(const Cam& c, const double (&d)[6], Cam& o) {
    const double th2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2], th = sqrt(th2);
    double A, B;   // sin(th)/th, (1 - cos(th))/th^2
    if (th < 1e-8) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; }
    else { A = sin(th) / th; B = (1.0 - cos(th)) / th2; }
    const double K[9] = {0.0, -d[2], d[1], d[2], 0.0, -d[0], -d[1], d[0], 0.0};
    double E[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double k2 = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) k2 += K[i * 3 + k] * K[k * 3 + j];
            E[i * 3 + j] = (i == j ? 1.0 : 0.0) + A * K[i * 3 + j] + B * k2;
        }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) o.R[i * 3 + j] = E[i * 3] * c.R[j] + E[i * 3 + 1] * c.R[3 + j] + E[i * 3 + 2] * c.R[6 + j];
        o.t[i] = E[i * 3] * c.t[0] + E[i * 3 + 1] * c.t[1] + E[i * 3 + 2] * c.t[2] + d[3 + i];
    }
}

// fixed-order block sum of NV values per thread -> s_out[NV] (same result for every launch)
template <int NV>
__device__ __forceinline__ void block_sum_vec(const double (&v)[NV], double (*s_part)[BA_THREADS / 32], double* s_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const double w = warp_sum_f64(v[i]);
        if (lane == 0) s_part[i][warp] = w;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < BA_THREADS / 32; ++w) t += s_part[threadIdx.x][w];
        s_out[threadIdx.x] = t;
    }
    __syncthreads();
}


}  // namespace ba
}  // namespace tvg
