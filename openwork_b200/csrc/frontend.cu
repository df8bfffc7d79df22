// Feature front-end  --  replaces extract_points of
// /root/reference/MonocularSlam/slam.py:112-122:
//     image = cv2.cvtColor(frame.image, cv2.COLOR_BGR2GRAY)
//     pts   = cv2.goodFeaturesToTrack(image, 3000, qualityLevel=0.01, minDistance=3)
//     kps, des = cv2.ORB_create().compute(image, [cv2.KeyPoint(x, y, size=20) ...])
// Every float operation below follows the order (and the fused / unfused form) that was found to reproduce
// opencv-python 4.13 bit for bit on the CPU (tests/test_frontend.py documents the probes):
//   gray      = (B*3735 + G*19235 + R*9798 + 2^14) >> 15
//   Sobel 3x3 with scale s = 1/(255*4*3), BORDER_REFLECT_101:
//       Dx = fma(s, r(y-1) + r(y+1), (2s) * r(y)),  r(y') = p[y'][x+1] - p[y'][x-1]
//       Dy = q(y+1) - q(y-1),  q(y') = fma(s, p[y'][x+1], fma(2s, p[y'][x], s * p[y'][x-1]))
//   cov       = 3x3 unnormalised box sums (accumulated in double) of Dx*Dx, Dx*Dy, Dy*Dy
//   minEig    = (a + c) - sqrt((a - c)^2 + b^2),  a = cov_xx/2, b = cov_xy, c = cov_yy/2   (no FMA)
//   corners   = pixels > max*quality that equal their 3x3 maximum, sorted by (value, address) descending,
//               greedily kept when no kept corner lies closer than minDistance (grid of cells), first 3000
//   ORB       = keypoints inside the 31-pixel border; image blurred 7x7 sigma 2 in fp32 (rows: sequential
//               FMAs, columns: centre-first symmetric FMAs), rounded to u8; 256 pairwise tests on the learned
//               pattern rotated by the KeyPoint default angle (-1 degree: no orientation is computed by
//               ORB.compute for user-supplied keypoints)
// (OpenCV's boxFilter keeps double running sums along rows and columns; a direct double 9-sum differs from it
//  in the last bit of cov_xy for ~1e-5 of the pixels, which has never changed a keypoint in the tests.)
#include "tvg_common.cuh"
#include "tvg_kernels.h"

namespace tvg {

__constant__ int8_t c_orb_pattern[512 * 2];
__device__ alignas(4) int8_t g_orb_pattern[512 * 2];   // same table in global memory: coalesced staging into shared memory
__constant__ float c_gauss7[7];

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

__global__ void gray_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ gray, size_t npix, int channels) {
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;    // four pixels per thread
    if (i4 >= npix) return;
    if (i4 + 4 <= npix) {
        if (channels == 1) {
            *reinterpret_cast<uint32_t*>(gray + i4) = *reinterpret_cast<const uint32_t*>(img + i4);
        } else {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(img + 3 * i4);   // 12 bytes = B G R x 4
            const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
            const uint32_t by[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24,
                                     w1 & 255u, (w1 >> 8) & 255u, (w1 >> 16) & 255u, w1 >> 24,
                                     w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
            uint32_t out = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                out |= ((by[3 * k] * 3735u + by[3 * k + 1] * 19235u + by[3 * k + 2] * 9798u + (1u << 14)) >> 15) << (8 * k);
            *reinterpret_cast<uint32_t*>(gray + i4) = out;
        }
    } else {
        for (size_t i = i4; i < npix; ++i) {
            if (channels == 1) gray[i] = img[i];
            else gray[i] = (uint8_t)((img[3 * i] * 3735u + img[3 * i + 1] * 19235u + img[3 * i + 2] * 9798u + (1u << 14)) >> 15);
        }
    }
}

// exact u8 -> float without the conversion unit: 2^23 + b has b in its low mantissa bits
__device__ __forceinline__ float u8f(unsigned b) { return __fsub_rn(__uint_as_float(0x4B000000u | b), 8388608.0f); }

// Sobel derivatives and their products at one pixel: (Dx*Dx, Dx*Dy, Dy*Dy)
__device__ __forceinline__ void cov_at(const uint8_t* __restrict__ g, int y, int x, int H, int W, float (&c)[3]) {
    const float s = (float)(1.0 / (255.0 * 4 * 3)), s2 = (float)(2.0 / (255.0 * 4 * 3));
    const int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
    float r[3], q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint8_t* row = g + (size_t)reflect101(y - 1 + k, H) * W;
        const float pm = u8f(row[xm]), p0 = u8f(row[x]), pp = u8f(row[xp]);
        r[k] = __fsub_rn(pp, pm);
        q[k] = __fmaf_rn(s, pp, __fmaf_rn(s2, p0, __fmul_rn(s, pm)));
    }
    const float dx = __fmaf_rn(s, __fadd_rn(r[0], r[2]), __fmul_rn(s2, r[1]));
    const float dy = __fsub_rn(q[2], q[0]);
    c[0] = __fmul_rn(dx, dx);
    c[1] = __fmul_rn(dx, dy);
    c[2] = __fmul_rn(dy, dy);
}

// cornerMinEigenVal fused: a CTA builds the three product planes of an (EIG_ROWS + 2) x 130 tile in shared memory
// (BORDER_REFLECT_101 of the PRODUCT planes, like cv::boxFilter), then every thread walks its pixel column with a
// rolling three-row window of horizontal 3-sums (double).  No product plane ever goes to global memory.  The tile
// is stored as double: one float->double conversion per tile entry instead of one per use (the conversion unit,
// 16 lanes/clk/SM, is what bounds this kernel).
constexpr int EIG_ROWS = 16;
constexpr int EIG_TW = 132;
using eig_tile_t = float;   // double halves the conversions but the 57 KB tile costs more in occupancy (measured)
constexpr int EIG_SMEM = 3 * (EIG_ROWS + 2) * EIG_TW * (int)sizeof(eig_tile_t);

__global__ void __launch_bounds__(128) eig_kernel(const uint8_t* __restrict__ gray, float* __restrict__ eig,
                                                  int* __restrict__ max_bits, int H, int W) {
    extern __shared__ eig_tile_t s_cov[];   // [3][EIG_ROWS + 2][EIG_TW]
    constexpr int PLANE = (EIG_ROWS + 2) * EIG_TW;
    const int x0 = blockIdx.x * 128, y0 = blockIdx.y * EIG_ROWS, b = blockIdx.z;
    const uint8_t* g = gray + (size_t)b * H * W;
    auto put = [&](int ty, int tx, const float (&c)[3]) {
        s_cov[ty * EIG_TW + tx] = (eig_tile_t)c[0];
        s_cov[PLANE + ty * EIG_TW + tx] = (eig_tile_t)c[1];
        s_cov[2 * PLANE + ty * EIG_TW + tx] = (eig_tile_t)c[2];
    };
    if (y0 >= 2 && y0 + EIG_ROWS + 1 <= H - 1) {
        // Rows away from the top/bottom border: a thread walks down tile column threadIdx.x keeping the horizontal
        // Sobel parts (difference r, smoothing q) of three gray rows in registers: 3 loads per entry instead of 9.
        const float s = (float)(1.0 / (255.0 * 4 * 3)), s2 = (float)(2.0 / (255.0 * 4 * 3));
        const int xx = reflect101(min(x0 - 1 + (int)threadIdx.x, W), W);
        const int xm = reflect101(xx - 1, W), xp = reflect101(xx + 1, W);
        const uint8_t* row = g + (size_t)(y0 - 2) * W;
        auto rq = [&](float& r, float& q) {
            const float pm = u8f(row[xm]), p0 = u8f(row[xx]), pp = u8f(row[xp]);
            r = __fsub_rn(pp, pm);
            q = __fmaf_rn(s, pp, __fmaf_rn(s2, p0, __fmul_rn(s, pm)));
            row += W;
        };
        float r0, q0, r1, q1;
        rq(r0, q0);
        rq(r1, q1);
#pragma unroll
        for (int ty = 0; ty < EIG_ROWS + 2; ++ty) {
            float r2, q2;
            rq(r2, q2);
            const float dx = __fmaf_rn(s, __fadd_rn(r0, r2), __fmul_rn(s2, r1));
            const float dy = __fsub_rn(q2, q0);
            const float c[3] = {__fmul_rn(dx, dx), __fmul_rn(dx, dy), __fmul_rn(dy, dy)};
            put(ty, threadIdx.x, c);
            r0 = r1; r1 = r2; q0 = q1; q1 = q2;
        }
        if (threadIdx.x < 2 * (EIG_ROWS + 2)) {   // the two halo columns on the right
            const int ty = threadIdx.x >> 1, tx = 128 + (threadIdx.x & 1);
            float c[3];
            cov_at(g, y0 - 1 + ty, reflect101(min(x0 - 1 + tx, W), W), H, W, c);
            put(ty, tx, c);
        }
    } else {
        for (int i = threadIdx.x; i < (EIG_ROWS + 2) * 130; i += 128) {
            const int ty = i / 130, tx = i % 130;
            const int yy = reflect101(min(y0 - 1 + ty, H), H), xx = reflect101(min(x0 - 1 + tx, W), W);
            float c[3];
            cov_at(g, yy, xx, H, W, c);
            put(ty, tx, c);
        }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    float m = 0.0f;
    if (x < W) {
        const int tx = threadIdx.x + 1;
        double h[3][3];   // h[row slot][plane]
        auto hsum = [&](int ty, double (&o)[3]) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const eig_tile_t* row = s_cov + k * PLANE + ty * EIG_TW + tx;
                o[k] = ((double)row[-1] + (double)row[0]) + (double)row[1];
            }
        };
        hsum(0, h[0]);
        hsum(1, h[1]);
        const int y1 = min(H, y0 + EIG_ROWS);
        for (int y = y0; y < y1; ++y) {
            hsum(y - y0 + 2, h[2]);
            const float a = __fmul_rn((float)((h[0][0] + h[1][0]) + h[2][0]), 0.5f);
            const float bb = (float)((h[0][1] + h[1][1]) + h[2][1]);
            const float cc = __fmul_rn((float)((h[0][2] + h[1][2]) + h[2][2]), 0.5f);
            const float d = __fsub_rn(a, cc);
            const float e = __fsub_rn(__fadd_rn(a, cc), __fsqrt_rn(__fadd_rn(__fmul_rn(d, d), __fmul_rn(bb, bb))));
            eig[(size_t)b * H * W + (size_t)y * W + x] = e;
            m = fmaxf(m, e);
#pragma unroll
            for (int k = 0; k < 3; ++k) { h[0][k] = h[1][k]; h[1][k] = h[2][k]; }
        }
    }
    // image maximum (cv::minMaxLoc): positive floats order like their bit patterns
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(max_bits + b, __float_as_int(m));
}

// threshold + 3x3 non-maximum suppression; candidates are appended as sortable 64-bit keys
__global__ void __launch_bounds__(128) collect_kernel(const float* __restrict__ eig, const int* __restrict__ max_bits,
                                                      double quality, unsigned long long* __restrict__ keys,
                                                      int* __restrict__ n_cand, int cap, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = blockIdx.y * EIG_ROWS, b = blockIdx.z;
    const bool live = x >= 1 && x < W - 1;
    const int xs = live ? x : 1;            // out-of-range lanes shadow a valid column and never emit
    const float* e = eig + (size_t)b * H * W;
    const float thr = (float)((double)__int_as_float(max_bits[b]) * quality);
    // rolling window over rows: horizontal maxima of the thresholded map (THRESH_TOZERO) and the centre values
    auto hmax = [&](int yy, float& centre) -> float {
        if (yy < 0 || yy >= H) { centre = 0.0f; return 0.0f; }
        const float* row = e + (size_t)yy * W;
        const float l = row[xs - 1], c = row[xs], r = row[xs + 1];
        centre = c > thr ? c : 0.0f;
        return fmaxf(fmaxf(l > thr ? l : 0.0f, centre), r > thr ? r : 0.0f);
    };
    // Candidates are staged in shared memory (a pixel column holds at most EIG_ROWS/2 local maxima in EIG_ROWS
    // rows); the CTA then reserves its output range with ONE global atomic, so no thread ever waits for an atomic's
    // return value inside the row loop.
    __shared__ unsigned long long s_keys[128 * (EIG_ROWS / 2)];
    __shared__ int s_n, s_base;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    float c0, c1, c2;
    float m0 = hmax(y0 - 1, c0), m1 = hmax(y0, c1);
    const int y1 = min(H - 1, y0 + EIG_ROWS);
    for (int y = y0; y < y1; ++y) {
        const float m2 = hmax(y + 1, c2);
        if (live && y >= 1 && c1 != 0.0f && c1 == fmaxf(fmaxf(m0, m1), m2))
            s_keys[atomicAdd(&s_n, 1)] = ((unsigned long long)__float_as_uint(c1) << 32) | (unsigned)(y * W + x);
        m0 = m1; m1 = m2; c0 = c1; c1 = c2;
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    if (threadIdx.x == 0) s_base = atomicAdd(n_cand + b, n);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int slot = s_base + i;
        if (slot < cap) keys[(size_t)b * cap + slot] = s_keys[i];
    }
}

// Bitonic sort (descending) of the candidate keys of one image, one CTA per image.  Compare-exchange levels whose
// stride is below SORT_CHUNK stay inside aligned SORT_CHUNK-key blocks and run in shared memory, up to three levels
// per pass with the 2/4/8 keys of a butterfly held in registers (35 passes instead of 91 for 8192 keys, each key
// read and written once per pass).  Only the few levels with larger strides touch global memory (L2 resident).
// Shared indices are padded by i/16 so the stride-1 pass (8 consecutive keys per thread) is bank-conflict free.
constexpr int SORT_CHUNK = 8192;
constexpr int SORT_SMEM = (SORT_CHUNK + SORT_CHUNK / 16) * (int)sizeof(unsigned long long);

__device__ __forceinline__ int spad(int i) { return i + (i >> 4); }

// strides low << (NLEV-1) ... low of the merge stage `size`, on the chunk starting at global index c0
template <int NLEV>
__device__ __forceinline__ void sort_pass(unsigned long long* s, int chunk, int c0, int size, int low) {
    constexpr int E = 1 << NLEV;
    for (int g = threadIdx.x; g < chunk / E; g += blockDim.x) {
        const int base = ((g & ~(low - 1)) << NLEV) | (g & (low - 1));
        unsigned long long r[E];
#pragma unroll
        for (int e = 0; e < E; ++e) r[e] = s[spad(base + e * low)];
        const bool desc = ((c0 + base) & size) == 0;
#pragma unroll
        for (int d = E >> 1; d > 0; d >>= 1)
#pragma unroll
            for (int e = 0; e < E; ++e)
                if (!(e & d)) {
                    const unsigned long long a = r[e], c = r[e | d];
                    const bool sw = desc ? (a < c) : (a > c);
                    r[e] = sw ? c : a;
                    r[e | d] = sw ? a : c;
                }
#pragma unroll
        for (int e = 0; e < E; ++e) s[spad(base + e * low)] = r[e];
    }
    __syncthreads();
}

// the `levels` lowest-stride levels (strides 2^(levels-1) ... 1) of merge stage `size`
__device__ __forceinline__ void sort_levels(unsigned long long* s, int chunk, int c0, int size, int levels) {
    int rem = levels;
    while (rem > 3) {
        const int take = rem - 3 >= 3 ? 3 : rem - 3;
        const int low = 1 << (rem - take);
        if (take == 3) sort_pass<3>(s, chunk, c0, size, low);
        else if (take == 2) sort_pass<2>(s, chunk, c0, size, low);
        else sort_pass<1>(s, chunk, c0, size, low);
        rem -= take;
    }
    if (rem == 3) sort_pass<3>(s, chunk, c0, size, 1);
    else if (rem == 2) sort_pass<2>(s, chunk, c0, size, 1);
    else if (rem == 1) sort_pass<1>(s, chunk, c0, size, 1);
}

__global__ void __launch_bounds__(1024) sort_kernel(unsigned long long* __restrict__ keys, const int* __restrict__ n_cand,
                                                    int cap) {
    extern __shared__ unsigned long long skeys[];
    const int b = blockIdx.x;
    unsigned long long* k = keys + (size_t)b * cap;
    const int n = min(n_cand[b], cap);
    int N = 1;
    while (N < n) N <<= 1;
    const int chunk = min(N, SORT_CHUNK);
    const int chunk_levels = 31 - __clz(chunk);
    // phase 1: every chunk fully sorted in shared memory (all sizes <= chunk), directions as in the full network
    for (int c0 = 0; c0 < N; c0 += chunk) {
        for (int i = threadIdx.x; i < chunk; i += blockDim.x) skeys[spad(i)] = c0 + i < n ? k[c0 + i] : 0ull;
        __syncthreads();                                                 // zero padding sorts last
        for (int lv = 1; lv <= chunk_levels; ++lv) sort_levels(skeys, chunk, c0, 1 << lv, lv);
        for (int i = threadIdx.x; i < chunk; i += blockDim.x) k[c0 + i] = skeys[spad(i)];
        __syncthreads();
    }
    // phase 2: remaining merge stages; strides >= chunk in global memory, the rest per chunk in shared memory
    for (int size = chunk << 1; size <= N; size <<= 1) {
        for (int stride = size >> 1; stride >= chunk; stride >>= 1) {
            for (int p = threadIdx.x; p < N / 2; p += blockDim.x) {
                const int i = ((p & ~(stride - 1)) << 1) | (p & (stride - 1)), j = i | stride;
                const unsigned long long a = k[i], c = k[j];
                if ((i & size) == 0 ? (a < c) : (a > c)) { k[i] = c; k[j] = a; }
            }
            __syncthreads();
        }
        for (int c0 = 0; c0 < N; c0 += chunk) {
            for (int i = threadIdx.x; i < chunk; i += blockDim.x) skeys[spad(i)] = k[c0 + i];
            __syncthreads();
            sort_levels(skeys, chunk, c0, size, chunk_levels);
            for (int i = threadIdx.x; i < chunk; i += blockDim.x) k[c0 + i] = skeys[spad(i)];
            __syncthreads();
        }
    }
}

// Greedy minimum-distance filter of goodFeaturesToTrack, in sorted order, one warp per image.
// With cell = round(minDistance) and (cell-1)^2 * 2 < minDistance^2 a grid cell can hold one corner at most,
// (slots = 1); otherwise up to four (corners of a cell of side ~minDistance).  The grid lives in shared memory:
// one byte per slot = in-cell offset oy*cell+ox, 0xFF = empty (cell <= 15).
__global__ void __launch_bounds__(256) suppress_kernel(const unsigned long long* __restrict__ keys,
                                                       const int* __restrict__ n_cand, int cap, int H, int W, int cell,
                                                       int slots,
                                                       float min_dist, int max_corners, float2* __restrict__ corners,
                                                       int* __restrict__ n_corners) {
    extern __shared__ uint8_t grid[];
    const int b = blockIdx.x;
    const int gw = (W + cell - 1) / cell, gh = (H + cell - 1) / cell;
    for (int i = threadIdx.x; i < gw * gh * slots; i += blockDim.x) grid[i] = 0xFFu;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const unsigned long long* k = keys + (size_t)b * cap;
    const int n = min(n_cand[b], cap);
    const float md2 = min_dist * min_dist;
    int kept = 0;
    for (int base = 0; base < n && kept < max_corners; base += 32) {
        // Each lane tests one candidate against the corners accepted so far (grid); the 32 candidates are then
        // resolved against each other strictly in sorted order, which is what the sequential reference does.
        const bool valid = base + lane < n;
        const unsigned addr = valid ? (unsigned)k[base + lane] : 0u;
        const int y = (int)(addr / (unsigned)W), x = (int)(addr % (unsigned)W);
        const int xc = x / cell, yc = y / cell;
        bool bad = !valid;
        if (valid) {
            for (int cy = max(0, yc - 1); cy <= min(gh - 1, yc + 1); ++cy)
                for (int cx = max(0, xc - 1); cx <= min(gw - 1, xc + 1); ++cx)
                    for (int sl = 0; sl < slots; ++sl) {
                        const int g = grid[(cy * gw + cx) * slots + sl];
                        if (g == 0xFF) break;
                        const float dx = (float)(x - (cx * cell + g % cell)), dy = (float)(y - (cy * cell + g / cell));
                        bad = bad || (dx * dx + dy * dy < md2);
                    }
        }
        unsigned open = __ballot_sync(0xffffffffu, !bad);
        while (open && kept < max_corners) {
            const int t = __ffs(open) - 1;                      // earliest candidate still alive: accepted
            const int xt = __shfl_sync(0xffffffffu, x, t), yt = __shfl_sync(0xffffffffu, y, t);
            if (lane == t) {
                uint8_t* cellp = grid + (yc * gw + xc) * slots;
                int sl = 0;
                while (sl < slots - 1 && cellp[sl] != 0xFF) ++sl;   // a cell never needs more than `slots`
                cellp[sl] = (uint8_t)((y - yc * cell) * cell + (x - xc * cell));
                corners[(size_t)b * max_corners + kept] = make_float2((float)x, (float)y);
            }
            ++kept;
            const float dx = (float)(x - xt), dy = (float)(y - yt);
            const bool dead = lane <= t || (dx * dx + dy * dy < md2);
            open &= ~__ballot_sync(0xffffffffu, dead);
        }
        __syncwarp();
    }
    if (lane == 0) n_corners[b] = kept;
}

// 7x7 Gaussian blur, sigma 2, fp32 separable, BORDER_REFLECT_101.  A CTA filters the rows of a
// (BLUR_TY + 6) x 128 tile into shared memory and then the columns out of it.
constexpr int BLUR_TY = 16;

__global__ void __launch_bounds__(128) blur_kernel(const uint8_t* __restrict__ gray, uint8_t* __restrict__ blur, int H,
                                                   int W) {
    __shared__ float s_rows[BLUR_TY + 6][128];
    const int x = blockIdx.x * 128 + threadIdx.x, y0 = blockIdx.y * BLUR_TY, b = blockIdx.z;
    const uint8_t* g = gray + (size_t)b * H * W;
    if (x < W) {
        int xi[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) xi[k] = reflect101(x - 3 + k, W);
        for (int r = 0; r < BLUR_TY + 6; ++r) {
            const uint8_t* row = g + (size_t)reflect101(y0 - 3 + r, H) * W;
            float s = __fmul_rn(c_gauss7[0], u8f(row[xi[0]]));
#pragma unroll
            for (int k = 1; k < 7; ++k) s = __fmaf_rn(u8f(row[xi[k]]), c_gauss7[k], s);
            s_rows[r][threadIdx.x] = s;
        }
        const int y1 = min(H, y0 + BLUR_TY);
        for (int y = y0; y < y1; ++y) {
            const int r = y - y0 + 3;
            float s = __fmul_rn(c_gauss7[3], s_rows[r][threadIdx.x]);
#pragma unroll
            for (int j = 1; j <= 3; ++j)
                s = __fmaf_rn(__fadd_rn(s_rows[r - j][threadIdx.x], s_rows[r + j][threadIdx.x]), c_gauss7[3 - j], s);
            // saturate_cast<uchar>(float): round half to even, then clamp.  0 <= s < 256 here, so adding 1.5 * 2^23
            // leaves rint(s) in the low mantissa bits without touching the conversion unit.
            const unsigned v = __float_as_uint(__fadd_rn(s, 12582912.0f)) & 0x1FFu;
            blur[(size_t)b * H * W + (size_t)y * W + x] = (uint8_t)min(v, 255u);
        }
    }
}

// KeyPointsFilter::runByImageBorder(31), order preserving; one CTA per image
__global__ void __launch_bounds__(256) border_filter_kernel(const float2* __restrict__ corners,
                                                            const int* __restrict__ n_corners, int max_corners, int H,
                                                            int W, int border, double* __restrict__ kps,
                                                            int* __restrict__ n_kps) {
    __shared__ int s_scan[256 / 32];
    __shared__ int s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = min(n_corners[b], max_corners);
    const float2* c = corners + (size_t)b * max_corners;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + tid;
        float2 p = make_float2(0.f, 0.f);
        bool keep = false;
        if (i < n) {
            p = c[i];
            keep = p.x >= (float)border && p.x < (float)(W - border) && p.y >= (float)border && p.y < (float)(H - border);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_scan[w];
        const int pos = off + __popc(bal & ((1u << lane) - 1u));
        if (keep) {
            kps[((size_t)b * max_corners + pos) * 2] = (double)p.x;
            kps[((size_t)b * max_corners + pos) * 2 + 1] = (double)p.y;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 256 / 32; ++w) tot += s_scan[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) n_kps[b] = s_base;
}

// The 256 rBRIEF tests.  One warp per keypoint: the 31x32 patch around it is fetched with coalesced row loads
// into shared memory, the pattern sits in shared memory too (lanes index it divergently, which constant memory
// would serialise), then lane = descriptor byte: 16 patch reads, 8 comparisons.
constexpr int DESC_WARPS = 8;

__global__ void __launch_bounds__(DESC_WARPS * 32) orb_desc_kernel(const uint8_t* __restrict__ blur,
                                                                   const double* __restrict__ kps,
                                                                   const int* __restrict__ n_kps, int max_corners, int H,
                                                                   int W, uint8_t* __restrict__ des) {
    __shared__ alignas(4) int8_t s_pat[1024];
    __shared__ uint8_t s_patch[DESC_WARPS][31][32];
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x * DESC_WARPS >= n_kps[b]) return;   // whole CTA beyond the last keypoint
    reinterpret_cast<int*>(s_pat)[threadIdx.x] = reinterpret_cast<const int*>(g_orb_pattern)[threadIdx.x];
    __syncthreads();
    const int kp = blockIdx.x * DESC_WARPS + warp;
    if (kp >= n_kps[b]) return;
    const size_t o = (size_t)b * max_corners + kp;
    const int cx = __double2int_rn(kps[o * 2]), cy = __double2int_rn(kps[o * 2 + 1]);
    // rows cy-15 .. cy+15, columns cx-15 .. cx+16 (keypoints are >= 31 pixels from every border)
    const uint8_t* src = blur + (size_t)b * H * W + (size_t)(cy - 15) * W + (cx - 15) + lane;
#pragma unroll
    for (int r = 0; r < 31; ++r) s_patch[warp][r][lane] = src[(size_t)r * W];
    __syncwarp();
    uint32_t v = 0;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const int q = (lane * 8 + bit) * 4;
        const int t0 = s_patch[warp][15 + s_pat[q + 1]][15 + s_pat[q]];
        const int t1 = s_patch[warp][15 + s_pat[q + 3]][15 + s_pat[q + 2]];
        v |= (uint32_t)(t0 < t1) << bit;
    }
    des[o * 32 + lane] = (uint8_t)v;
}

cudaError_t frontend_set_tables(const int8_t* pattern_host, const float* gauss7_host, cudaStream_t stream) {
    cudaError_t e = cudaMemcpyToSymbolAsync(c_orb_pattern, pattern_host, 1024, 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbolAsync(g_orb_pattern, pattern_host, 1024, 0, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbolAsync(c_gauss7, gauss7_host, 7 * sizeof(float), 0, cudaMemcpyHostToDevice, stream);
}

cudaError_t launch_frontend(const FrontendParams& p, cudaStream_t stream) {
    const int H = p.H, W = p.W, B = p.batch;
    const size_t npix = (size_t)B * H * W;
    gray_kernel<<<(unsigned)((npix / 4 + 256) / 256), 256, 0, stream>>>(p.images, p.gray, npix, p.channels);
    cudaMemsetAsync(p.max_bits, 0, (size_t)B * sizeof(int), stream);
    cudaMemsetAsync(p.n_cand, 0, (size_t)B * sizeof(int), stream);
    dim3 gr((W + 127) / 128, (H + EIG_ROWS - 1) / EIG_ROWS, B);
    cudaError_t e = cudaFuncSetAttribute(eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EIG_SMEM);
    if (e != cudaSuccess) return e;
    eig_kernel<<<gr, 128, EIG_SMEM, stream>>>(p.gray, p.eig, p.max_bits, H, W);
    collect_kernel<<<gr, 128, 0, stream>>>(p.eig, p.max_bits, p.quality, p.keys, p.n_cand, p.cap, H, W);
    e = cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMEM);
    if (e != cudaSuccess) return e;
    sort_kernel<<<B, 1024, SORT_SMEM, stream>>>(p.keys, p.n_cand, p.cap);
    const int gw = (W + p.cell - 1) / p.cell, gh = (H + p.cell - 1) / p.cell;
    const size_t smem = (size_t)gw * gh * p.slots;
    e = cudaFuncSetAttribute(suppress_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    suppress_kernel<<<B, 256, smem, stream>>>(p.keys, p.n_cand, p.cap, H, W, p.cell, p.slots, p.min_distance, p.max_corners,
                                              p.corners, p.n_corners);
    blur_kernel<<<dim3((W + 127) / 128, (H + BLUR_TY - 1) / BLUR_TY, B), 128, 0, stream>>>(p.gray, p.blur, H, W);
    border_filter_kernel<<<B, 256, 0, stream>>>(p.corners, p.n_corners, p.max_corners, H, W, 31, p.kps, p.n_kps);
    orb_desc_kernel<<<dim3((p.max_corners + DESC_WARPS - 1) / DESC_WARPS, B), DESC_WARPS * 32, 0, stream>>>(
        p.blur, p.kps, p.n_kps, p.max_corners, H, W, p.des);
    return cudaGetLastError();
}

}  // namespace tvg
