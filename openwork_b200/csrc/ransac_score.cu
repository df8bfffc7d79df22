// RANSAC hypothesis scoring in fp32 (the FP32-pipe-bound stage of skimage.measure.ransac's loop: `residuals` +
// `< residual_threshold` per trial, /root/reference/MonocularSlam/slam.py:155-159; algorithm SURVEY.md Appendix A.2/A.3).
//
// This file is compiled with `-Xptxas -O1` (openwork_b200/build.py): at that level ptxas keeps the instruction order of
// the source, and the order below is chosen for the register file (see score_row).
#include <cstdlib>
#include "tvg_common.cuh"
#include "tvg_kernels.h"

namespace tvg {

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
// fma.rn.f32x2 over two hypotheses: score_row_pairs below -- same time, half the issue slots.)
// Register-file bandwidth shapes the instruction order (measured: tools/microbench/ffma_operands.cu -- 0.46 of the
// FFMA peak with three fresh register operands per instruction, 0.93 with two): a B200 sub-partition reads its
// registers through two banks (even / odd index), one distinct register per bank per clock, so an FFMA whose three
// sources are three different registers issues every OTHER clock unless one of them comes from the operand-reuse
// cache (same register in the same source slot as the previous instruction of the warp).  The 12 FFMAs per
// evaluation that multiply by a coordinate of the correspondence are therefore issued coordinate by coordinate for
// all HPT hypotheses of the thread: runs y, x, v, u of 3 * HPT consecutive FFMAs that share slot a; only the first
// of a run pays the second clock.  The band tests form one more run (shared -band).  volatile asm + ptxas -O1 keep
// this order (at -O3 ptxas interleaves the rows and 47-62 % of the FFMAs read three fresh registers:
// tools/sass_rf_model.py).  The arithmetic, and every rounding, is that of the plain nested-fmaf form
//     f_i = fma(e_i0, x, fma(e_i1, y, e_i2)),  g_j = fma(e_0j, u, fma(e_1j, v, e_2j)),  num = fma(u, f0, fma(v, f1, f2)).
#define TVG_FMA(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))

// One correspondence c = (x, y, u, v)  (x = src / last, u = dst / curr) against the HPT hypotheses of the thread:
// 19 FFMA per evaluation (+ 1 MUFU.RCP + 1 FFMA when sum r^2 is wanted) + 2 integer-pipe instructions for BOTH
// counters: the sign bits of t (inlier) and bt (inside the fp32 uncertainty band) are replicated into the two 16-bit
// lanes of one word by a single PRMT (selector 0xFFBB: bytes 0,1 = sign of t, bytes 2,3 = sign of bt), and one IADD
// subtracts that word from the packed accumulator:
//     acc - 0x0000FFFF  ==  acc + 1 - 0x10000   -> low lane += 1, high lane -= 1
//     acc - 0xFFFF0000  ==  acc + 0x10000       -> high lane += 1
// so after <= 32768 rows  low = count, high = (band - count) mod 2^16.
//
// too close for fp32 to call:  |t| < band*den + floor.  The relative part covers rounding of num^2 and thr^2*den; the
// absolute floor covers cancellation in num for points near the epipole (den -> 0):
// 2*thr*sqrt(den)*delta <= (band_rel/2)*thr^2*den + delta^2/(band_rel/2)  with  delta ~ 3e-7 (|E|_F = 1, |x| <~ 1);
// the other half of band_rel is the rounding allowance.  The floor costs no instruction: with c1 = floor/band added
// to den and thr^2*c1 added to num^2 (two FMULs become FFMAs), t is unchanged and band*den' = band*den + floor.
template <int HPT, bool SSE>
__device__ __forceinline__ void score_row(const float (&e)[HPT][9], const float4 c, float nthr2, float nband, float c1,
                                          float k1, unsigned (&acc)[HPT], float (&sse)[HPT]) {
    float t0[HPT], t1[HPT], t2[HPT], f0[HPT], f1[HPT], f2[HPT], s0[HPT], s1[HPT], nn[HPT], g0[HPT], g1[HPT], num[HPT];
    float den[HPT], n2[HPT], t[HPT], bt[HPT];
#pragma unroll
    for (int k = 0; k < HPT; ++k) {
        TVG_FMA(t0[k], c.y, e[k][1], e[k][2]);
        TVG_FMA(t1[k], c.y, e[k][4], e[k][5]);
        TVG_FMA(t2[k], c.y, e[k][7], e[k][8]);
    }
#pragma unroll
    for (int k = 0; k < HPT; ++k) {
        TVG_FMA(f0[k], c.x, e[k][0], t0[k]);
        TVG_FMA(f1[k], c.x, e[k][3], t1[k]);
        TVG_FMA(f2[k], c.x, e[k][6], t2[k]);
    }
#pragma unroll
    for (int k = 0; k < HPT; ++k) {
        TVG_FMA(s0[k], c.w, e[k][3], e[k][6]);
        TVG_FMA(s1[k], c.w, e[k][4], e[k][7]);
        TVG_FMA(nn[k], c.w, f1[k], f2[k]);
    }
#pragma unroll
    for (int k = 0; k < HPT; ++k) {
        TVG_FMA(g0[k], c.z, e[k][0], s0[k]);
        TVG_FMA(g1[k], c.z, e[k][1], s1[k]);
        TVG_FMA(num[k], c.z, f0[k], nn[k]);
    }
    // den = f0^2 + (f1^2 + (g0^2 + (g1^2 + c1))): the HPT chains interleaved stage by stage.  Plain fmaf here: c1, k1
    // and thr^2 are warp-uniform and the compiler feeds them from uniform registers (no register-file read).
#pragma unroll
    for (int k = 0; k < HPT; ++k) den[k] = fmaf(g1[k], g1[k], c1);
#pragma unroll
    for (int k = 0; k < HPT; ++k) n2[k] = fmaf(num[k], num[k], k1);
#pragma unroll
    for (int k = 0; k < HPT; ++k) den[k] = fmaf(g0[k], g0[k], den[k]);
#pragma unroll
    for (int k = 0; k < HPT; ++k) den[k] = fmaf(f1[k], f1[k], den[k]);
#pragma unroll
    for (int k = 0; k < HPT; ++k) den[k] = fmaf(f0[k], f0[k], den[k]);
#pragma unroll
    for (int k = 0; k < HPT; ++k) t[k] = fmaf(nthr2, den[k], n2[k]);          // r < thr  <=>  num^2 - thr^2 den < 0
#pragma unroll
    for (int k = 0; k < HPT; ++k) bt[k] = fmaf(nband, den[k], fabsf(t[k]));   // |t| - band den (one run: -band reused)
#pragma unroll
    for (int k = 0; k < HPT; ++k) {
        unsigned w;
        asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(w) : "r"(__float_as_uint(t[k])), "r"(__float_as_uint(bt[k])));
        acc[k] -= w;
        if (SSE) {
            float rinv;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(den[k] - c1));
            sse[k] = fmaf(n2[k] - k1, rinv, sse[k]);
        }
    }
}

// ---- the same evaluation on PAIRS of hypotheses: fma.rn.f32x2 (FFMA2, new with sm_100) performs the two FMAs of a
// pair (hypothesis 2j in the low half, 2j+1 in the high half of a 64-bit register pair) in ONE instruction.  The FP32
// pipe does not get faster -- FFMA2 occupies it for two clocks (tools/microbench/ffma2_operands.cu: 0.91-0.98 of 128
// FMA/clk/SM with up to two fresh pair operands, 0.66 with three) -- but an evaluation needs half the issue slots, and
// the coordinate is a SCALAR operand broadcast to both halves (`R.F32` in SASS, what ptxas makes of mov.b64 {x, x}).
// Every FMA is the same IEEE fp32 operation as in score_row, so counts and bands are bit-identical.
// Measured (profiles/r2_kernels.md): 1.10 ms per C2 batch, the SAME as the scalar form, with the issue slots down from
// 88 % to 52 % and the instruction count from 1.09 G to 0.64 G -- so the scoring stage is bound by operand delivery
// to the FP32 pipe (pipe 76 % active, warps waiting on math-pipe throttle), not by instruction issue.  Opt-in
// (TVG_SCORE_PACKED=1); the scalar form stays the default.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pair_of(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 both(float v) { f32x2 r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v)); return r; }
__device__ __forceinline__ void halves(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
#define TVG_FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))

template <int NP>
__device__ __forceinline__ void score_row_pairs(const f32x2 (&e)[NP][9], const float4 c, float nthr2, float nband, float c1,
                                                float k1, unsigned (&acc)[2 * NP]) {
    const f32x2 cx = both(c.x), cy = both(c.y), cu = both(c.z), cv = both(c.w);
    const f32x2 pc1 = both(c1), pk1 = both(k1), pnthr2 = both(nthr2);
    f32x2 t0[NP], t1[NP], t2[NP], f0[NP], f1[NP], f2[NP], s0[NP], s1[NP], nn[NP], g0[NP], g1[NP], num[NP];
    f32x2 den[NP], n2[NP], t[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        TVG_FMA2(t0[k], e[k][1], cy, e[k][2]);
        TVG_FMA2(t1[k], e[k][4], cy, e[k][5]);
        TVG_FMA2(t2[k], e[k][7], cy, e[k][8]);
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        TVG_FMA2(f0[k], e[k][0], cx, t0[k]);
        TVG_FMA2(f1[k], e[k][3], cx, t1[k]);
        TVG_FMA2(f2[k], e[k][6], cx, t2[k]);
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        TVG_FMA2(s0[k], e[k][3], cv, e[k][6]);
        TVG_FMA2(s1[k], e[k][4], cv, e[k][7]);
        TVG_FMA2(nn[k], f1[k], cv, f2[k]);
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        TVG_FMA2(g0[k], e[k][0], cu, s0[k]);
        TVG_FMA2(g1[k], e[k][1], cu, s1[k]);
        TVG_FMA2(num[k], f0[k], cu, nn[k]);
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) TVG_FMA2(den[k], g1[k], g1[k], pc1);
#pragma unroll
    for (int k = 0; k < NP; ++k) TVG_FMA2(n2[k], num[k], num[k], pk1);
#pragma unroll
    for (int k = 0; k < NP; ++k) TVG_FMA2(den[k], g0[k], g0[k], den[k]);
#pragma unroll
    for (int k = 0; k < NP; ++k) TVG_FMA2(den[k], f1[k], f1[k], den[k]);
#pragma unroll
    for (int k = 0; k < NP; ++k) TVG_FMA2(den[k], f0[k], f0[k], den[k]);
#pragma unroll
    for (int k = 0; k < NP; ++k) TVG_FMA2(t[k], den[k], pnthr2, n2[k]);
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        float ta, tb, da, db;
        halves(t[k], ta, tb);
        halves(den[k], da, db);
        // |t| - band den: scalar FFMAs (the |.| is an operand modifier there; FFMA2 has none)
        const float bta = fmaf(nband, da, fabsf(ta)), btb = fmaf(nband, db, fabsf(tb));
        unsigned wa, wb;
        asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(wa) : "r"(__float_as_uint(ta)), "r"(__float_as_uint(bta)));
        asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(wb) : "r"(__float_as_uint(tb)), "r"(__float_as_uint(btb)));
        acc[2 * k] -= wa;
        acc[2 * k + 1] -= wb;
    }
}

// Persistent CTAs pull work items from a device-side queue (atomic counter): item = (pair, row slot, block of
// HPT * 128 hypotheses).  A thread owns HPT hypotheses (E in registers); the item's correspondences stream through a
// two-stage shared-memory ring filled by TMA bulk copies; one broadcast LDS.128 feeds HPT evaluations.  Items whose
// row slot lies beyond the pair's match count (known only on the device) cost one queue pop.
template <int HPT, bool SSE, bool PK>
__global__ void __launch_bounds__(SCORE_THREADS) ransac_score_kernel(const RansacParams p, int items, int hblocks) {
    static_assert(!(PK && SSE) && (!PK || HPT % 2 == 0), "the packed form scores pairs of hypotheses, counts only");
    extern __shared__ __align__(128) unsigned char score_smem[];
    // a stage holds SCORE_CHUNK rows + one spare row, so the row prefetch of the inner loop needs no bounds check
    float4(*corr)[SCORE_CHUNK + 1] = reinterpret_cast<float4(*)[SCORE_CHUNK + 1]>(score_smem);
    __shared__ alignas(8) uint64_t bar[2];
    __shared__ int s_item;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    const float nthr2 = p.s_nthr2, nband = p.s_nband, c1 = p.s_c1, k1 = p.s_k1;   // constant-bank FFMA operands
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
        f32x2 e2[PK ? HPT / 2 : 1][9];
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
        if (PK) {
#pragma unroll
            for (int k = 0; k < HPT / 2; ++k)
#pragma unroll
                for (int i = 0; i < 9; ++i) e2[k][i] = pair_of(e[2 * k][i], e[2 * k + 1][i]);
        }
        for (int i = 0; i < n; ++i) {
            const uint32_t gi = g + (uint32_t)i;
            mbar_wait(&bar[gi & 1u], (gi >> 1) & 1u);
            const int rows = min(SCORE_CHUNK, r_end - r_begin - i * SCORE_CHUNK);
            const float4* __restrict__ cp = corr[gi & 1u];
            float4 c = cp[0];
#pragma unroll(8 / HPT)
            for (int j = 0; j < rows; ++j) {
                const float4 cn = cp[j + 1];   // next row in flight while this one is evaluated (spare row at the end)
                if (PK) score_row_pairs<PK ? HPT / 2 : 1>(e2, c, nthr2, nband, c1, k1, reinterpret_cast<unsigned (&)[PK ? HPT : 2]>(acc));
                else score_row<HPT, SSE>(e, c, nthr2, nband, c1, k1, acc, sse);
                c = cn;
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

void ransac_score_constants(RansacParams* p) {
    const float thr2 = (float)(p->thr * p->thr);
    const float band = p->band_rel * thr2;
    p->s_c1 = p->band_floor / band;   // den offset that realises the absolute floor
    p->s_k1 = thr2 * p->s_c1;         // matching num^2 offset: t stays num^2 - thr^2 den
    p->s_nthr2 = -thr2;
    p->s_nband = -band;
}

// Work decomposition of the scoring launch.  hpt hypotheses per thread; `slots` row ranges of `rows_per_slot`
// correspondences each (partial sums merged by the reduce kernel); persistent grid of `ctas_per_sm` CTAs per SM
// (dynamic shared memory is padded so that no more fit), chosen so that the item count fills whole rounds.
static bool score_packed_default() {
    static int v = -1;
    if (v < 0) {
        const char* s = getenv("TVG_SCORE_PACKED");   // 1 selects the FFMA2 form (A/B timing, tools/score_sweep.py)
        v = (s && s[0] == '1') ? 1 : 0;
    }
    return v != 0;
}

static int score_max_ctas(int hpt, bool packed) {
    static int cached[2][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
    int& c = cached[packed ? 1 : 0][hpt];
    if (!c) {
        int n = 0;
        const int smem = 2 * (SCORE_CHUNK + 1) * 16;
        cudaError_t e;
        if (packed)
            e = (hpt == 4) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ransac_score_kernel<4, false, true>, SCORE_THREADS, smem)
                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ransac_score_kernel<2, false, true>, SCORE_THREADS, smem);
        else
            e = (hpt == 4) ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ransac_score_kernel<4, false, false>, SCORE_THREADS, smem)
                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ransac_score_kernel<2, false, false>, SCORE_THREADS, smem);
        if (e != cudaSuccess || n < 1) { cudaGetLastError(); n = hpt == 4 ? 5 : 8; }
        c = n;
    }
    return c;
}

void ransac_score_plan(int batch, int n_max, int H, int sm_count, int cfg_hpt, int cfg_ctas, ScorePlan* out) {
    ScorePlan pl{};
    pl.hpt = (cfg_hpt == 4) ? 4 : 2;
    pl.packed = score_packed_default() ? 1 : 0;
    const int per_block = pl.hpt * SCORE_THREADS;
    pl.hblocks = (H + per_block - 1) / per_block;
    const int chunks = (n_max + SCORE_CHUNK - 1) / SCORE_CHUNK > 0 ? (n_max + SCORE_CHUNK - 1) / SCORE_CHUNK : 1;
    int max_c = score_max_ctas(pl.hpt, pl.packed != 0);
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
    // CTAs per SM: measured on the 1024-pair C2 batch (tools/score_sweep.py): 6 resident CTAs of 128 threads for two
    // hypotheses per thread (5: +0.5 %, 7-9: +3.5 %, the register-limited maximum of 10: +44 %), 4 for four.  The
    // dynamic queue evens out the last round, so filling whole rounds matters less than the issue mix.
    int best_c = pl.hpt == 4 ? 4 : 6;
    if (best_c > max_c) best_c = max_c;
    if (cfg_ctas > 0) best_c = cfg_ctas < score_max_ctas(pl.hpt, pl.packed != 0) ? cfg_ctas : score_max_ctas(pl.hpt, pl.packed != 0);
    pl.ctas_per_sm = best_c;
    const long long g = (long long)sm_count * best_c;
    pl.grid = (int)(pl.items < g ? pl.items : g);
    if (pl.grid < 1) pl.grid = 1;
    // pad the dynamic shared memory so that exactly ctas_per_sm CTAs fit on an SM (228 KB, 1 KB reserved per CTA)
    int smem = (int)((228 * 1024) / best_c) - 1024 - 256;
    smem &= ~127;
    if (smem < 2 * (SCORE_CHUNK + 1) * 16) smem = 2 * (SCORE_CHUNK + 1) * 16;
    if (smem > 200 * 1024) smem = 200 * 1024;
    pl.smem = smem;
    *out = pl;
}

cudaError_t launch_ransac_score(const RansacParams& p, int batch, const ScorePlan& pl, cudaStream_t stream) {
    // the fp32 sum of squared residuals is diagnostic (selection uses the exact fp64 re-score): computed
    // only when the caller asked for it
    (void)batch;
#define TVG_SCORE_LAUNCH(HPT, SSE, PK)                                                                                \
    do {                                                                                                              \
        cudaError_t e__ = cudaFuncSetAttribute(ransac_score_kernel<HPT, SSE, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem); \
        if (e__ != cudaSuccess) return e__;                                                                           \
        ransac_score_kernel<HPT, SSE, PK><<<pl.grid, SCORE_THREADS, pl.smem, stream>>>(p, pl.items, pl.hblocks);     \
    } while (0)
    const bool pk = pl.packed != 0 && !p.want_sse;
    if (pl.hpt == 4) {
        if (p.want_sse) TVG_SCORE_LAUNCH(4, true, false); else if (pk) TVG_SCORE_LAUNCH(4, false, true); else TVG_SCORE_LAUNCH(4, false, false);
    } else {
        if (p.want_sse) TVG_SCORE_LAUNCH(2, true, false); else if (pk) TVG_SCORE_LAUNCH(2, false, true); else TVG_SCORE_LAUNCH(2, false, false);
    }
#undef TVG_SCORE_LAUNCH
    return cudaGetLastError();
}

}  // namespace tvg
