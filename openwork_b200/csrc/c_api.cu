// C ABI (include/tvg_b200.h) over the sm_100a kernels.  No CPU compute path exists here: every entry
// point launches CUDA kernels or fails with an error code.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "tvg_ctx.h"

namespace {
thread_local std::string g_err;
}  // namespace

namespace tvgapi {
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return (int)e > 0 ? (int)e : 1;
}
}  // namespace tvgapi

namespace {

__global__ void make_pairs_kernel(const int32_t* seq_off, int nseq, int batch, int32_t* pair_q, int32_t* pair_t) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    int s = 0;
    if (seq_off)   // NULL = one sequence holding every pair
        while (s + 1 < nseq && b >= seq_off[s + 1]) ++s;
    pair_t[b] = b + s;      // last frame of the pair
    pair_q[b] = b + s + 1;  // current frame
}

}  // namespace

extern "C" {

const char* tvg_last_error(void) { return g_err.c_str(); }
int tvg_abi_version(void) { return 1; }

int tvg_ctx_create(tvg_ctx** out, int device) {
    if (!out) return fail(-1, "tvg_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount (no CUDA device: this library has no CPU path)");
    if (device < 0 || device >= n) return fail(-2, "tvg_ctx_create: bad device ordinal");
    cudaDeviceProp prop;
    TVG_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(-3, "tvg_ctx_create: sm_100a (Blackwell B200) required");
    tvg_ctx* c = new tvg_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    *out = c;
    return 0;
}

void tvg_ctx_destroy(tvg_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    tvg_comm_release(ctx);
    for (int i = 0; i < S_NUM; ++i)
        if (ctx->slot[i]) cudaFree(ctx->slot[i]);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    if (ctx->done_event) cudaEventDestroy(ctx->done_event);
    for (int i = 0; i < TVG_PROF_MARKS; ++i)
        if (ctx->prof[i]) cudaEventDestroy(ctx->prof[i]);
    for (int i = 0; i < 4; ++i)
        if (ctx->xprof[i]) cudaEventDestroy(ctx->xprof[i]);
    delete ctx;
}

long long tvg_launch_count(const tvg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int tvg_set_matcher_config(tvg_ctx* ctx, int r, int variant, int ctas_per_sm) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (variant == -1) {   // back to the automatic choice (integer pipe for small calls, tensor cores otherwise)
        if (ctas_per_sm < 1 || ctas_per_sm > 16) return fail(-2, "tvg_set_matcher_config: ctas_per_sm in [1,16]");
        ctx->knn_auto = true;
        ctx->knn_r = 4;
        ctx->knn_variant = 11;
        ctx->knn_ctas_per_sm = ctas_per_sm;
        return 0;
    }
    if (!(r == 2 || r == 4 || r == 8) || !((variant >= 0 && variant <= 3) || variant == 5 || variant == 7 || variant == 9 || variant == 11 || variant == 16 || variant == 17) || ctas_per_sm < 1 || ctas_per_sm > 16)
        return fail(-2, "tvg_set_matcher_config: r in {2,4,8}, variant in {0,1,2,3,5,7,9,11,16,17}, ctas_per_sm in [1,16]");
    ctx->knn_auto = false;
    ctx->knn_r = r;
    ctx->knn_variant = variant;
    ctx->knn_ctas_per_sm = ctas_per_sm;
    return 0;
}

int tvg_set_score_config(tvg_ctx* ctx, int hyp_per_thread, int ctas_per_sm) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (!(hyp_per_thread == 0 || hyp_per_thread == 2 || hyp_per_thread == 4) || ctas_per_sm < 0 || ctas_per_sm > 16)
        return fail(-2, "tvg_set_score_config: hyp_per_thread in {0, 2, 4}, ctas_per_sm in [0, 16]");
    ctx->score_hpt = hyp_per_thread;
    ctx->score_ctas = ctas_per_sm;
    return 0;
}

int tvg_set_matcher_l2_budget(tvg_ctx* ctx, long long bytes) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (bytes < 0) return fail(-2, "tvg_set_matcher_l2_budget: bytes must be >= 0");
    ctx->knn_tc_l2_budget = (size_t)bytes;
    return 0;
}

int tvg_knn2_hamming256(tvg_ctx* ctx, const uint8_t* q_des, const uint8_t* t_des, int batch, int n_max,
                        int nq, int nt, const int32_t* pair_q, const int32_t* pair_t,
                        const int32_t* q_count, const int32_t* t_count, double ratio, int32_t* idx,
                        int32_t* dist, uint8_t* keep, void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (batch < 0 || n_max < 0 || nq < 0 || nt < 0 || nq > n_max || nt > n_max)
        return fail(-2, "tvg_knn2_hamming256: bad sizes");
    if (n_max >= (1 << TVG_KEY_SHIFT)) return fail(-2, "tvg_knn2_hamming256: n_max must be < 2^23");
    if (batch == 0 || n_max == 0) return 0;
    if (!q_des || !t_des || !idx || !dist) return fail(-3, "tvg_knn2_hamming256: NULL array");
    if (!aligned16(q_des) || !aligned16(t_des)) return fail(-4, "tvg_knn2_hamming256: descriptors must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    tvg::KnnParams p{};
    p.q_base = q_des; p.t_base = t_des; p.pair_q = pair_q; p.pair_t = pair_t;
    p.q_count = q_count; p.t_count = t_count;
    p.n_max = n_max; p.nq = nq; p.nt = nt;
    p.idx = idx; p.dist = dist; p.keep = keep; p.ratio = ratio;
    p.key_mul = 1u << TVG_KEY_SHIFT; p.w2 = 2u; p.w128 = 128u; p.w65536 = 65536u;
    const int q_rows_tc = q_count ? n_max : nq, t_rows_tc = t_count ? n_max : nt;
    // Tensor-core path (variant 16): e4m3 +-1 operands, exact f16 dot products, same results.  Chosen automatically
    // once the call holds enough comparisons to amortise the operand expansion (a single 2k x 2k pair stays on the
    // split integer-pipe kernel, which has the lower latency).
    // Tensor-core paths, same results: variant 17 = CTA-pair form (tcgen05.mma.cta_group::2, train rows expanded inside
    // the kernel, offset folded into a ninth MMA) -- the automatic choice once the call holds enough comparisons
    // to amortise the operand expansion (a single 2k x 2k pair stays on the split integer-pipe kernel, which has the
    // lower latency); variant 16 = the single-CTA form with a separate unpack kernel (round 1).
    const bool big = (double)batch * q_rows_tc * t_rows_tc >= (double)TVG_TC_MIN_CMP;
    const bool use_tc = !ctx->knn_auto && ctx->knn_variant == 16;
    if (ctx->knn_auto ? big : ctx->knn_variant == 17) {
        const int s2 = tvg::knn2_tc2_splits(batch, q_rows_tc, n_max, ctx->sm_count);
        if (s2 > 1) TVG_GET(ctx, S_PARTIAL, (size_t)batch * n_max * s2 * sizeof(uint2), p.partial, uint2*);
        int nl = 0;
        TVG_CUDA(tvg::launch_knn2_tc2(p, batch, q_rows_tc, ctx->sm_count, &nl, (cudaStream_t)stream));
        ctx->launches += nl;
        return 0;
    }
    if (use_tc) {
        uint8_t* ws = nullptr;
        TVG_GET(ctx, S_KNN_TC, tvg::knn2_tc_workspace_bytes(batch, n_max, ctx->knn_tc_l2_budget), ws, uint8_t*);
        const int tc_splits = tvg::knn2_tc_splits(batch, q_rows_tc, n_max, ctx->sm_count);
        if (tc_splits > 1) TVG_GET(ctx, S_PARTIAL, (size_t)batch * n_max * tc_splits * sizeof(uint2), p.partial, uint2*);
        int nl = 0;
        TVG_CUDA(tvg::launch_knn2_tc(p, batch, q_rows_tc, ws, ctx->sm_count, ctx->knn_tc_l2_budget, &nl, (cudaStream_t)stream));
        ctx->launches += nl;
        return 0;
    }
    // queries per thread: 4 for big batches; when the grid has to be filled by splitting the train rows
    // (one or few pairs) 2 queries per thread give twice the CTAs per query tile and measure ~6 % faster
    int knn_r = ctx->knn_r;
    const int q_rows = q_count ? n_max : nq;   // query tiles from nq when it is known on the host
    const long long slots = (long long)ctx->sm_count * ctx->knn_ctas_per_sm;
    if (ctx->knn_auto) {
        const int bq4 = tvg::knn2_block_queries(4);
        const long long base4 = (long long)((q_rows + bq4 - 1) / bq4) * batch;
        knn_r = (slots / (base4 > 0 ? base4 : 1) >= 2) ? 2 : 4;
    }
    const int bq = tvg::knn2_block_queries(knn_r);
    const int qtiles = (q_rows + bq - 1) / bq > 0 ? (q_rows + bq - 1) / bq : 1;
    const long long base = (long long)qtiles * batch;
    int splits = (int)(slots / base);
    const int max_splits = (n_max + 63) / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.splits = splits;
    p.rows_per_split = (n_max + splits - 1) / splits;
    if (splits > 1) {
        TVG_GET(ctx, S_PARTIAL, (size_t)batch * n_max * splits * sizeof(uint2), p.partial, uint2*);
        void* v = nullptr;
        int rc = ctx->get(S_COUNTERS, (size_t)batch * qtiles * sizeof(unsigned int), &v, true);
        if (rc) return rc;
        p.tile_counters = reinterpret_cast<unsigned int*>(v);
    }
    TVG_CUDA(tvg::launch_knn2(p, batch, qtiles, knn_r, ctx->knn_variant, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

int tvg_ratio_compact(tvg_ctx* ctx, const uint8_t* keep, const int32_t* idx, int batch, int n_max, int nq,
                      const int32_t* pair_q, const int32_t* pair_t, const int32_t* q_count,
                      const double* kps_q, const double* kps_t, const double* kinv_host, int32_t* qidx,
                      int32_t* tidx, int32_t* m_count, double* src, double* dst, float* corr32,
                      void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (batch < 0 || n_max < 0 || nq < 0 || nq > n_max) return fail(-2, "tvg_ratio_compact: bad sizes");
    if (batch == 0) return 0;
    if (!keep || !idx || !qidx || !tidx || !m_count) return fail(-3, "tvg_ratio_compact: NULL array");
    const bool gather = kps_q || kps_t;
    if (gather && (!kps_q || !kps_t || !kinv_host || !src || !dst || !corr32))
        return fail(-3, "tvg_ratio_compact: gather needs kps_q, kps_t, kinv, src, dst, corr32");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    tvg::CompactParams p{};
    p.keep = keep; p.idx = idx; p.pair_q = pair_q; p.pair_t = pair_t; p.q_count = q_count;
    p.n_max = n_max; p.nq = nq;
    p.kps_q = gather ? kps_q : nullptr; p.kps_t = gather ? kps_t : nullptr;
    if (gather) memcpy(p.kinv, kinv_host, sizeof(p.kinv));
    p.qidx = qidx; p.tidx = tidx; p.m_count = m_count; p.src = src; p.dst = dst;
    p.corr32 = reinterpret_cast<float4*>(corr32);
    if (tvg::compact_big_blocks(n_max) > 0)
        TVG_GET(ctx, S_COMPACTBIG, (size_t)batch * tvg::compact_big_blocks(n_max) * sizeof(int32_t), p.blk_cnt, int32_t*);
    TVG_CUDA(tvg::launch_compact(p, batch, (cudaStream_t)stream));
    ctx->launches += p.blk_cnt ? 2 : 1;
    return 0;
}

}  // extern "C"

// Fills the launch parameters and scratch of one RANSAC call.  force_big selects the grid-wide re-score / select /
// refit kernels whatever n_max is (tvg_huge_pair: results must not depend on the number of ranks); rec != NULL makes
// the selection publish this rank's record for the exchange instead of p.best / p.E_best; h_off = global index of
// hypothesis 0 of this call.
int tvg_ransac_prepare(tvg_ctx* ctx, const tvg_ransac_io* io, bool force_big, int h_off, tvg::BestRec* rec, RansacRun* run) {
    if (!ctx || !io) return fail(-1, "ctx / io is NULL");
    if (io->batch < 0 || io->n_max <= 0 || io->H <= 0 || io->m < 0 || io->m > io->n_max)
        return fail(-2, "tvg_ransac_essential: bad sizes");
    if (!io->src || !io->dst || !io->corr32) return fail(-3, "tvg_ransac_essential: NULL correspondences");
    if (!io->samples && !io->words) return fail(-3, "tvg_ransac_essential: need samples or words");
    if (!(io->thr > 0.0)) return fail(-2, "tvg_ransac_essential: residual threshold must be > 0");
    if (!aligned16(io->src) || !aligned16(io->dst) || !aligned16(io->corr32))
        return fail(-4, "tvg_ransac_essential: correspondences must be 16-byte aligned");
    const size_t B = (size_t)io->batch, N = (size_t)io->n_max, H = (size_t)io->H;
    tvg::RansacParams& p = run->p;
    p = tvg::RansacParams{};
    p.src = io->src; p.dst = io->dst; p.corr32 = reinterpret_cast<const float4*>(io->corr32);
    p.m_count = io->m_count; p.n_max = io->n_max; p.m = io->m; p.H = io->H;
    p.samples = io->samples; p.words = io->samples ? nullptr : io->words;
    p.sample_stride = io->sample_stride; p.samples_out = io->samples_out;
    p.thr = io->thr;
    p.band_rel = 2e-3f;     // |num^2 - thr^2 den| <= band_rel * thr^2 den  <=>  |r - thr| <~ 1e-3 thr = 5e-6
    p.want_sse = io->sse32 ? 1 : 0;
    p.band_floor = 9e-11f;  // delta^2 / (band_rel / 2), delta = 3e-7: fp32 cancellation error of num
    tvg::ransac_score_constants(&p);
    tvg::ransac_score_plan(io->batch, io->n_max, io->H, ctx->sm_count, ctx->score_hpt, ctx->score_ctas, &run->plan);
    p.chunks = run->plan.slots; p.chunk = run->plan.rows_per_slot;
    {
        void* v = nullptr;
        int rcq = ctx->get(S_SCOREQ, sizeof(unsigned int), &v, true);
        if (rcq) return rcq;
        p.score_queue = reinterpret_cast<unsigned int*>(v);
    }
    TVG_OUT(ctx, io->models, S_E64, B * H * 9 * sizeof(double), p.E64, double*);
    TVG_GET(ctx, S_E32, B * H * 9 * sizeof(float), p.E32, float*);
    TVG_GET(ctx, S_VALID, B * H * sizeof(int32_t), p.valid, int32_t*);
    TVG_GET(ctx, S_PARTCNT, B * p.chunks * H * sizeof(int2), p.part_cnt, int2*);
    TVG_GET(ctx, S_PARTSSE, B * p.chunks * H * sizeof(float), p.part_sse, float*);
    TVG_OUT(ctx, io->counts, S_COUNTS, B * H * sizeof(int32_t), p.counts, int32_t*);
    TVG_OUT(ctx, io->band, S_BAND, B * H * sizeof(int32_t), p.band, int32_t*);
    TVG_OUT(ctx, io->sse32, S_SSE32, B * H * sizeof(float), p.sse32, float*);
    TVG_GET(ctx, S_CAND, B * H * sizeof(int32_t), p.cand, int32_t*);
    TVG_OUT(ctx, io->n_cand, S_NCAND, B * sizeof(int32_t), p.n_cand, int32_t*);
    TVG_GET(ctx, S_CANDSSE, B * H * sizeof(double), p.cand_sse, double*);
    TVG_GET(ctx, S_CANDCNT, B * H * sizeof(int32_t), p.cand_cnt, int32_t*);
    p.qidx = io->qidx; p.tidx = io->tidx; p.pair_q = io->pair_q; p.pair_t = io->pair_t;
    const bool gather = io->kps_q && io->kps_t && io->qidx && io->tidx;
    p.kps_q = gather ? io->kps_q : nullptr; p.kps_t = gather ? io->kps_t : nullptr;
    TVG_OUT(ctx, io->best, S_BEST, B * sizeof(int32_t), p.best, int32_t*);
    TVG_OUT(ctx, io->best_sse, S_BESTSSE, B * sizeof(double), p.best_sse, double*);
    TVG_OUT(ctx, io->E_best, S_EBEST, B * 9 * sizeof(double), p.E_best, double*);
    TVG_OUT(ctx, io->mask, S_MASK, B * N, p.mask, uint8_t*);
    TVG_OUT(ctx, io->n_inl, S_NINL, B * sizeof(int32_t), p.n_inl, int32_t*);
    TVG_OUT(ctx, io->inl_idx, S_INLIDX, B * N * sizeof(int32_t), p.inl_idx, int32_t*);
    TVG_OUT(ctx, io->pix_last, S_PIXLAST, B * N * 2 * sizeof(double), p.pix_last, double*);
    TVG_OUT(ctx, io->pix_curr, S_PIXCURR, B * N * 2 * sizeof(double), p.pix_curr, double*);
    TVG_OUT(ctx, io->E, S_E, B * 9 * sizeof(double), p.E, double*);
    TVG_OUT(ctx, io->Rt, S_RT, B * 16 * sizeof(double), p.Rt, double*);
    TVG_OUT(ctx, io->status, S_STATUS, B * sizeof(int32_t), p.status, int32_t*);
    TVG_GET(ctx, S_REFITWS, B * tvg::ransac_refit_ws_doubles() * sizeof(double), p.refit_ws, double*);
    const int nblk = tvg::ransac_refit_big_blocks(io->n_max, force_big);
    run->big = nblk > 0;
    if (run->big) {
        TVG_GET(ctx, S_REFITBIG, B * nblk * tvg::ransac_refit_big_stride() * sizeof(double), p.refit_big, double*);
        TVG_GET(ctx, S_CANDPART, B * H * nblk * sizeof(double2), p.cand_part, double2*);
    }
    // few pairs with many hypotheses: grid-wide reduce instead of one CTA per pair
    if (force_big || (io->batch < 32 && io->H >= 4096)) TVG_GET(ctx, S_LB, B * sizeof(int32_t), p.lb, int32_t*);
    p.rec = rec;
    p.h_off = h_off;
    return 0;
}

int tvg_ransac_phase_a(tvg_ctx* ctx, RansacRun* run, int batch, cudaStream_t st) {
    const tvg::RansacParams& p = run->p;
    int mrc = ctx->mark(2, st);
    if (mrc) return mrc;
    TVG_CUDA(tvg::launch_ransac_fit(p, batch, st));
    if ((mrc = ctx->mark(3, st))) return mrc;
    TVG_CUDA(tvg::launch_ransac_score(p, batch, run->plan, st));
    if ((mrc = ctx->mark(4, st))) return mrc;
    TVG_CUDA(tvg::launch_ransac_reduce(p, batch, st));
    if ((mrc = ctx->mark(5, st))) return mrc;
    TVG_CUDA(tvg::launch_ransac_rescore(p, batch, st));
    ctx->launches += 3 + tvg::ransac_reduce_launches(p);
    if (run->big) {
        TVG_CUDA(tvg::launch_ransac_select_big(p, batch, st));
        ctx->launches += 1;
    }
    return ctx->mark(6, st);
}

int tvg_ransac_phase_b(tvg_ctx* ctx, RansacRun* run, int batch, cudaStream_t st) {
    TVG_CUDA(tvg::launch_ransac_refit(run->p, batch, st));
    ctx->launches += run->big ? 5 : 2;
    return ctx->mark(7, st);
}

extern "C" {

int tvg_ransac_essential(tvg_ctx* ctx, const tvg_ransac_io* io, void* stream) {
    if (!ctx || !io) return fail(-1, "ctx / io is NULL");
    if (io->batch == 0) return 0;
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    RansacRun run;
    int rc = tvg_ransac_prepare(ctx, io, false, 0, nullptr, &run);
    if (rc) return rc;
    if ((rc = tvg_ransac_phase_a(ctx, &run, io->batch, (cudaStream_t)stream))) return rc;
    return tvg_ransac_phase_b(ctx, &run, io->batch, (cudaStream_t)stream);
}

int tvg_extract_rt(tvg_ctx* ctx, const double* E, int batch, double* Rt, void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (batch < 0) return fail(-2, "tvg_extract_rt: bad batch");
    if (batch == 0) return 0;
    if (!E || !Rt) return fail(-3, "tvg_extract_rt: NULL array");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    tvg::ExtractRtParams p{E, Rt};
    TVG_CUDA(tvg::launch_extract_rt(p, batch, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

int tvg_pose_chain(tvg_ctx* ctx, const double* Rt, const int32_t* seq_off, int nseq,
                   const double* init_pose, double* poses, double* pair_pinv, void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (nseq < 0) return fail(-2, "tvg_pose_chain: bad nseq");
    if (nseq == 0) return 0;
    if (!Rt || !seq_off || !poses) return fail(-3, "tvg_pose_chain: NULL array");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    tvg::PoseChainParams p{Rt, seq_off, init_pose, poses, pair_pinv, nullptr, 0};
    TVG_CUDA(tvg::launch_pose_chain(p, nseq, 0, (cudaStream_t)stream));
    ctx->launches += pair_pinv ? 2 : 1;
    return 0;
}

int tvg_invert_poses(tvg_ctx* ctx, const double* poses, int n, double* poses_inv, void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (n < 0) return fail(-2, "tvg_invert_poses: bad n");
    if (n == 0) return 0;
    if (!poses || !poses_inv) return fail(-3, "tvg_invert_poses: NULL array");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    TVG_CUDA(tvg::launch_invert4(poses, poses_inv, n, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

int tvg_triangulate(tvg_ctx* ctx, const double* pix1, const double* pix2, int batch, int n_max, int n,
                    const int32_t* n_pts, const double* P1, long long p1_stride, const double* P2,
                    long long p2_stride, const double* kinv_host, double* points4d, uint8_t* good,
                    void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (batch < 0 || n_max < 0 || n < 0 || n > n_max) return fail(-2, "tvg_triangulate: bad sizes");
    if (batch == 0 || n_max == 0) return 0;
    if (!pix1 || !pix2 || !P1 || !P2 || !kinv_host || !points4d) return fail(-3, "tvg_triangulate: NULL array");
    if (!aligned16(pix1) || !aligned16(pix2) || (reinterpret_cast<uintptr_t>(points4d) & 31u))
        return fail(-4, "tvg_triangulate: pix arrays must be 16-byte, points4d 32-byte aligned");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    tvg::TriParams p{};
    p.pix1 = pix1; p.pix2 = pix2; p.n_pts = n_pts; p.n_max = n_max; p.n = n;
    p.P1 = P1; p.P2 = P2; p.p1_stride = p1_stride; p.p2_stride = p2_stride;
    memcpy(p.kinv, kinv_host, sizeof(p.kinv));
    p.points4d = points4d; p.good = good;
    {
        void* v = nullptr;
        int rcf = ctx->get(S_TRI_FALLBACK, sizeof(int), &v, true);
        if (rcf) return rcf;
        p.n_fallback = reinterpret_cast<int*>(v);
    }
    TVG_CUDA(tvg::launch_triangulate(p, batch, (cudaStream_t)stream));
    ctx->launches += 3;
    return 0;
}

int tvg_point_quality(tvg_ctx* ctx, const double* pix1, const double* pix2, int batch, int n_max, int n,
                      const int32_t* n_pts, const double* P1, long long p1_stride, const double* P2,
                      long long p2_stride, const double* kinv_host, double focal, double* points4d, uint8_t* good,
                      double* reproj, double* parallax_cos, void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (batch < 0 || n_max < 0 || n < 0 || n > n_max) return fail(-2, "tvg_point_quality: bad sizes");
    if (batch == 0 || n_max == 0) return 0;
    if (!pix1 || !pix2 || !P1 || !P2 || !kinv_host || !points4d) return fail(-3, "tvg_point_quality: NULL array");
    if (!aligned16(pix1) || !aligned16(pix2) || (reinterpret_cast<uintptr_t>(points4d) & 31u))
        return fail(-4, "tvg_point_quality: pix arrays must be 16-byte, points4d 32-byte aligned");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    tvg::TriParams p{};
    p.pix1 = pix1; p.pix2 = pix2; p.n_pts = n_pts; p.n_max = n_max; p.n = n;
    p.P1 = P1; p.P2 = P2; p.p1_stride = p1_stride; p.p2_stride = p2_stride;
    memcpy(p.kinv, kinv_host, sizeof(p.kinv));
    p.points4d = points4d; p.good = good; p.reproj = reproj; p.parallax_cos = parallax_cos; p.focal = focal;
    {
        void* v = nullptr;
        int rcf = ctx->get(S_TRI_FALLBACK, sizeof(int), &v, true);
        if (rcf) return rcf;
        p.n_fallback = reinterpret_cast<int*>(v);
    }
    TVG_CUDA(tvg::launch_triangulate(p, batch, (cudaStream_t)stream));
    ctx->launches += 3;
    return 0;
}

int tvg_bundle_adjust_pairs(tvg_ctx* ctx, const double* pix1, const double* pix2, int batch, int n_max, int n,
                            const int32_t* n_pts, const uint8_t* use, const double* kinv_host, const double* pose1,
                            long long pose1_stride, double* pose2, long long pose2_stride, double* points4d,
                            int iters, double lambda0, double* cost, int32_t* accepted, void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (batch < 0 || n_max < 0 || n < 0 || n > n_max || iters < 0 || !(lambda0 > 0.0))
        return fail(-2, "tvg_bundle_adjust_pairs: bad sizes / iters / lambda0");
    if (batch == 0) return 0;
    if (!pix1 || !pix2 || !kinv_host || !pose1 || !pose2 || !points4d || !cost)
        return fail(-3, "tvg_bundle_adjust_pairs: NULL array");
    if (!aligned16(pix1) || !aligned16(pix2)) return fail(-4, "tvg_bundle_adjust_pairs: pix arrays must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    tvg::BundleParams p{};
    p.pix1 = pix1; p.pix2 = pix2; p.n_pts = n_pts; p.use = use; p.n_max = n_max; p.n = n; p.iters = iters;
    p.lambda0 = lambda0;
    memcpy(p.kinv, kinv_host, sizeof(p.kinv));
    p.pose1 = pose1; p.pose2 = pose2; p.pose1_stride = pose1_stride; p.pose2_stride = pose2_stride;
    p.points4d = points4d; p.cost = cost; p.accepted = accepted;
    TVG_GET(ctx, S_BA_TRIAL, (size_t)batch * (n_max > 0 ? n_max : 1) * 3 * sizeof(double), p.trial, double*);
    TVG_CUDA(tvg::launch_bundle_pairs(p, batch, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

int tvg_bundle_adjust_window(tvg_ctx* ctx, const double* obs, const uint8_t* vis, int batch, int n_max, int n,
                             const int32_t* n_pts, int W, int n_fixed, const double* kinv_host, double* poses,
                             double* points4d, int iters, double lambda0, double* cost, int32_t* accepted, void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (batch < 0 || n_max < 0 || n < 0 || n > n_max || iters < 0 || !(lambda0 > 0.0))
        return fail(-2, "tvg_bundle_adjust_window: bad sizes / iters / lambda0");
    if (W < 2 || W > tvg::bundle_window_max_cameras() || n_fixed < 1 || n_fixed >= W)
        return fail(-2, "tvg_bundle_adjust_window: 2 <= W <= 8 cameras, 1 <= n_fixed < W");
    if (batch == 0) return 0;
    if (!obs || !vis || !kinv_host || !poses || !points4d || !cost) return fail(-3, "tvg_bundle_adjust_window: NULL array");
    if (!aligned16(obs)) return fail(-4, "tvg_bundle_adjust_window: obs must be 16-byte aligned");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    tvg::BundleWindowParams p{};
    p.obs = obs; p.vis = vis; p.n_pts = n_pts; p.n_max = n_max; p.n = n; p.W = W; p.n_fixed = n_fixed; p.iters = iters;
    p.lambda0 = lambda0;
    memcpy(p.kinv, kinv_host, sizeof(p.kinv));
    p.poses = poses; p.points4d = points4d; p.cost = cost; p.accepted = accepted;
    TVG_GET(ctx, S_BA_TRIAL, tvg::bundle_window_scratch_doubles(batch, n_max) * sizeof(double), p.scratch, double*);
    TVG_CUDA(tvg::launch_bundle_window(p, batch, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

int tvg_map_accumulate(tvg_ctx* ctx, const double* points4d, const uint8_t* good, const int32_t* n_pts,
                       const int32_t* seq_off, int nseq, int n_max, long long cap, int include_origin,
                       double* map_points, int32_t* map_count, int32_t* pair_off, void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (nseq < 0 || n_max < 0 || cap < 0) return fail(-2, "tvg_map_accumulate: bad sizes");
    if (nseq == 0) return 0;
    if (!points4d || !good || !n_pts || !seq_off || !map_points || !map_count)
        return fail(-3, "tvg_map_accumulate: NULL array");
    if ((reinterpret_cast<uintptr_t>(points4d) & 31u) || (reinterpret_cast<uintptr_t>(map_points) & 31u))
        return fail(-4, "tvg_map_accumulate: point arrays must be 32-byte aligned");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    tvg::MapParams p{points4d, good, n_pts, seq_off, n_max, cap, include_origin, map_points, map_count, pair_off};
    TVG_CUDA(tvg::launch_map_accumulate(p, nseq, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

int tvg_extract_points(tvg_ctx* ctx, const uint8_t* images, int batch, int height, int width, int channels,
                       int max_corners, double quality_level, double min_distance, const int8_t* pattern_host,
                       double* kps, uint8_t* des, int32_t* n_kps, int32_t* n_candidates, void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (batch < 0 || height < 64 || width < 64 || height > 32767 || width > 32767 || (channels != 1 && channels != 3) ||
        max_corners < 1)
        return fail(-2, "tvg_extract_points: bad sizes (64 <= H, W <= 32767; channels 1 or 3; max_corners >= 1)");
    if (batch == 0) return 0;
    if (!images || !pattern_host || !kps || !des || !n_kps) return fail(-3, "tvg_extract_points: NULL array");
    if (reinterpret_cast<uintptr_t>(images) & 3u) return fail(-4, "tvg_extract_points: images must be 4-byte aligned");
    if (!(quality_level > 0.0) || !(min_distance >= 1.0)) return fail(-2, "tvg_extract_points: quality_level > 0, min_distance >= 1");
    const int cell = (int)__builtin_rint(min_distance);
    // one kept corner per grid cell when even the cell diagonal is shorter than min_distance; else at most 4
    // (min_distance >= cell - 0.5 >= 2/3 of the cell side for cell >= 2: five points that far apart do not fit)
    const int slots = ((double)(cell - 1) * (cell - 1) * 2.0 < min_distance * min_distance) ? 1 : 4;
    const int gw = (width + cell - 1) / cell, gh = (height + cell - 1) / cell;
    if (cell > 15 || (size_t)gw * gh * slots > 200u * 1024u)
        return fail(-2, "tvg_extract_points: min_distance > 15.5 or image too large for the shared-memory corner grid");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = (size_t)batch, HW = (size_t)height * width;
    tvg::FrontendParams p{};
    p.images = images; p.batch = batch; p.H = height; p.W = width; p.channels = channels;
    p.max_corners = max_corners; p.cell = cell; p.slots = slots; p.cap = (int)(HW / 2); p.quality = quality_level;
    p.min_distance = (float)min_distance;
    TVG_GET(ctx, S_FE_GRAY, B * HW, p.gray, uint8_t*);
    TVG_GET(ctx, S_FE_EIG, B * HW * sizeof(float), p.eig, float*);
    TVG_GET(ctx, S_FE_MAX, B * sizeof(int), p.max_bits, int*);
    TVG_GET(ctx, S_FE_KEYS, B * (size_t)p.cap * sizeof(unsigned long long), p.keys, unsigned long long*);
    TVG_OUT(ctx, n_candidates, S_FE_NCAND, B * sizeof(int), p.n_cand, int*);
    TVG_GET(ctx, S_FE_CORNERS, B * (size_t)max_corners * sizeof(float2), p.corners, float2*);
    TVG_GET(ctx, S_FE_NCORNERS, B * sizeof(int), p.n_corners, int*);
    TVG_GET(ctx, S_FE_BLUR, B * HW, p.blur, uint8_t*);
    p.kps = kps; p.des = des; p.n_kps = n_kps;
    // cv::getGaussianKernel(7, 2, CV_32F): exp(-x^2/8) normalised in double, rounded to float
    static const uint32_t gauss_bits[7] = {0x3d8fafb1u, 0x3e06387eu, 0x3e434a39u, 0x3e5d4ae0u, 0x3e434a39u, 0x3e06387eu,
                                           0x3d8fafb1u};
    float gauss[7];
    memcpy(gauss, gauss_bits, sizeof(gauss));
    TVG_CUDA(tvg::frontend_set_tables(pattern_host, gauss, st));
    TVG_CUDA(tvg::launch_frontend(p, st));
    ctx->launches += 9;
    return 0;
}

int tvg_cheirality(tvg_ctx* ctx, const double* E, const double* src, const double* dst, const uint8_t* mask,
                   const int32_t* m_count, int batch, int n_max, int m, int32_t* votes, double* Rts,
                   int32_t* choice, void* stream) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (batch < 0 || n_max < 0 || m < 0 || m > n_max) return fail(-2, "tvg_cheirality: bad sizes");
    if (batch == 0) return 0;
    if (!E || !src || !dst || !votes || !Rts || !choice) return fail(-3, "tvg_cheirality: NULL array");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    tvg::CheiralityParams p{E, src, dst, mask, m_count, n_max, m, votes, Rts, choice, nullptr};
    TVG_CUDA(tvg::launch_cheirality(p, batch, (cudaStream_t)stream));
    ctx->launches += 1;
    return 0;
}

static int pairs_device_impl(tvg_ctx* ctx, const tvg_pairs_io* io, const double* first_init, void* stream) {
    if (!ctx || !io) return fail(-1, "ctx / io is NULL");
    if (io->batch < 0 || io->n_max <= 0 || io->H <= 0) return fail(-2, "tvg_pairs_device: bad sizes");
    if (io->batch == 0) return 0;
    if (!io->des || !io->kps) return fail(-3, "tvg_pairs_device: NULL frame pools");
    if (!io->words && !io->samples) return fail(-3, "tvg_pairs_device: need words or samples");
    if ((io->pair_q == nullptr) != (io->pair_t == nullptr)) return fail(-3, "tvg_pairs_device: give both pair_q and pair_t or neither");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = (size_t)io->batch, N = (size_t)io->n_max;
    const int nseq = io->seq_off ? io->nseq : 1;
    if (nseq < 1) return fail(-2, "tvg_pairs_device: nseq must be >= 1");

    // sequence offsets + default consecutive-frame pairing
    // (seq_off == NULL is one sequence [0, batch): the kernels handle it, no host-side table is uploaded, so the
    // call stays capturable into a CUDA graph)
    const int32_t* seq_off = io->seq_off;
    const int32_t *pair_q = io->pair_q, *pair_t = io->pair_t;
    if (!pair_q) {
        int32_t *pq, *pt;
        TVG_GET(ctx, S_PAIRQ, B * sizeof(int32_t), pq, int32_t*);
        TVG_GET(ctx, S_PAIRT, B * sizeof(int32_t), pt, int32_t*);
        make_pairs_kernel<<<(io->batch + 255) / 256, 256, 0, st>>>(seq_off, nseq, io->batch, pq, pt);
        TVG_CUDA(cudaGetLastError());
        ctx->launches += 1;
        pair_q = pq; pair_t = pt;
    }

    int32_t *idx, *dist, *qidx, *tidx, *m_count;
    uint8_t* keep;
    double *src, *dst;
    float* corr32;
    TVG_OUT(ctx, io->idx, S_IDX, B * N * 2 * sizeof(int32_t), idx, int32_t*);
    TVG_OUT(ctx, io->dist, S_DIST, B * N * 2 * sizeof(int32_t), dist, int32_t*);
    TVG_GET(ctx, S_KEEP, B * N, keep, uint8_t*);
    TVG_OUT(ctx, io->qidx, S_QIDX, B * N * sizeof(int32_t), qidx, int32_t*);
    TVG_OUT(ctx, io->tidx, S_TIDX, B * N * sizeof(int32_t), tidx, int32_t*);
    TVG_OUT(ctx, io->m_count, S_MCOUNT, B * sizeof(int32_t), m_count, int32_t*);
    TVG_GET(ctx, S_SRC, B * N * 2 * sizeof(double), src, double*);
    TVG_GET(ctx, S_DST, B * N * 2 * sizeof(double), dst, double*);
    TVG_GET(ctx, S_CORR32, B * N * 4 * sizeof(float), corr32, float*);

    int rc = ctx->mark(0, st);
    if (rc) return rc;
    rc = tvg_knn2_hamming256(ctx, io->des, io->des, io->batch, io->n_max, io->n_max, io->n_max, pair_q, pair_t,
                             io->frame_count, io->frame_count, io->ratio, idx, dist, keep, stream);
    if (rc) return rc;
    if ((rc = ctx->mark(1, st))) return rc;
    rc = tvg_ratio_compact(ctx, keep, idx, io->batch, io->n_max, io->n_max, pair_q, pair_t, io->frame_count,
                           io->kps, io->kps, io->kinv, qidx, tidx, m_count, src, dst, corr32, stream);
    if (rc) return rc;

    tvg_ransac_io r{};
    r.src = src; r.dst = dst; r.corr32 = corr32; r.m_count = m_count;
    r.batch = io->batch; r.n_max = io->n_max; r.m = io->n_max; r.H = io->H;
    r.samples = io->samples; r.words = io->words; r.sample_stride = io->sample_stride; r.thr = io->thr;
    r.qidx = qidx; r.tidx = tidx; r.pair_q = pair_q; r.pair_t = pair_t; r.kps_q = io->kps; r.kps_t = io->kps;
    r.E = io->E; r.mask = io->mask; r.n_inl = io->n_inl; r.best = io->best; r.E_best = io->E_best;
    r.counts = io->counts; r.Rt = io->Rt; r.status = io->status; r.inl_idx = io->inl_idx;
    int32_t* n_inl;
    double *Rt, *pix_last, *pix_curr;
    TVG_OUT(ctx, io->n_inl, S_NINL, B * sizeof(int32_t), n_inl, int32_t*);
    TVG_OUT(ctx, io->Rt, S_RT, B * 16 * sizeof(double), Rt, double*);
    TVG_GET(ctx, S_PIXLAST, B * N * 2 * sizeof(double), pix_last, double*);
    TVG_GET(ctx, S_PIXCURR, B * N * 2 * sizeof(double), pix_curr, double*);
    r.n_inl = n_inl; r.Rt = Rt; r.pix_last = pix_last; r.pix_curr = pix_curr;
    rc = tvg_ransac_essential(ctx, &r, stream);
    if (rc) return rc;
    if (io->pose_mode == 1) {
        // north-star option: pick (R, t) among the four decompositions of E by a cheirality vote over the
        // inliers instead of the reference's trace / t_z heuristics (slam.py:174-181)
        int32_t *votes, *choice;
        double *Rts, *Eref;
        uint8_t* mask;
        TVG_OUT(ctx, io->pose_votes, S_VOTES, B * 4 * sizeof(int32_t), votes, int32_t*);
        TVG_GET(ctx, S_RTS, B * 64 * sizeof(double), Rts, double*);
        TVG_GET(ctx, S_CHOICE, B * sizeof(int32_t), choice, int32_t*);
        TVG_OUT(ctx, io->E, S_E, B * 9 * sizeof(double), Eref, double*);
        TVG_OUT(ctx, io->mask, S_MASK, B * N, mask, uint8_t*);
        tvg::CheiralityParams cp{Eref, src, dst, mask, m_count, io->n_max, io->n_max, votes, Rts, choice, Rt};
        TVG_CUDA(tvg::launch_cheirality(cp, io->batch, st));
        ctx->launches += 1;
    } else if (io->pose_mode != 0) {
        return fail(-2, "tvg_pairs_device: pose_mode must be 0 (reference heuristics) or 1 (cheirality vote)");
    }

    double *poses, *pinv, *points;
    uint8_t* good;
    TVG_OUT(ctx, io->poses, S_POSES, (B + nseq) * 16 * sizeof(double), poses, double*);
    TVG_GET(ctx, S_POSESINV, B * 32 * sizeof(double), pinv, double*);
    TVG_OUT(ctx, io->points4d, S_POINTS, B * N * 4 * sizeof(double), points, double*);
    TVG_OUT(ctx, io->good, S_GOOD, B * N, good, uint8_t*);
    {
        tvg::PoseChainParams pc{Rt, seq_off, io->init_pose, poses, pinv, first_init, io->batch};
        TVG_CUDA(tvg::launch_pose_chain(pc, nseq, io->batch, st));
        ctx->launches += 2;
    }
    if ((rc = ctx->mark(8, st))) return rc;
    // triangulate(frame.last_kps, frame.curr_kps, frame.last_pose, frame.curr_pose)  (slam.py:263)
    rc = tvg_triangulate(ctx, pix_last, pix_curr, io->batch, io->n_max, io->n_max, n_inl, pinv, 32, pinv + 16, 32,
                         io->kinv, points, good, stream);
    if (rc) return rc;
    return ctx->mark(9, st);
}

int tvg_pairs_device(tvg_ctx* ctx, const tvg_pairs_io* io, void* stream) {
    return pairs_device_impl(ctx, io, nullptr, stream);
}

int tvg_set_profile(tvg_ctx* ctx, int on) {
    if (!ctx) return fail(-1, "ctx is NULL");
    ctx->profile = on != 0;
    return 0;
}

int tvg_get_stage_ms(tvg_ctx* ctx, float* ms, int n) {
    if (!ctx || !ms) return fail(-1, "ctx / ms is NULL");
    if (n < TVG_PROF_MARKS - 1) return fail(-2, "tvg_get_stage_ms: need room for 9 stages");
    for (int i = 0; i < TVG_PROF_MARKS; ++i)
        if (!ctx->prof[i]) return fail(-5, "tvg_get_stage_ms: no profiled tvg_pairs_device call yet");
    DeviceGuard g(ctx->device);
    TVG_CUDA(cudaEventSynchronize(ctx->prof[TVG_PROF_MARKS - 1]));
    for (int i = 0; i + 1 < TVG_PROF_MARKS; ++i) TVG_CUDA(cudaEventElapsedTime(&ms[i], ctx->prof[i], ctx->prof[i + 1]));
    return 0;
}

int tvg_set_host_chunk(tvg_ctx* ctx, int pairs) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (pairs < 1) return fail(-2, "tvg_set_host_chunk: pairs must be >= 1");
    ctx->host_chunk_pairs = pairs;
    return 0;
}

int tvg_pairs_host_submit(tvg_ctx* ctx, const tvg_pairs_io* hio, int frames, long long* bytes_h2d,
                          long long* bytes_d2h, void* stream) {
    if (!ctx || !hio) return fail(-1, "ctx / io is NULL");
    if (hio->batch < 0 || hio->n_max <= 0 || hio->H <= 0 || frames <= 0) return fail(-2, "tvg_pairs_host: bad sizes");
    if (!hio->des || !hio->kps) return fail(-3, "tvg_pairs_host: NULL frame pools");
    if (!hio->words && !hio->samples) return fail(-3, "tvg_pairs_host: need words or samples");
    if (ctx->in_flight)
        return fail(-7, "tvg_pairs_host_submit: a batch is already in flight on this context (call tvg_pairs_host_wait "
                        "first, or alternate two contexts)");
    if (hio->sample_stride != 0 && hio->sample_stride != (long long)hio->H * 8)
        return fail(-2, "tvg_pairs_host: sample_stride must be 0 (shared table) or H*8 (one table per pair)");
    DeviceGuard g(ctx->device);
    ctx->cur_stream = (cudaStream_t)stream;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = (size_t)hio->batch, N = (size_t)hio->n_max, H = (size_t)hio->H, F = (size_t)frames;
    const int nseq = hio->seq_off ? hio->nseq : 1;
    long long h2d = 0, d2h = 0;
    tvg_pairs_io d = *hio;
    // Small arrays go through pinned staging so that cudaMemcpyAsync never has to synchronise with the
    // caller's stream (it may for pageable memory), which would serialise consecutive batches.
    const size_t tab = (hio->sample_stride ? B : 1) * H * 8 * 4;
    const size_t small_limit = 1u << 20;
    {
        size_t want = (F + 2 * B + 4 * ((size_t)nseq + 1 + B / 16 + 64)) * sizeof(int32_t) + (size_t)nseq * 16 * 8 + 8192;
        if (tab <= small_limit) want += 2 * tab;
        int rc0 = ctx->reserve_pinned(want);
        if (rc0) return rc0;
    }
    // every host->device copy of this call goes through s_in (in order: small arrays, then the frame chunks), so a
    // copy never queues behind the previous batch's kernels on the caller's stream
    if (!ctx->s_in) TVG_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    if (!ctx->s_out) TVG_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
#define TVG_UP(slot, hptr, bytes, field, type)                                              \
    do {                                                                                    \
        if (hptr) {                                                                         \
            type dp__;                                                                      \
            TVG_GET(ctx, slot, (bytes), dp__, type);                                        \
            const void* src__ = (hptr);                                                     \
            if ((size_t)(bytes) <= small_limit) {                                           \
                void* st__ = ctx->stage(src__, (bytes));                                    \
                if (st__) src__ = st__;                                                     \
            }                                                                               \
            TVG_CUDA(cudaMemcpyAsync((void*)dp__, src__, (bytes), cudaMemcpyHostToDevice, ctx->s_in)); \
            h2d += (long long)(bytes);                                                      \
            field = dp__;                                                                   \
        }                                                                                   \
    } while (0)
    // big frame pools: device buffers only; the copies are scheduled per chunk below
    uint8_t* des_d;
    double* kps_d;
    TVG_GET(ctx, S_H_DES, F * N * 32, des_d, uint8_t*);
    TVG_GET(ctx, S_H_KPS, F * N * 2 * sizeof(double), kps_d, double*);
    d.des = des_d;
    d.kps = kps_d;
    TVG_UP(S_H_FCOUNT, hio->frame_count, F * sizeof(int32_t), d.frame_count, int32_t*);
    TVG_UP(S_H_PAIRQ, hio->pair_q, B * sizeof(int32_t), d.pair_q, int32_t*);
    TVG_UP(S_H_PAIRT, hio->pair_t, B * sizeof(int32_t), d.pair_t, int32_t*);
    TVG_UP(S_H_INIT, hio->init_pose, (size_t)nseq * 16 * sizeof(double), d.init_pose, double*);
    TVG_UP(S_H_WORDS, hio->words, tab, d.words, uint32_t*);
    TVG_UP(S_H_SAMPLES, hio->samples, tab, d.samples, int32_t*);
#undef TVG_UP
    // device-side outputs always go to scratch; requested ones are copied back
    d.idx = d.dist = d.qidx = d.tidx = d.m_count = d.best = d.counts = d.n_inl = d.inl_idx = d.status = nullptr;
    d.E = d.E_best = d.Rt = d.poses = d.points4d = nullptr;
    d.mask = d.good = nullptr;
    // route scratch-backed outputs through fixed slots so we know where they live
    int32_t *idx, *dist, *qidx, *tidx, *m_count, *best, *counts, *n_inl, *inl_idx, *status;
    double *E, *E_best, *Rt, *poses, *points;
    uint8_t *mask, *good;
    TVG_GET(ctx, S_IDX, B * N * 2 * sizeof(int32_t), idx, int32_t*);
    TVG_GET(ctx, S_DIST, B * N * 2 * sizeof(int32_t), dist, int32_t*);
    TVG_GET(ctx, S_QIDX, B * N * sizeof(int32_t), qidx, int32_t*);
    TVG_GET(ctx, S_TIDX, B * N * sizeof(int32_t), tidx, int32_t*);
    TVG_GET(ctx, S_MCOUNT, B * sizeof(int32_t), m_count, int32_t*);
    TVG_GET(ctx, S_BEST, B * sizeof(int32_t), best, int32_t*);
    TVG_GET(ctx, S_COUNTS, B * H * sizeof(int32_t), counts, int32_t*);
    TVG_GET(ctx, S_NINL, B * sizeof(int32_t), n_inl, int32_t*);
    TVG_GET(ctx, S_INLIDX, B * N * sizeof(int32_t), inl_idx, int32_t*);
    TVG_GET(ctx, S_STATUS, B * sizeof(int32_t), status, int32_t*);
    TVG_GET(ctx, S_E, B * 9 * sizeof(double), E, double*);
    TVG_GET(ctx, S_EBEST, B * 9 * sizeof(double), E_best, double*);
    TVG_GET(ctx, S_RT, B * 16 * sizeof(double), Rt, double*);
    TVG_GET(ctx, S_POSES, (B + nseq) * 16 * sizeof(double), poses, double*);
    TVG_GET(ctx, S_POINTS, B * N * 4 * sizeof(double), points, double*);
    TVG_GET(ctx, S_MASK, B * N, mask, uint8_t*);
    TVG_GET(ctx, S_GOOD, B * N, good, uint8_t*);
    // ---- chunked schedule: frame copies (stream s_in), kernels (caller's stream), result copies (s_out)
    // chunk boundaries: ramp up (C/4, C/2, C, C, ...) so the first kernels start after a short first copy
    const int C = ctx->host_chunk_pairs;
    std::vector<int> cb(1, 0);
    for (int sz = C >= 128 ? C / 4 : C; cb.back() < hio->batch; sz = (sz * 2 < C) ? sz * 2 : C)
        cb.push_back(cb.back() + sz < hio->batch ? cb.back() + sz : hio->batch);
    const int nchunks = (int)cb.size() - 1;
    while ((int)ctx->events.size() < 2 * nchunks) {
        cudaEvent_t e;
        TVG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->events.push_back(e);
    }
    // host view of the sequence offsets (one sequence holding everything when absent)
    std::vector<int32_t> so_h(nseq + 1);
    if (hio->seq_off) memcpy(so_h.data(), hio->seq_off, (size_t)(nseq + 1) * sizeof(int32_t));
    else { so_h[0] = 0; so_h[1] = hio->batch; }
    if (so_h[0] != 0 || so_h[nseq] != hio->batch) return fail(-2, "tvg_pairs_host: seq_off must run from 0 to batch");
    // chunk-local sequence offsets, all chunks concatenated, one upload
    std::vector<int32_t> so_all;
    std::vector<int> so_pos(nchunks), s_first(nchunks), n_local(nchunks);
    {
        int s = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int b0 = cb[c], b1 = cb[c + 1];
            while (so_h[s + 1] <= b0) ++s;
            int s1 = s;
            while (so_h[s1 + 1] < b1) ++s1;
            s_first[c] = s; n_local[c] = s1 - s + 1; so_pos[c] = (int)so_all.size();
            for (int k = s; k <= s1 + 1; ++k) {
                int v = so_h[k] < b0 ? b0 : (so_h[k] > b1 ? b1 : so_h[k]);
                so_all.push_back(v - b0);
            }
        }
    }
    int32_t* so_d;
    TVG_GET(ctx, S_H_SEQOFF, so_all.size() * sizeof(int32_t), so_d, int32_t*);
    {
        const void* src = ctx->stage(so_all.data(), so_all.size() * sizeof(int32_t));
        TVG_CUDA(cudaMemcpyAsync(so_d, src ? src : so_all.data(), so_all.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->s_in));
    }
    h2d += (long long)((size_t)(nseq + 1) * sizeof(int32_t));
    // The frame copies read HOST buffers and write this context's private staging pool, which nothing else
    // touches while a batch is in flight (one batch per context), so they need not wait for earlier work on
    // the caller's stream: with two contexts alternating they overlap the previous batch's kernels.
    long long copied_upto = 0;   // frames [0, copied_upto) are on the device (or queued on s_in)
    int rc = 0;
    for (int c = 0; c < nchunks; ++c) {
        const int b0 = cb[c], b1 = cb[c + 1], nb = b1 - b0;
        // frames this chunk needs
        long long need = 0;
        if (hio->pair_q) {
            for (int b = b0; b < b1; ++b) {
                if (hio->pair_q[b] < 0 || hio->pair_q[b] >= frames || hio->pair_t[b] < 0 || hio->pair_t[b] >= frames)
                    return fail(-2, "tvg_pairs_host: pair frame index out of range");
                if (hio->pair_q[b] + 1 > need) need = hio->pair_q[b] + 1;
                if (hio->pair_t[b] + 1 > need) need = hio->pair_t[b] + 1;
            }
        } else {
            need = (long long)(b1 - 1) + (s_first[c] + n_local[c] - 1) + 2;   // curr frame of the last pair, +1
            if (need > frames) return fail(-2, "tvg_pairs_host: not enough frames for the consecutive pairing");
        }
        if (need > copied_upto) {
            const size_t f0 = (size_t)copied_upto, nf = (size_t)(need - copied_upto);
            TVG_CUDA(cudaMemcpyAsync(des_d + f0 * N * 32, hio->des + f0 * N * 32, nf * N * 32, cudaMemcpyHostToDevice, ctx->s_in));
            TVG_CUDA(cudaMemcpyAsync(kps_d + f0 * N * 2, hio->kps + f0 * N * 2, nf * N * 2 * sizeof(double),
                                     cudaMemcpyHostToDevice, ctx->s_in));
            h2d += (long long)(nf * N * (32 + 2 * sizeof(double)));
            copied_upto = need;
        }
        TVG_CUDA(cudaEventRecord(ctx->events[2 * c], ctx->s_in));
        TVG_CUDA(cudaStreamWaitEvent(st, ctx->events[2 * c], 0));

        tvg_pairs_io k = d;
        k.batch = nb;
        k.seq_off = so_d + so_pos[c];
        k.nseq = n_local[c];
        const size_t row0 = (size_t)b0 + s_first[c];          // pose row / frame shift of this chunk
        if (hio->pair_q) { k.pair_q = d.pair_q + b0; k.pair_t = d.pair_t + b0; }
        else {
            k.des = des_d + row0 * N * 32;
            k.kps = kps_d + row0 * N * 2;
            if (d.frame_count) k.frame_count = d.frame_count + row0;
        }
        k.init_pose = d.init_pose ? d.init_pose + (size_t)s_first[c] * 16 : nullptr;
        if (hio->sample_stride) {
            if (d.words) k.words = d.words + (size_t)b0 * hio->sample_stride;
            if (d.samples) k.samples = d.samples + (size_t)b0 * hio->sample_stride;
        }
        k.idx = idx + (size_t)b0 * N * 2; k.dist = dist + (size_t)b0 * N * 2;
        k.qidx = qidx + (size_t)b0 * N; k.tidx = tidx + (size_t)b0 * N; k.m_count = m_count + b0;
        k.best = best + b0; k.counts = counts + (size_t)b0 * H; k.n_inl = n_inl + b0;
        k.inl_idx = inl_idx + (size_t)b0 * N; k.status = status + b0; k.E = E + (size_t)b0 * 9;
        k.E_best = E_best + (size_t)b0 * 9; k.Rt = Rt + (size_t)b0 * 16; k.poses = poses + row0 * 16;
        k.points4d = points + (size_t)b0 * N * 4; k.mask = mask + (size_t)b0 * N; k.good = good + (size_t)b0 * N;
        // a sequence continuing from the previous chunk starts from that chunk's last pose (already at `row0`)
        const double* first_init = (so_h[s_first[c]] < b0) ? poses + row0 * 16 : nullptr;
        rc = pairs_device_impl(ctx, &k, first_init, stream);
        if (rc) return rc;
        TVG_CUDA(cudaEventRecord(ctx->events[2 * c + 1], st));
        TVG_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->events[2 * c + 1], 0));
        cudaStream_t so = ctx->s_out;
#define TVG_DOWN(hptr, dptr, off, bytes)                                                               \
    do {                                                                                               \
        if (hptr) {                                                                                    \
            TVG_CUDA(cudaMemcpyAsync((hptr) + (off), (dptr) + (off), (bytes), cudaMemcpyDeviceToHost, so)); \
            d2h += (long long)(bytes);                                                                 \
        }                                                                                              \
    } while (0)
        const size_t nbz = (size_t)nb;
        TVG_DOWN(hio->idx, idx, (size_t)b0 * N * 2, nbz * N * 2 * sizeof(int32_t));
        TVG_DOWN(hio->dist, dist, (size_t)b0 * N * 2, nbz * N * 2 * sizeof(int32_t));
        TVG_DOWN(hio->qidx, qidx, (size_t)b0 * N, nbz * N * sizeof(int32_t));
        TVG_DOWN(hio->tidx, tidx, (size_t)b0 * N, nbz * N * sizeof(int32_t));
        TVG_DOWN(hio->m_count, m_count, (size_t)b0, nbz * sizeof(int32_t));
        TVG_DOWN(hio->best, best, (size_t)b0, nbz * sizeof(int32_t));
        TVG_DOWN(hio->counts, counts, (size_t)b0 * H, nbz * H * sizeof(int32_t));
        TVG_DOWN(hio->n_inl, n_inl, (size_t)b0, nbz * sizeof(int32_t));
        TVG_DOWN(hio->inl_idx, inl_idx, (size_t)b0 * N, nbz * N * sizeof(int32_t));
        TVG_DOWN(hio->status, status, (size_t)b0, nbz * sizeof(int32_t));
        TVG_DOWN(hio->E, E, (size_t)b0 * 9, nbz * 9 * sizeof(double));
        TVG_DOWN(hio->E_best, E_best, (size_t)b0 * 9, nbz * 9 * sizeof(double));
        TVG_DOWN(hio->Rt, Rt, (size_t)b0 * 16, nbz * 16 * sizeof(double));
        TVG_DOWN(hio->points4d, points, (size_t)b0 * N * 4, nbz * N * 4 * sizeof(double));
        TVG_DOWN(hio->mask, mask, (size_t)b0 * N, nbz * N);
        TVG_DOWN(hio->good, good, (size_t)b0 * N, nbz * N);
        // pose rows of this chunk: [row0, row0 + nb + n_local)
        const size_t prow = (c == 0 || !first_init) ? row0 : row0 + 1, pend = row0 + nbz + (size_t)n_local[c];
        TVG_DOWN(hio->poses, poses, prow * 16, (pend - prow) * 16 * sizeof(double));
#undef TVG_DOWN
    }
    // completion marker of THIS batch on the caller's stream (later work on that stream is not waited for)
    if (!ctx->done_event) TVG_CUDA(cudaEventCreateWithFlags(&ctx->done_event, cudaEventDisableTiming));
    TVG_CUDA(cudaEventRecord(ctx->done_event, st));
    if (bytes_h2d) *bytes_h2d = h2d;
    if (bytes_d2h) *bytes_d2h = d2h;
    ctx->in_flight = true;
    return 0;
}

int tvg_pairs_host_wait(tvg_ctx* ctx) {
    if (!ctx) return fail(-1, "ctx is NULL");
    DeviceGuard g(ctx->device);
    ctx->in_flight = false;
    if (ctx->s_out) TVG_CUDA(cudaStreamSynchronize(ctx->s_out));
    if (ctx->done_event) TVG_CUDA(cudaEventSynchronize(ctx->done_event));
    return 0;
}

int tvg_pairs_host(tvg_ctx* ctx, const tvg_pairs_io* hio, int frames, long long* bytes_h2d,
                   long long* bytes_d2h, void* stream) {
    const int rc = tvg_pairs_host_submit(ctx, hio, frames, bytes_h2d, bytes_d2h, stream);
    return rc ? rc : tvg_pairs_host_wait(ctx);
}

}  // extern "C"
