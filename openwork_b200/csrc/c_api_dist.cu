// Multi-GPU part of the C ABI: an NCCL communicator bound to a context, and the single-huge-pair pipeline
// (BASELINE.json config C5) whose query rows and hypotheses are split over the ranks of that communicator.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): inside a PyTorch process that is the copy torch already
// loaded, in a plain C program the system library.  Nothing here is needed by the single-GPU entry points.
#include <dlfcn.h>
#include <nccl.h>

#include "tvg_ctx.h"

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (api.lib) break;
    }
    if (!api.lib) {
        api.error = std::string("cannot load NCCL: ") + dlerror();
        return &api;
    }
#define TVG_SYM(field, name)                                                  \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));  \
    if (!api.field) api.error = std::string("NCCL symbol missing: ") + name
    TVG_SYM(GetUniqueId, "ncclGetUniqueId");
    TVG_SYM(CommInitRank, "ncclCommInitRank");
    TVG_SYM(CommDestroy, "ncclCommDestroy");
    TVG_SYM(AllGather, "ncclAllGather");
    TVG_SYM(GroupStart, "ncclGroupStart");
    TVG_SYM(GroupEnd, "ncclGroupEnd");
    TVG_SYM(GetErrorString, "ncclGetErrorString");
#undef TVG_SYM
    return &api;
}

int nccl_fail(ncclResult_t r, const char* where) {
    NcclApi* a = nccl_api();
    return fail(1000 + (int)r, std::string(where) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error"));
}
#define TVG_NCCL(x)                                        \
    do {                                                   \
        ncclResult_t r__ = (x);                            \
        if (r__ != ncclSuccess) return nccl_fail(r__, #x); \
    } while (0)

}  // namespace

struct tvg_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

void tvg_comm_release(tvg_ctx* ctx) {
    if (!ctx || !ctx->comm) return;
    NcclApi* a = nccl_api();
    if (ctx->comm->comm && a->CommDestroy) a->CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
}

extern "C" {

int tvg_comm_unique_id(void* id_out) {
    if (!id_out) return fail(-1, "tvg_comm_unique_id: id_out is NULL");
    static_assert(sizeof(ncclUniqueId) == TVG_COMM_ID_BYTES, "ncclUniqueId size");
    NcclApi* a = nccl_api();
    if (!a->error.empty()) return fail(-8, a->error);
    ncclUniqueId id;
    TVG_NCCL(a->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

int tvg_comm_init(tvg_ctx* ctx, const void* id_bytes, int rank, int world) {
    if (!ctx || !id_bytes) return fail(-1, "tvg_comm_init: ctx / id is NULL");
    if (world < 1 || rank < 0 || rank >= world) return fail(-2, "tvg_comm_init: bad rank / world");
    if (ctx->comm) return fail(-2, "tvg_comm_init: the context already has a communicator");
    NcclApi* a = nccl_api();
    if (!a->error.empty()) return fail(-8, a->error);
    DeviceGuard g(ctx->device);
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    tvg_comm* c = new tvg_comm();
    c->rank = rank;
    c->world = world;
    ncclResult_t r = a->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return nccl_fail(r, "ncclCommInitRank");
    }
    ctx->comm = c;
    return 0;
}

int tvg_comm_destroy(tvg_ctx* ctx) {
    if (!ctx) return fail(-1, "ctx is NULL");
    DeviceGuard g(ctx->device);
    cudaDeviceSynchronize();
    tvg_comm_release(ctx);
    return 0;
}

int tvg_comm_info(const tvg_ctx* ctx, int* rank, int* world) {
    if (!ctx) return fail(-1, "ctx is NULL");
    if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
    if (world) *world = ctx->comm ? ctx->comm->world : 1;
    return 0;
}

int tvg_huge_pair_rows(int nq, int nt, int world) {
    if (world < 1) world = 1;
    const int per = (nq + world - 1) / world;
    int n = per * world;
    if (nt > n) n = nt;
    return n > 0 ? n : 1;
}

void tvg_shard_range(int n, int rank, int world, int* lo, int* hi) {
    const int base = n / world, extra = n % world;
    const int l = rank * base + (rank < extra ? rank : extra);
    if (lo) *lo = l;
    if (hi) *hi = l + base + (rank < extra ? 1 : 0);
}

int tvg_huge_pair(tvg_ctx* ctx, const tvg_huge_pair_io* io, void* stream) {
    if (!ctx || !io) return fail(-1, "ctx / io is NULL");
    if (io->nq <= 0 || io->nt < 0 || io->H <= 0) return fail(-2, "tvg_huge_pair: bad sizes");
    if (!io->curr_des || !io->last_des || !io->curr_kps || !io->last_kps) return fail(-3, "tvg_huge_pair: NULL inputs");
    if (!io->words && !io->samples) return fail(-3, "tvg_huge_pair: need words or samples");
    const int rank = ctx->comm ? ctx->comm->rank : 0, world = ctx->comm ? ctx->comm->world : 1;
    if (io->H < world) return fail(-2, "tvg_huge_pair: fewer hypotheses than ranks");
    NcclApi* a = nccl_api();
    if (world > 1 && !a->error.empty()) return fail(-8, a->error);
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    ctx->cur_stream = st;
    const int per = (io->nq + world - 1) / world;
    const int n_max = tvg_huge_pair_rows(io->nq, io->nt, world);
    const size_t N = (size_t)n_max;
    const int q_lo = rank * per < io->nq ? rank * per : io->nq;
    const int q_n = (q_lo + per < io->nq ? q_lo + per : io->nq) - q_lo;

    int32_t *idx, *dist, *qidx, *tidx, *m_count, *n_inl;
    uint8_t* keep;
    double *src, *dst, *Rt, *pix_last, *pix_curr;
    float* corr32;
    TVG_OUT(ctx, io->idx, S_IDX, N * 2 * sizeof(int32_t), idx, int32_t*);
    TVG_OUT(ctx, io->dist, S_DIST, N * 2 * sizeof(int32_t), dist, int32_t*);
    TVG_GET(ctx, S_KEEP, N, keep, uint8_t*);
    TVG_OUT(ctx, io->qidx, S_QIDX, N * sizeof(int32_t), qidx, int32_t*);
    TVG_OUT(ctx, io->tidx, S_TIDX, N * sizeof(int32_t), tidx, int32_t*);
    TVG_OUT(ctx, io->m_count, S_MCOUNT, sizeof(int32_t), m_count, int32_t*);
    TVG_GET(ctx, S_SRC, N * 2 * sizeof(double), src, double*);
    TVG_GET(ctx, S_DST, N * 2 * sizeof(double), dst, double*);
    TVG_GET(ctx, S_CORR32, N * 4 * sizeof(float), corr32, float*);
    TVG_OUT(ctx, io->n_inl, S_NINL, sizeof(int32_t), n_inl, int32_t*);
    TVG_OUT(ctx, io->Rt, S_RT, 16 * sizeof(double), Rt, double*);
    TVG_GET(ctx, S_PIXLAST, N * 2 * sizeof(double), pix_last, double*);
    TVG_GET(ctx, S_PIXCURR, N * 2 * sizeof(double), pix_curr, double*);

    // 1. matcher: this rank's block of query rows against the whole train set, written at its place in the full arrays
    // (stage marks as in tvg_pairs_device: "knn2" includes exchange 1, "refit_solve" exchange 2 and the merge)
    int rc = ctx->mark(0, st);
    if (rc) return rc;
    if (q_n > 0) {
        rc = tvg_knn2_hamming256(ctx, io->curr_des + (size_t)q_lo * 32, io->last_des, 1, n_max, q_n, io->nt, nullptr, nullptr,
                                 nullptr, nullptr, io->ratio, idx + (size_t)q_lo * 2, dist + (size_t)q_lo * 2, keep + q_lo, stream);
        if (rc) return rc;
    }
    // 2. exchange step 1: fixed-size all-gather of the matcher outputs (rank order == ascending query order); no
    //    size exchange, no host synchronisation.  One NCCL group = one launch.
    if ((rc = ctx->xmark(0, st))) return rc;
    if (world > 1) {
        TVG_NCCL(a->GroupStart());
        TVG_NCCL(a->AllGather(idx + (size_t)rank * per * 2, idx, (size_t)per * 2, ncclInt32, ctx->comm->comm, st));
        TVG_NCCL(a->AllGather(dist + (size_t)rank * per * 2, dist, (size_t)per * 2, ncclInt32, ctx->comm->comm, st));
        TVG_NCCL(a->AllGather(keep + (size_t)rank * per, keep, (size_t)per, ncclUint8, ctx->comm->comm, st));
        TVG_NCCL(a->GroupEnd());
        ctx->launches += 1;
    }
    // 3. ratio compaction + gathers + K-normalisation over ALL query rows, on every rank (identical results)
    if ((rc = ctx->xmark(1, st))) return rc;
    if ((rc = ctx->mark(1, st))) return rc;
    rc = tvg_ratio_compact(ctx, keep, idx, 1, n_max, io->nq, nullptr, nullptr, nullptr, io->curr_kps, io->last_kps, io->kinv,
                           qidx, tidx, m_count, src, dst, corr32, stream);
    if (rc) return rc;
    // 4. RANSAC over this rank's block of hypotheses (rows of the same global table on every rank)
    int h_lo, h_hi;
    tvg_shard_range(io->H, rank, world, &h_lo, &h_hi);
    tvg::BestRec* recs;
    TVG_GET(ctx, S_HP_RECS, (size_t)world * sizeof(tvg::BestRec), recs, tvg::BestRec*);
    tvg_ransac_io r{};
    r.src = src; r.dst = dst; r.corr32 = corr32; r.m_count = m_count;
    r.batch = 1; r.n_max = n_max; r.m = n_max; r.H = h_hi - h_lo;
    r.samples = io->samples ? io->samples + (size_t)h_lo * 8 : nullptr;
    r.words = io->words ? io->words + (size_t)h_lo * 8 : nullptr;
    r.sample_stride = 0; r.thr = io->thr;
    r.qidx = qidx; r.tidx = tidx; r.kps_q = io->curr_kps; r.kps_t = io->last_kps;
    r.E = io->E; r.mask = io->mask; r.n_inl = n_inl; r.best = io->best; r.E_best = io->E_best; r.best_sse = io->best_sse;
    r.counts = io->counts ? io->counts + h_lo : nullptr;
    r.Rt = Rt; r.status = io->status; r.inl_idx = io->inl_idx; r.pix_last = pix_last; r.pix_curr = pix_curr;
    r.n_cand = io->n_cand;
    RansacRun run;
    if ((rc = tvg_ransac_prepare(ctx, &r, true, h_lo, recs + rank, &run))) return rc;
    if ((rc = tvg_ransac_phase_a(ctx, &run, 1, st))) return rc;
    // 5. exchange step 2: ONE all-gather of the ranks' 96-byte records {count, hypothesis, exact sum r^2, E}; every
    //    rank then applies the reference's selection rule (most inliers, smallest sum r^2, earliest trial) to the
    //    same list -- exact fp64 keys, so G ranks pick what one rank picks.
    if ((rc = ctx->xmark(2, st))) return rc;
    if (world > 1) {
        TVG_NCCL(a->AllGather(recs + rank, recs, sizeof(tvg::BestRec), ncclUint8, ctx->comm->comm, st));
        ctx->launches += 1;
    }
    if ((rc = ctx->xmark(3, st))) return rc;
    TVG_CUDA(tvg::launch_ransac_merge_records(run.p, recs, world, st));
    ctx->launches += 1;
    // 6. mask, refit, extract_Rt on every rank (identical), then pose and triangulation of the inliers
    if ((rc = tvg_ransac_phase_b(ctx, &run, 1, st))) return rc;
    double *poses, *pinv, *points;
    uint8_t* good;
    TVG_OUT(ctx, io->poses, S_POSES, 2 * 16 * sizeof(double), poses, double*);
    TVG_GET(ctx, S_POSESINV, 32 * sizeof(double), pinv, double*);
    TVG_OUT(ctx, io->points4d, S_POINTS, N * 4 * sizeof(double), points, double*);
    TVG_OUT(ctx, io->good, S_GOOD, N, good, uint8_t*);
    {
        tvg::PoseChainParams pc{Rt, nullptr, nullptr, poses, pinv, nullptr, 1};
        TVG_CUDA(tvg::launch_pose_chain(pc, 1, 1, st));
        ctx->launches += 2;
    }
    if ((rc = ctx->mark(8, st))) return rc;
    rc = tvg_triangulate(ctx, pix_last, pix_curr, 1, n_max, n_max, n_inl, pinv, 32, pinv + 16, 32, io->kinv, points, good,
                         stream);
    if (rc) return rc;
    return ctx->mark(9, st);
}

int tvg_get_exchange_ms(tvg_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(-1, "ctx / ms is NULL");
    for (int i = 0; i < 4; ++i)
        if (!ctx->xprof[i]) return fail(-5, "tvg_get_exchange_ms: no profiled tvg_huge_pair call yet");
    DeviceGuard g(ctx->device);
    TVG_CUDA(cudaEventSynchronize(ctx->xprof[3]));
    TVG_CUDA(cudaEventElapsedTime(&ms[0], ctx->xprof[0], ctx->xprof[1]));
    TVG_CUDA(cudaEventElapsedTime(&ms[1], ctx->xprof[2], ctx->xprof[3]));
    return 0;
}

}  // extern "C"
