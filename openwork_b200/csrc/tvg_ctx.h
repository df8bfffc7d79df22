// Private to the C-ABI translation units (c_api.cu, c_api_dist.cu): error helpers, the context object, scratch slots.
#pragma once
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/tvg_b200.h"
#include "tvg_common.cuh"
#include "tvg_kernels.h"

namespace tvgapi {
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* where);
}  // namespace tvgapi
using tvgapi::cuda_fail;
using tvgapi::fail;

#define TVG_CUDA(x)                                   \
    do {                                              \
        cudaError_t e__ = (x);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__, #x); \
    } while (0)

enum Slot {
    S_PARTIAL, S_COUNTERS, S_IDX, S_DIST, S_KEEP, S_QIDX, S_TIDX, S_MCOUNT, S_SRC, S_DST, S_CORR32,
    S_E64, S_E32, S_VALID, S_PARTCNT, S_PARTSSE, S_COUNTS, S_BAND, S_SSE32, S_CAND, S_NCAND, S_CANDSSE,
    S_CANDCNT, S_BEST, S_BESTSSE, S_EBEST, S_MASK, S_NINL, S_INLIDX, S_PIXLAST, S_PIXCURR, S_E, S_RT,
    S_STATUS, S_POSES, S_POSESINV, S_POINTS, S_GOOD, S_PAIRQ, S_PAIRT, S_SEQOFF, S_VOTES, S_RTS, S_CHOICE, S_REFITWS, S_FE_IMG, S_FE_GRAY, S_FE_COV, S_FE_EIG, S_FE_MAX, S_FE_KEYS, S_FE_NCAND,
    S_FE_CORNERS, S_FE_NCORNERS, S_FE_BLURTMP, S_FE_BLUR,
    // host-entry staging
    S_H_DES, S_H_KPS, S_H_FCOUNT, S_H_PAIRQ, S_H_PAIRT, S_H_SEQOFF, S_H_INIT, S_H_WORDS, S_H_SAMPLES,
    S_KNN_TC, S_BA_TRIAL, S_REFITBIG, S_COMPACTBIG, S_SCOREQ, S_CANDPART, S_LB,
    // tvg_huge_pair
    S_HP_RECS, S_HP_POSE, S_TRI_FALLBACK,
    S_NUM
};

// stage marks of the pair pipeline: before knn, compact, fit, score, reduce, rescore, refit+solve,
// pose stage (cheirality + pose chain), triangulate, and after triangulate
#define TVG_PROF_MARKS 10
// tvg_knn2_hamming256 switches to the tensor-core matcher from this many descriptor comparisons per call
#define TVG_TC_MIN_CMP (16ll << 20)

struct tvg_comm;   // NCCL communicator state (c_api_dist.cu)
void tvg_comm_release(tvg_ctx* ctx);   // destroys ctx->comm (no-op when absent)

struct tvg_ctx {
    int device = 0;
    int sm_count = 148;
    int knn_r = 4, knn_variant = 11, knn_ctas_per_sm = 6;
    int score_hpt = 0, score_ctas = 0;   // scoring kernel tuning (0 = automatic), tvg_set_score_config
    size_t knn_tc_l2_budget = 0;   // 0 = one unpack + one match launch per call; > 0 = L2-resident sub-batches
    bool knn_auto = true;      // until tvg_set_matcher_config is called: r = 2 when the train rows get split
    long long launches = 0;
    int host_chunk_pairs = 256;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t done_event = nullptr;
    // optional stage timing of tvg_pairs_device (tvg_set_profile): events at the stage boundaries
    bool profile = false;
    cudaEvent_t prof[TVG_PROF_MARKS] = {};
    cudaEvent_t xprof[4] = {};   // tvg_huge_pair: start / end of exchange 1 and of exchange 2 (same profiling switch)
    int xmark(int i, cudaStream_t st) {
        if (!profile) return 0;
        if (!xprof[i]) {
            cudaError_t e = cudaEventCreate(&xprof[i]);
            if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
        }
        cudaError_t e = cudaEventRecord(xprof[i], st);
        return e == cudaSuccess ? 0 : cuda_fail(e, "cudaEventRecord");
    }
    int mark(int i, cudaStream_t st) {
        if (!profile) return 0;
        if (!prof[i]) {
            cudaError_t e = cudaEventCreate(&prof[i]);
            if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
        }
        cudaError_t e = cudaEventRecord(prof[i], st);
        return e == cudaSuccess ? 0 : cuda_fail(e, "cudaEventRecord");
    }
    std::vector<cudaEvent_t> events;
    unsigned char* pin = nullptr;   // pinned host staging for the small uploads of the host entry
    size_t pin_bytes = 0, pin_used = 0;

    // copy `bytes` of pageable caller memory into pinned staging; returns the pinned address (or nullptr)
    void* stage(const void* src, size_t bytes) {
        const size_t need = (pin_used + bytes + 255) & ~size_t(255);
        if (need > pin_bytes) return nullptr;
        void* dst = pin + pin_used;
        memcpy(dst, src, bytes);
        pin_used = need;
        return dst;
    }
    int reserve_pinned(size_t bytes) {
        pin_used = 0;
        if (bytes <= pin_bytes) return 0;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceSynchronize");
        if (pin) cudaFreeHost(pin);
        pin = nullptr; pin_bytes = 0;
        e = cudaMallocHost((void**)&pin, bytes + bytes / 2 + 4096);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost(staging)");
        pin_bytes = bytes + bytes / 2 + 4096;
        return 0;
    }
    void* slot[S_NUM] = {};
    size_t slot_bytes[S_NUM] = {};
    cudaStream_t cur_stream = nullptr;   // stream of the entry point being served (set by every stream-taking call)
    bool in_flight = false;              // tvg_pairs_host_submit without its tvg_pairs_host_wait yet
    struct tvg_comm* comm = nullptr;     // NCCL communicator of this context (tvg_comm_init), c_api_dist.cu

    // grow-only scratch; zero_on_grow for buffers the kernels expect zeroed (tile counters): the memset is ordered
    // on the caller's stream, like the kernels that read the buffer.  Growing synchronises the device and
    // (re)allocates, which is illegal while `cur_stream` is being captured into a CUDA graph: warm the context up
    // at the same sizes before capturing.
    int get(Slot s, size_t bytes, void** out, bool zero_on_grow = false) {
        if (bytes == 0) bytes = 16;
        if (slot_bytes[s] < bytes) {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(cur_stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
                return fail(-6, "scratch must grow while the stream is being captured: run the call once eagerly at "
                                "the same sizes before capturing it into a CUDA graph");
            if (slot[s]) {
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceSynchronize");
                cudaFree(slot[s]);
                slot[s] = nullptr;
                slot_bytes[s] = 0;
            }
            const size_t want = bytes + bytes / 4 + 256;
            cudaError_t e = cudaMalloc(&slot[s], want);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(scratch)");
            if (zero_on_grow) {
                e = cudaMemsetAsync(slot[s], 0, want, cur_stream);
                if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(scratch)");
            }
            slot_bytes[s] = want;
        }
        *out = slot[s];
        return 0;
    }
};

#define TVG_GET(ctx, s, bytes, ptr, type)                              \
    do {                                                               \
        void* v__ = nullptr;                                           \
        int rc__ = (ctx)->get(s, (bytes), &v__);                       \
        if (rc__) return rc__;                                         \
        ptr = reinterpret_cast<type>(v__);                             \
    } while (0)
// output pointer if the caller gave one, scratch otherwise
#define TVG_OUT(ctx, user, s, bytes, ptr, type)                        \
    do {                                                               \
        if (user) ptr = reinterpret_cast<type>(user);                  \
        else TVG_GET(ctx, s, bytes, ptr, type);                        \
    } while (0)


inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};


// RANSAC in two phases around an optional exchange step (tvg_huge_pair): prepare fills the launch parameters and
// scratch; phase A = fit, score, reduce, re-score (+ selection on the grid-wide path); phase B = mask, refit, solve.
struct RansacRun {
    tvg::RansacParams p;
    tvg::ScorePlan plan;
    bool big;          // grid-wide rescore / select / refit kernels
};
int tvg_ransac_prepare(tvg_ctx* ctx, const tvg_ransac_io* io, bool force_big, int h_off, tvg::BestRec* rec, RansacRun* run);
int tvg_ransac_phase_a(tvg_ctx* ctx, RansacRun* run, int batch, cudaStream_t st);
int tvg_ransac_phase_b(tvg_ctx* ctx, RansacRun* run, int batch, cudaStream_t st);
