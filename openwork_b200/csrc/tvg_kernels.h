// Internal launch interface between the C-ABI (c_api.cu) and the kernels.  Every array has a leading
// pair ("batch") dimension with stride n_max rows (or H hypotheses); per-pair valid counts live on the
// device so the whole pipeline runs without host synchronisation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tvg {

// ---------------------------------------------------------------- matcher (hamming_knn2.cu)
struct KnnParams {
    const uint8_t* q_base;     // query descriptor pool  u8[frames][n_max][32]
    const uint8_t* t_base;     // train descriptor pool  u8[frames][n_max][32]
    const int32_t* pair_q;     // frame index of the query side per pair (NULL -> b)
    const int32_t* pair_t;     // frame index of the train side per pair (NULL -> b)
    const int32_t* q_count;    // valid rows per FRAME (NULL -> nq)
    const int32_t* t_count;    // valid rows per FRAME (NULL -> nt)
    int n_max, nq, nt;
    int splits, rows_per_split;
    uint2* partial;            // [batch][n_max][splits] packed top-2 keys
    unsigned int* tile_counters;  // [batch][qtiles], zero on entry, zero on exit
    int32_t* idx;              // [batch][n_max][2]
    int32_t* dist;             // [batch][n_max][2]
    uint8_t* keep;             // [batch][n_max] ratio-test flag (may be NULL)
    double ratio;
    uint32_t key_mul, w2, w128, w65536;  // 1<<23, 2, 128, 65536 passed as data so the multiplies stay IMADs (FMA pipe)
};
cudaError_t launch_knn2(const KnnParams& p, int batch, int qtiles, int r, int variant, cudaStream_t stream);
int knn2_block_queries(int r);
// tensor-core matcher (hamming_tc.cu): workspace = e4m3 train rows, 256 B per row, rows padded to 256 per pair
size_t knn2_tc_workspace_bytes(int batch, int n_max, size_t l2_budget);
int knn2_tc_splits(int batch, int q_rows, int n_max, int sm_count);
cudaError_t launch_knn2_tc(const KnnParams& p, int batch, int q_rows, uint8_t* workspace, int sm_count, size_t l2_budget,
                           int* launches, cudaStream_t stream);

// CTA-pair tensor-core matcher (hamming_tc2.cu): tcgen05.mma.cta_group::2, train rows expanded inside the kernel (no workspace)
int knn2_tc2_splits(int batch, int q_rows, int n_max, int sm_count);
cudaError_t launch_knn2_tc2(const KnnParams& p, int batch, int q_rows, int sm_count, int* launches, cudaStream_t stream);

// ---------------------------------------------------------------- ratio compaction (compact.cu)
struct CompactParams {
    const uint8_t* keep;       // [batch][n_max]
    const int32_t* idx;        // [batch][n_max][2]
    const int32_t* pair_q;
    const int32_t* pair_t;
    const int32_t* q_count;    // per frame
    int n_max, nq;
    const double* kps_q;       // keypoint pool f64[frames][n_max][2] (NULL -> no gather)
    const double* kps_t;
    double kinv[9];
    int32_t* qidx;             // [batch][n_max]
    int32_t* tidx;             // [batch][n_max]
    int32_t* m_count;          // [batch]
    double* src;               // [batch][n_max][2]  K-normalised LAST (train) points
    double* dst;               // [batch][n_max][2]  K-normalised CURR (query) points
    float4* corr32;            // [batch][n_max]     (src.x, src.y, dst.x, dst.y)
    int32_t* blk_cnt;          // [batch][compact_big_blocks(n_max)] scratch (NULL -> single-CTA form)
};
int compact_big_blocks(int n_max);   // 0 when the single-CTA form is used
cudaError_t launch_compact(const CompactParams& p, int batch, cudaStream_t stream);

// ---------------------------------------------------------------- RANSAC (ransac.cu)
// One rank's best hypothesis of a pair whose hypotheses are shared between ranks (tvg_huge_pair): what the ranks
// exchange.  96 bytes.
struct BestRec {
    int32_t count;             // exact (fp64) inlier count
    int32_t hyp;               // GLOBAL hypothesis index, -1 = this rank has no model
    double sse;                // exact sum of squared residuals over all correspondences
    double E[9];               // the hypothesis' model (pre-refit)
    double pad;
};

struct RansacParams {
    const double* src;         // [batch][n_max][2]
    const double* dst;         // [batch][n_max][2]
    const float4* corr32;      // [batch][n_max]
    const int32_t* m_count;    // [batch] (NULL -> m)
    int n_max, m, H;
    const int32_t* samples;    // int32[.][H][8] explicit indices (NULL -> words)
    const uint32_t* words;     // u32[.][H][8] random words, idx = (w*M)>>32, +1 mod M on clash
    long long sample_stride;   // elements between pairs (0 = one table shared by all pairs)
    int32_t* samples_out;      // [batch][H][8] indices actually used (may be NULL)
    double thr;
    float band_rel;            // fp32 uncertainty band, relative to thr^2*den
    int want_sse;              // compute the fp32 sum r^2 per hypothesis (diagnostic output)
    float band_floor;          // ... plus an absolute floor (cancellation in num when den -> 0)
    // scoring constants derived from the three above on the host (ransac_score_constants): as kernel parameters they
    // are constant-bank operands of the FFMAs and cost no register-file read
    float s_nthr2, s_nband, s_c1, s_k1;
    // outputs / scratch
    double* E64;               // [batch][H][9]
    float* E32;                // [batch][9][H]  Frobenius-normalised, SoA
    int32_t* valid;            // [batch][H]
    int chunks, chunk;         // partial-sum slots per hypothesis, correspondences covered by one slot
    unsigned int* score_queue; // work-item counter of the scoring kernel: zero on entry, zeroed again by the reduce kernel
    int2* part_cnt;            // [batch][chunks][H] (count, band)
    float* part_sse;           // [batch][chunks][H]
    int32_t* counts;           // [batch][H]  (-1 = failed fit)
    int32_t* band;             // [batch][H]
    float* sse32;              // [batch][H]
    int32_t* cand;             // [batch][H]
    int32_t* n_cand;           // [batch]
    double* cand_sse;          // [batch][H]  exact fp64 sum r^2 (indexed by hypothesis)
    int32_t* cand_cnt;         // [batch][H]  exact fp64 count   (indexed by hypothesis)
    // final
    const int32_t* qidx;       // [batch][n_max]
    const int32_t* tidx;
    const int32_t* pair_q;
    const int32_t* pair_t;
    const double* kps_q;       // pixel keypoints pool (for the inlier gather; may be NULL)
    const double* kps_t;
    int32_t* best;             // [batch]
    double* best_sse;          // [batch]
    double* E_best;            // [batch][9]
    uint8_t* mask;             // [batch][n_max]
    int32_t* n_inl;            // [batch]
    int32_t* inl_idx;          // [batch][n_max] match indices of the inliers, ascending
    double* pix_last;          // [batch][n_max][2] pixel coords of inliers (last / train side)
    double* pix_curr;          // [batch][n_max][2]
    double* E;                 // [batch][9]  refit on the inliers
    double* Rt;                // [batch][16]
    int32_t* status;           // [batch] 0 ok, 1 = fewer than 9 correspondences
    double* refit_ws;          // [batch][ransac_refit_ws_doubles()] accumulate -> solve hand-off
    double* refit_big;         // [batch][ransac_refit_big_blocks(n_max)][ransac_refit_big_stride()] (NULL -> single-CTA refit)
    // grid-wide forms (few pairs, many hypotheses / many correspondences)
    int32_t* lb;               // [batch] best lower bound of the count (NULL -> single-CTA reduce per pair)
    double2* cand_part;        // [batch][H][blocks] per-range re-score partials; set together with refit_big
    BestRec* rec;              // [batch] this rank's record instead of p.best / p.E_best (NULL unless distributed)
    int h_off;                 // global index of this rank's hypothesis 0
};
cudaError_t launch_ransac_fit(const RansacParams& p, int batch, cudaStream_t stream);
// scoring launch plan: hypotheses per thread, row slots (p.chunks) x rows per slot (p.chunk), persistent grid
struct ScorePlan {
    int hpt, hblocks, slots, rows_per_slot, items, ctas_per_sm, grid, smem;
    int packed;   // 1: pairs of hypotheses in 64-bit registers, fma.rn.f32x2 (FFMA2); 0: scalar FFMA form
};
// cfg_hpt in {0 (default 2), 2, 4}; cfg_ctas > 0 forces the resident CTAs per SM
void ransac_score_plan(int batch, int n_max, int H, int sm_count, int cfg_hpt, int cfg_ctas, ScorePlan* out);
// queue: one zeroed unsigned int (the work-item counter; the reduce kernel zeroes it again)
cudaError_t launch_ransac_score(const RansacParams& p, int batch, const ScorePlan& plan, cudaStream_t stream);
void ransac_score_constants(RansacParams* p);   // fills s_nthr2 / s_nband / s_c1 / s_k1 from thr, band_rel, band_floor
cudaError_t launch_ransac_reduce(const RansacParams& p, int batch, cudaStream_t stream);
cudaError_t launch_ransac_rescore(const RansacParams& p, int batch, cudaStream_t stream);
int ransac_reduce_launches(const RansacParams& p);
cudaError_t launch_ransac_select_big(const RansacParams& p, int batch, cudaStream_t stream);
cudaError_t launch_ransac_merge_records(const RansacParams& p, const BestRec* recs, int world, cudaStream_t stream);
cudaError_t launch_ransac_refit(const RansacParams& p, int batch, cudaStream_t stream);
int ransac_refit_ws_doubles();
int ransac_refit_big_blocks(int n_max, bool force = false);   // 0 when the single-CTA refit is used
int ransac_refit_big_stride();

// ---------------------------------------------------------------- pose + triangulation (pose_tri.cu)
struct PoseChainParams {
    const double* Rt;          // [batch][16]
    const int32_t* seq_off;    // [nseq+1] pair offsets of each sequence (NULL -> {0, batch})
    const double* init_pose;   // [nseq][16] (NULL -> identity)
    double* poses;             // [batch + nseq][16]  pose of pair b's LAST frame at [b + s], CURR at [b + s + 1]
    double* pair_pinv;         // [batch][2][16] inv(last pose), inv(curr pose) per pair (may be NULL)
    const double* first_init;  // device 4x4 overriding the initial pose of sequence 0 of this launch (may be NULL)
    int batch;                 // used when seq_off == NULL: one sequence holding pairs [0, batch)
};
// batch_hint = total pairs when the host knows it (0: unknown); two launches when pair_pinv is wanted
cudaError_t launch_pose_chain(const PoseChainParams& p, int nseq, int batch_hint, cudaStream_t stream);

struct TriParams {
    const double* pix1;        // [batch][n_max][2] pixel coords in the first view
    const double* pix2;
    const int32_t* n_pts;      // [batch] (NULL -> n)
    int n_max, n;
    const double* P1;          // [batch] x 16 stride p_stride doubles: inverse pose of view 1
    const double* P2;
    long long p1_stride, p2_stride;  // doubles between pairs
    double kinv[9];
    double* points4d;          // [batch][n_max][4]
    uint8_t* good;             // [batch][n_max]  z > 0   (may be NULL)
    double* reproj;            // [batch][n_max] max reprojection error over both views, pixels (may be NULL)
    double* parallax_cos;      // [batch][n_max] cosine of the angle between the two viewing rays (may be NULL)
    double focal;              // pixels per normalised unit for `reproj` (K[0][0])
    int* n_fallback;           // one zeroed int: points left to the Jacobi kernel (zeroed again by the launch)
};
cudaError_t launch_triangulate(const TriParams& p, int batch, cudaStream_t stream);
cudaError_t launch_invert4(const double* in, double* out, int n, cudaStream_t stream);

// ---------------------------------------------------------------- two-view bundle adjustment (bundle.cu)
struct BundleParams {
    const double* pix1;        // [batch][n_max][2] pixel coords in the first / second view
    const double* pix2;
    const int32_t* n_pts;      // [batch] (NULL -> n)
    const uint8_t* use;        // [batch][n_max] points that take part (NULL -> all)
    int n_max, n, iters;
    double lambda0;
    double kinv[9];
    const double* pose1;       // 4x4 camera-to-world pose of the first view per pair (fixed)
    double* pose2;             // second view, refined in place
    long long pose1_stride, pose2_stride;
    double* points4d;          // [batch][n_max][4], refined in place (w stays 1)
    double* trial;             // [batch][n_max][3] scratch
    double* cost;              // [batch][2] initial / final sum of squared reprojection errors (normalised coords)
    int32_t* accepted;         // [batch] accepted steps (may be NULL)
};
cudaError_t launch_bundle_pairs(const BundleParams& p, int batch, cudaStream_t stream);

// ---------------------------------------------------------------- sliding-window bundle adjustment (bundle_window.cu)
struct BundleWindowParams {
    const double* obs;         // [batch][n_max][W][2] pixel observations of point j in camera c
    const uint8_t* vis;        // [batch][n_max][W] visibility mask
    const int32_t* n_pts;      // [batch] (NULL -> n)
    int n_max, n, W, n_fixed, iters;
    double lambda0;
    double kinv[9];
    double* poses;             // [batch][W][16] camera-to-world, cameras >= n_fixed refined in place
    double* points4d;          // [batch][n_max][4], refined in place (w stays 1)
    double* scratch;           // bundle_window_scratch_doubles(batch, n_max)
    double* cost;              // [batch][2] initial / final cost (normalised coordinates)
    int32_t* accepted;         // [batch] accepted steps (may be NULL)
};
size_t bundle_window_scratch_doubles(int batch, int n_max);
int bundle_window_max_cameras();
cudaError_t launch_bundle_window(const BundleWindowParams& p, int batch, cudaStream_t stream);

struct MapParams {
    const double* points4d;    // [batch][n_max][4]
    const uint8_t* good;       // [batch][n_max]
    const int32_t* n_pts;      // [batch]
    const int32_t* seq_off;    // [nseq+1]
    int n_max;
    long long cap;             // rows available per sequence in map_points
    int include_origin;        // prepend (0,0,0,1): what the first frame hands to the map (slam.py:248)
    double* map_points;        // [nseq][cap][4]
    int32_t* map_count;        // [nseq] rows the sequence needs (may exceed cap: the excess is dropped)
    int32_t* pair_off;         // [batch] row of each pair's first point inside its sequence (may be NULL)
};
cudaError_t launch_map_accumulate(const MapParams& p, int nseq, cudaStream_t stream);

// ---------------------------------------------------------------- feature front-end (frontend.cu)
struct FrontendParams {
    const uint8_t* images;     // [batch][H][W][channels] (BGR when channels == 3)
    int batch, H, W, channels;
    int max_corners, cell, slots, cap;
    double quality;
    float min_distance;
    uint8_t* gray;             // [batch][H][W]
    float* eig;                // [batch][H][W]
    int* max_bits;             // [batch]
    unsigned long long* keys;  // [batch][cap]
    int* n_cand;               // [batch]
    float2* corners;           // [batch][max_corners]
    int* n_corners;            // [batch]
    uint8_t* blur;             // [batch][H][W]
    double* kps;               // [batch][max_corners][2]
    uint8_t* des;              // [batch][max_corners][32]
    int* n_kps;                // [batch]
};
cudaError_t frontend_set_tables(const int8_t* pattern_host, const float* gauss7_host, cudaStream_t stream);
cudaError_t launch_frontend(const FrontendParams& p, cudaStream_t stream);

struct ExtractRtParams {
    const double* E;           // [batch][9]
    double* Rt;                // [batch][16]
};
cudaError_t launch_extract_rt(const ExtractRtParams& p, int batch, cudaStream_t stream);

struct CheiralityParams {
    const double* E;           // [batch][9]
    const double* src;         // [batch][n_max][2] normalised
    const double* dst;
    const uint8_t* mask;       // [batch][n_max] (NULL -> all)
    const int32_t* m_count;    // [batch] (NULL -> m)
    int n_max, m;
    int32_t* votes;            // [batch][4]
    double* Rts;               // [batch][4][16]
    int32_t* choice;           // [batch]
    double* Rt_out;            // [batch][16] the chosen candidate (may be NULL)
};
cudaError_t launch_cheirality(const CheiralityParams& p, int batch, cudaStream_t stream);

}  // namespace tvg
