/* tvg_b200.h -- C ABI of the B200-native two-view geometry hot path.
 *
 * Drop-in boundary for the path  match_points -> fit_essential_matrix -> extract_Rt -> triangulate ->
 * check_points  of OpenWork's MonocularSlam/slam.py (/root/reference/MonocularSlam/slam.py:125-237).
 * The reference has no FFI of its own: its boundary is the Python function surface plus three
 * third-party calls (cv2.BFMatcher.knnMatch, skimage.measure.ransac, numpy.linalg.svd).  Each entry
 * point below names the reference call it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes; every array pointer is a DEVICE pointer unless the name says host;
 *   - arrays carry a leading pair ("batch") dimension with stride n_max rows; per-pair valid counts
 *     are int32 device arrays so a whole batch runs with no host synchronisation;
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *   - return 0 on success, <0 for a validation error, >0 for a cudaError_t; tvg_last_error() returns
 *     a thread-local message; there is NO CPU fallback anywhere behind this ABI;
 *   - descriptor / keypoint pointers must be 16-byte aligned (TMA bulk copies, 128-bit loads).
 *   - layouts are the reference's: des u8[N][32] C-contiguous, kps f64[N][2] pixel (x,y),
 *     pose / Rt f64[4][4] row-major, E f64[3][3] row-major, points4d f64[N][4] = (x,y,z,1).
 */
#ifndef TVG_B200_H
#define TVG_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TVG_API __attribute__((visibility("default")))
#else
#define TVG_API
#endif

typedef struct tvg_ctx tvg_ctx;

/* Context = device + grow-only scratch arena (closest precedent in the reference repo, off-path:
 * CreateSession / ReleaseSession, sort/src/sort.h:34-43). One context per host thread / stream. */
TVG_API int tvg_ctx_create(tvg_ctx** out, int device);
TVG_API void tvg_ctx_destroy(tvg_ctx* ctx);
TVG_API const char* tvg_last_error(void);
TVG_API int tvg_abi_version(void);
/* Number of CUDA kernels this library has launched through `ctx` since creation (bench.py: gpu_launches). */
TVG_API long long tvg_launch_count(const tvg_ctx* ctx);
/* Matcher tuning.  variant -1 (the default state) = automatic: calls with >= 16 Mi descriptor comparisons run on
 * the tensor cores (variant 17: e4m3 +-1 / -+0.5 operands, tcgen05.mma.cta_group::2 over CTA pairs, train rows expanded
 * inside the kernel, exact f16 accumulators holding 1024 + hamming, top-2 in the TMEM epilogue), smaller calls on the
 * integer pipe.  Variant 16 = the single-CTA tensor-core form with a separate operand-expansion kernel (round 1).
 * Integer-pipe variants: queries per thread r in {2,4,8}; popcount
 * variant 0 = 8xPOPC, 1 = Harley-Seal 4xPOPC, 2 = 5xPOPC, 3 = Harley-Seal with query/train-shared XOR combos
 * (14 LOP3 + 4 POPC); +4 (variants 5, 7) = lazy top-2 insertion, +8 (variants 9, 11) = two queries' chunk-local
 * 16-bit keys packed per register (VIMNMX.U16x2 top-2 updates); ctas_per_sm sizes the split-Nt grid.
 * Results are bit-identical for every setting. */
TVG_API int tvg_set_matcher_config(tvg_ctx* ctx, int r, int variant, int ctas_per_sm);
/* Tensor-core matcher variant 16 only.  bytes = 0 (default): the e4m3 expansion of all train frames of a call is written once and
 * read back (fastest; 7x the algorithmic DRAM traffic, 20 % of HBM peak).  bytes > 0: the call is processed in
 * sub-batches whose expansion fits in `bytes` of L2 (40 MiB measured: DRAM traffic = the algorithmic bytes, +8 % time). */
TVG_API int tvg_set_matcher_l2_budget(tvg_ctx* ctx, long long bytes);

/* RANSAC scoring kernel tuning (results do not depend on it): hypotheses per thread (2 or 4; 0 = default) and resident
 * CTAs per SM of the persistent grid (0 = automatic: the count whose rounds the work items fill best). */
TVG_API int tvg_set_score_config(tvg_ctx* ctx, int hyp_per_thread, int ctas_per_sm);

/* cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)  (slam.py:126-127) + the Lowe ratio flag
 * m.distance < ratio * n.distance (slam.py:131).
 * Pair b matches query frame (pair_q ? pair_q[b] : b) of the pool q_des[frames][n_max][32] against train
 * frame (pair_t ? pair_t[b] : b) of t_des.  q_count / t_count: valid rows per FRAME (NULL -> nq / nt).
 * Outputs idx, dist int32[batch][n_max][2] (-1 where fewer than k train rows exist), keep u8[batch][n_max]
 * (may be NULL).  Ties break to the lowest train index: bit-exact with OpenCV. */
TVG_API int tvg_knn2_hamming256(tvg_ctx* ctx, const uint8_t* q_des, const uint8_t* t_des, int batch, int n_max,
                        int nq, int nt, const int32_t* pair_q, const int32_t* pair_t,
                        const int32_t* q_count, const int32_t* t_count, double ratio, int32_t* idx,
                        int32_t* dist, uint8_t* keep, void* stream);

/* The ratio loop + gathers of match_points (slam.py:130-141) and the two normalize(K, .) calls of
 * fit_essential_matrix (slam.py:151-152; normalize at :87-92).  Emits the surviving (queryIdx, trainIdx)
 * in ascending query order, their count, and (when keypoint pools are given) the K-normalised
 * correspondences: src = LAST (train) side, dst = CURR (query) side, f64[batch][n_max][2] each,
 * plus an fp32 copy corr32 float4[batch][n_max] = (src.x, src.y, dst.x, dst.y) for scoring.
 * kinv = inv(K) row-major 3x3 (HOST pointer, 9 doubles). kps/src/dst/corr32 may all be NULL. */
TVG_API int tvg_ratio_compact(tvg_ctx* ctx, const uint8_t* keep, const int32_t* idx, int batch, int n_max, int nq,
                      const int32_t* pair_q, const int32_t* pair_t, const int32_t* q_count,
                      const double* kps_q, const double* kps_t, const double* kinv_host, int32_t* qidx,
                      int32_t* tidx, int32_t* m_count, double* src, double* dst, float* corr32,
                      void* stream);

/* skimage.measure.ransac((src, dst), EssentialMatrixTransform, min_samples=8, residual_threshold=thr,
 * max_trials=H)  (slam.py:155-159), driven by a host-supplied sample table: either explicit indices
 * samples int32[H][8] or random words u32[H][8] mapped on device by idx = (w*M)>>32 with in-row clashes
 * bumped +1 mod M.  sample_stride = elements between pairs (0: one table shared by the batch).
 * All outputs are optional (NULL -> internal scratch):
 *   E f64[batch][9] (refit on the inliers), mask u8[batch][n_max] (PRE-refit best model's inliers),
 *   n_inl, best int32[batch], E_best f64[batch][9], counts int32[batch][H] (-1 = failed fit; exact fp64
 *   count for the finalists, fp32 count otherwise), sse32 f32[batch][H] (fp32 sum r^2, diagnostic: computed only
 *   when requested), models f64[batch][H][9],
 *   Rt f64[batch][16] (extract_Rt of the refit E), status int32[batch] (0 ok, 1: M <= 8, 2: no model),
 *   inl_idx int32[batch][n_max] (match indices of the inliers), samples_out int32[batch][H][8].
 * Optional inlier gather for triangulation: when qidx/tidx and pixel keypoint pools are given,
 *   pix_last / pix_curr f64[batch][n_max][2] receive frame.last_kps / frame.curr_kps after the
 *   inlier filter of slam.py:161-162. */
typedef struct tvg_ransac_io {
    const double* src;
    const double* dst;
    const float* corr32;
    const int32_t* m_count; /* NULL -> m */
    int batch, n_max, m, H;
    const int32_t* samples;
    const uint32_t* words;
    long long sample_stride;
    double thr;
    /* optional gather inputs */
    const int32_t* qidx;
    const int32_t* tidx;
    const int32_t* pair_q;
    const int32_t* pair_t;
    const double* kps_q;
    const double* kps_t;
    /* outputs (all optional) */
    double* E;
    uint8_t* mask;
    int32_t* n_inl;
    int32_t* best;
    double* E_best;
    int32_t* counts;
    float* sse32;
    double* models;
    double* Rt;
    int32_t* status;
    int32_t* inl_idx;
    int32_t* samples_out;
    double* pix_last;
    double* pix_curr;
    int32_t* band;      /* fp32 uncertainty-band count per hypothesis */
    int32_t* n_cand;    /* finalists re-scored in fp64 */
    double* best_sse;   /* [batch] exact sum of squared residuals of the winner (all points) */
} tvg_ransac_io;
TVG_API int tvg_ransac_essential(tvg_ctx* ctx, const tvg_ransac_io* io, void* stream);

/* extract_Rt(E)  (slam.py:167-186): E f64[batch][9] -> Rt f64[batch][16]. */
TVG_API int tvg_extract_rt(tvg_ctx* ctx, const double* E, int batch, double* Rt, void* stream);

/* frame.curr_pose = np.dot(Rt, frame.last_pose)  (slam.py:261) chained over sequences.
 * Pairs of sequence s are b in [seq_off[s], seq_off[s+1]); poses f64[batch + nseq][16]: the pose of
 * pair b's LAST frame sits at row b + s, its CURR frame at row b + s + 1.
 * init_pose f64[nseq][16] (NULL -> identity, slam.py:247).  pair_pinv f64[batch][2][16] (may be NULL)
 * receives np.linalg.inv(last_pose), np.linalg.inv(curr_pose) per pair -- what triangulate needs. */
TVG_API int tvg_pose_chain(tvg_ctx* ctx, const double* Rt, const int32_t* seq_off, int nseq,
                   const double* init_pose, double* poses, double* pair_pinv, void* stream);

/* np.linalg.inv(pose) (slam.py:204-205) for n row-major 4x4 matrices (pivoted Gauss-Jordan, fp64). */
TVG_API int tvg_invert_poses(tvg_ctx* ctx, const double* poses, int n, double* poses_inv, void* stream);

/* triangulate(pts1, pts2, pose1, pose2) (slam.py:202-221) + check_points (slam.py:233-237).
 * pix1 / pix2 f64[batch][n_max][2] pixels; P1 / P2 = inv(pose1) / inv(pose2), one 4x4 per pair at
 * strides p1_stride / p2_stride doubles; n_pts int32[batch] (NULL -> n); kinv_host = inv(K).
 * points4d f64[batch][n_max][4], good u8[batch][n_max] (z > 0; may be NULL). */
TVG_API int tvg_triangulate(tvg_ctx* ctx, const double* pix1, const double* pix2, int batch, int n_max, int n,
                    const int32_t* n_pts, const double* P1, long long p1_stride, const double* P2,
                    long long p2_stride, const double* kinv_host, double* points4d, uint8_t* good,
                    void* stream);

/* tvg_triangulate plus the two point filters slam.py:236 leaves as TODO ("parallax, reprojection error"):
 * reproj f64[batch][n_max] = max over both views of the reprojection error in pixels (focal = K[0][0]),
 * parallax_cos f64[batch][n_max] = cosine of the angle between the two viewing rays.  No reference behaviour
 * exists for these; their spec is oracle/two_view.py:point_quality.  Either may be NULL. */
TVG_API int tvg_point_quality(tvg_ctx* ctx, const double* pix1, const double* pix2, int batch, int n_max, int n,
                      const int32_t* n_pts, const double* P1, long long p1_stride, const double* P2,
                      long long p2_stride, const double* kinv_host, double focal, double* points4d, uint8_t* good,
                      double* reproj, double* parallax_cos, void* stream);

/* Map.add_observation's point-cloud append (slam.py:37-40) for whole sequences: per sequence s,
 * map_points[s] = [(0,0,0,1) if include_origin (the first frame's entry, slam.py:248)] followed by every pair's
 * points4d rows with good != 0, in pair order -- the `points` half of the (poses, points) tuple Map.display()
 * sends to the viewer (slam.py:81-84; the `poses` half is tvg_pose_chain's output).
 * map_points f64[nseq][cap][4]; map_count int32[nseq] = rows needed (rows beyond cap are dropped);
 * pair_off int32[batch] (optional) = first row of each pair inside its sequence. */
TVG_API int tvg_map_accumulate(tvg_ctx* ctx, const double* points4d, const uint8_t* good, const int32_t* n_pts,
                       const int32_t* seq_off, int nseq, int n_max, long long cap, int include_origin,
                       double* map_points, int32_t* map_count, int32_t* pair_off, void* stream);

/* extract_points(frame) (slam.py:112-122):  cv2.cvtColor(BGR2GRAY) -> cv2.goodFeaturesToTrack(image, max_corners,
 * qualityLevel, minDistance) -> cv2.ORB_create().compute(image, [KeyPoint(x, y, size=20)]).  SURVEY §8f rank 1; outside
 * the north-star path (the benchmark feeds synthetic keypoints), provided so that a whole frame can stay on the GPU.
 * images u8[batch][H][W][channels] (device; channels 3 = BGR, 1 = already gray).  pattern_host = int8[512][2], the ORB
 * test pattern rotated by the KeyPoint default angle and rounded (HOST pointer; openwork_b200/orb_pattern.npy).
 * Outputs (device): kps f64[batch][max_corners][2] pixel (x, y) of the keypoints that survive ORB's 31-pixel border
 * filter, in goodFeaturesToTrack order; des u8[batch][max_corners][32]; n_kps int32[batch];
 * n_candidates int32[batch] (optional) = local maxima found before the min-distance filter. */
TVG_API int tvg_extract_points(tvg_ctx* ctx, const uint8_t* images, int batch, int height, int width, int channels,
                       int max_corners, double quality_level, double min_distance, const int8_t* pattern_host,
                       double* kps, uint8_t* des, int32_t* n_kps, int32_t* n_candidates, void* stream);

/* Two-view bundle adjustment per pair -- the back-end step slam.py:267 leaves as a TODO ("g2o bundle adjustment");
 * no reference behaviour, specification = oracle/bundle.py.  Levenberg-Marquardt (Marquardt scaling, lambda0, a step is
 * accepted iff the cost decreases: lambda /= 10, else *= 10; exactly `iters` iterations) on the squared reprojection
 * error in normalised image coordinates over the SECOND view's pose (6-DoF left twist; the first view is fixed) and
 * the points (eliminated by the Schur complement).  pose1 / pose2: camera-to-world 4x4 per pair (the reference's
 * `pose`, slam.py:261) at strides pose*_stride doubles; pose2 and points4d f64[batch][n_max][4] are refined in place.
 * pix1 / pix2 f64[batch][n_max][2] pixels; n_pts int32[batch] (NULL -> n); use u8[batch][n_max] (NULL -> all points).
 * cost f64[batch][2] = initial / final cost; accepted int32[batch] (may be NULL) = accepted steps. */
TVG_API int tvg_bundle_adjust_pairs(tvg_ctx* ctx, const double* pix1, const double* pix2, int batch, int n_max, int n,
                            const int32_t* n_pts, const uint8_t* use, const double* kinv_host, const double* pose1,
                            long long pose1_stride, double* pose2, long long pose2_stride, double* points4d,
                            int iters, double lambda0, double* cost, int32_t* accepted, void* stream);

/* Sliding-window bundle adjustment over W <= 8 consecutive frames, batched over windows -- the multi-frame form of the
 * same TODO (slam.py:267); no reference behaviour, specification = oracle/bundle_window.py.  Same Levenberg-Marquardt
 * scheme as tvg_bundle_adjust_pairs over the poses of the cameras c >= n_fixed (n_fixed >= 1 leading cameras stay put:
 * 1 leaves the scale gauge to the damping, 2 fixes it) and the points, points eliminated by the Schur complement
 * (reduced camera system of size 6 (W - n_fixed)), one CTA per window, fp64, fixed-order sums.
 * obs f64[batch][n_max][W][2] pixels of point j in camera c, vis u8[batch][n_max][W]; poses f64[batch][W][16]
 * camera-to-world and points4d f64[batch][n_max][4] are refined in place; n_pts int32[batch] (NULL -> n);
 * cost f64[batch][2] = initial / final; accepted int32[batch] (may be NULL). */
TVG_API int tvg_bundle_adjust_window(tvg_ctx* ctx, const double* obs, const uint8_t* vis, int batch, int n_max, int n,
                             const int32_t* n_pts, int W, int n_fixed, const double* kinv_host, double* poses,
                             double* points4d, int iters, double lambda0, double* cost, int32_t* accepted, void* stream);

/* Cheirality vote over the four (R, +-t) decompositions of E (north-star extra; the reference uses the
 * trace / t_z heuristics of slam.py:174-181 instead).  votes int32[batch][4], Rts f64[batch][4][16],
 * choice int32[batch].  mask may be NULL. */
TVG_API int tvg_cheirality(tvg_ctx* ctx, const double* E, const double* src, const double* dst, const uint8_t* mask,
                   const int32_t* m_count, int batch, int n_max, int m, int32_t* votes, double* Rts,
                   int32_t* choice, void* stream);

/* The `else` branch of process_frame (slam.py:249-266) for a batch of frame pairs, device resident,
 * no host synchronisation: match -> ratio/compact -> RANSAC -> extract_Rt -> pose chain ->
 * triangulate -> z>0.  Frame pools: des u8[frames][n_max][32], kps f64[frames][n_max][2],
 * frame_count int32[frames] (NULL -> all n_max rows valid).  Pair b: curr = pair_q[b], last = pair_t[b]
 * (NULL -> curr = b + 1 + s, last = b + s for sequence s, i.e. consecutive frames of each sequence).
 * Outputs are optional (NULL -> scratch). */
typedef struct tvg_pairs_io {
    const uint8_t* des;
    const double* kps;
    const int32_t* frame_count;
    const int32_t* pair_q;
    const int32_t* pair_t;
    int batch, n_max, H;
    const int32_t* seq_off;   /* int32[nseq+1]; NULL -> one sequence holding all pairs */
    int nseq;
    const double* init_pose;  /* f64[nseq][16] or NULL */
    const uint32_t* words;    /* u32[.][H][8] */
    const int32_t* samples;   /* int32[.][H][8] (alternative to words) */
    long long sample_stride;
    double kinv[9];
    double ratio, thr;
    /* outputs */
    int32_t* idx;       /* [batch][n_max][2] */
    int32_t* dist;      /* [batch][n_max][2] */
    int32_t* qidx;      /* [batch][n_max] */
    int32_t* tidx;      /* [batch][n_max] */
    int32_t* m_count;   /* [batch] */
    double* E;          /* [batch][9] */
    double* E_best;     /* [batch][9] */
    int32_t* best;      /* [batch] */
    int32_t* counts;    /* [batch][H] */
    uint8_t* mask;      /* [batch][n_max] */
    int32_t* n_inl;     /* [batch] */
    int32_t* inl_idx;   /* [batch][n_max] */
    double* Rt;         /* [batch][16] */
    double* poses;      /* [batch + nseq][16] */
    double* points4d;   /* [batch][n_max][4] */
    uint8_t* good;      /* [batch][n_max] */
    int32_t* status;    /* [batch] */
    /* 0: Rt by the reference's heuristics (extract_Rt, slam.py:167-186; default, parity path);
     * 1: Rt by a cheirality vote over the inliers among the four decompositions of E (north-star extra) */
    int pose_mode;
    int32_t* pose_votes; /* [batch][4], pose_mode 1 only (optional) */
} tvg_pairs_io;
TVG_API int tvg_pairs_device(tvg_ctx* ctx, const tvg_pairs_io* io, void* stream);

/* Stage timing of tvg_pairs_device: with profiling on, CUDA events are recorded at the stage boundaries on the
 * caller's stream; tvg_get_stage_ms waits for the last profiled call and returns the 9 stage durations in ms:
 * knn2, compact, fit, score, reduce, rescore, refit+solve, pose (cheirality + pose chain), triangulate. */
TVG_API int tvg_set_profile(tvg_ctx* ctx, int on);
TVG_API int tvg_get_stage_ms(tvg_ctx* ctx, float* ms, int n);

/* Pairs per chunk of tvg_pairs_host's copy/compute pipeline (default 256). */
TVG_API int tvg_set_host_chunk(tvg_ctx* ctx, int pairs);

/* Same pipeline behind HOST buffers (what a numpy caller hands over).  Every pointer in `io` is a HOST
 * pointer here (pinned memory makes the copies asynchronous).  The batch is cut into chunks of pairs:
 * frame copies host->device (own stream), kernels (caller's stream) and result copies device->host (own
 * stream) of successive chunks overlap; the call returns after synchronising.  Frames are copied in
 * index order up to the highest frame a chunk references.  bytes_h2d / bytes_d2h (may be NULL) receive
 * the bytes moved. */
TVG_API int tvg_pairs_host(tvg_ctx* ctx, const tvg_pairs_io* io, int frames, long long* bytes_h2d,
                   long long* bytes_d2h, void* stream);

/* Streaming form of tvg_pairs_host: _submit enqueues the copies and kernels and returns; _wait blocks until
 * every requested result is in the host buffers.  One batch may be in flight per context: alternate two
 * contexts to overlap the copies of batch k+1 with the kernels of batch k.  Host buffers must stay
 * untouched between _submit and _wait. */
TVG_API int tvg_pairs_host_submit(tvg_ctx* ctx, const tvg_pairs_io* io, int frames, long long* bytes_h2d,
                          long long* bytes_d2h, void* stream);
TVG_API int tvg_pairs_host_wait(tvg_ctx* ctx);

/* Measurement aid (not on the path): the e4m3 tensor-core peak of this device in TFLOP/s, from `iters` rounds of 8
 * back-to-back tcgen05.mma kind::f8f6f4 (M=128, N=256, K=32, f16 accumulate) per SM with resident operands; best of 3.
 * bench.py uses it as the roofline denominator of the tensor-core matcher. */
TVG_API int tvg_measure_fp8_peak(tvg_ctx* ctx, int iters, double* tflops, double* ms);

/* ---- multi-GPU (one process per GPU) ---------------------------------------------------------------------------
 * An NCCL communicator bound to a context ("context create/destroy for workspace + NCCL communicator").
 * Rank 0 calls tvg_comm_unique_id and hands the 128 bytes to the other ranks over any host channel (the Python
 * layer broadcasts them through torch.distributed); every rank then calls tvg_comm_init (collective).  NCCL is
 * resolved at run time (libnccl.so.2); the single-GPU entry points never need it. */
#define TVG_COMM_ID_BYTES 128
TVG_API int tvg_comm_unique_id(void* id_out);
TVG_API int tvg_comm_init(tvg_ctx* ctx, const void* id_bytes, int rank, int world);
TVG_API int tvg_comm_destroy(tvg_ctx* ctx);
TVG_API int tvg_comm_info(const tvg_ctx* ctx, int* rank, int* world);

/* Contiguous block [lo, hi) of n items owned by `rank` (the first n % world ranks own one more). */
TVG_API void tvg_shard_range(int n, int rank, int world, int* lo, int* hi);
/* Rows every per-row array of tvg_huge_pair must hold: max(ceil(nq / world) * world, nt). */
TVG_API int tvg_huge_pair_rows(int nq, int nt, int world);

/* ONE huge keyframe pair (BASELINE.json config C5: 64k x 64k descriptors, 65536 hypotheses) split over the ranks of
 * the context's communicator; the `else` branch of process_frame (slam.py:249-266) for that pair, device resident, no
 * host synchronisation.  Inputs are replicated on every rank (descriptor sets are <= 2 MiB).
 *   matcher     rank r matches query rows [r * per, (r + 1) * per), per = ceil(nq / world), against the whole train set
 *   exchange 1  fixed-size all-gather of idx / dist / keep (rank order = ascending query order)
 *   compaction  ratio loop + gathers + K-normalisation over all rows, on every rank (identical)
 *   RANSAC      rank r fits and scores hypotheses tvg_shard_range(H, r, world) of the SAME table; exact fp64
 *               re-score of its finalists; its best as a 96-byte record {count, hypothesis, sum r^2, E}
 *   exchange 2  all-gather of the records; every rank applies the reference's selection rule (most inliers, then
 *               smallest sum of squared residuals, then earliest trial; skimage ransac) to the same
 *               list with exact keys: world ranks return bit for bit what one rank returns
 *   refit ...   inlier mask, refit, extract_Rt, pose, triangulation, z > 0 on every rank (identical)
 * Without a communicator (or world = 1) the same code runs with no exchange.  Every output is optional and, except
 * `counts`, identical on all ranks.  Per-row arrays hold tvg_huge_pair_rows(nq, nt, world) rows. */
typedef struct tvg_huge_pair_io {
    const uint8_t* curr_des;   /* u8[nq][32]  query side (frame.curr_des) */
    const uint8_t* last_des;   /* u8[nt][32]  train side (frame.last_des) */
    const double* curr_kps;    /* f64[nq][2] pixels */
    const double* last_kps;    /* f64[nt][2] pixels */
    int nq, nt, H;
    const uint32_t* words;     /* u32[H][8] the whole table, same on every rank (or samples) */
    const int32_t* samples;    /* int32[H][8] */
    double kinv[9];
    double ratio, thr;
    /* outputs */
    int32_t* idx;        /* [rows][2] */
    int32_t* dist;       /* [rows][2] */
    int32_t* qidx;       /* [rows] */
    int32_t* tidx;       /* [rows] */
    int32_t* m_count;    /* [1] */
    double* E;           /* [9] refit */
    double* E_best;      /* [9] */
    int32_t* best;       /* [1] GLOBAL hypothesis index */
    double* best_sse;    /* [1] */
    int32_t* counts;     /* [H]: this rank writes its own block only */
    int32_t* n_cand;     /* [1] finalists this rank re-scored in fp64 */
    uint8_t* mask;       /* [rows] */
    int32_t* n_inl;      /* [1] */
    int32_t* inl_idx;    /* [rows] */
    double* Rt;          /* [16] */
    double* poses;       /* [2][16]: identity, Rt @ identity */
    double* points4d;    /* [rows][4] */
    uint8_t* good;       /* [rows] */
    int32_t* status;     /* [1] */
} tvg_huge_pair_io;
TVG_API int tvg_huge_pair(tvg_ctx* ctx, const tvg_huge_pair_io* io, void* stream);
/* With profiling on (tvg_set_profile): durations in ms of the two exchange steps of the last tvg_huge_pair call on THIS
 * rank -- ms[0] = the all-gather of the matcher outputs, ms[1] = the all-gather of the best records.  Each includes the
 * wait for the slowest rank to arrive; the minimum over the ranks is the transfer itself. */
TVG_API int tvg_get_exchange_ms(tvg_ctx* ctx, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* TVG_B200_H */
