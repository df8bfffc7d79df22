"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the two-view geometry hot path.

This module is a plain numpy restatement of what OpenWork's ``MonocularSlam/slam.py``
computes on the path ``match_points -> fit_essential_matrix -> extract_Rt -> triangulate
-> check_points``.  It exists to CHECK the CUDA path (``tests/``, ``__graft_entry__.smoke``)
and to be TIMED as the CPU baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``).
Nothing in ``openwork_b200/`` imports it; the product path has no CPU fallback.

Where the arithmetic lives
--------------------------
The reference file holds no arithmetic of its own for the heavy steps; it calls third-party
libraries that are NOT vendored under ``/root/reference`` and whose versions it does not pin:

* ``cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)``  (slam.py:126-127)
  -- opencv-python, unpinned; this image has 4.13.0 and the oracle calls it directly
  (``knn2_cv2``); ``knn2_numpy`` restates its published behaviour (popcount of XOR, the two
  smallest distances scanning train rows ascending with strict ``<`` -> ties break to the
  lowest train index).
* ``skimage.measure.ransac`` + ``skimage.transform.EssentialMatrixTransform``
  (slam.py:20-21, 155-159) -- scikit-image, unpinned (sibling ``sort/README.md:10`` pins
  0.17.2), ABSENT from this image and not installable.  ``center_and_normalize_points``,
  ``essential_estimate``, ``essential_residuals`` and ``ransac_essential`` restate the
  published algorithm of ``skimage/transform/_geometric.py`` and ``skimage/measure/fit.py``
  (0.17-0.25, arithmetic unchanged) as recorded in SURVEY.md Appendix A.
* ``numpy.linalg`` (svd / inv / det) -- called directly, as the reference does.

Pinning status
--------------
The reference has NO tests, golden vectors or fixtures for this path.
* matcher / ratio / normalize / extract_Rt / triangulate / check_points: pinned against the
  reference's own functions imported from ``/root/reference/MonocularSlam/slam.py`` in the
  build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
* RANSAC + essential estimate/residuals: **parity unpinned** -- scikit-image cannot be run
  here, so these follow the restated algorithm only; cross-checked against
  ``cv2.sampsonDistance`` (r**2) and closed-form ``E = [t]x R`` scenes.
"""
from __future__ import annotations

import numpy as np

RATIO = 0.75               # slam.py:131
RESIDUAL_THRESHOLD = 0.005  # slam.py:158
MIN_SAMPLES = 8            # slam.py:157
MAX_TRIALS = 200           # slam.py:159


# --------------------------------------------------------------------------- matcher
def hamming_matrix(q, t):
    """dist[i,j] = popcount(q[i] XOR t[j]) over 32 bytes -> int32[Nq,Nt] (Appendix A.1)."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    q64 = q.view(np.uint64)          # [Nq,4]
    t64 = t.view(np.uint64)          # [Nt,4]
    d = np.zeros((q.shape[0], t.shape[0]), dtype=np.int32)
    for w in range(4):
        d += np.bitwise_count(q64[:, w, None] ^ t64[None, :, w]).astype(np.int32)
    return d


def knn2_numpy(q, t, block=1024):
    """Two nearest train rows per query row; ties -> lowest train index.

    Restates ``cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)`` (slam.py:126-127):
    equals ``np.argsort(dist, kind='stable')[:, :2]``.  Returns (idx int32[Nq,2],
    dist int32[Nq,2]); with Nt == 1 the second column is (-1, -1).
    """
    nq, nt = len(q), len(t)
    idx = np.full((nq, 2), -1, dtype=np.int32)
    dist = np.full((nq, 2), -1, dtype=np.int32)
    if nq == 0 or nt == 0:
        return idx, dist
    for s in range(0, nq, block):
        d = hamming_matrix(q[s:s + block], t).astype(np.int64)
        key = d * nt + np.arange(nt)[None, :]          # (distance, train index) lexicographic
        r = np.arange(len(d))
        k1 = key.argmin(axis=1)
        idx[s:s + block, 0] = k1
        dist[s:s + block, 0] = d[r, k1]
        if nt > 1:
            key[r, k1] = np.iinfo(np.int64).max
            k2 = key.argmin(axis=1)
            idx[s:s + block, 1] = k2
            dist[s:s + block, 1] = d[r, k2]
    return idx, dist


def knn2_cv2(q, t):
    """The real third-party call the reference makes (slam.py:126-127), as arrays."""
    import cv2
    nq = len(q)
    idx = np.full((nq, 2), -1, dtype=np.int32)
    dist = np.full((nq, 2), -1, dtype=np.int32)
    if nq == 0 or len(t) == 0:
        return idx, dist
    matches = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(
        np.ascontiguousarray(q, dtype=np.uint8), np.ascontiguousarray(t, dtype=np.uint8), k=2)
    for i, row in enumerate(matches):
        for k, m in enumerate(row[:2]):
            idx[i, k] = m.trainIdx
            dist[i, k] = int(m.distance)
    return idx, dist


def ratio_filter(idx, dist, ratio=RATIO):
    """Lowe ratio loop (slam.py:130-137): keep iff ``d1 < ratio*d2`` (float compare, strict),
    ascending query order.  Returns (qidx int32[M], tidx int32[M])."""
    d1 = dist[:, 0].astype(np.float32)
    d2 = dist[:, 1].astype(np.float32)
    keep = (d1 < ratio * d2.astype(np.float64)) & (idx[:, 1] >= 0)
    qidx = np.flatnonzero(keep).astype(np.int32)
    return qidx, idx[keep, 0].astype(np.int32)


def match_points(curr_kps, curr_des, last_kps, last_des, ratio=RATIO, use_cv2=True):
    """Array form of ``match_points(frame)`` (slam.py:125-143).

    Returns (match_kps f64[M,2,2] with [:,0]=curr xy and [:,1]=last xy, idx1, idx2).
    Raises ValueError when the train set has < 2 rows (the reference's ``for m,n in``
    unpack, slam.py:130) and AssertionError when < 8 matches survive (slam.py:138).
    """
    if len(last_des) < 2:
        raise ValueError("not enough values to unpack (expected 2, got %d)" % len(last_des))
    idx, dist = (knn2_cv2 if use_cv2 else knn2_numpy)(curr_des, last_des)
    idx1, idx2 = ratio_filter(idx, dist, ratio)
    assert len(idx1) >= 8
    match_kps = np.stack([curr_kps[idx1], last_kps[idx2]], axis=1)
    return match_kps, idx1, idx2


# --------------------------------------------------------------------------- geometry
def normalize(K, pts):
    """``normalize(K, pts)`` (slam.py:87-92): (inv(K) @ [x,y,1]^T)^T[:, :2]."""
    Kinv = np.linalg.inv(K)
    ph = np.concatenate([pts, np.ones((pts.shape[0], 1))], axis=1)
    return (Kinv @ ph.T).T[:, 0:2]


def center_and_normalize_points(P):
    """skimage ``_center_and_normalize_points`` (Appendix A.2): centroid -> 0, RMS distance
    -> sqrt(2).  Returns (T[3,3], Pn[n,2]); NaNs when rms == 0."""
    P = np.asarray(P, dtype=np.float64)
    n = P.shape[0]
    c = np.mean(P, axis=0)
    rms = np.sqrt(np.sum((P - c) ** 2) / n)
    if rms == 0:
        return np.full((3, 3), np.nan), np.full_like(P, np.nan)
    s = np.sqrt(2.0) / rms
    T = np.array([[s, 0, -s * c[0]], [0, s, -s * c[1]], [0, 0, 1]])
    ph = np.vstack([P.T, np.ones(n)])
    nh = (T @ ph).T
    out = nh[:, :2].copy()
    out[:, 0] /= nh[:, 2]
    out[:, 1] /= nh[:, 2]
    return T, out


def essential_estimate(src, dst, full_matrices=True):
    """skimage ``EssentialMatrixTransform.estimate`` (Appendix A.2; call site slam.py:156).

    Hartley-normalise, 8-point DLT null vector, (s,s,0) projection, de-normalise.
    Returns E f64[3,3] or None when the fit is not finite (estimate -> False).
    ``full_matrices=False`` is only valid for n >= 9 rows (the inlier refit at scale).
    """
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    Ts, s = center_and_normalize_points(src)
    Td, d = center_and_normalize_points(dst)
    if not (np.all(np.isfinite(Ts)) and np.all(np.isfinite(Td))):
        return None
    A = np.ones((s.shape[0], 9))
    A[:, 0] = d[:, 0] * s[:, 0]
    A[:, 1] = d[:, 0] * s[:, 1]
    A[:, 2] = d[:, 0]
    A[:, 3] = d[:, 1] * s[:, 0]
    A[:, 4] = d[:, 1] * s[:, 1]
    A[:, 5] = d[:, 1]
    A[:, 6] = s[:, 0]
    A[:, 7] = s[:, 1]
    _, _, V = np.linalg.svd(A, full_matrices=full_matrices)
    En = V[-1, :].reshape(3, 3)
    U, S, Vt = np.linalg.svd(En)
    S[0] = (S[0] + S[1]) / 2.0
    S[1] = S[0]
    S[2] = 0
    E = U @ np.diag(S) @ Vt
    E = Td.T @ E @ Ts
    if not np.all(np.isfinite(E)):
        return None
    return E


def essential_residuals(E, src, dst):
    """skimage ``FundamentalMatrixTransform.residuals`` (Appendix A.2): first-order Sampson
    distance ``|x'^T E x| / sqrt((Ex)_0^2 + (Ex)_1^2 + (E^T x')_0^2 + (E^T x')_1^2)``."""
    sh = np.column_stack([src, np.ones(len(src))])
    dh = np.column_stack([dst, np.ones(len(dst))])
    Fs = E @ sh.T
    Ftd = E.T @ dh.T
    num = np.sum(dh * Fs.T, axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.abs(num) / np.sqrt(Fs[0] ** 2 + Fs[1] ** 2 + Ftd[0] ** 2 + Ftd[1] ** 2)


def sample_conditioning(src, dst, samples):
    """sigma_8/sigma_1 of each hypothesis' normalised 8x9 system (to flag near-degenerate
    samples whose LAPACK null vector is arbitrary; SURVEY.md §7 'hard parts' 2)."""
    out = np.zeros(len(samples))
    for h, s in enumerate(samples):
        _, a = center_and_normalize_points(src[s])
        _, b = center_and_normalize_points(dst[s])
        if not (np.all(np.isfinite(a)) and np.all(np.isfinite(b))):
            continue
        A = np.ones((8, 9))
        A[:, 0] = b[:, 0] * a[:, 0]; A[:, 1] = b[:, 0] * a[:, 1]; A[:, 2] = b[:, 0]
        A[:, 3] = b[:, 1] * a[:, 0]; A[:, 4] = b[:, 1] * a[:, 1]; A[:, 5] = b[:, 1]
        A[:, 6] = a[:, 0]; A[:, 7] = a[:, 1]
        sv = np.linalg.svd(A, compute_uv=False)
        out[h] = sv[7] / sv[0]
    return out


def ransac_essential(src, dst, samples, residual_threshold=RESIDUAL_THRESHOLD,
                     return_models=False):
    """skimage ``ransac((src,dst), EssentialMatrixTransform, 8, thr, max_trials=H)``
    (Appendix A.3; call site slam.py:155-159) driven by an explicit sample table.

    ``samples int32[H,8]`` replaces the per-trial ``rng.choice`` -- the one deliberate,
    north-star-mandated deviation.  With the reference's arguments there is no early exit,
    so every row is a trial.  Best = most inliers, then smallest sum of squared residuals
    over ALL points, then earliest.  The returned mask is the PRE-refit best model's; the
    returned E is the refit on that mask.

    Returns dict(E, mask, counts int32[H], sse f64[H], best, E_best, [models f64[H,3,3]]).
    Hypotheses whose estimate fails have count -1 / sse inf and are skipped.
    """
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    M = len(src)
    if not (0 < MIN_SAMPLES < M):
        raise ValueError("`min_samples` must be in range (0, <number-of-samples>)")
    H = len(samples)
    counts = np.full(H, -1, dtype=np.int32)
    sse = np.full(H, np.inf)
    models = np.full((H, 3, 3), np.nan) if return_models else None
    best, best_n, best_rss, best_E, best_mask = -1, 0, np.inf, None, None
    for h in range(H):
        s = samples[h]
        E = essential_estimate(src[s], dst[s])
        if E is None:
            continue
        res = np.abs(essential_residuals(E, src, dst))
        inl = res < residual_threshold
        rss = float(np.sum(res ** 2))
        n = int(np.sum(inl))
        counts[h] = n
        sse[h] = rss
        if return_models:
            models[h] = E
        if n > best_n or (n == best_n and rss < best_rss):
            best, best_n, best_rss, best_E, best_mask = h, n, rss, E, inl
    out = dict(E=None, mask=best_mask, counts=counts, sse=sse, best=best, E_best=best_E)
    if best_mask is not None:
        n_in = int(best_mask.sum())
        if n_in >= 8:
            out["E"] = essential_estimate(src[best_mask], dst[best_mask],
                                          full_matrices=(n_in < 9))
        else:
            # skimage's refit would raise for < 8 points; report the un-refit model
            out["E"] = best_E
    if return_models:
        out["models"] = models
    return out


def fit_essential_matrix(match_kps, K, samples, residual_threshold=RESIDUAL_THRESHOLD):
    """Array form of ``fit_essential_matrix(match_kps)`` (slam.py:146-164): K-normalise both
    sides, RANSAC with src=last, dst=curr.  Returns (E, inliers bool[M], full result dict)."""
    match_kps = np.asarray(match_kps)
    norm_curr = normalize(K, match_kps[:, 0])
    norm_last = normalize(K, match_kps[:, 1])
    r = ransac_essential(norm_last, norm_curr, samples, residual_threshold)
    return r["E"], r["mask"], r


def extract_Rt(E):
    """``extract_Rt(E)`` (slam.py:167-186): SVD, det sign fixes, W / W^T by trace(R) >= 0,
    t = U[:,2] flipped so t_z >= 0.  Returns Rt f64[4,4]."""
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], dtype=float)
    U, _, Vt = np.linalg.svd(np.asarray(E, dtype=np.float64))
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    R = U @ W @ Vt
    if np.trace(R) < 0:
        R = U @ W.T @ Vt
    t = U[:, 2].copy()
    if t[2] < 0:
        t = -t
    Rt = np.eye(4)
    Rt[:3, :3] = R
    Rt[:3, 3] = t
    return Rt


def triangulate(pts1, pts2, pose1, pose2, K):
    """``triangulate(pts1, pts2, pose1, pose2)`` (slam.py:202-221): P = inv(pose); per point
    the 4x4 DLT system, ``vt[3]`` of its SVD, divided by w.  Same per-point LAPACK loop as
    the reference so that it can also serve as the timed CPU baseline."""
    P1 = np.linalg.inv(pose1)
    P2 = np.linalg.inv(pose2)
    n1 = normalize(K, pts1)
    n2 = normalize(K, pts2)
    X = np.zeros((n1.shape[0], 4))
    A = np.zeros((4, 4))
    for i in range(n1.shape[0]):
        A[0] = n1[i, 0] * P1[2] - P1[0]
        A[1] = n1[i, 1] * P1[2] - P1[1]
        A[2] = n2[i, 0] * P2[2] - P2[0]
        A[3] = n2[i, 1] * P2[2] - P2[1]
        X[i] = np.linalg.svd(A)[2][3]
    with np.errstate(divide="ignore", invalid="ignore"):
        X /= X[:, 3:]
    return X


def triangulate_batched(pts1, pts2, pose1, pose2, K):
    """Same result as :func:`triangulate` with one batched LAPACK call (for large M)."""
    P1 = np.linalg.inv(pose1)
    P2 = np.linalg.inv(pose2)
    n1 = normalize(K, pts1)
    n2 = normalize(K, pts2)
    A = np.empty((n1.shape[0], 4, 4))
    A[:, 0] = n1[:, 0:1] * P1[2] - P1[0]
    A[:, 1] = n1[:, 1:2] * P1[2] - P1[1]
    A[:, 2] = n2[:, 0:1] * P2[2] - P2[0]
    A[:, 3] = n2[:, 1:2] * P2[2] - P2[1]
    X = np.linalg.svd(A)[2][:, 3, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        return X / X[:, 3:]


def check_points(points4d):
    """``check_points`` (slam.py:233-237): world-frame z > 0."""
    return points4d[:, 2] > 0


def point_quality(pts1, pts2, pose1, pose2, K):
    """SPEC (no reference behaviour: slam.py:236 only says "TODO: parallax, reprojection error filters").
    For each triangulated point: the larger of the two reprojection errors, in pixels (focal = K[0,0]), and
    the cosine of the angle between the two viewing rays (parallax).  Returns (points4d, reproj_px, cos)."""
    X = triangulate_batched(pts1, pts2, pose1, pose2, K)
    P1, P2 = np.linalg.inv(pose1), np.linalg.inv(pose2)
    n1, n2 = normalize(K, pts1), normalize(K, pts2)
    c1 = X @ P1[:3].T
    c2 = X @ P2[:3].T
    with np.errstate(divide="ignore", invalid="ignore"):
        e1 = np.hypot(c1[:, 0] / c1[:, 2] - n1[:, 0], c1[:, 1] / c1[:, 2] - n1[:, 1])
        e2 = np.hypot(c2[:, 0] / c2[:, 2] - n2[:, 0], c2[:, 1] / c2[:, 2] - n2[:, 1])
        r1 = c1 @ P1[:3, :3]          # R1^T c1 : ray in the world frame
        r2 = c2 @ P2[:3, :3]
        cos = np.sum(r1 * r2, axis=1) / (np.linalg.norm(r1, axis=1) * np.linalg.norm(r2, axis=1))
    return X, float(np.asarray(K)[0, 0]) * np.maximum(e1, e2), cos


def map_points(pairs_points4d, pairs_good):
    """Map.add_observation's append (slam.py:37-40, 248, 266): origin for the first frame, then each pair's
    z>0 points in order.  Returns f64[P,4]."""
    rows = [np.array([[0.0, 0.0, 0.0, 1.0]])]
    for X, g in zip(pairs_points4d, pairs_good):
        rows.append(X[g])
    return np.concatenate(rows, axis=0)


def cheirality_votes(E, src, dst):
    """Votes for the four (R, +-t) decompositions of E: number of correspondences whose
    DLT-triangulated point has positive depth in BOTH cameras (x_dst = R x_src + t).

    Not in the reference (it uses the trace / t_z heuristics of slam.py:174-181); this is
    the north-star's 'cheirality-based pose selection', spec'd here.  Candidate order:
    (U W Vt, +u3), (U W Vt, -u3), (U W^T Vt, +u3), (U W^T Vt, -u3) after the det fixes.
    Returns (votes int32[4], Rts f64[4,4,4]).
    """
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], dtype=float)
    U, _, Vt = np.linalg.svd(np.asarray(E, dtype=np.float64))
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    cands = []
    for Wk in (W, W.T):
        R = U @ Wk @ Vt
        for sg in (1.0, -1.0):
            cands.append((R, sg * U[:, 2]))
    votes = np.zeros(4, dtype=np.int32)
    Rts = np.zeros((4, 4, 4))
    for c, (R, t) in enumerate(cands):
        Rt = np.eye(4); Rt[:3, :3] = R; Rt[:3, 3] = t
        Rts[c] = Rt
        P1 = np.eye(4)[:3]
        P2 = Rt[:3]
        A = np.empty((len(src), 4, 4))
        A[:, 0] = src[:, 0:1] * P1[2] - P1[0]
        A[:, 1] = src[:, 1:2] * P1[2] - P1[1]
        A[:, 2] = dst[:, 0:1] * P2[2] - P2[0]
        A[:, 3] = dst[:, 1:2] * P2[2] - P2[1]
        X = np.linalg.svd(A)[2][:, 3, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            X = X / X[:, 3:]
        z1 = X[:, 2]
        z2 = (X[:, :3] @ R.T + t)[:, 2]
        votes[c] = int(np.sum((z1 > 0) & (z2 > 0)))
    return votes, Rts


# --------------------------------------------------------------------------- one pair
def process_pair(last_kps, last_des, curr_kps, curr_des, last_pose, K, samples=None,
                 words=None, ratio=RATIO, residual_threshold=RESIDUAL_THRESHOLD, use_cv2=True):
    """The ``else`` branch of ``process_frame`` (slam.py:249-266) on arrays.

    ``samples`` is either an int32[H,8] table or, if None, built from ``words`` u32[H,8] via
    the sync-free mapping (``openwork_b200.synth.samples_from_words``).
    """
    match_kps, idx1, idx2 = match_points(curr_kps, curr_des, last_kps, last_des, ratio, use_cv2)
    if samples is None:
        from openwork_b200.synth import samples_from_words
        samples = samples_from_words(words, len(match_kps))
    E, inl, r = fit_essential_matrix(match_kps, K, samples, residual_threshold)
    ck = curr_kps[idx1][inl]
    lk = last_kps[idx2][inl]
    Rt = extract_Rt(E)
    curr_pose = Rt @ last_pose
    X = triangulate(lk, ck, last_pose, curr_pose, K)
    good = check_points(X)
    return dict(idx1=idx1, idx2=idx2, match_kps=match_kps, E=E, inliers=inl, ransac=r,
                Rt=Rt, curr_pose=curr_pose, points4d=X, good=good)
