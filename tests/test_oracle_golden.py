"""CPU tests: the oracle (oracle/two_view.py) against the committed golden fixtures generated from the
REFERENCE's own functions (tests/golden/make_golden.py), against the live reference when
/root/reference is mounted, and against closed-form / independent checks for the restated RANSAC."""
import os

import numpy as np
import pytest

from openwork_b200 import synth
from oracle import ref_loader, two_view as tv

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name", ["pair_small", "pair_medium"])
def test_matcher_and_ratio_against_reference_fixture(name):
    g = _load(name)
    for knn in (tv.knn2_numpy, tv.knn2_cv2):
        idx, dist = knn(g["curr_des"], g["last_des"])
        assert np.array_equal(idx, g["knn_idx"]) and np.array_equal(dist, g["knn_dist"])
    mk, i1, i2 = tv.match_points(g["curr_kps"], g["curr_des"], g["last_kps"], g["last_des"], use_cv2=False)
    assert np.array_equal(mk, g["match_kps"])
    assert np.array_equal(tv.normalize(g["K"], mk[:, 0]), g["norm_curr"])
    assert np.array_equal(tv.normalize(g["K"], mk[:, 1]), g["norm_last"])


def test_knn_tie_break_fixture():
    g = _load("knn_ties")
    idx, dist = tv.knn2_numpy(g["q"], g["t"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    d = tv.hamming_matrix(g["q"], g["t"])
    assert np.array_equal(np.argsort(d, axis=1, kind="stable")[:, :2], g["idx"])
    # the fixture really contains exact top-2 ties
    assert (g["dist"][:, 0] == g["dist"][:, 1]).sum() > 10


@pytest.mark.parametrize("name", ["pair_small", "pair_medium"])
def test_pose_and_triangulation_against_reference_fixture(name):
    g = _load(name)
    np.testing.assert_allclose(tv.extract_Rt(g["E"]), g["Rt"], atol=1e-12)
    inl = g["ransac_mask"]
    mk, i1, i2 = tv.match_points(g["curr_kps"], g["curr_des"], g["last_kps"], g["last_des"], use_cv2=False)
    X = tv.triangulate(g["last_kps"][i2][inl], g["curr_kps"][i1][inl], g["last_pose"], g["curr_pose"], g["K"])
    np.testing.assert_allclose(X, g["points4d"], rtol=1e-9, atol=1e-9)
    Xb = tv.triangulate_batched(g["last_kps"][i2][inl], g["curr_kps"][i1][inl], g["last_pose"], g["curr_pose"], g["K"])
    np.testing.assert_allclose(Xb, g["points4d"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(tv.check_points(X), g["good"])


def test_extract_rt_fixture():
    g = _load("extract_rt")
    for E, Rt in zip(g["E"], g["Rt"]):
        np.testing.assert_allclose(tv.extract_Rt(E), Rt, atol=1e-12)
        np.testing.assert_allclose(tv.extract_Rt(-3.7 * E), Rt, atol=1e-9)      # invariant to sign / scale


@pytest.mark.parametrize("name", ["pair_small", "pair_medium"])
def test_ransac_restatement_is_deterministic_against_fixture(name):
    """NOT pinned by the reference (scikit-image is not installable here): guards against drift only."""
    g = _load(name)
    assert np.array_equal(synth.samples_from_words(g["words"], len(g["match_kps"])), g["samples"])
    r = tv.ransac_essential(g["norm_last"], g["norm_curr"], g["samples"], 0.005)
    assert np.array_equal(r["counts"], g["ransac_counts"])
    assert r["best"] == int(g["ransac_best"])
    assert np.array_equal(r["mask"], g["ransac_mask"])
    np.testing.assert_allclose(r["E"], g["E"], atol=1e-10)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_equals_live_reference_functions():
    ref = ref_loader.load()
    K = synth.K_640x480
    ref.K = K
    lk, ld, ck, cd = synth.make_pair(600, 4242, forced_ties=25)

    class F:
        pass
    f = F()
    f.curr_kps, f.curr_des, f.last_kps, f.last_des = ck, cd, lk, ld
    mk_ref = np.array(ref.match_points(f))
    mk, i1, i2 = tv.match_points(ck, cd, lk, ld)
    assert np.array_equal(mk, mk_ref)
    assert np.array_equal(f.curr_kps, ck[i1]) and np.array_equal(f.last_kps, lk[i2])
    assert np.array_equal(ref.normalize(K, lk), tv.normalize(K, lk))
    samples = synth.sample_table(len(mk), 64, 1)
    E, inl, _ = tv.fit_essential_matrix(mk, K, samples)
    assert np.array_equal(np.asarray(ref.extract_Rt(E)), tv.extract_Rt(E))
    pose0 = np.eye(4)
    pose1 = tv.extract_Rt(E) @ pose0
    Xr = ref.triangulate(lk[i2][inl], ck[i1][inl], pose0, pose1)
    X = tv.triangulate(lk[i2][inl], ck[i1][inl], pose0, pose1, K)
    assert np.array_equal(X, Xr)
    assert np.array_equal(ref.check_points(Xr), tv.check_points(X))


def test_match_points_error_conventions():
    lk, ld, ck, cd = synth.make_pair(64, 1)
    with pytest.raises(ValueError):                      # `for m, n in matches` with one train row (slam.py:130)
        tv.match_points(ck, cd, lk[:1], ld[:1])
    rnd = synth.random_descriptors(64, 3)
    with pytest.raises(AssertionError):                  # fewer than 8 survivors (slam.py:138)
        tv.match_points(ck, rnd, lk, synth.random_descriptors(64, 4))
    with pytest.raises(ValueError):                      # skimage: min_samples must be < number of samples
        tv.ransac_essential(np.zeros((8, 2)), np.zeros((8, 2)), np.arange(8)[None])


def _closed_form_scene(n, seed, noise=0.0):
    rng = np.random.default_rng(seed)
    X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(4, 20, n)], axis=1)
    R = synth._small_rotation(rng, 0.05)
    t = np.array([0.1, -0.05, -0.6])
    x1 = X[:, :2] / X[:, 2:]
    Xc = X @ R.T + t
    x2 = Xc[:, :2] / Xc[:, 2:] + rng.normal(0, noise, (n, 2))
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    return x1, x2, tx @ R, R, t, X


def test_essential_estimate_on_closed_form_scene():
    """skimage applies the (s,s,0) constraint in HARTLEY-NORMALISED coordinates, so even noise-free data
    is not reproduced exactly (the restatement keeps that behaviour): E is close to [t]x R, not equal."""
    x1, x2, E_true, R, t, X = _closed_form_scene(200, 0)
    E = tv.essential_estimate(x1, x2)
    a, b = E / np.linalg.norm(E), E_true / np.linalg.norm(E_true)
    assert min(np.linalg.norm(a - b), np.linalg.norm(a + b)) < 2e-2
    assert tv.essential_residuals(E, x1, x2).max() < 0.005
    # before the projection the DLT null vector IS exact: check through the fundamental-style residual
    Ts, s = tv.center_and_normalize_points(x1)
    Td, d = tv.center_and_normalize_points(x2)
    A = np.column_stack([d[:, 0] * s[:, 0], d[:, 0] * s[:, 1], d[:, 0], d[:, 1] * s[:, 0], d[:, 1] * s[:, 1],
                         d[:, 1], s[:, 0], s[:, 1], np.ones(len(s))])
    F = Td.T @ np.linalg.svd(A)[2][-1].reshape(3, 3) @ Ts
    a = F / np.linalg.norm(F)
    assert min(np.linalg.norm(a - b), np.linalg.norm(a + b)) < 1e-9
    # projected spectrum is (s, s, 0) in normalised coordinates
    En = np.linalg.inv(Td.T) @ E @ np.linalg.inv(Ts)
    sv = np.linalg.svd(En, compute_uv=False)
    assert abs(sv[0] - sv[1]) < 1e-12 * sv[0] and sv[2] < 1e-12 * sv[0]


def _dlt_fundamental(src, dst):
    """The oracle's 8-point DLT up to (not including) the (s,s,0) step: T_d^T . null(A) . T_s."""
    Ts, s = tv.center_and_normalize_points(src)
    Td, d = tv.center_and_normalize_points(dst)
    A = np.column_stack([d[:, 0] * s[:, 0], d[:, 0] * s[:, 1], d[:, 0], d[:, 1] * s[:, 0], d[:, 1] * s[:, 1],
                         d[:, 1], s[:, 0], s[:, 1], np.ones(len(s))])
    return Td.T @ np.linalg.svd(A)[2][-1].reshape(3, 3) @ Ts


def _unit(F):
    return F / np.linalg.norm(F)


def test_minimal_dlt_null_vector_matches_cv2_8point():
    """Independent (importable) check of essential_estimate's DLT stage -- the part skimage and OpenCV must agree on.
    For exactly 8 correspondences the null vector of the 8x9 system does not depend on the coordinate normalisation,
    so OpenCV's own normalised 8-point algorithm (cv2.findFundamentalMat(FM_8POINT): different normalisation,
    eigen-solve of A^T A instead of an SVD of A) must return the same matrix wherever its rank-2 step changes
    nothing, i.e. on exact correspondences; with noise the two differ only by that step (|sigma_3| of the null
    vector), and both annihilate the 8 epipolar constraints before it."""
    import cv2
    for seed in range(6):
        x1, x2, E_true, *_ = _closed_form_scene(8, 100 + seed)
        F_cv, _ = cv2.findFundamentalMat(x1, x2, cv2.FM_8POINT)
        F = _dlt_fundamental(x1, x2)
        a, b, c = _unit(F), _unit(F_cv), _unit(E_true)
        # oracle DLT == OpenCV (which solves the squared system A^T A: ~1e-6 is its own accuracy) == the true E
        assert min(np.linalg.norm(a - b), np.linalg.norm(a + b)) < 1e-5, seed
        assert min(np.linalg.norm(a - c), np.linalg.norm(a + c)) < 1e-8, seed
    for seed in range(6):
        x1, x2, *_ = _closed_form_scene(8, 200 + seed, noise=1e-4)
        F = _dlt_fundamental(x1, x2)
        F_cv, _ = cv2.findFundamentalMat(x1, x2, cv2.FM_8POINT)
        x1h, x2h = np.column_stack([x1, np.ones(8)]), np.column_stack([x2, np.ones(8)])
        assert np.abs(np.sum(x2h * (x1h @ _unit(F).T), axis=1)).max() < 1e-12       # exact null vector (x'^T F x = 0)
        sv = np.linalg.svd(_unit(F), compute_uv=False)
        a, b = _unit(F), _unit(F_cv)
        assert min(np.linalg.norm(a - b), np.linalg.norm(a + b)) < 20 * sv[2] + 1e-5, (seed, sv)


def test_residual_squared_is_cv2_sampson_distance():
    import cv2
    x1, x2, E_true, *_ = _closed_form_scene(50, 1, noise=0.01)
    r = tv.essential_residuals(E_true, x1, x2)
    sd = np.array([cv2.sampsonDistance(np.array([*x1[i], 1.0]), np.array([*x2[i], 1.0]), E_true) for i in range(50)])
    np.testing.assert_allclose(r ** 2, sd, rtol=1e-10, atol=1e-18)


def test_ransac_selection_rule_and_refit_mask_semantics():
    x1, x2, E_true, *_ = _closed_form_scene(120, 2, noise=0.002)
    rng = np.random.default_rng(0)
    x2[:30] += rng.normal(0, 0.2, (30, 2))               # gross outliers
    samples = synth.sample_table(120, 50, 3)
    r = tv.ransac_essential(x1, x2, samples, 0.005, return_models=True)
    # best = max count, then min sse, then earliest
    order = sorted(range(50), key=lambda h: (-r["counts"][h], r["sse"][h], h))
    assert r["best"] == order[0]
    # mask is the PRE-refit model's, E is the refit on that mask
    res = tv.essential_residuals(r["models"][r["best"]], x1, x2)
    assert np.array_equal(r["mask"], res < 0.005)
    np.testing.assert_allclose(r["E"], tv.essential_estimate(x1[r["mask"]], x2[r["mask"]]), atol=1e-12)
    # duplicated hypothesis rows tie exactly: the earliest wins
    samples2 = np.concatenate([samples[r["best"]][None], samples])
    assert tv.ransac_essential(x1, x2, samples2, 0.005)["best"] == 0


def test_cheirality_votes_pick_the_true_motion():
    x1, x2, E_true, R, t, X = _closed_form_scene(150, 5)
    votes, Rts = tv.cheirality_votes(E_true, x1, x2)
    c = int(np.argmax(votes))
    assert votes[c] == 150 and sorted(votes)[-2] == 0
    np.testing.assert_allclose(Rts[c][:3, :3], R, atol=1e-9)
    np.testing.assert_allclose(Rts[c][:3, 3], t / np.linalg.norm(t), atol=1e-9)


def test_triangulate_reproduces_ground_truth_points():
    x1, x2, E_true, R, t, X = _closed_form_scene(100, 6)
    K = synth.K_640x480
    Kf = K.astype(float)
    p1 = x1 * Kf[0, 0] + Kf[:2, 2]
    p2 = x2 * Kf[0, 0] + Kf[:2, 2]
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    Xh = tv.triangulate(p1, p2, np.eye(4), np.linalg.inv(T), K)   # pose = camera->world = inv(world->camera)
    np.testing.assert_allclose(Xh[:, :3], X, rtol=1e-8, atol=1e-8)


def test_samples_from_words_properties():
    rng = np.random.default_rng(0)
    for M in (8, 9, 17, 1000, 65536):
        w = rng.integers(0, 2 ** 32, (500, 8), dtype=np.uint64).astype(np.uint32)
        if M == 9:
            w[:50] = 0                                    # worst case: every slot maps to index 0
        s = synth.samples_from_words(w, M)
        assert s.min() >= 0 and s.max() < M
        assert all(len(set(row)) == 8 for row in s.tolist())
        first = ((w[:, 0].astype(np.uint64) * np.uint64(M)) >> np.uint64(32)).astype(np.int32)
        assert np.array_equal(s[:, 0], first)
    assert np.array_equal(synth.samples_from_words(np.zeros((1, 8), np.uint32), 20)[0], np.arange(8))


def test_ratio_test_equals_integer_form():
    """d1 < 0.75*d2 (float, slam.py:131) <=> 4*d1 < 3*d2 for integer distances in [0,256]."""
    d1, d2 = np.meshgrid(np.arange(257), np.arange(257), indexing="ij")
    f = d1.astype(np.float32) < 0.75 * d2.astype(np.float32)
    assert np.array_equal(f, 4 * d1 < 3 * d2)
    idx = np.stack([np.zeros(257 * 257, np.int32), np.ones(257 * 257, np.int32)], axis=1)
    dist = np.stack([d1.ravel(), d2.ravel()], axis=1).astype(np.int32)
    q, _ = tv.ratio_filter(idx, dist)
    assert np.array_equal(q, np.flatnonzero((4 * d1 < 3 * d2).ravel()))


def test_knn_numpy_matches_cv2_on_ragged_shapes():
    rng = np.random.default_rng(9)
    for nq, nt in ((1, 2), (3, 1), (65, 130), (257, 31)):
        q = rng.integers(0, 4, (nq, 32), dtype=np.uint8)          # low entropy -> many ties
        t = rng.integers(0, 4, (nt, 32), dtype=np.uint8)
        a, b = tv.knn2_numpy(q, t)
        c, d = tv.knn2_cv2(q, t)
        assert np.array_equal(a, c) and np.array_equal(b, d)
