"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded
inputs.  Bars (BASELINE.json north_star): matcher indices/distances bit-exact; per-hypothesis inlier
counts bit-exact except points within 1e-5 of the threshold; E within 1e-4 Frobenius after
normalisation + sign alignment; triangulated points within 1e-4 relative."""
import numpy as np
import pytest
import torch

from openwork_b200 import synth
from oracle import two_view as tv

pytestmark = pytest.mark.gpu

THR = 0.005
BAND = 1e-5          # north star: counts exact except points within 1e-5 of the residual threshold
E_TOL = 1e-4         # Frobenius, after normalisation and sign alignment
X_TOL = 1e-4         # relative error of triangulated points
K = synth.K_640x480


@pytest.fixture(scope="module")
def ops():
    from openwork_b200 import ops as _ops
    assert torch.cuda.is_available(), "GPU tests need a B200"
    _ops.context(0)   # raises loudly if the extension is missing
    return _ops


def _fro_diff(A, B):
    A = A / np.linalg.norm(A)
    B = B / np.linalg.norm(B)
    return min(np.linalg.norm(A - B), np.linalg.norm(A + B))


# ------------------------------------------------------------------------------ matcher
@pytest.mark.parametrize("nq,nt", [(1, 2), (5, 1), (7, 3), (33, 129), (500, 700), (2048, 2048), (1000, 4099),
                                   (4100, 257)])
@pytest.mark.parametrize("variant", [-1, 11, 16, 17])
def test_knn2_bit_exact(ops, nq, nt, variant):
    ops.context(0).set_matcher_config(4, variant, 6)
    rng = np.random.default_rng(nq * 10007 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    if nt > 40:   # forced exact ties: duplicated train rows and queries equal to train rows
        d = rng.choice(nt, nt // 5, replace=False)
        t[d] = t[rng.choice(nt, nt // 5, replace=False)]
        qq = rng.choice(nq, max(1, nq // 7), replace=False)
        q[qq] = t[rng.integers(0, nt, len(qq))]
    idx, dist = ops.knn2_hamming256(q, t)
    ri, rd = tv.knn2_cv2(q, t)
    assert np.array_equal(dist.cpu().numpy(), rd)
    assert np.array_equal(idx.cpu().numpy(), ri)
    ops.context(0).set_matcher_config()
    ni, nd = tv.knn2_numpy(q, t)
    assert np.array_equal(idx.cpu().numpy(), ni) and np.array_equal(dist.cpu().numpy(), nd)


@pytest.mark.parametrize("variant", [11, 16, 17])
def test_knn2_all_equal_descriptors(ops, variant):
    q = np.full((300, 32), 0xA5, dtype=np.uint8)
    t = np.full((513, 32), 0xA5, dtype=np.uint8)
    ops.context(0).set_matcher_config(4, variant, 6)
    try:
        idx, dist, keep = ops.knn2_hamming256(q, t, return_keep=True)
    finally:
        ops.context(0).set_matcher_config()
    assert (idx.cpu().numpy() == np.array([0, 1])).all() and (dist.cpu().numpy() == 0).all()
    assert not keep.any()           # 0 < 0.75*0 is False


@pytest.mark.parametrize("variant", [11, 16, 17])
def test_knn2_extreme_distances(ops, variant):
    """Distances 0, 1, 255 and 256 (complemented descriptors) and 300 identical train rows: key packing limits."""
    rng = np.random.default_rng(256)
    r = rng.integers(0, 256, 32, dtype=np.uint8)
    t = np.tile(r, (300, 1))
    one = r.copy(); one[5] ^= 0x10
    q = np.stack([~r, r, one, ~one] * 70).astype(np.uint8)
    ops.context(0).set_matcher_config(4, variant, 6)
    try:
        idx, dist = ops.knn2_hamming256(q, t)
    finally:
        ops.context(0).set_matcher_config()
    ri, rd = tv.knn2_numpy(q, t)
    assert set(rd[:, 0].tolist()) == {0, 1, 255, 256}
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(dist.cpu().numpy(), rd)


@pytest.mark.parametrize("variant", [11, 16, 17])
def test_knn2_batched_with_empty_frames(ops, variant):
    """Frames with zero rows on either side of a pair (the persistent tensor-core kernel skips such items)."""
    F, N = 6, 600
    rng = np.random.default_rng(77)
    des = rng.integers(0, 256, (F, N, 32), dtype=np.uint8)
    cnt = np.array([600, 0, 1, 300, 0, 599], dtype=np.int32)
    pq = np.array([0, 1, 2, 3, 5, 4, 0, 3], dtype=np.int32)
    pt = np.array([1, 0, 0, 2, 4, 5, 5, 0], dtype=np.int32)
    ops.context(0).set_matcher_config(4, variant, 6)
    try:
        idx, dist, keep = ops.knn2_hamming256_batched(des, pq, pt, frame_count=cnt)
    finally:
        ops.context(0).set_matcher_config()
    idx, dist, keep = idx.cpu().numpy(), dist.cpu().numpy(), keep.cpu().numpy()
    for b in range(len(pq)):
        nq, nt = cnt[pq[b]], cnt[pt[b]]
        if nq == 0:
            continue
        if nt == 0:
            assert (idx[b, :nq] == -1).all() and (dist[b, :nq] == -1).all() and not keep[b, :nq].any()
            continue
        ri, rd = tv.knn2_numpy(des[pq[b], :nq], des[pt[b], :nt])
        assert np.array_equal(idx[b, :nq], ri) and np.array_equal(dist[b, :nq], rd), b


def test_knn2_empty(ops):
    q = np.zeros((0, 32), dtype=np.uint8)
    t = np.zeros((10, 32), dtype=np.uint8)
    idx, dist = ops.knn2_hamming256(q, t)
    assert idx.shape == (0, 2)
    idx, dist = ops.knn2_hamming256(t, q)
    assert (idx.cpu().numpy() == -1).all() and (dist.cpu().numpy() == -1).all()


@pytest.mark.parametrize("r,variant,ctas", [(2, 0, 6), (4, 0, 1), (4, 1, 6), (4, 2, 6), (8, 0, 4), (8, 1, 8),
                                            (2, 1, 16), (2, 2, 3), (8, 2, 2), (4, 3, 6), (2, 3, 2), (8, 3, 8),
                                            (4, 5, 6), (2, 5, 1), (8, 5, 8), (4, 7, 6), (2, 7, 16), (8, 7, 3),
                                            (4, 9, 6), (2, 9, 3), (8, 9, 8), (4, 11, 6), (2, 11, 16), (8, 11, 1),
                                            (4, 16, 6), (4, 17, 6)])
def test_knn2_every_config_is_bit_identical(ops, r, variant, ctas):
    q, t = synth.correlated_descriptors(3000, 2500, seed=r * 100 + variant * 10 + ctas)
    t[100:200] = t[300:400]
    ctx = ops.context(0)
    try:
        ctx.set_matcher_config(r, variant, ctas)
        idx, dist, keep = ops.knn2_hamming256(q, t, return_keep=True)
    finally:
        ctx.set_matcher_config()
    ri, rd = tv.knn2_cv2(q, t)
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(dist.cpu().numpy(), rd)
    qi, ti = tv.ratio_filter(ri, rd)
    assert np.array_equal(np.flatnonzero(keep.cpu().numpy()), qi)


def test_knn2_8k_matches_cv2(ops):
    q, t = synth.correlated_descriptors(8192, 8192, seed=3)
    idx, dist = ops.knn2_hamming256(q, t)
    ri, rd = tv.knn2_cv2(q, t)
    assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(dist.cpu().numpy(), rd)


def test_knn2_tensor_l2_resident_sub_batches(ops):
    """L2-resident mode of the tensor-core matcher: several sub-batch launches, default and explicit pairing."""
    F, N = 9, 520
    rng = np.random.default_rng(40)
    des = rng.integers(0, 256, (F, N, 32), dtype=np.uint8)
    cnt = np.array([520, 400, 7, 519, 300, 520, 64, 257, 511], dtype=np.int32)
    pq = np.array([1, 2, 3, 4, 5, 6, 7, 8, 0, 4], dtype=np.int32)
    pt = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 4], dtype=np.int32)
    ctx = ops.context(0)
    ctx.set_matcher_config(4, 16, 6)
    ctx.set_matcher_l2_budget(3 * 768 * 256)       # room for 3 pairs (768 padded rows x 256 B each) -> 4 launches
    import ctypes as C
    import torch
    qd = torch.from_numpy(des).cuda()
    td = torch.from_numpy(np.roll(des, 1, axis=0).copy()).cuda()
    qc = torch.from_numpy(cnt).cuda()
    tc = torch.from_numpy(np.roll(cnt, 1).copy()).cuda()
    i2 = torch.full((F, N, 2), -7, dtype=torch.int32, device="cuda")
    d2 = torch.full((F, N, 2), -7, dtype=torch.int32, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    try:
        idx, dist, keep = ops.knn2_hamming256_batched(des, pq, pt, frame_count=cnt)
        # default pairing (pair b = frame b of either pool), per-frame counts, straight through the C ABI
        ops.check(ctx.lib.tvg_knn2_hamming256(ctx.handle, P(qd), P(td), F, N, N, N, None, None, P(qc), P(tc), 0.75,
                                              P(i2), P(d2), None, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                  "tvg_knn2_hamming256")
    finally:
        ctx.set_matcher_l2_budget(0)
        ctx.set_matcher_config()
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    for b in range(len(pq)):
        nq, nt = cnt[pq[b]], cnt[pt[b]]
        ri, rd = tv.knn2_numpy(des[pq[b], :nq], des[pt[b], :nt])
        assert np.array_equal(idx[b, :nq], ri) and np.array_equal(dist[b, :nq], rd), b
    i2, d2 = i2.cpu().numpy(), d2.cpu().numpy()
    for b in range(F):
        nq, nt = cnt[b], cnt[(b - 1) % F]
        ri, rd = tv.knn2_numpy(des[b, :nq], des[(b - 1) % F, :nt])
        assert np.array_equal(i2[b, :nq], ri) and np.array_equal(d2[b, :nq], rd), b
        assert (i2[b, nq:] == -7).all()


@pytest.mark.parametrize("variant", [11, 16, 17])
def test_knn2_batched_ragged(ops, variant):
    F, N = 5, 700
    rng = np.random.default_rng(5)
    des = rng.integers(0, 256, (F, N, 32), dtype=np.uint8)
    cnt = np.array([700, 650, 3, 512, 699], dtype=np.int32)
    pq = np.array([1, 2, 3, 4, 0, 2], dtype=np.int32)
    pt = np.array([0, 1, 2, 3, 4, 2], dtype=np.int32)
    ops.context(0).set_matcher_config(4, variant, 6)
    try:
        idx, dist, keep = ops.knn2_hamming256_batched(des, pq, pt, frame_count=cnt)
    finally:
        ops.context(0).set_matcher_config()
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    for b in range(len(pq)):
        nq, nt = cnt[pq[b]], cnt[pt[b]]
        ri, rd = tv.knn2_numpy(des[pq[b], :nq], des[pt[b], :nt])
        assert np.array_equal(idx[b, :nq], ri) and np.array_equal(dist[b, :nq], rd)


# ------------------------------------------------------------------------------ ratio + gather
def test_ratio_compact_matches_reference_loop(ops):
    lk, ld, ck, cd = synth.make_pair(1500, 77, forced_ties=40)
    idx, dist, keep = ops.knn2_hamming256(cd, ld, return_keep=True)
    out = ops.ratio_compact(idx, keep, ck, lk, K)
    mk, i1, i2 = tv.match_points(ck, cd, lk, ld)
    assert out["M"] == len(i1)
    assert np.array_equal(out["qidx"].cpu().numpy(), i1) and np.array_equal(out["tidx"].cpu().numpy(), i2)
    np.testing.assert_allclose(out["dst"].cpu().numpy(), tv.normalize(K, mk[:, 0]), rtol=0, atol=1e-14)
    np.testing.assert_allclose(out["src"].cpu().numpy(), tv.normalize(K, mk[:, 1]), rtol=0, atol=1e-14)


# ------------------------------------------------------------------------------ RANSAC
def _pair_correspondences(n, seed):
    lk, ld, ck, cd = synth.make_pair(n, seed)
    mk, i1, i2 = tv.match_points(ck, cd, lk, ld)
    return tv.normalize(K, mk[:, 1]), tv.normalize(K, mk[:, 0]), (lk, ck, i1, i2)


@pytest.mark.parametrize("n,H,seed,use_words", [(2048, 512, 11, False), (700, 300, 12, True), (2048, 1024, 13, True)])
def test_ransac_counts_best_mask_E(ops, n, H, seed, use_words):
    src, dst, _ = _pair_correspondences(n, seed)
    M = len(src)
    if use_words:
        words = synth.sample_words(H, seed)
        samples = synth.samples_from_words(words, M)
        g = ops.ransac_essential(src, dst, words=words, residual_threshold=THR, return_models=True)
        assert np.array_equal(g["samples"].cpu().numpy(), samples)      # device mapping == oracle mapping
    else:
        samples = synth.sample_table(M, H, seed)
        g = ops.ransac_essential(src, dst, samples=samples, residual_threshold=THR, return_models=True)
    ref = tv.ransac_essential(src, dst, samples, THR, return_models=True)
    cond = tv.sample_conditioning(src, dst, samples)
    ok = cond > 1e-7                      # near-degenerate 8-samples: LAPACK's null vector is arbitrary
    counts = g["counts"].cpu().numpy()
    models = g["models"].cpu().numpy()
    n_checked = n_exact = 0
    for h in range(H):
        if not ok[h] or ref["counts"][h] < 0:
            continue
        r = tv.essential_residuals(ref["models"][h], src, dst)
        band = np.abs(r - THR) < BAND
        lo = ref["counts"][h] - int(np.sum(band & (r < THR)))
        hi = ref["counts"][h] + int(np.sum(band & (r >= THR)))
        assert lo <= counts[h] <= hi, (h, counts[h], ref["counts"][h], lo, hi)
        n_exact += counts[h] == ref["counts"][h]
        n_checked += 1
        assert _fro_diff(models[h], ref["models"][h]) < E_TOL
    assert n_checked >= 0.95 * H
    assert n_exact >= 0.98 * n_checked
    assert g["best"] == ref["best"]
    assert g["n_inl"] == int(ref["mask"].sum())
    rb = tv.essential_residuals(ref["E_best"], src, dst)
    differs = g["mask"].cpu().numpy() != ref["mask"]
    assert not (differs & (np.abs(rb - THR) >= BAND)).any()
    assert _fro_diff(g["E"].cpu().numpy(), ref["E"]) < E_TOL
    assert _fro_diff(g["E_best"].cpu().numpy(), ref["E_best"]) < E_TOL
    np.testing.assert_allclose(g["Rt"].cpu().numpy(), tv.extract_Rt(ref["E"]), atol=1e-6)


def test_ransac_packed_scoring_form_counts_identical(ops, tmp_path):
    """The opt-in FFMA2 form of the scoring kernel (pairs of hypotheses in 64-bit registers, TVG_SCORE_PACKED=1: read
    once per process, hence the child process) performs the same fp32 operations: every count, the winner and the mask
    are bit-identical to the default scalar form."""
    import os
    import subprocess
    import sys
    src, dst, _ = _pair_correspondences(2048, 21)
    words = synth.sample_words(1024, 21)
    g = ops.ransac_essential(src, dst, words=words, residual_threshold=THR, return_models=True)
    np.savez(tmp_path / "in.npz", src=src, dst=dst, words=words)
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from openwork_b200 import ops\n"
        "d = np.load(%r)\n"
        "g = ops.ransac_essential(d['src'], d['dst'], words=d['words'], residual_threshold=%r, return_models=True)\n"
        "np.savez(%r, counts=g['counts'].cpu().numpy(), mask=g['mask'].cpu().numpy(), best=g['best'], E=g['E'].cpu().numpy())\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "in.npz"), THR, str(tmp_path / "out.npz"))
    env = dict(os.environ, TVG_SCORE_PACKED="1")
    subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=600)
    o = np.load(tmp_path / "out.npz")
    assert np.array_equal(o["counts"], g["counts"].cpu().numpy())
    assert int(o["best"]) == g["best"]
    assert np.array_equal(o["mask"], g["mask"].cpu().numpy())
    assert np.array_equal(o["E"], g["E"].cpu().numpy())


def test_ransac_noise_free_scene_all_inliers(ops):
    """Closed form: noise-free correspondences of E = [t]x R -> every hypothesis explains every point."""
    rng = np.random.default_rng(0)
    n = 400
    X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(4, 20, n)], axis=1)
    R = synth._small_rotation(rng, 0.02)
    t = np.array([0.05, -0.02, -0.5])
    x1 = X[:, :2] / X[:, 2:]
    Xc = X @ R.T + t
    x2 = Xc[:, :2] / Xc[:, 2:]
    samples = synth.sample_table(n, 64, 1)
    g = ops.ransac_essential(x1, x2, samples=samples)
    counts = g["counts"].cpu().numpy()
    ref = tv.ransac_essential(x1, x2, samples)
    cond = tv.sample_conditioning(x1, x2, samples)
    ok = cond > 1e-6
    # skimage projects to (s,s,0) in Hartley-normalised coordinates, so a few models miss a few points even
    # without noise -- the GPU must reproduce exactly that
    assert np.array_equal(counts[ok], ref["counts"][ok])
    assert (counts == n).sum() >= 48
    assert g["n_cand"] >= (counts == n).sum()     # every full-count hypothesis ties -> re-scored in fp64
    assert g["best"] == ref["best"] and g["n_inl"] == n
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    assert _fro_diff(g["E"].cpu().numpy(), tx @ R) < 2e-2
    assert _fro_diff(g["E"].cpu().numpy(), ref["E"]) < E_TOL


def test_ransac_too_few_correspondences_raises(ops):
    src = np.random.default_rng(0).normal(size=(8, 2))
    with pytest.raises(ValueError):
        ops.ransac_essential(src, src, samples=np.arange(8, dtype=np.int32)[None])


# ------------------------------------------------------------------------------ pose + triangulation
def test_extract_rt_matches_reference_conventions(ops):
    rng = np.random.default_rng(4)
    Es, refs = [], []
    for _ in range(64):
        R = synth._small_rotation(rng, 0.3)
        t = rng.normal(size=3)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        E = (tx @ R) * rng.uniform(0.1, 10) * rng.choice([-1, 1])
        Es.append(E)
        refs.append(tv.extract_Rt(E))
    got = ops.extract_rt(np.stack(Es)).cpu().numpy()
    np.testing.assert_allclose(got, np.stack(refs), atol=1e-9)


def test_extract_rt_fast_and_fallback_paths(ops):
    """extract_Rt takes U, V from a non-iterative rank-two SVD when E is an essential matrix to 1e-10 and from the
    Jacobi SVD otherwise: 2000 exact essential matrices (fast path), the same matrices pushed off rank two by 1e-13
    (still the fast path), by 1e-7 (fallback) and with well-separated, unequal singular values (fast path, the 2x2
    rotation is not trivial) -- all against numpy's SVD through the oracle."""
    rng = np.random.default_rng(40)
    Es = []
    for i in range(2000):
        R = synth._small_rotation(rng, 0.5)
        t = rng.normal(size=3)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        E = (tx @ R) * 10.0 ** rng.uniform(-3, 3) * rng.choice([-1, 1])
        kind = i % 4
        if kind == 1:
            E = E + 1e-13 * np.linalg.norm(E) * rng.normal(size=(3, 3))
        elif kind == 2:
            E = E + 1e-7 * np.linalg.norm(E) * rng.normal(size=(3, 3))
        elif kind == 3:      # T_d^T E T_s as after the refit's de-normalisation: rank two, s0 != s1
            Ts = np.array([[rng.uniform(0.5, 3), 0, rng.normal()], [0, rng.uniform(0.5, 3), rng.normal()], [0, 0, 1]])
            E = Ts.T @ E @ np.array([[rng.uniform(0.5, 3), 0, rng.normal()], [0, rng.uniform(0.5, 3), rng.normal()], [0, 0, 1]])
        Es.append(E)
    Es = np.stack(Es)
    refs = np.stack([tv.extract_Rt(E) for E in Es])
    got = ops.extract_rt(Es).cpu().numpy()
    # the trace rule is discontinuous: leave out the (rare) matrices whose two candidate rotations both have a trace
    # within 1e-6 of zero
    U, _, Vt = np.linalg.svd(Es)
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], dtype=float)
    du = np.sign(np.linalg.det(U))[:, None, None]
    dv = np.sign(np.linalg.det(Vt))[:, None, None]
    tra = np.trace((U * du) @ W @ (Vt * dv), axis1=1, axis2=2)
    trb = np.trace((U * du) @ W.T @ (Vt * dv), axis1=1, axis2=2)
    # ... and those where BOTH candidates have a positive trace: there the reference's answer is whichever sign
    # LAPACK happened to give the singular vectors (no camera motion between consecutive frames looks like that)
    ok = (np.abs(tra) > 1e-6) & (np.minimum(tra, trb) < 0)
    assert ok.mean() > 0.9
    err = np.abs(got - refs).reshape(len(Es), -1).max(axis=1)
    tol = np.where(np.arange(len(Es)) % 4 == 2, 1e-7, 1e-9)
    assert (err[ok] < tol[ok]).all(), (int(np.argmax(err * ok)), float((err * ok).max()))


def test_triangulate_matches_reference(ops):
    src, dst, (lk, ck, i1, i2) = _pair_correspondences(2048, 21)
    samples = synth.sample_table(len(src), 256, 3)
    ref = tv.ransac_essential(src, dst, samples, THR)
    inl = ref["mask"]
    Rt = tv.extract_Rt(ref["E"])
    pose0 = np.eye(4)
    pose0[:3, 3] = [0.3, -0.2, 1.5]
    pose1 = Rt @ pose0
    p1, p2 = lk[i2][inl], ck[i1][inl]
    Xr = tv.triangulate(p1, p2, pose0, pose1, K)
    X, good = ops.triangulate(p1, p2, pose0, pose1, K, return_good=True)
    X = X.cpu().numpy()
    rel = np.linalg.norm(X[:, :3] - Xr[:, :3], axis=1) / np.maximum(np.linalg.norm(Xr[:, :3], axis=1), 1e-12)
    assert rel.max() < X_TOL, rel.max()
    assert (X[:, 3] == 1.0).all()
    gr = tv.check_points(Xr)
    safe = np.abs(Xr[:, 2]) > 1e-4 * np.linalg.norm(Xr[:, :3], axis=1)
    assert np.array_equal(good.cpu().numpy()[safe], gr[safe])


def test_triangulate_points_without_parallax_take_the_jacobi_path(ops):
    """The fast solve (adjugate + repeated squaring) hands points it cannot separate to the Jacobi SVD kernel: with
    identical poses and identical observations the 4x4 system has rank 2 (no unique answer, also for LAPACK), so
    only A X = 0 can be asked of those; ordinary points in the same call must still equal the reference's."""
    src, dst, (lk, ck, i1, i2) = _pair_correspondences(1024, 23)
    samples = synth.sample_table(len(src), 128, 3)
    ref = tv.ransac_essential(src, dst, samples, THR)
    inl = ref["mask"]
    pose0 = np.eye(4)
    pose1 = tv.extract_Rt(ref["E"]) @ pose0
    p1, p2 = lk[i2][inl], ck[i1][inl]
    Xr = tv.triangulate(p1, p2, pose0, pose1, K)
    X = ops.triangulate(p1, p2, pose0, pose1, K).cpu().numpy()
    rel = np.linalg.norm(X[:, :3] - Xr[:, :3], axis=1) / np.maximum(np.linalg.norm(Xr[:, :3], axis=1), 1e-12)
    assert rel.max() < X_TOL
    # the same observations in both views, same pose: rank-2 systems
    Xd = ops.triangulate(p1, p1, pose0, pose0, K).cpu().numpy()
    n1 = tv.normalize(K, p1)
    P = np.linalg.inv(pose0)
    fin = np.isfinite(Xd).all(axis=1)
    assert fin.any()          # (a null vector with w = 0 divides to inf, exactly like the reference's X /= X[3])
    for x, (u, v) in zip(Xd[fin][:200], n1[fin][:200]):
        A = np.stack([u * P[2] - P[0], v * P[2] - P[1], u * P[2] - P[0], v * P[2] - P[1]])
        assert np.linalg.norm(A @ x) < 1e-9 * max(np.linalg.norm(x), 1.0)
    # and the next ordinary call is unaffected (the fallback counter was reset)
    X2 = ops.triangulate(p1, p2, pose0, pose1, K).cpu().numpy()
    assert np.array_equal(X2, X)


def test_cheirality_votes_match_oracle(ops):
    src, dst, _ = _pair_correspondences(1200, 31)
    samples = synth.sample_table(len(src), 128, 3)
    ref = tv.ransac_essential(src, dst, samples, THR)
    votes, Rts, choice = ops.cheirality(ref["E"], src, dst, mask=ref["mask"])
    rv, rRts = tv.cheirality_votes(ref["E"], src[ref["mask"]], dst[ref["mask"]])
    # candidate ORDER depends on the SVD's sign choices (LAPACK vs Jacobi); the SET and the winner do not
    v = votes.cpu().numpy()
    assert np.abs(np.sort(v) - np.sort(rv)).max() <= 2      # points at |z| ~ 0 may flip
    assert v[choice] == v.max()
    np.testing.assert_allclose(Rts.cpu().numpy()[choice], rRts[int(np.argmax(rv))], atol=1e-8)


# ------------------------------------------------------------------------------ whole pairs
def _check_sequence(ops, out, frames, words, H, seq_pairs, pose0=None):
    pose = np.eye(4) if pose0 is None else pose0
    for b, (fl, fc) in enumerate(seq_pairs):
        (lk, ld), (ck, cd) = frames[fl], frames[fc]
        ref = tv.process_pair(lk, ld, ck, cd, pose, K, words=words, residual_threshold=THR)
        M = len(ref["idx1"])
        assert int(out["m_count"][b]) == M
        assert np.array_equal(out["qidx"][b, :M], ref["idx1"]) and np.array_equal(out["tidx"][b, :M], ref["idx2"])
        assert int(out["best"][b]) == ref["ransac"]["best"]
        n_in = int(ref["inliers"].sum())
        assert int(out["n_inl"][b]) == n_in
        assert np.array_equal(out["mask"][b, :M].astype(bool), ref["inliers"])
        assert _fro_diff(out["E"][b], ref["E"]) < E_TOL
        np.testing.assert_allclose(out["Rt"][b], ref["Rt"], atol=1e-6)
        np.testing.assert_allclose(out["poses"][b + 1], ref["curr_pose"], atol=1e-5)
        X, Xr = out["points4d"][b, :n_in], ref["points4d"]
        rel = np.linalg.norm(X[:, :3] - Xr[:, :3], axis=1) / np.maximum(np.linalg.norm(Xr[:, :3], axis=1), 1e-12)
        # the GPU triangulates with ITS pose chain; pose differences ~1e-9 feed through
        assert np.median(rel) < X_TOL and (rel < 1e-3).mean() > 0.99
        pose = ref["curr_pose"]


ALL_OUT = ("idx", "dist", "qidx", "tidx", "m_count", "E", "E_best", "best", "counts", "mask", "n_inl", "inl_idx",
           "Rt", "poses", "points4d", "good", "status")


def test_pairs_device_sequence_matches_oracle(ops):
    F, N, H = 5, 1024, 256
    frames, _ = synth.make_sequence(F, N, seed=99)
    des = np.stack([f[1] for f in frames])
    kps = np.stack([f[0] for f in frames])
    words = synth.sample_words(H, 7)
    out = ops.process_pairs_device(torch.from_numpy(des).cuda(), torch.from_numpy(kps).cuda(), K, words=words,
                                   residual_threshold=THR, outputs=ALL_OUT)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    assert (out["status"] == 0).all()
    _check_sequence(ops, out, frames, words, H, [(i, i + 1) for i in range(F - 1)])


@pytest.mark.parametrize("chunk,seq_off", [(128, None), (2, None), (1, None), (2, [0, 3, 4, 7]), (3, [0, 1, 7])])
def test_pairs_host_entry_matches_device_entry(ops, chunk, seq_off):
    """tvg_pairs_host (chunked copy/compute pipeline, sequences continuing across chunks) == device entry."""
    npairs, N, H = 7, 512, 128
    nseq = 1 if seq_off is None else len(seq_off) - 1
    F = npairs + nseq
    frames, _ = synth.make_sequence(F, N, seed=5)
    des = np.stack([f[1] for f in frames])
    kps = np.stack([f[0] for f in frames])
    words = synth.sample_words(H, 8)
    so = None if seq_off is None else np.array(seq_off, dtype=np.int32)
    outs = ("m_count", "E", "n_inl", "Rt", "poses", "points4d", "good", "status", "best", "mask", "qidx")
    dev = ops.process_pairs_device(torch.from_numpy(des).cuda(), torch.from_numpy(kps).cuda(), K, words=words,
                                   seq_off=so, outputs=outs)
    ctx = ops.context(0)
    try:
        ctx.set_host_chunk(chunk)
        host, h2d, d2h = ops.process_pairs_host(des, kps, K, words=words, seq_off=so, outputs=outs)
    finally:
        ctx.set_host_chunk(256)
    assert h2d == des.nbytes + kps.nbytes + words.nbytes + 4 * (nseq + 1)
    assert d2h == sum(host[k].numel() * host[k].element_size() for k in outs)
    assert (host["status"].numpy() == 0).all()
    for k in outs:
        a, b = dev[k].cpu().numpy(), host[k].numpy()
        if k in ("points4d", "good"):
            for i in range(npairs):
                n = int(host["n_inl"][i])
                assert np.array_equal(a[i, :n], b[i, :n]), (k, i)
        elif k in ("mask", "qidx"):
            for i in range(npairs):
                n = int(host["m_count"][i])
                assert np.array_equal(a[i, :n], b[i, :n]), (k, i)
        else:
            assert np.array_equal(a, b, equal_nan=True), k


def test_pairs_host_explicit_pair_lists(ops):
    """Arbitrary (curr, last) frame lists, each pair its own sequence, chunked."""
    N, H = 400, 64
    frames, _ = synth.make_sequence(6, N, seed=17)
    des = np.stack([f[1] for f in frames])
    kps = np.stack([f[0] for f in frames])
    words = synth.sample_words(H, 3)
    pq = np.array([1, 3, 5, 2, 4], dtype=np.int32)
    pt = np.array([0, 2, 4, 1, 3], dtype=np.int32)
    so = np.arange(6, dtype=np.int32)
    outs = ("m_count", "E", "n_inl", "Rt", "poses", "status", "best")
    dev = ops.process_pairs_device(torch.from_numpy(des).cuda(), torch.from_numpy(kps).cuda(), K, words=words,
                                   pair_q=pq, pair_t=pt, seq_off=so, outputs=outs)
    ctx = ops.context(0)
    try:
        ctx.set_host_chunk(2)
        host, h2d, d2h = ops.process_pairs_host(des, kps, K, words=words, pair_q=pq, pair_t=pt, seq_off=so,
                                                outputs=outs)
    finally:
        ctx.set_host_chunk(256)
    for k in outs:
        assert np.array_equal(dev[k].cpu().numpy(), host[k].numpy(), equal_nan=True), k
    for b in range(5):
        (lk, ld), (ck, cd) = frames[pt[b]], frames[pq[b]]
        ref = tv.process_pair(lk, ld, ck, cd, np.eye(4), K, words=words)
        assert int(host["best"][b]) == ref["ransac"]["best"] and int(host["n_inl"][b]) == int(ref["inliers"].sum())


def test_pairs_two_sequences(ops):
    N, H = 600, 128
    fa, _ = synth.make_sequence(3, N, seed=1)
    fb, _ = synth.make_sequence(4, N, seed=2)
    frames = fa + fb
    des = np.stack([f[1] for f in frames])
    kps = np.stack([f[0] for f in frames])
    words = synth.sample_words(H, 9)
    seq_off = np.array([0, 2, 5], dtype=np.int32)
    out = ops.process_pairs_device(torch.from_numpy(des).cuda(), torch.from_numpy(kps).cuda(), K, words=words,
                                   seq_off=seq_off, outputs=ALL_OUT)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    # sequence 0: pairs 0,1 use frames (0,1),(1,2), poses rows 0..2 ; sequence 1: pairs 2..4, frames 3..6, poses rows 3..6
    o0 = {k: (v[0:3] if k == "poses" else v[0:2]) for k, v in out.items()}
    o1 = {k: (v[3:7] if k == "poses" else v[2:5]) for k, v in out.items()}
    _check_sequence(ops, o0, frames, words, H, [(0, 1), (1, 2)])
    _check_sequence(ops, o1, frames, words, H, [(3, 4), (4, 5), (5, 6)])


# ------------------------------------------------------------------------------ BASELINE.json configs
def test_config_c1_reference_sequence_100_frames(ops):
    """C1: synthetic 640x480 sequence, 100 frames, 2k keypoints, the reference's max_trials=200 -- every one
    of the 99 pairs against the oracle (matches bit-exact, same best hypothesis and mask, E, pose chain)."""
    F, N, H = 100, 2048, 200
    frames, _ = synth.make_sequence(F, N, seed=1000)
    des = np.stack([f[1] for f in frames])
    kps = np.stack([f[0] for f in frames])
    words = synth.sample_words(H, 1000)
    out = ops.process_pairs_device(torch.from_numpy(des).cuda(), torch.from_numpy(kps).cuda(), K, words=words,
                                   residual_threshold=THR, outputs=ALL_OUT)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    assert (out["status"] == 0).all()
    _check_sequence(ops, out, frames, words, H, [(i, i + 1) for i in range(F - 1)])


def test_config_c3_keyframe_pair_8k_4096_hypotheses(ops):
    """C3: 8192 x 8192 descriptors, 4096 hypotheses, full pipeline against the oracle."""
    N, H = 8192, 4096
    lk, ld, ck, cd = synth.make_pair_large(N, 3003)
    words = synth.sample_words(H, 3003)
    des = torch.from_numpy(np.stack([ld, cd])).cuda()
    kps = torch.from_numpy(np.stack([lk, ck])).cuda()
    out = ops.process_pairs_device(des, kps, K, words=words, residual_threshold=THR, outputs=ALL_OUT)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    ref = tv.process_pair(lk, ld, ck, cd, np.eye(4), K, words=words, residual_threshold=THR)
    M = len(ref["idx1"])
    ri, rd = tv.knn2_cv2(cd, ld)
    assert np.array_equal(out["idx"][0], ri) and np.array_equal(out["dist"][0], rd)
    assert int(out["m_count"][0]) == M
    assert np.array_equal(out["qidx"][0, :M], ref["idx1"]) and np.array_equal(out["tidx"][0, :M], ref["idx2"])
    assert int(out["best"][0]) == ref["ransac"]["best"]
    # per-hypothesis counts: exact except points within 1e-5 of the threshold
    src, dst = tv.normalize(K, ref["match_kps"][:, 1]), tv.normalize(K, ref["match_kps"][:, 0])
    samples = synth.samples_from_words(words, M)
    cond = tv.sample_conditioning(src, dst, samples)
    rc, gc = ref["ransac"]["counts"], out["counts"][0]
    ok = (cond > 1e-7) & (rc >= 0)
    assert ok.mean() > 0.95
    assert (gc[ok] == rc[ok]).mean() > 0.98
    # the north star's bound, per hypothesis: the GPU count may differ from the oracle's only by correspondences
    # whose residual lies within 1e-5 of the threshold (oracle model of the same 8 samples)
    for h in np.flatnonzero(ok):
        if gc[h] == rc[h]:
            continue
        r = np.abs(tv.essential_residuals(tv.essential_estimate(src[samples[h]], dst[samples[h]]), src, dst))
        band = np.abs(r - THR) < BAND
        lo, hi = rc[h] - int(np.sum(band & (r < THR))), rc[h] + int(np.sum(band & (r >= THR)))
        assert lo <= gc[h] <= hi, (h, gc[h], rc[h], lo, hi)
    assert np.array_equal(out["mask"][0, :M].astype(bool), ref["inliers"])
    assert _fro_diff(out["E"][0], ref["E"]) < E_TOL
    np.testing.assert_allclose(out["Rt"][0], ref["Rt"], atol=1e-6)
    n_in = int(ref["inliers"].sum())
    X, Xr = out["points4d"][0, :n_in], ref["points4d"]
    rel = np.linalg.norm(X[:, :3] - Xr[:, :3], axis=1) / np.maximum(np.linalg.norm(Xr[:, :3], axis=1), 1e-12)
    assert np.median(rel) < X_TOL and (rel < 1e-3).mean() > 0.99


def test_matcher_64k_properties(ops):
    """C5-size matcher (65536 x 65536): too big for the oracle in seconds, so check size-independent
    properties: planted true matches are found at distance <= 49 with a strictly larger runner-up, results do
    not depend on the split / variant configuration, and a 4096-row slice equals cv2 bit-for-bit."""
    N = 65536
    q, t = synth.correlated_descriptors(N, N, seed=64, clutter=0.0)
    idx, dist = ops.knn2_hamming256(q, t)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    assert (dist[:, 0] <= 49).all() and (dist[:, 1] >= dist[:, 0]).all()
    assert len(np.unique(idx[:, 0])) == N                      # planted matches are a permutation
    d_true = tv.hamming_matrix(q[:512], t[idx[:512, 0]]).diagonal()
    assert np.array_equal(d_true, dist[:512, 0])
    ri, rd = tv.knn2_cv2(q[1000:5096], t)
    assert np.array_equal(idx[1000:5096], ri) and np.array_equal(dist[1000:5096], rd)
    ctx = ops.context(0)
    try:
        ctx.set_matcher_config(8, 1, 2)
        i2, d2 = ops.knn2_hamming256(q, t)
    finally:
        ctx.set_matcher_config()
    assert np.array_equal(i2.cpu().numpy(), idx) and np.array_equal(d2.cpu().numpy(), dist)


def test_pairs_pose_mode_cheirality(ops):
    """pose_mode=1: Rt comes from the cheirality vote (true motion has t_z < 0 here, which the reference's
    t_z >= 0 heuristic flips); everything before the pose is unchanged."""
    F, N, H = 3, 1024, 128
    frames, Ts = synth.make_sequence(F, N, seed=808)
    des = np.stack([f[1] for f in frames])
    kps = np.stack([f[0] for f in frames])
    words = synth.sample_words(H, 2)
    outs = ("E", "Rt", "mask", "m_count", "n_inl", "pose_votes", "poses", "points4d", "good", "qidx", "tidx")
    a = ops.process_pairs_device(torch.from_numpy(des).cuda(), torch.from_numpy(kps).cuda(), K, words=words,
                                 outputs=outs, pose_mode=1)
    b = ops.process_pairs_device(torch.from_numpy(des).cuda(), torch.from_numpy(kps).cuda(), K, words=words,
                                 outputs=("E", "Rt", "mask"), pose_mode=0)
    a = {k: v.cpu().numpy() for k, v in a.items()}
    assert np.array_equal(a["E"], b["E"].cpu().numpy())
    for i in range(F - 1):
        M = int(a["m_count"][i])
        assert np.array_equal(a["mask"][i, :M], b["mask"].cpu().numpy()[i, :M])
        (lk, ld), (ck, cd) = frames[i], frames[i + 1]
        src = tv.normalize(K, lk[a["tidx"][i, :M]])
        dst = tv.normalize(K, ck[a["qidx"][i, :M]])
        inl = a["mask"][i, :M].astype(bool)
        rv, rRts = tv.cheirality_votes(a["E"][i], src[inl], dst[inl])
        assert np.abs(np.sort(a["pose_votes"][i]) - np.sort(rv)).max() <= 2
        np.testing.assert_allclose(a["Rt"][i], rRts[int(np.argmax(rv))], atol=1e-8)
        # ground truth relative motion x_curr = R x_last + t (up to the scale of t)
        T = Ts[i + 1] @ np.linalg.inv(Ts[i])
        np.testing.assert_allclose(a["Rt"][i][:3, :3], T[:3, :3], atol=2e-2)
        t = T[:3, 3] / np.linalg.norm(T[:3, 3])
        assert a["Rt"][i][:3, 3] @ t > 0.8          # right sign (the t_z heuristic would give < 0), roughly right direction
        assert a["pose_votes"][i].max() > 0.9 * int(a["n_inl"][i])


def test_pairs_host_submit_wait_two_contexts(ops):
    """Streaming form: two contexts alternate submit / wait; results equal the blocking call."""
    from openwork_b200 import _lib
    N, H = 512, 64
    batches = []
    for s in range(3):
        frames, _ = synth.make_sequence(5, N, seed=40 + s)
        batches.append((np.stack([f[1] for f in frames]), np.stack([f[0] for f in frames])))
    words = synth.sample_words(H, 4)
    outs = ("m_count", "E", "n_inl", "Rt", "poses", "status", "best")
    ref = [ops.process_pairs_host(d, k, K, words=words, outputs=outs)[0] for d, k in batches]
    ref = [{k: v.clone() for k, v in r.items()} for r in ref]
    ctxs = [_lib.Context(0), _lib.Context(0)]
    for c in ctxs:
        c.set_host_chunk(2)
    got = []
    pending = None
    for i, (d, k) in enumerate(batches):
        o, _, _ = ops.process_pairs_host(d, k, K, words=words, outputs=outs, ctx=ctxs[i % 2], wait=False)
        if pending is not None:
            ops.wait_pairs_host(ctxs[(i - 1) % 2])
            got.append({kk: v.clone() for kk, v in pending.items()})
        pending = o
    ops.wait_pairs_host(ctxs[(len(batches) - 1) % 2])
    got.append({kk: v.clone() for kk, v in pending.items()})
    for r, g in zip(ref, got):
        for kk in outs:
            assert torch.equal(r[kk], g[kk]) or (kk in ("E", "Rt", "poses") and
                                                 np.array_equal(r[kk].numpy(), g[kk].numpy(), equal_nan=True)), kk


# ------------------------------------------------------------------------------ "next" rows
def test_point_quality_matches_spec(ops):
    src, dst, (lk, ck, i1, i2) = _pair_correspondences(1500, 61)
    ref = tv.ransac_essential(src, dst, synth.sample_table(len(src), 128, 3), THR)
    inl = ref["mask"]
    pose0 = np.eye(4)
    pose1 = tv.extract_Rt(ref["E"]) @ pose0
    p1, p2 = lk[i2][inl], ck[i1][inl]
    Xr, er, cr = tv.point_quality(p1, p2, pose0, pose1, K)
    X, e, c = ops.point_quality(p1, p2, pose0, pose1, K)
    X, e, c = X.cpu().numpy(), e.cpu().numpy(), c.cpu().numpy()
    rel = np.linalg.norm(X[:, :3] - Xr[:, :3], axis=1) / np.maximum(np.linalg.norm(Xr[:, :3], axis=1), 1e-12)
    ok = rel < 1e-6
    assert ok.mean() > 0.99
    np.testing.assert_allclose(e[ok], er[ok], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(c[ok], cr[ok], atol=1e-7)
    assert (c <= 1.0 + 1e-12).all() and np.median(e) < 5.0


def test_run_sequences_map_accumulation(ops):
    """Device-resident sequence driver: (poses, points) per sequence == what the reference's loop would have
    accumulated in Map (origin point, then every pair's z>0 points in order)."""
    N, H = 700, 96
    fa, _ = synth.make_sequence(4, N, seed=71)
    fb, _ = synth.make_sequence(3, N, seed=72)
    frames = fa + fb
    des = np.stack([f[1] for f in frames])
    kps = np.stack([f[0] for f in frames])
    words = synth.sample_words(H, 6)
    seq_off = np.array([0, 3, 5], dtype=np.int32)
    res = ops.run_sequences(torch.from_numpy(des).cuda(), torch.from_numpy(kps).cuda(), K, words, seq_off=seq_off)
    for s, fr in enumerate((fa, fb)):
        poses, pts = res[s][0].cpu().numpy(), res[s][1].cpu().numpy()
        pose = np.eye(4)
        Xs, gs, ref_poses = [], [], [pose]
        for i in range(len(fr) - 1):
            r = tv.process_pair(fr[i][0], fr[i][1], fr[i + 1][0], fr[i + 1][1], pose, K, words=words)
            Xs.append(r["points4d"]); gs.append(r["good"]); pose = r["curr_pose"]; ref_poses.append(pose)
        ref_pts = tv.map_points(Xs, gs)
        assert poses.shape == (len(fr), 4, 4) and pts.shape == ref_pts.shape
        np.testing.assert_allclose(poses, np.stack(ref_poses), atol=1e-5)
        assert np.array_equal(pts[0], [0, 0, 0, 1])
        rel = np.linalg.norm(pts[1:, :3] - ref_pts[1:, :3], axis=1) / np.linalg.norm(ref_pts[1:, :3], axis=1)
        assert np.median(rel) < X_TOL and (rel < 1e-3).mean() > 0.99


def test_pairs_with_degenerate_members_do_not_disturb_the_batch(ops):
    """A pair with < 9 surviving matches (the reference would raise AssertionError / ValueError, slam.py:138 and
    skimage) reports status 1 with NaN E and no points; an empty frame likewise; healthy pairs are untouched."""
    N, H = 600, 64
    fa, _ = synth.make_sequence(2, N, seed=91)
    fb, _ = synth.make_sequence(2, N, seed=92)
    rnd = (np.random.default_rng(0).uniform(0, 480, (N, 2)), synth.random_descriptors(N, 5))
    rnd2 = (np.random.default_rng(1).uniform(0, 480, (N, 2)), synth.random_descriptors(N, 6))
    frames = [fa[0], fa[1], rnd, rnd2, fb[0], fb[1], fb[0], fb[1]]
    des = np.stack([f[1] for f in frames])
    kps = np.stack([f[0] for f in frames])
    cnt = np.array([N, N, N, N, N, N, 0, N], dtype=np.int32)       # frame 6 is empty
    words = synth.sample_words(H, 1)
    pq = np.array([1, 3, 5, 7], dtype=np.int32)
    pt = np.array([0, 2, 4, 6], dtype=np.int32)
    so = np.arange(5, dtype=np.int32)
    outs = ("m_count", "E", "n_inl", "Rt", "poses", "status", "best", "points4d", "good")
    out = ops.process_pairs_device(torch.from_numpy(des).cuda(), torch.from_numpy(kps).cuda(), K, words=words,
                                   frame_count=cnt, pair_q=pq, pair_t=pt, seq_off=so, outputs=outs)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    assert out["status"].tolist() == [0, 1, 0, 1]
    assert out["m_count"][1] <= 8 and out["m_count"][3] == 0
    assert out["n_inl"][1] == 0 and out["n_inl"][3] == 0 and np.isnan(out["E"][1]).all() and np.isnan(out["E"][3]).all()
    for b, fr in ((0, fa), (2, fb)):
        ref = tv.process_pair(fr[0][0], fr[0][1], fr[1][0], fr[1][1], np.eye(4), K, words=words)
        assert int(out["best"][b]) == ref["ransac"]["best"] and int(out["n_inl"][b]) == int(ref["inliers"].sum())
        assert _fro_diff(out["E"][b], ref["E"]) < E_TOL
        np.testing.assert_allclose(out["poses"][2 * b + 1], ref["curr_pose"], atol=1e-5)
    host, _, _ = ops.process_pairs_host(des, kps, K, words=words, frame_count=cnt, pair_q=pq, pair_t=pt, seq_off=so,
                                        outputs=("status", "n_inl", "m_count"))
    assert host["status"].tolist() == [0, 1, 0, 1]


def test_knn2_randomised_shapes_and_configs(ops):
    """Seeded sweep over ragged (Nq, Nt), low-entropy descriptors (many exact ties) and matcher configurations:
    always equal to the stable-argsort oracle."""
    rng = np.random.default_rng(20261019)
    ctx = ops.context(0)
    variants = [0, 1, 2, 3, 5, 7, 9, 11, 16, 16, 16]
    try:
        for trial in range(24):
            nq = int(rng.integers(1, 1400))
            nt = int(rng.integers(2, 3000))
            bits = int(rng.choice([2, 16, 256]))
            q = rng.integers(0, bits, (nq, 32), dtype=np.uint8)
            t = rng.integers(0, bits, (nt, 32), dtype=np.uint8)
            ctx.set_matcher_config(int(rng.choice([2, 4, 8])), int(rng.choice(variants)), int(rng.integers(1, 17)))
            idx, dist, keep = ops.knn2_hamming256(q, t, return_keep=True)
            ri, rd = tv.knn2_numpy(q, t)
            assert np.array_equal(idx.cpu().numpy(), ri) and np.array_equal(dist.cpu().numpy(), rd), (trial, nq, nt)
            assert np.array_equal(np.flatnonzero(keep.cpu().numpy()), tv.ratio_filter(ri, rd)[0])
    finally:
        ctx.set_matcher_config()
