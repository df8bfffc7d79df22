"""Two-view bundle adjustment (slam.py:267 TODO; spec = oracle/bundle.py): the oracle against scipy on the CPU, the CUDA
kernel against the oracle on the GPU."""
import numpy as np
import pytest

from oracle import bundle as ba
from openwork_b200 import synth

K = synth.K_640x480


def _scene(seed, n=300, noise=0.5, pose_err=0.02, depth_err=0.05):
    rng = np.random.default_rng(seed)
    X = np.c_[rng.uniform(-6, 6, n), rng.uniform(-3, 3, n), rng.uniform(6, 30, n)]
    pose1 = np.eye(4)
    th = rng.normal(0, 0.01, 3)
    R = ba._exp_so3(th)
    pose2 = np.eye(4)
    pose2[:3, :3] = R
    pose2[:3, 3] = np.array([0.05, -0.02, 0.6]) + rng.normal(0, 0.02, 3)

    def pix(pose):
        P = np.linalg.inv(pose)
        Xc = X @ P[:3, :3].T + P[:3, 3]
        u = Xc[:, :2] / Xc[:, 2:3]
        return (np.c_[u, np.ones(n)] @ np.asarray(K, dtype=np.float64).T)[:, :2]
    p1 = pix(pose1) + rng.normal(0, noise, (n, 2))
    p2 = pix(pose2) + rng.normal(0, noise, (n, 2))
    # perturbed start: pose and depths
    start = pose2.copy()
    start[:3, :3] = ba._exp_so3(rng.normal(0, pose_err, 3)) @ start[:3, :3]
    start[:3, 3] += rng.normal(0, pose_err, 3)
    X0 = X * (1.0 + rng.normal(0, depth_err, (n, 1)))
    pts = np.c_[X0, np.ones(n)]
    return p1, p2, pose1, start, pts, pose2, X


def test_oracle_reduces_cost_and_recovers_pose():
    p1, p2, pose1, start, pts, true_pose, X = _scene(1)
    pose, out, (c0, c1), acc = ba.bundle_adjust_pair(p1, p2, pose1, start, pts, K, iters=15)
    assert acc >= 3 and c1 < 0.05 * c0
    # rotation recovered far better than the perturbed start (translation scale is a gauge freedom)
    err = lambda P: np.linalg.norm(P[:3, :3] - true_pose[:3, :3])
    assert err(pose) < 0.2 * err(start)


def test_oracle_reaches_a_minimum_scipy_cannot_improve():
    """scipy.optimize.least_squares on the same residual function: from the same start it must not find a lower cost
    than the oracle (it stops earlier in the flat scale-gauge valley), and started AT the oracle's solution it must
    not improve it."""
    scipy_opt = pytest.importorskip("scipy.optimize")
    p1, p2, pose1, start, pts, _, _ = _scene(2, n=60)
    pose, out, (c0, c1), acc = ba.bundle_adjust_pair(p1, p2, pose1, start, pts, K, iters=80)
    Kinv = np.linalg.inv(np.asarray(K, dtype=np.float64))
    h = lambda px: (np.c_[px, np.ones(len(px))] @ Kinv.T)[:, :2]
    x1, x2 = h(p1), h(p2)
    R1, t1 = ba._cam(pose1)

    def make_fun(pose_ref):
        R2, t2 = ba._cam(pose_ref)

        def fun(z):
            E = ba._exp_so3(z[:3])
            R, t = E @ R2, E @ t2 + z[3:6]
            X = z[6:].reshape(-1, 3)
            a, _ = ba._proj(R1, t1, X)
            b, _ = ba._proj(R, t, X)
            return np.r_[(a - x1).ravel(), (b - x2).ravel()]
        return fun
    kw = dict(method="trf", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=400)
    res = scipy_opt.least_squares(make_fun(start), np.r_[np.zeros(6), pts[:, :3].ravel()], **kw)
    assert c1 <= 2 * res.cost * (1 + 1e-9)
    res2 = scipy_opt.least_squares(make_fun(pose), np.r_[np.zeros(6), out[:, :3].ravel()], **kw)
    assert abs(2 * res2.cost - c1) <= 1e-7 * c1
    assert np.linalg.norm(res2.grad) <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,iters", [(3, 300, 10), (4, 2048, 8), (5, 17, 12)])
def test_bundle_kernel_matches_oracle(seed, n, iters):
    import torch
    from openwork_b200 import ops
    B = 5
    scenes = [_scene(seed * 10 + b, n=n) for b in range(B)]
    p1 = np.stack([s[0] for s in scenes]); p2 = np.stack([s[1] for s in scenes])
    pose1 = np.stack([s[2] for s in scenes]); pose2 = np.stack([s[3] for s in scenes])
    pts = np.stack([s[4] for s in scenes])
    cnt = np.array([n, n, max(9, n // 2), n, n - 1], dtype=np.int32)
    use = np.ones((B, n), dtype=np.uint8)
    use[1, ::3] = 0
    pose_g, pts_g, cost_g, acc_g = ops.bundle_adjust_pairs(p1, p2, pose1, pose2, pts, K, n_pts=cnt, use=use, iters=iters)
    pose_g, pts_g, cost_g, acc_g = [t.cpu().numpy() for t in (pose_g, pts_g, cost_g, acc_g)]
    for b in range(B):
        m = cnt[b]
        pose_o, pts_o, (c0, c1), acc = ba.bundle_adjust_pair(p1[b, :m], p2[b, :m], pose1[b], pose2[b], pts[b, :m], K,
                                                             use=use[b, :m], iters=iters)
        assert abs(cost_g[b, 0] - c0) <= 1e-10 * c0
        # tolerances: the two fp64 evaluation orders differ at 1e-16 per operation and the scale-gauge valley of the
        # two-view problem amplifies that over the iterations: final cost 1e-6 relative, pose 1e-6, points 1e-5
        assert abs(cost_g[b, 1] - c1) <= 1e-6 * c1, (b, cost_g[b], c1)
        assert abs(int(acc_g[b]) - acc) <= 1
        # two-view BA leaves the global scale (about the fixed first camera) undetermined: compare the rotation, the
        # baseline direction and the structure up to that one common factor
        assert np.allclose(pose_g[b][:3, :3], pose_o[:3, :3], rtol=0, atol=1e-7)
        c1w = pose1[b][:3, 3]
        bg, bo = pose_g[b][:3, 3] - c1w, pose_o[:3, 3] - c1w
        sc = np.linalg.norm(bg) / np.linalg.norm(bo)
        assert abs(sc - 1.0) < 1e-4
        assert np.allclose(bg / np.linalg.norm(bg), bo / np.linalg.norm(bo), rtol=0, atol=1e-6)
        live = use[b, :m] != 0
        Xg, Xo = pts_g[b, :m][live, :3], pts_o[live, :3]
        near = np.linalg.norm(Xo - c1w, axis=1) < 100.0          # depth of (near-)infinite points is unobservable
        assert np.allclose(Xg[near] - c1w, sc * (Xo[near] - c1w), rtol=1e-4, atol=1e-6)
        for pg, po in ((pose1[b], pose1[b]), (pose_g[b], pose_o)):   # what the cost sees: the projections
            ug, _ = ba._proj(*ba._cam(pg), Xg)
            uo, _ = ba._proj(*ba._cam(po), Xo)
            assert np.allclose(ug, uo, rtol=0, atol=1e-7)
        assert np.array_equal(pts_g[b, m:], pts[b, m:])                     # rows beyond n_pts untouched
        assert np.array_equal(pts_g[b, :m][use[b, :m] == 0], pts[b, :m][use[b, :m] == 0])
        assert cost_g[b, 1] < cost_g[b, 0]


@pytest.mark.gpu
def test_bundle_after_pipeline_reduces_reprojection_error():
    """End to end: pairs through the device pipeline, then bundle adjustment on their inliers lowers the cost."""
    import torch
    from openwork_b200 import ops
    seq, _ = synth.make_sequence(4, 1024, seed=9)
    des = torch.from_numpy(np.stack([f[1] for f in seq])).cuda()
    kps = torch.from_numpy(np.stack([f[0] for f in seq])).cuda()
    out = ops.process_pairs_device(des, kps, K, words=synth.sample_words(512, 3),
                                   outputs=("n_inl", "inl_idx", "qidx", "tidx", "poses", "points4d", "good", "status"))
    assert (out["status"].cpu().numpy() == 0).all()
    B, N = out["points4d"].shape[0], out["points4d"].shape[1]
    inl = out["inl_idx"].long().clamp(0, N - 1)           # rows beyond n_inl / m_count are unspecified
    qi = torch.gather(out["qidx"].long().clamp(0, N - 1), 1, inl)
    ti = torch.gather(out["tidx"].long().clamp(0, N - 1), 1, inl)
    pix_curr = torch.stack([kps[b + 1][qi[b]] for b in range(B)])
    pix_last = torch.stack([kps[b][ti[b]] for b in range(B)])
    poses = out["poses"]
    pose2, pts, cost, acc = ops.bundle_adjust_pairs(pix_last, pix_curr, poses[:B], poses[1:B + 1], out["points4d"], K,
                                                    n_pts=out["n_inl"], use=out["good"], iters=8)
    cost = cost.cpu().numpy()
    assert (cost[:, 1] <= cost[:, 0]).all() and (cost[:, 1] < 0.9 * cost[:, 0]).any()
    assert np.isfinite(pose2.cpu().numpy()).all()
