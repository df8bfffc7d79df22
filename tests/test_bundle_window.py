"""Sliding-window bundle adjustment (slam.py:267 TODO; spec = oracle/bundle_window.py): the oracle against the two-view
spec and against scipy on the CPU, the CUDA kernel against the oracle on the GPU."""
import numpy as np
import pytest

from oracle import bundle as ba2
from oracle import bundle_window as baw
from openwork_b200 import synth

K = synth.K_640x480
Kf = np.asarray(K, dtype=np.float64)


def _window(seed, W=5, n=200, noise=0.4, pose_err=0.01, depth_err=0.04, drop=0.25):
    """W frames moving forward, n points; each point visible in a random contiguous run of >= 2 frames."""
    rng = np.random.default_rng(seed)
    X = np.c_[rng.uniform(-6, 6, n), rng.uniform(-3, 3, n), rng.uniform(8, 30, n)]
    poses = np.empty((W, 4, 4))
    T = np.eye(4)
    for k in range(W):
        poses[k] = T
        step = np.eye(4)
        step[:3, :3] = ba2._exp_so3(rng.normal(0, 0.01, 3))
        step[:3, 3] = np.array([0.05, -0.02, 0.5]) + rng.normal(0, 0.02, 3)
        T = T @ step
    obs = np.empty((n, W, 2))
    for k in range(W):
        P = np.linalg.inv(poses[k])
        Xc = X @ P[:3, :3].T + P[:3, 3]
        obs[:, k] = (np.c_[Xc[:, :2] / Xc[:, 2:3], np.ones(n)] @ Kf.T)[:, :2] + rng.normal(0, noise, (n, 2))
    vis = np.zeros((n, W), dtype=bool)
    for j in range(n):
        if rng.random() < drop:
            a = rng.integers(0, W - 1)
            b = rng.integers(a + 2, W + 1)
            vis[j, a:b] = True
        else:
            vis[j] = True
    start = poses.copy()
    for k in range(1, W):
        start[k, :3, :3] = ba2._exp_so3(rng.normal(0, pose_err, 3)) @ start[k, :3, :3]
        start[k, :3, 3] += rng.normal(0, pose_err, 3)
    pts = np.c_[X * (1.0 + rng.normal(0, depth_err, (n, 1))), np.ones(n)]
    return obs, vis, start, pts, poses, X


def test_window_of_two_frames_is_the_two_view_spec():
    obs, vis, start, pts, _, _ = _window(3, W=2, n=120, drop=0.0)
    pw, xw, cw, aw = baw.bundle_adjust_window(obs, vis, start, pts, K, n_fixed=1, iters=8)
    p2, x2, c2, a2 = ba2.bundle_adjust_pair(obs[:, 0], obs[:, 1], start[0], start[1], pts, K, iters=8)
    assert aw == a2
    np.testing.assert_allclose(cw, c2, rtol=1e-9)
    np.testing.assert_allclose(pw[1], p2, atol=1e-9)
    np.testing.assert_allclose(xw, x2, rtol=1e-8, atol=1e-8)


def test_window_oracle_reduces_cost_and_recovers_rotations():
    obs, vis, start, pts, true_poses, _ = _window(4, W=6, n=250)
    poses, out, (c0, c1), acc = baw.bundle_adjust_window(obs, vis, start, pts, K, n_fixed=1, iters=15)
    assert acc >= 3 and c1 < 0.1 * c0
    assert np.array_equal(poses[0], start[0])
    err = lambda P: max(np.linalg.norm(P[k, :3, :3] - true_poses[k, :3, :3]) for k in range(1, 6))
    assert err(poses) < 0.3 * err(start)
    # points nobody observes twice still come back finite; unobserved points unchanged
    assert np.isfinite(out).all()


def test_window_oracle_reaches_a_minimum_scipy_cannot_improve():
    scipy_opt = pytest.importorskip("scipy.optimize")
    obs, vis, start, pts, _, _ = _window(5, W=4, n=40, drop=0.3)
    n, W = vis.shape
    poses, out, (c0, c1), acc = baw.bundle_adjust_window(obs, vis, start, pts, K, n_fixed=2, iters=60)
    Kinv = np.linalg.inv(Kf)
    x = np.stack([(np.c_[obs[:, k], np.ones(n)] @ Kinv.T)[:, :2] for k in range(W)], axis=1)

    def make_fun(pose_ref):
        cams = [ba2._cam(p) for p in pose_ref]

        def fun(z):
            res = []
            X = z[6 * (W - 2):].reshape(n, 3)
            for k in range(W):
                R, t = cams[k]
                if k >= 2:
                    d = z[6 * (k - 2):6 * (k - 2) + 6]
                    E = ba2._exp_so3(d[:3])
                    R, t = E @ R, E @ t + d[3:]
                p, _ = ba2._proj(R, t, X)
                res.append(((p - x[:, k]) * vis[:, k, None]).ravel())
            return np.concatenate(res)
        return fun
    z0 = np.concatenate([np.zeros(6 * (W - 2)), pts[:, :3].ravel()])
    sp = scipy_opt.least_squares(make_fun(start), z0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=20000)
    assert c1 <= 2.0 * sp.cost * (1 + 1e-6) + 1e-12          # least_squares reports 0.5 * sum r^2
    z1 = np.concatenate([np.zeros(6 * (W - 2)), out[:, :3].ravel()])
    sp2 = scipy_opt.least_squares(make_fun(poses), z1, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=20000)
    assert 2.0 * sp2.cost >= c1 * (1 - 1e-6)


def test_window_tracks_chains_inlier_matches_across_frames():
    """ops.window_tracks on hand-written matches (CPU tensors): a track is extended while its keypoint is matched
    again, a train keypoint claimed by two queries goes to the lowest query, new tracks start where none ends."""
    import torch
    from openwork_b200 import ops
    N, W = 6, 3
    kps = torch.arange(3 * N * 2, dtype=torch.float64).reshape(3, N, 2)
    # pair 0 (frame 0 -> 1): rows (q,t) = (0,3) (1,4) (2,4) (5,0); row 2 is an outlier; rows 1 and 3.. : inliers 0,1,3
    # pair 1 (frame 1 -> 2): rows (q,t) = (2,0) (3,1) (4,5); all inliers.  q of pair 0 = t of pair 1 continues a track.
    qidx = torch.tensor([[0, 1, 2, 5, 0, 0], [2, 3, 4, 0, 0, 0]], dtype=torch.int32)
    tidx = torch.tensor([[3, 4, 4, 0, 0, 0], [0, 1, 5, 0, 0, 0]], dtype=torch.int32)
    inl_idx = torch.tensor([[0, 1, 3, 0, 0, 0], [0, 1, 2, 0, 0, 0]], dtype=torch.int32)
    n_inl = torch.tensor([3, 3], dtype=torch.int32)
    pts = torch.arange(2 * N * 4, dtype=torch.float64).reshape(2, N, 4)
    out = dict(qidx=qidx, tidx=tidx, inl_idx=inl_idx, n_inl=n_inl, points4d=pts)
    obs, vis, p4 = ops.window_tracks(out, kps, 0, W)
    # pair 0 sorted by train keypoint: t=0 (row 3, q=5), t=3 (row 0, q=0), t=4 (row 1, q=1) -> tracks 0, 1, 2
    # pair 1: t=0 -> q=2 continues track 1 (frame-1 keypoint 0); t=1 -> q=3 continues track 2; t=5 -> q=4 continues 0
    assert obs.shape == (3, W, 2) and vis.shape == (3, W)
    assert vis.numpy().tolist() == [[1, 1, 1]] * 3
    k = kps.numpy()
    np.testing.assert_array_equal(obs[0].numpy(), [k[0, 0], k[1, 5], k[2, 4]])
    np.testing.assert_array_equal(obs[1].numpy(), [k[0, 3], k[1, 0], k[2, 2]])
    np.testing.assert_array_equal(obs[2].numpy(), [k[0, 4], k[1, 1], k[2, 3]])
    # the 3-D start of a track is the triangulation of the pair (compacted inlier order) that started it
    np.testing.assert_array_equal(p4.numpy(), pts[0, [2, 0, 1]].numpy())
    # a track that is not matched again stops, and a fresh one starts on an unclaimed keypoint
    tidx2 = tidx.clone(); tidx2[1, 2] = 2                         # frame-1 keypoint 2 carries no track
    obs, vis, p4 = ops.window_tracks(dict(out, tidx=tidx2), kps, 0, W)
    assert vis.numpy().tolist() == [[1, 1, 0], [1, 1, 1], [1, 1, 1], [0, 1, 1]]
    np.testing.assert_array_equal(obs[3, 1:].numpy(), [k[1, 2], k[2, 4]])
    np.testing.assert_array_equal(p4[3].numpy(), pts[1, 2].numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("W,n_fixed,n", [(2, 1, 300), (5, 1, 700), (8, 2, 1500), (3, 1, 40)])
def test_gpu_window_matches_oracle(W, n_fixed, n):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    from openwork_b200 import ops
    B = 3
    scenes = [_window(10 + b, W=W, n=n - 17 * b) for b in range(B)]
    n_max = n
    obs = np.zeros((B, n_max, W, 2))
    vis = np.zeros((B, n_max, W), dtype=np.uint8)
    pts = np.zeros((B, n_max, 4))
    poses = np.zeros((B, W, 4, 4))
    cnt = np.zeros(B, dtype=np.int32)
    for b, (o, v, s, p, _, _) in enumerate(scenes):
        m = len(o)
        obs[b, :m], vis[b, :m], pts[b, :m], poses[b], cnt[b] = o, v, p, s, m
    iters = 8
    gp, gx, gc, ga = ops.bundle_adjust_window(obs, vis, poses, pts, K, n_pts=cnt, n_fixed=n_fixed, iters=iters)
    gp, gx, gc, ga = gp.cpu().numpy(), gx.cpu().numpy(), gc.cpu().numpy(), ga.cpu().numpy()
    for b, (o, v, s, p, _, _) in enumerate(scenes):
        rp, rx, (c0, c1), acc = baw.bundle_adjust_window(o, v, s, p, K, n_fixed=n_fixed, iters=iters)
        m = len(o)
        assert ga[b] == acc
        np.testing.assert_allclose(gc[b], [c0, c1], rtol=1e-6)
        np.testing.assert_allclose(gp[b], rp, atol=1e-6)
        np.testing.assert_allclose(gx[b, :m], rx, rtol=1e-5, atol=1e-5)
        assert np.array_equal(gp[b, :n_fixed], s[:n_fixed])


@pytest.mark.gpu
def test_window_after_pipeline_reduces_reprojection_error():
    """End to end: a sequence through the device pipeline, tracks chained from its inlier matches, then the window
    adjustment over five frames lowers the multi-frame reprojection cost (the pair pipeline leaves every pair at its
    own unit-baseline scale; the window makes them consistent)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    from openwork_b200 import ops
    W = 5
    seq, _ = synth.make_sequence(W, 1024, seed=11)
    des = torch.from_numpy(np.stack([f[1] for f in seq])).cuda()
    kps = torch.from_numpy(np.stack([f[0] for f in seq])).cuda()
    out = ops.process_pairs_device(des, kps, K, words=synth.sample_words(512, 3),
                                   outputs=("n_inl", "inl_idx", "qidx", "tidx", "poses", "points4d", "good", "status"))
    assert (out["status"].cpu().numpy() == 0).all()
    obs, vis, pts = ops.window_tracks(out, kps, 0, W)
    v = vis.cpu().numpy()
    assert (v.sum(axis=1) >= 2).all() and (v.sum(axis=1) >= 3).sum() > 100          # tracks really span frames
    runs = np.diff(np.pad(v.astype(int), ((0, 0), (1, 1))), axis=1)
    assert ((runs == 1).sum(axis=1) == 1).all()                                      # each track one contiguous run
    poses, X, cost, acc = ops.bundle_adjust_window(obs[None], vis[None], out["poses"][None, :W], pts[None], K,
                                                   n_fixed=1, iters=12)
    cost = cost.cpu().numpy()[0]
    assert int(acc[0]) >= 2 and cost[1] < 0.5 * cost[0]
    assert np.isfinite(poses.cpu().numpy()).all() and np.isfinite(X.cpu().numpy()).all()
    # the same window through the specification
    rp, rx, (c0, c1), racc = baw.bundle_adjust_window(obs.cpu().numpy(), v, out["poses"][:W].cpu().numpy(),
                                                      pts.cpu().numpy(), K, n_fixed=1, iters=12)
    np.testing.assert_allclose(cost, [c0, c1], rtol=1e-5)
