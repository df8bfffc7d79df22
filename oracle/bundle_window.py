"""CPU specification of the sliding-window bundle adjustment (TEST INFRASTRUCTURE ONLY -- imported by tests/, never by
the product).

The reference has no bundle adjustment: MonocularSlam/slam.py:267 is a TODO ("g2o bundle adjustment").  Parity is
therefore UNPINNED by the reference; this file is the specification the CUDA kernel (csrc/bundle_window.cu) is tested
against, and it is itself cross-checked against scipy.optimize.least_squares in tests/test_bundle_window.py.

Problem: W consecutive frames of a sequence (camera-to-world poses, the reference's `pose`, slam.py:261), n tracked
points, a dense observation table obs[j, c] (pixels) with a visibility mask vis[j, c].
    P_c = inv(pose_c)[:3]   (world -> camera, slam.py:204-205),   x = (inv(K) [u, v, 1])[:2]   (slam.py:87-92)
    cost = sum over visible (j, c) of | proj(P_c X_j) - x_cj |^2
Variables: a left-multiplied twist (w, v) on every P_c with c >= n_fixed (R <- exp([w]x) R, t <- exp([w]x) t + v; the
first n_fixed cameras stay put: n_fixed = 1 leaves the scale gauge to the damping, n_fixed = 2 fixes it) and the points.
Method: Levenberg-Marquardt, Marquardt scaling of the diagonals (1 + lambda), points eliminated by the Schur complement
(reduced camera system of size 6 (W - n_fixed)), exactly `iters` iterations, a step is accepted iff the cost
decreases (lambda /= 10, else lambda *= 10).  With W = 2, n_fixed = 1 this is oracle/bundle.py.
"""
import numpy as np

from .bundle import _blocks, _cam, _exp_so3, _proj


def window_cost(Rs, ts, X, x, vis):
    c = 0.0
    for k in range(len(Rs)):
        p, _ = _proj(Rs[k], ts[k], X)
        c += float(np.sum(((p - x[:, k]) ** 2)[vis[:, k]]))
    return c


def bundle_adjust_window(obs, vis, poses, points4d, K, n_fixed=1, iters=10, lambda0=1e-3):
    """obs f64[n,W,2] pixels, vis bool[n,W], poses f64[W,4,4], points4d f64[n,4].
    Returns (poses_refined f64[W,4,4], points4d_refined f64[n,4], (cost0, cost1), accepted)."""
    Kinv = np.linalg.inv(np.asarray(K, dtype=np.float64))
    obs = np.asarray(obs, dtype=np.float64)
    vis = np.asarray(vis).astype(bool)
    n, W = vis.shape
    x = np.empty_like(obs)
    for k in range(W):
        x[:, k] = (np.c_[obs[:, k], np.ones(n)] @ Kinv.T)[:, :2]
    cams = [_cam(p) for p in np.asarray(poses, dtype=np.float64)]
    Rs, ts = [c[0] for c in cams], [c[1] for c in cams]
    pts = np.array(points4d, dtype=np.float64)
    X = pts[:, :3].copy()
    nv = W - n_fixed
    lam = float(lambda0)
    c_now = window_cost(Rs, ts, X, x, vis)
    c0, accepted = c_now, 0
    for _ in range(iters):
        V = np.zeros((n, 3, 3))
        gp = np.zeros((n, 3))
        Wb = np.zeros((n, nv, 6, 3))
        U = np.zeros((nv, 6, 6))
        gc = np.zeros((nv, 6))
        for k in range(W):
            r, Jp, Jc = _blocks(Rs[k], ts[k], X, x[:, k])
            m = vis[:, k].astype(np.float64)
            r, Jp, Jc = r * m[:, None], Jp * m[:, None, None], Jc * m[:, None, None]
            V += np.einsum("nki,nkj->nij", Jp, Jp)
            gp += np.einsum("nki,nk->ni", Jp, r)
            if k >= n_fixed:
                a = k - n_fixed
                Wb[:, a] = np.einsum("nka,nkj->naj", Jc, Jp)
                U[a] = np.einsum("nka,nkb->ab", Jc, Jc)
                gc[a] = np.einsum("nka,nk->a", Jc, r)
        idx = np.arange(3)
        V[:, idx, idx] *= (1.0 + lam)
        seen = vis.any(axis=1)
        V[~seen] = np.eye(3)                      # points nobody sees: neutral blocks
        Vi = np.linalg.inv(V)
        Y = np.einsum("ncak,nkj->ncaj", Wb, Vi)   # W V^-1
        S = np.zeros((6 * nv, 6 * nv))
        rhs = np.zeros(6 * nv)
        for a in range(nv):
            Ua = U[a].copy()
            Ua[np.arange(6), np.arange(6)] *= (1.0 + lam)
            S[6 * a:6 * a + 6, 6 * a:6 * a + 6] += Ua
            rhs[6 * a:6 * a + 6] = -(gc[a] - np.einsum("nak,nk->a", Y[:, a], gp))
            for b in range(nv):
                S[6 * a:6 * a + 6, 6 * b:6 * b + 6] -= np.einsum("nak,nbk->ab", Y[:, a], Wb[:, b])
        try:
            d = np.linalg.solve(S, rhs)
        except np.linalg.LinAlgError:
            d = np.full(6 * nv, np.nan)
        ok = np.isfinite(d).all()
        if ok:
            Rn, tn = list(Rs), list(ts)
            for a in range(nv):
                E = _exp_so3(d[6 * a:6 * a + 3])
                Rn[n_fixed + a], tn[n_fixed + a] = E @ Rs[n_fixed + a], E @ ts[n_fixed + a] + d[6 * a + 3:6 * a + 6]
            e = gp + np.einsum("ncak,ca->nk", Wb, d.reshape(nv, 6))
            Xn = X - np.einsum("nij,nj->ni", Vi, e)
            Xn[~seen] = X[~seen]
            c_new = window_cost(Rn, tn, Xn, x, vis)
        if ok and c_new < c_now:
            Rs, ts, X, c_now = Rn, tn, Xn, c_new
            lam = max(lam * 0.1, 1e-15)
            accepted += 1
        else:
            lam = min(lam * 10.0, 1e12)
    out = np.empty((W, 4, 4))
    for k in range(W):
        P = np.eye(4)
        P[:3, :3], P[:3, 3] = Rs[k], ts[k]
        out[k] = np.linalg.inv(P)
    pts[:, :3] = X
    return out, pts, (c0, c_now), accepted
