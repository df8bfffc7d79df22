"""CPU specification of the two-view bundle adjustment (TEST INFRASTRUCTURE ONLY -- imported by tests/, never by the
product).

The reference has no bundle adjustment: MonocularSlam/slam.py:267 is a TODO ("g2o bundle adjustment").  Parity is
therefore UNPINNED by the reference; this file is the specification the CUDA kernel (csrc/bundle.cu) is tested
against, and it is itself cross-checked against scipy.optimize.least_squares in tests/test_bundle.py.

Problem: two views of one frame pair, P_i = inv(pose_i)[:3] (world -> camera, the reference's convention at
slam.py:204-205), normalised image points x_i = (inv(K) [u, v, 1])[:2] (slam.py:87-92), points X_j.
    cost = sum_j |proj(P1 X_j) - x1_j|^2 + |proj(P2 X_j) - x2_j|^2
Variables: a left-multiplied twist (w, v) on P2 (R <- exp([w]x) R, t <- exp([w]x) t + v; view 1 fixed) and the points.
Method: Levenberg-Marquardt, Marquardt scaling of the diagonals (1 + lambda), points eliminated by the Schur
complement, exactly `iters` iterations, a step is accepted iff the cost decreases (lambda /= 10, else lambda *= 10).
"""
import numpy as np


def _cam(pose):
    P = np.linalg.inv(np.asarray(pose, dtype=np.float64))
    return P[:3, :3].copy(), P[:3, 3].copy()


def _proj(R, t, X):
    Xc = X @ R.T + t
    return Xc[:, :2] / Xc[:, 2:3], Xc


def _exp_so3(w):
    th2 = float(w @ w)
    th = np.sqrt(th2)
    if th < 1e-8:
        A, B = 1.0 - th2 / 6.0, 0.5 - th2 / 24.0
    else:
        A, B = np.sin(th) / th, (1.0 - np.cos(th)) / th2
    Kx = np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])
    return np.eye(3) + A * Kx + B * (Kx @ Kx)


def cost(R1, t1, R2, t2, X, x1, x2):
    p1, _ = _proj(R1, t1, X)
    p2, _ = _proj(R2, t2, X)
    return float(np.sum((p1 - x1) ** 2) + np.sum((p2 - x2) ** 2))


def _blocks(R, t, X, x):
    """residuals r[n,2], d r / d X  [n,2,3], d r / d twist [n,2,6]"""
    p, Xc = _proj(R, t, X)
    iz = 1.0 / Xc[:, 2]
    n = len(X)
    Jpi = np.zeros((n, 2, 3))
    Jpi[:, 0, 0] = iz
    Jpi[:, 1, 1] = iz
    Jpi[:, 0, 2] = -p[:, 0] * iz
    Jpi[:, 1, 2] = -p[:, 1] * iz
    Jp = Jpi @ R
    mx = np.zeros((n, 3, 3))          # -[Xc]x
    mx[:, 0, 1], mx[:, 0, 2] = Xc[:, 2], -Xc[:, 1]
    mx[:, 1, 0], mx[:, 1, 2] = -Xc[:, 2], Xc[:, 0]
    mx[:, 2, 0], mx[:, 2, 1] = Xc[:, 1], -Xc[:, 0]
    Jc = np.concatenate([Jpi @ mx, Jpi], axis=2)
    return p - x, Jp, Jc


def bundle_adjust_pair(pix1, pix2, pose1, pose2, points4d, K, use=None, iters=10, lambda0=1e-3):
    """Returns (pose2_refined f64[4,4], points4d_refined f64[n,4], (cost0, cost1), accepted)."""
    Kinv = np.linalg.inv(np.asarray(K, dtype=np.float64))
    pix1, pix2 = np.asarray(pix1, dtype=np.float64), np.asarray(pix2, dtype=np.float64)
    sel = np.ones(len(pix1), dtype=bool) if use is None else np.asarray(use).astype(bool)
    h = lambda px: (np.c_[px, np.ones(len(px))] @ Kinv.T)[:, :2]
    x1, x2 = h(pix1[sel]), h(pix2[sel])
    R1, t1 = _cam(pose1)
    R2, t2 = _cam(pose2)
    pts = np.array(points4d, dtype=np.float64)
    X = pts[sel, :3].copy()
    lam = float(lambda0)
    c_now = cost(R1, t1, R2, t2, X, x1, x2)
    c0, accepted = c_now, 0
    for _ in range(iters):
        r1, Jp1, _ = _blocks(R1, t1, X, x1)
        r2, Jp2, Jc = _blocks(R2, t2, X, x2)
        V = np.einsum("nki,nkj->nij", Jp1, Jp1) + np.einsum("nki,nkj->nij", Jp2, Jp2)
        gp = np.einsum("nki,nk->ni", Jp1, r1) + np.einsum("nki,nk->ni", Jp2, r2)
        idx = np.arange(3)
        V[:, idx, idx] *= (1.0 + lam)
        Vi = np.linalg.inv(V)
        W = np.einsum("nka,nkj->naj", Jc, Jp2)
        U = np.einsum("nka,nkb->ab", Jc, Jc)
        gc = np.einsum("nka,nk->a", Jc, r2)
        WV = W @ Vi
        S = U.copy()
        S[np.arange(6), np.arange(6)] *= (1.0 + lam)
        S -= np.einsum("nak,nbk->ab", WV, W)
        rhs = -(gc - np.einsum("nak,nk->a", WV, gp))
        try:
            d = np.linalg.solve(S, rhs)
        except np.linalg.LinAlgError:
            d = np.full(6, np.nan)
        ok = np.isfinite(d).all()
        if ok:
            E = _exp_so3(d[:3])
            R2n, t2n = E @ R2, E @ t2 + d[3:]
            e = gp + np.einsum("nak,a->nk", W, d)
            Xn = X - np.einsum("nij,nj->ni", Vi, e)
            c_new = cost(R1, t1, R2n, t2n, Xn, x1, x2)
        if ok and c_new < c_now:
            R2, t2, X, c_now = R2n, t2n, Xn, c_new
            lam = max(lam * 0.1, 1e-15)
            accepted += 1
        else:
            lam = min(lam * 10.0, 1e12)
    P = np.eye(4)
    P[:3, :3], P[:3, 3] = R2, t2
    pts[sel, :3] = X
    return np.linalg.inv(P), pts, (c0, c_now), accepted
