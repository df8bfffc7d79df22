"""Seeded synthetic two-view inputs (keypoints + 256-bit descriptors + sample tables).

ORB detection is outside the hot path (BASELINE.json north_star), so every input to
the path is synthetic: a 3-D scene observed by a forward-moving camera, one 256-bit
descriptor per scene point with per-observation bit flips, clutter keypoints, and a
seeded RANSAC sample table.  The recipe follows SURVEY.md §8(d).  The same bytes are
fed to the CUDA path and to the CPU oracle, which is what makes parity checkable.

Array layouts are the reference's (``/root/reference/MonocularSlam/slam.py:121-122``):
``kps f64[N,2]`` pixel (x, y), ``des u8[N,32]`` C-contiguous, ``pose f64[4,4]``.
"""
from __future__ import annotations

import numpy as np

# K for the 640x480 configs; F=270 is the reference's focal (slam.py:278).
K_640x480 = np.array([[270, 0, 320], [0, 270, 240], [0, 0, 1]], dtype=np.int64)


def _small_rotation(rng, max_angle):
    w = rng.uniform(-max_angle, max_angle, 3)
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)


def _flip_masks(rng, n, max_flips):
    """u8[n,32] masks with k ~ U{0..max_flips} distinct random bits set per row."""
    if n == 0:
        return np.zeros((0, 32), dtype=np.uint8)
    k = rng.integers(0, max_flips + 1, n)
    order = np.argsort(rng.random((n, 256)), axis=1)            # a random permutation of the 256 bits per row
    rank = np.empty_like(order)
    np.put_along_axis(rank, order, np.arange(256)[None, :].repeat(n, 0), axis=1)
    bits = (rank < k[:, None]).astype(np.uint8)                 # the first k bits of the permutation
    return np.packbits(bits, axis=1, bitorder="little")


def camera_track(n_frames, seed, step=(0.3, 0.7), max_angle=0.01):
    """World->camera transforms T_i (x_cam = R_i x_world + t_i) of a forward-moving camera."""
    rng = np.random.default_rng(seed)
    Ts = [np.eye(4)]
    for _ in range(n_frames - 1):
        R = _small_rotation(rng, max_angle)
        # forward motion in the camera frame: the scene moves towards the camera (-z)
        t = np.array([rng.uniform(-0.03, 0.03), rng.uniform(-0.03, 0.03), -rng.uniform(*step)])
        d = np.eye(4)
        d[:3, :3] = R
        d[:3, 3] = t
        Ts.append(d @ Ts[-1])
    return Ts


def make_sequence(n_frames, n_kpts, seed, K=K_640x480, width=640, height=480,
                  clutter=0.20, outlier=0.30, pix_sigma=0.3, outlier_sigma=25.0,
                  max_flips=49, forced_ties=0):
    """One synthetic sequence: list of (kps f64[N,2], des u8[N,32]) per frame.

    Every frame observes the same pool of scene points (re-sampled as the camera moves
    so each frame has exactly ``n_kpts`` rows), with descriptor bit flips, pixel noise,
    geometric outliers and clutter rows, shuffled per frame.  ``forced_ties`` duplicates
    that many descriptor rows inside a frame (exact Hamming ties) -- used by parity
    inputs, never by timing inputs.
    """
    rng = np.random.default_rng(seed)
    Kf = np.asarray(K, dtype=np.float64)
    Ts = camera_track(n_frames, seed + 7919)
    n_true = int(round(n_kpts * (1.0 - clutter)))
    frames = []
    # a scene deep enough that points stay in view for the whole track
    depth_hi = 40.0 + 0.7 * n_frames
    n_pool = (8 + n_frames // 6) * n_kpts
    pts = np.empty((n_pool, 3))
    pts[:, 2] = rng.uniform(4.0, depth_hi, n_pool)
    pts[:, 0] = rng.uniform(-1.1, 1.1, n_pool) * pts[:, 2] * (width / 2) / Kf[0, 0]
    pts[:, 1] = rng.uniform(-1.1, 1.1, n_pool) * pts[:, 2] * (height / 2) / Kf[1, 1]
    base_des = rng.integers(0, 256, (n_pool, 32), dtype=np.uint8)
    prev_ids = None
    for T in Ts:
        pc = pts @ T[:3, :3].T + T[:3, 3]
        z = pc[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = Kf[0, 0] * pc[:, 0] / z + Kf[0, 2]
            v = Kf[1, 1] * pc[:, 1] / z + Kf[1, 2]
        vis = (z > 1.0) & (u >= 0) & (u < width) & (v >= 0) & (v < height)
        ids = np.flatnonzero(vis)
        # prefer points tracked from the previous frame so consecutive frames overlap
        if prev_ids is not None:
            keep = np.intersect1d(ids, prev_ids)
            rest = np.setdiff1d(ids, keep)
            rng.shuffle(rest)
            ids = np.concatenate([keep, rest])
        else:
            rng.shuffle(ids)
        ids = ids[:n_true]
        prev_ids = ids
        n_t = len(ids)
        kps = np.stack([u[ids], v[ids]], axis=1) + rng.normal(0, pix_sigma, (n_t, 2))
        bad = rng.random(n_t) < outlier
        kps[bad] += rng.normal(0, outlier_sigma, (int(bad.sum()), 2))
        des = base_des[ids] ^ _flip_masks(rng, n_t, max_flips)
        n_c = n_kpts - n_t
        ckps = np.stack([rng.uniform(0, width, n_c), rng.uniform(0, height, n_c)], axis=1)
        cdes = rng.integers(0, 256, (n_c, 32), dtype=np.uint8)
        kps = np.concatenate([kps, ckps])
        des = np.concatenate([des, cdes])
        if forced_ties:
            src = rng.choice(n_kpts, forced_ties, replace=False)
            dst = rng.choice(n_kpts, forced_ties, replace=False)
            des[dst] = des[src]
        perm = rng.permutation(n_kpts)
        frames.append((np.ascontiguousarray(kps[perm]), np.ascontiguousarray(des[perm])))
    return frames, Ts


def make_pair(n_kpts, seed, **kw):
    """One frame pair: (last_kps, last_des, curr_kps, curr_des)."""
    frames, _ = make_sequence(2, n_kpts, seed, **kw)
    (lk, ld), (ck, cd) = frames
    return lk, ld, ck, cd


def make_pair_large(n, seed):
    """(last_kps, last_des, curr_kps, curr_des) with n keypoints per side: one two-view scene like
    make_pair, sized for the 8k / 64k keyframe configs (80 % true points, 30 % geometric outliers)."""
    rng = np.random.default_rng(seed)
    K = K_640x480.astype(float)
    n_true = int(n * 0.8)
    X = np.stack([rng.uniform(-1, 1, n_true), rng.uniform(-1, 1, n_true), rng.uniform(4, 40, n_true)], axis=1)
    X[:, 0] *= X[:, 2] * 320 / 270
    X[:, 1] *= X[:, 2] * 240 / 270
    R = _small_rotation(rng, 0.01)
    t = np.array([0.02, -0.01, -0.5])
    def proj(P):
        return np.stack([K[0, 0] * P[:, 0] / P[:, 2] + K[0, 2], K[1, 1] * P[:, 1] / P[:, 2] + K[1, 2]], axis=1)
    k1 = proj(X) + rng.normal(0, 0.3, (n_true, 2))
    k2 = proj(X @ R.T + t) + rng.normal(0, 0.3, (n_true, 2))
    bad = rng.random(n_true) < 0.3
    k2[bad] += rng.normal(0, 25, (int(bad.sum()), 2))
    d1 = rng.integers(0, 256, (n_true, 32), dtype=np.uint8)
    d2 = d1 ^ _flip_masks(rng, n_true, 49)
    nc = n - n_true
    k1 = np.concatenate([k1, rng.uniform(0, 640, (nc, 2))])
    k2 = np.concatenate([k2, rng.uniform(0, 640, (nc, 2))])
    d1 = np.concatenate([d1, rng.integers(0, 256, (nc, 32), dtype=np.uint8)])
    d2 = np.concatenate([d2, rng.integers(0, 256, (nc, 32), dtype=np.uint8)])
    p1, p2 = rng.permutation(n), rng.permutation(n)
    return (np.ascontiguousarray(k1[p1]), np.ascontiguousarray(d1[p1]), np.ascontiguousarray(k2[p2]),
            np.ascontiguousarray(d2[p2]))


def random_descriptors(n, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, 32), dtype=np.uint8)


def correlated_descriptors(nq, nt, seed, max_flips=49, clutter=0.2):
    """Query/train descriptor sets with true matches (for matcher-only configs at 8k/64k)."""
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    src = rng.permutation(nt)[:nq] if nq <= nt else rng.integers(0, nt, nq)
    q = t[src] ^ _flip_masks(rng, nq, max_flips)
    c = rng.random(nq) < clutter
    q[c] = rng.integers(0, 256, (int(c.sum()), 32), dtype=np.uint8)
    return np.ascontiguousarray(q), np.ascontiguousarray(t)


def sample_table(M, H, seed):
    """Seeded ``int32[H,8]`` table of 8 distinct indices in ``[0,M)`` per hypothesis.

    Row h plays the role of skimage's per-trial ``rng.choice(num_samples, 8, replace=False)``
    (SURVEY.md Appendix A.3); supplying it from the host makes GPU and oracle hypotheses
    identical (north_star (2)).
    """
    if M < 8:
        raise ValueError("need at least 8 correspondences")
    g = np.random.default_rng(seed)
    # argsort of random keys = H independent permutations; take the first 8 of each
    if M <= 4096:
        return np.argsort(g.random((H, M)), axis=1)[:, :8].astype(np.int32)
    out = np.empty((H, 8), dtype=np.int32)
    for h in range(H):
        out[h] = g.choice(M, 8, replace=False)
    return out


def sample_words(H, seed):
    """Seeded ``u32[H,8]`` random words for the sync-free sample form (mapped on device
    and in the oracle by ``idx = (word*M)>>32`` with in-row duplicates bumped ``+1 mod M``)."""
    return np.random.default_rng(seed).integers(0, 2**32, (H, 8), dtype=np.uint64).astype(np.uint32)


def samples_from_words(words, M):
    """Map random words to 8 distinct indices in [0,M) (M >= 8).  Device does the same."""
    words = np.asarray(words, dtype=np.uint64)
    idx = ((words * np.uint64(M)) >> np.uint64(32)).astype(np.int64)
    H = idx.shape[0]
    for k in range(1, 8):
        # bump slot k until it differs from all earlier slots (slot order, +1 mod M)
        while True:
            clash = np.zeros(H, dtype=bool)
            for j in range(k):
                clash |= idx[:, k] == idx[:, j]
            if not clash.any():
                break
            idx[clash, k] = (idx[clash, k] + 1) % M
    return idx.astype(np.int32)
