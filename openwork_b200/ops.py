"""Array-level operators of the two-view path (explicit arguments, no globals).

Each function is the array form of one reference call and goes straight to the
sm_100a kernels through the C ABI (``include/tvg_b200.h``).  Inputs may be numpy arrays or torch
CUDA tensors; outputs are torch CUDA tensors (use ``.cpu().numpy()`` for numpy).  Nothing here
computes on the CPU: without the built library and a B200 every call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import HugePairIO, PairsIO, RansacIO, check

RATIO = 0.75                # slam.py:131
RESIDUAL_THRESHOLD = 0.005  # slam.py:158

_contexts = {}


def context(device=None):
    """Per-device default context (created on first use; raises without a B200)."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    device = int(device)
    if device not in _contexts:
        _contexts[device] = _lib.Context(device)
    return _contexts[device]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(x, dtype, device):
    """numpy / tensor -> contiguous CUDA tensor of dtype (uint32 is carried as int32 bits)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if x.dtype == np.uint32:
            x = x.view(np.int32)
        x = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None and x.dtype != dtype:
        x = x.to(dtype)
    return x.to(device, non_blocking=True).contiguous()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _kinv(K):
    """inv(K) exactly as the reference computes it (np.linalg.inv on the int64 K, slam.py:88)."""
    kinv = np.ascontiguousarray(np.linalg.inv(np.asarray(K)), dtype=np.float64)
    return kinv, kinv.ctypes.data_as(C.POINTER(C.c_double))


def _device_of(*xs):
    for x in xs:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    return torch.device("cuda", torch.cuda.current_device())


# ------------------------------------------------------------------------------- matcher
def knn2_hamming256(q, t, ratio=RATIO, return_keep=False, ctx=None):
    """``cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)`` (slam.py:126-127) as arrays.

    q u8[Nq,32], t u8[Nt,32] -> idx int32[Nq,2], dist int32[Nq,2] (-1 where Nt < 2), and with
    ``return_keep`` the Lowe ratio flags ``dist0 < ratio*dist1`` (slam.py:131).
    """
    dev = _device_of(q, t)
    q = _dev(q, torch.uint8, dev)
    t = _dev(t, torch.uint8, dev)
    if q.dim() != 2 or t.dim() != 2 or q.shape[1] != 32 or t.shape[1] != 32:
        raise ValueError("descriptors must be u8[N,32]")
    nq, nt = q.shape[0], t.shape[0]
    ctx = ctx or context(dev.index)
    n_max = max(nq, nt, 1)
    idx = torch.full((n_max, 2), -1, dtype=torch.int32, device=dev)
    dist = torch.full((n_max, 2), -1, dtype=torch.int32, device=dev)
    keep = torch.zeros(n_max, dtype=torch.uint8, device=dev)
    if nq > 0:
        if nt == 0:
            t = torch.zeros((1, 32), dtype=torch.uint8, device=dev)
        check(ctx.lib.tvg_knn2_hamming256(ctx.handle, _p(q), _p(t), 1, n_max, nq, nt, None, None, None, None,
                                          float(ratio), _p(idx), _p(dist), _p(keep), _stream()),
              "tvg_knn2_hamming256")
    if return_keep:
        return idx[:nq], dist[:nq], keep[:nq].bool()
    return idx[:nq], dist[:nq]


def knn2_hamming256_batched(des, pair_q, pair_t, frame_count=None, ratio=RATIO, ctx=None):
    """Batched matcher over a frame pool des u8[F,N,32]: pair b = (query frame pair_q[b], train frame
    pair_t[b]).  Returns idx, dist int32[B,N,2] and keep bool[B,N]."""
    dev = _device_of(des)
    des = _dev(des, torch.uint8, dev)
    pq = _dev(pair_q, torch.int32, dev)
    pt = _dev(pair_t, torch.int32, dev)
    fc = _dev(frame_count, torch.int32, dev)
    F, N, _ = des.shape
    B = pq.shape[0]
    ctx = ctx or context(dev.index)
    idx = torch.empty((B, N, 2), dtype=torch.int32, device=dev)
    dist = torch.empty((B, N, 2), dtype=torch.int32, device=dev)
    keep = torch.zeros((B, N), dtype=torch.uint8, device=dev)
    check(ctx.lib.tvg_knn2_hamming256(ctx.handle, _p(des), _p(des), B, N, N, N, _p(pq), _p(pt), _p(fc), _p(fc),
                                      float(ratio), _p(idx), _p(dist), _p(keep), _stream()),
          "tvg_knn2_hamming256")
    return idx, dist, keep.bool()


def ratio_compact(idx, keep, curr_kps=None, last_kps=None, K=None, ctx=None):
    """Ratio loop + gathers of match_points (slam.py:130-141) and the K-normalisation of
    fit_essential_matrix (slam.py:151-152) for ONE pair.

    Returns dict(qidx, tidx int32[M], M, and when keypoints are given src, dst f64[M,2] (src = last,
    dst = curr, K-normalised) and corr32 f32[M,4]).  Synchronises once to learn M.
    """
    dev = _device_of(idx)
    idx = _dev(idx, torch.int32, dev)
    keep_u8 = _dev(keep, None, dev).to(torch.uint8).contiguous()
    nq = idx.shape[0]
    ctx = ctx or context(dev.index)
    gather = curr_kps is not None
    n_max = max(nq, 1)
    if gather:
        ck = _dev(curr_kps, torch.float64, dev)
        lk = _dev(last_kps, torch.float64, dev)
        kinv, kinv_p = _kinv(K)
        src = torch.empty((n_max, 2), dtype=torch.float64, device=dev)
        dst = torch.empty((n_max, 2), dtype=torch.float64, device=dev)
        corr = torch.empty((n_max, 4), dtype=torch.float32, device=dev)
    else:
        ck = lk = src = dst = corr = None
        kinv_p = None
    qidx = torch.empty(n_max, dtype=torch.int32, device=dev)
    tidx = torch.empty(n_max, dtype=torch.int32, device=dev)
    mcnt = torch.zeros(1, dtype=torch.int32, device=dev)
    if nq > 0:
        check(ctx.lib.tvg_ratio_compact(ctx.handle, _p(keep_u8), _p(idx), 1, n_max, nq, None, None, None, _p(ck),
                                        _p(lk), kinv_p, _p(qidx), _p(tidx), _p(mcnt), _p(src), _p(dst), _p(corr),
                                        _stream()), "tvg_ratio_compact")
    M = int(mcnt.item())
    out = dict(qidx=qidx[:M], tidx=tidx[:M], M=M)
    if gather:
        out.update(src=src[:M], dst=dst[:M], corr32=corr[:M])
    return out


# ------------------------------------------------------------------------------- RANSAC
def ransac_essential(src, dst, samples=None, words=None, residual_threshold=RESIDUAL_THRESHOLD,
                     return_models=False, ctx=None):
    """``skimage.measure.ransac((src, dst), EssentialMatrixTransform, min_samples=8,
    residual_threshold, max_trials=H)`` (slam.py:155-159) with a host-supplied sample table
    (``samples int32[H,8]`` or ``words u32[H,8]``).  ONE pair; src/dst f64[M,2] K-normalised.

    Returns dict(E f64[3,3] refit, mask bool[M] of the PRE-refit best model, counts int32[H], sse32,
    best, E_best, n_inl, Rt f64[4,4], band, n_cand, samples int32[H,8], [models f64[H,3,3]]).
    Raises ValueError when M <= 8 (skimage: ``min_samples`` must be < number of samples).
    """
    dev = _device_of(src, dst)
    src = _dev(src, torch.float64, dev)
    dst = _dev(dst, torch.float64, dev)
    M = src.shape[0]
    if not (0 < 8 < M):
        raise ValueError("`min_samples` must be in range (0, <number-of-samples>)")
    ctx = ctx or context(dev.index)
    corr = torch.cat([src, dst], dim=1).to(torch.float32).contiguous()
    if samples is not None:
        tab = _dev(np.asarray(samples, dtype=np.int32) if not isinstance(samples, torch.Tensor) else samples,
                   torch.int32, dev)
    else:
        tab = _dev(np.asarray(words, dtype=np.uint32) if not isinstance(words, torch.Tensor) else words,
                   torch.int32, dev)
    H = tab.shape[0]
    o = dict(
        E=torch.empty(9, dtype=torch.float64, device=dev), mask=torch.zeros(M, dtype=torch.uint8, device=dev),
        n_inl=torch.zeros(1, dtype=torch.int32, device=dev), best=torch.zeros(1, dtype=torch.int32, device=dev),
        E_best=torch.empty(9, dtype=torch.float64, device=dev), counts=torch.empty(H, dtype=torch.int32, device=dev),
        sse32=torch.empty(H, dtype=torch.float32, device=dev), Rt=torch.empty(16, dtype=torch.float64, device=dev),
        status=torch.zeros(1, dtype=torch.int32, device=dev), inl_idx=torch.empty(M, dtype=torch.int32, device=dev),
        samples_out=torch.empty((H, 8), dtype=torch.int32, device=dev),
        band=torch.empty(H, dtype=torch.int32, device=dev), n_cand=torch.zeros(1, dtype=torch.int32, device=dev),
        best_sse=torch.zeros(1, dtype=torch.float64, device=dev))
    if return_models:
        o["models"] = torch.empty((H, 3, 3), dtype=torch.float64, device=dev)
    io = RansacIO()
    io.src, io.dst, io.corr32 = src.data_ptr(), dst.data_ptr(), corr.data_ptr()
    io.batch, io.n_max, io.m, io.H = 1, M, M, H
    if samples is not None:
        io.samples = tab.data_ptr()
    else:
        io.words = tab.data_ptr()
    io.sample_stride = 0
    io.thr = float(residual_threshold)
    for k, v in o.items():
        setattr(io, k, v.data_ptr())
    check(ctx.lib.tvg_ransac_essential(ctx.handle, C.byref(io), _stream()), "tvg_ransac_essential")
    n_inl = int(o["n_inl"].item())
    res = dict(E=o["E"].view(3, 3), mask=o["mask"].bool(), counts=o["counts"], sse32=o["sse32"],
               best=int(o["best"].item()), E_best=o["E_best"].view(3, 3), n_inl=n_inl, Rt=o["Rt"].view(4, 4),
               status=int(o["status"].item()), inl_idx=o["inl_idx"][:n_inl], samples=o["samples_out"],
               band=o["band"], n_cand=int(o["n_cand"].item()), best_sse=float(o["best_sse"].item()))
    if return_models:
        res["models"] = o["models"]
    return res


# ------------------------------------------------------------------------------- pose + triangulation
def extract_rt(E, ctx=None):
    """``extract_Rt(E)`` (slam.py:167-186).  E f64[3,3] or [B,3,3] -> Rt f64[4,4] or [B,4,4]."""
    dev = _device_of(E)
    E = _dev(E, torch.float64, dev)
    single = E.dim() == 2
    Eb = E.reshape(-1, 9).contiguous()
    B = Eb.shape[0]
    ctx = ctx or context(dev.index)
    Rt = torch.empty((B, 16), dtype=torch.float64, device=dev)
    check(ctx.lib.tvg_extract_rt(ctx.handle, _p(Eb), B, _p(Rt), _stream()), "tvg_extract_rt")
    return Rt.view(4, 4) if single else Rt.view(B, 4, 4)


def pose_chain(Rt, seq_off=None, init_pose=None, ctx=None):
    """``curr_pose = Rt @ last_pose`` (slam.py:261) chained over sequences.
    Rt f64[B,4,4]; returns (poses f64[B+nseq,4,4], pair_pinv f64[B,2,4,4])."""
    dev = _device_of(Rt)
    Rt = _dev(Rt, torch.float64, dev).reshape(-1, 16).contiguous()
    B = Rt.shape[0]
    if seq_off is None:
        seq_off = np.array([0, B], dtype=np.int32)
    so = _dev(np.asarray(seq_off, dtype=np.int32) if not isinstance(seq_off, torch.Tensor) else seq_off,
              torch.int32, dev)
    nseq = so.shape[0] - 1
    ip = _dev(init_pose, torch.float64, dev)
    ctx = ctx or context(dev.index)
    poses = torch.empty((B + nseq, 16), dtype=torch.float64, device=dev)
    pinv = torch.empty((B, 32), dtype=torch.float64, device=dev)
    check(ctx.lib.tvg_pose_chain(ctx.handle, _p(Rt), _p(so), nseq, _p(ip), _p(poses), _p(pinv), _stream()),
          "tvg_pose_chain")
    return poses.view(-1, 4, 4), pinv.view(B, 2, 4, 4)


def _invert_poses(poses, dev, ctx):
    """np.linalg.inv of a stack of 4x4 poses, on the device (tvg_invert_poses)."""
    P = _dev(np.ascontiguousarray(poses, dtype=np.float64), torch.float64, dev).reshape(-1, 16).contiguous()
    out = torch.empty_like(P)
    check(ctx.lib.tvg_invert_poses(ctx.handle, _p(P), P.shape[0], _p(out), _stream()), "tvg_invert_poses")
    return out


def triangulate(pts1, pts2, pose1, pose2, K, return_good=False, ctx=None):
    """``triangulate(pts1, pts2, pose1, pose2)`` (slam.py:202-221) with K explicit.

    pts f64[M,2] pixels, pose f64[4,4] (camera -> world; inverted on the device like slam.py:204-205).
    Returns points4d f64[M,4] (and z>0 flags)."""
    dev = _device_of(pts1, pts2)
    p1 = _dev(pts1, torch.float64, dev)
    p2 = _dev(pts2, torch.float64, dev)
    M = p1.shape[0]
    ctx = ctx or context(dev.index)
    Pd = _invert_poses(np.stack([np.asarray(_np(pose1), dtype=np.float64), np.asarray(_np(pose2), dtype=np.float64)]),
                       dev, ctx)
    kinv, kinv_p = _kinv(K)
    X = torch.empty((max(M, 1), 4), dtype=torch.float64, device=dev)
    good = torch.zeros(max(M, 1), dtype=torch.uint8, device=dev)
    if M > 0:
        check(ctx.lib.tvg_triangulate(ctx.handle, _p(p1), _p(p2), 1, M, M, None, C.c_void_p(Pd.data_ptr()), 32,
                                      C.c_void_p(Pd.data_ptr() + 128), 32, kinv_p, _p(X), _p(good), _stream()),
              "tvg_triangulate")
    if return_good:
        return X[:M], good[:M].bool()
    return X[:M]


def triangulate_batched(pix1, pix2, n_pts, pair_pinv, K, ctx=None):
    """Batched triangulation: pix f64[B,N,2], n_pts int32[B], pair_pinv f64[B,2,4,4]."""
    dev = _device_of(pix1)
    pix1 = _dev(pix1, torch.float64, dev)
    pix2 = _dev(pix2, torch.float64, dev)
    n_pts = _dev(n_pts, torch.int32, dev)
    pinv = _dev(pair_pinv, torch.float64, dev).reshape(-1, 32).contiguous()
    B, N, _ = pix1.shape
    ctx = ctx or context(dev.index)
    kinv, kinv_p = _kinv(K)
    X = torch.empty((B, N, 4), dtype=torch.float64, device=dev)
    good = torch.zeros((B, N), dtype=torch.uint8, device=dev)
    check(ctx.lib.tvg_triangulate(ctx.handle, _p(pix1), _p(pix2), B, N, N, _p(n_pts), C.c_void_p(pinv.data_ptr()),
                                  32, C.c_void_p(pinv.data_ptr() + 128), 32, kinv_p, _p(X), _p(good), _stream()),
          "tvg_triangulate")
    return X, good.bool()


def point_quality(pts1, pts2, pose1, pose2, K, ctx=None):
    """Triangulation plus the filters slam.py:236 leaves as TODO (spec: oracle.two_view.point_quality):
    returns (points4d f64[M,4], reproj_px f64[M], parallax_cos f64[M])."""
    dev = _device_of(pts1, pts2)
    p1 = _dev(pts1, torch.float64, dev)
    p2 = _dev(pts2, torch.float64, dev)
    M = p1.shape[0]
    ctx = ctx or context(dev.index)
    Pd = _invert_poses(np.stack([np.asarray(_np(pose1), dtype=np.float64), np.asarray(_np(pose2), dtype=np.float64)]),
                       dev, ctx)
    kinv, kinv_p = _kinv(K)
    X = torch.empty((max(M, 1), 4), dtype=torch.float64, device=dev)
    rep = torch.empty(max(M, 1), dtype=torch.float64, device=dev)
    par = torch.empty(max(M, 1), dtype=torch.float64, device=dev)
    if M > 0:
        check(ctx.lib.tvg_point_quality(ctx.handle, _p(p1), _p(p2), 1, M, M, None, C.c_void_p(Pd.data_ptr()), 32,
                                        C.c_void_p(Pd.data_ptr() + 128), 32, kinv_p, float(np.asarray(K)[0, 0]),
                                        _p(X), None, _p(rep), _p(par), _stream()), "tvg_point_quality")
    return X[:M], rep[:M], par[:M]


def map_accumulate(points4d, good, n_pts, seq_off=None, cap=None, include_origin=True, ctx=None):
    """Point-cloud append of ``Map.add_observation`` (slam.py:37-40) for whole sequences on the device.

    points4d f64[B,N,4], good bool/u8[B,N], n_pts int32[B] (outputs of process_pairs_device).  Returns
    (map_points f64[nseq,cap,4], map_count int32[nseq], pair_off int32[B]): per sequence the origin point of
    the first frame (slam.py:248) followed by every pair's z>0 points in pair order."""
    dev = _device_of(points4d)
    pts = _dev(points4d, torch.float64, dev)
    B, N, _ = pts.shape
    g8 = _dev(good, None, dev).to(torch.uint8).contiguous()
    n = _dev(n_pts, torch.int32, dev)
    if seq_off is None:
        seq_off = np.array([0, B], dtype=np.int32)
    so = seq_off if isinstance(seq_off, torch.Tensor) else np.asarray(seq_off, dtype=np.int32)
    so = _dev(so, torch.int32, dev)
    nseq = so.shape[0] - 1
    if cap is None:
        so_h = _np(so)
        cap = int(max(so_h[1:] - so_h[:-1])) * N + 1
    ctx = ctx or context(dev.index)
    mp = torch.empty((nseq, cap, 4), dtype=torch.float64, device=dev)
    cnt = torch.zeros(nseq, dtype=torch.int32, device=dev)
    off = torch.zeros(B, dtype=torch.int32, device=dev)
    check(ctx.lib.tvg_map_accumulate(ctx.handle, _p(pts), _p(g8), _p(n), _p(so), nseq, N, cap,
                                     1 if include_origin else 0, _p(mp), _p(cnt), _p(off), _stream()),
          "tvg_map_accumulate")
    return mp, cnt, off


def run_sequences(des, kps, K, words, seq_off=None, residual_threshold=RESIDUAL_THRESHOLD, ratio=RATIO, ctx=None):
    """Whole sequences without host round-trips: what the reference's ``__main__`` loop accumulates in ``Map``
    (slam.py:283-293, 37-40, 81-84).  des u8[F,N,32], kps f64[F,N,2] with the frames of sequence s stored
    consecutively.  Returns a list of (poses f64[F_s,4,4], points f64[P_s,4]) per sequence (CUDA tensors)."""
    out = process_pairs_device(des, kps, K, words=words, seq_off=seq_off, ratio=ratio,
                               residual_threshold=residual_threshold,
                               outputs=("n_inl", "poses", "points4d", "good", "status"), ctx=ctx)
    B = out["n_inl"].shape[0]
    so = np.array([0, B], dtype=np.int32) if seq_off is None else np.asarray(_np(seq_off), dtype=np.int32)
    mp, cnt, _ = map_accumulate(out["points4d"], out["good"], out["n_inl"], so, ctx=ctx)
    cnt = cnt.cpu().numpy()
    res = []
    for s in range(len(so) - 1):
        p0 = so[s] + s
        res.append((out["poses"][p0:p0 + (so[s + 1] - so[s]) + 1], mp[s, :cnt[s]]))
    return res


_ORB_PATTERN = None


def orb_pattern():
    """int8[512,2] rBRIEF test points rotated by cv2.KeyPoint's default angle (-1 degree) and rounded, exactly as
    cv2's computeOrbDescriptors does for keypoints without orientation (ORB.compute never computes one)."""
    global _ORB_PATTERN
    if _ORB_PATTERN is None:
        import os
        raw = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "orb_pattern.npy")).astype(np.float32)
        ang = np.float32(-1.0) * np.float32(np.pi / 180.0)
        a, b = np.float32(np.cos(ang)), np.float32(np.sin(ang))
        x = raw[:, 0] * a - raw[:, 1] * b
        y = raw[:, 0] * b + raw[:, 1] * a
        _ORB_PATTERN = np.ascontiguousarray(np.stack([np.rint(x), np.rint(y)], axis=1).astype(np.int8))
    return _ORB_PATTERN


def extract_points(images, max_corners=3000, quality_level=0.01, min_distance=3, ctx=None):
    """``extract_points(frame)`` (slam.py:112-122) for a batch of images on the device: BGR2GRAY,
    goodFeaturesToTrack(max_corners, quality_level, min_distance), ORB.compute at the found corners.

    images u8[B,H,W,3] (BGR) or u8[B,H,W] (gray); a single image may omit B.  Returns (kps f64[B,max,2],
    des u8[B,max,32], n int32[B]) as CUDA tensors; rows beyond n[b] are unspecified."""
    dev = _device_of(images)
    img = _dev(images, torch.uint8, dev)
    single = img.dim() == 2 or (img.dim() == 3 and img.shape[-1] == 3)
    if single:
        img = img.unsqueeze(0)
    channels = 3 if (img.dim() == 4) else 1
    B, H, W = img.shape[0], img.shape[1], img.shape[2]
    ctx = ctx or context(dev.index)
    kps = torch.empty((B, max_corners, 2), dtype=torch.float64, device=dev)
    des = torch.empty((B, max_corners, 32), dtype=torch.uint8, device=dev)
    n = torch.zeros(B, dtype=torch.int32, device=dev)
    pat = orb_pattern()
    check(ctx.lib.tvg_extract_points(ctx.handle, _p(img), B, H, W, channels, int(max_corners), float(quality_level),
                                     float(min_distance), C.c_void_p(pat.ctypes.data), _p(kps), _p(des), _p(n), None,
                                     _stream()), "tvg_extract_points")
    return kps, des, n


def run_sequences_from_images(images, K, words, seq_off=None, max_corners=3000, quality_level=0.01, min_distance=3,
                              residual_threshold=RESIDUAL_THRESHOLD, ratio=RATIO, ctx=None):
    """The reference's whole per-frame loop (slam.py:239-272 for every frame of every sequence) on the device,
    from raw images: extract_points -> match -> RANSAC -> extract_Rt -> pose chain -> triangulate -> z>0 -> map.
    images u8[F,H,W,3]; frames of sequence s consecutive.  Returns [(poses, points)] like :func:`run_sequences`,
    plus the per-frame keypoint counts."""
    kps, des, n = extract_points(images, max_corners, quality_level, min_distance, ctx=ctx)
    F = kps.shape[0]
    out = process_pairs_device(des, kps, K, words=words, frame_count=n, seq_off=seq_off, ratio=ratio,
                               residual_threshold=residual_threshold,
                               outputs=("n_inl", "poses", "points4d", "good", "status"), ctx=ctx)
    B = out["n_inl"].shape[0]
    so = np.array([0, B], dtype=np.int32) if seq_off is None else np.asarray(_np(seq_off), dtype=np.int32)
    mp, cnt, _ = map_accumulate(out["points4d"], out["good"], out["n_inl"], so, ctx=ctx)
    cnt = cnt.cpu().numpy()
    res = []
    for s in range(len(so) - 1):
        p0 = so[s] + s
        res.append((out["poses"][p0:p0 + (so[s + 1] - so[s]) + 1], mp[s, :cnt[s]]))
    return res, n, out["status"]


def check_points(points4d):
    """``check_points`` (slam.py:233-237): world z > 0 (elementwise on the device tensor)."""
    return points4d[:, 2] > 0


def cheirality(E, src, dst, mask=None, ctx=None):
    """Cheirality vote over the four (R, +-t) decompositions of E (north-star extra, not in the
    reference).  Returns (votes int32[4], Rts f64[4,4,4], choice)."""
    dev = _device_of(E, src)
    E = _dev(E, torch.float64, dev).reshape(9).contiguous()
    src = _dev(src, torch.float64, dev)
    dst = _dev(dst, torch.float64, dev)
    m8 = None if mask is None else _dev(mask, None, dev).to(torch.uint8).contiguous()
    M = src.shape[0]
    ctx = ctx or context(dev.index)
    votes = torch.zeros(4, dtype=torch.int32, device=dev)
    Rts = torch.empty((4, 16), dtype=torch.float64, device=dev)
    choice = torch.zeros(1, dtype=torch.int32, device=dev)
    check(ctx.lib.tvg_cheirality(ctx.handle, _p(E), _p(src), _p(dst), _p(m8), None, 1, max(M, 1), M, _p(votes),
                                 _p(Rts), _p(choice), _stream()), "tvg_cheirality")
    return votes, Rts.view(4, 4, 4), int(choice.item())


def bundle_adjust_pairs(pix1, pix2, pose1, pose2, points4d, K, n_pts=None, use=None, iters=10, lambda0=1e-3, ctx=None):
    """Two-view bundle adjustment per pair (the `g2o bundle adjustment` TODO at slam.py:267; spec: oracle/bundle.py).

    pix1 / pix2 f64[B,N,2] pixels of the same points in the first / second view, pose1 / pose2 f64[B,4,4]
    camera-to-world (the reference's `pose`), points4d f64[B,N,4].  Returns (pose2_refined, points4d_refined,
    cost f64[B,2] = initial / final, accepted int32[B]); inputs are not modified."""
    dev = _device_of(pix1, pix2, points4d)
    pix1 = _dev(pix1, torch.float64, dev).contiguous()
    pix2 = _dev(pix2, torch.float64, dev).contiguous()
    B, N = pix1.shape[0], pix1.shape[1]
    pose1 = _dev(pose1, torch.float64, dev).reshape(B, 16).contiguous()
    pose2 = _dev(pose2, torch.float64, dev).reshape(B, 16).clone()
    pts = _dev(points4d, torch.float64, dev).reshape(B, N, 4).clone()
    cnt = None if n_pts is None else _dev(n_pts, torch.int32, dev).contiguous()
    u8 = None if use is None else _dev(use, None, dev).to(torch.uint8).contiguous()
    kinv = np.ascontiguousarray(np.linalg.inv(np.asarray(K, dtype=np.float64)))
    ctx = ctx or context(dev.index)
    cost = torch.zeros((B, 2), dtype=torch.float64, device=dev)
    acc = torch.zeros(B, dtype=torch.int32, device=dev)
    check(ctx.lib.tvg_bundle_adjust_pairs(ctx.handle, _p(pix1), _p(pix2), B, N, N, _p(cnt), _p(u8),
                                          kinv.ctypes.data_as(C.c_void_p), _p(pose1), 16, _p(pose2), 16, _p(pts),
                                          int(iters), float(lambda0), _p(cost), _p(acc), _stream()),
          "tvg_bundle_adjust_pairs")
    return pose2.view(B, 4, 4), pts, cost, acc


def bundle_adjust_window(obs, vis, poses, points4d, K, n_pts=None, n_fixed=1, iters=10, lambda0=1e-3, ctx=None):
    """Sliding-window bundle adjustment over W <= 8 consecutive frames (the multi-frame form of the `g2o bundle
    adjustment` TODO at slam.py:267; spec: oracle/bundle_window.py), batched over windows.

    obs f64[B,N,W,2] pixels of point j in camera c, vis bool/u8[B,N,W], poses f64[B,W,4,4] camera-to-world (the
    reference's `pose`), points4d f64[B,N,4]; the first ``n_fixed`` cameras stay put.  Returns (poses_refined,
    points4d_refined, cost f64[B,2] = initial / final, accepted int32[B]); inputs are not modified."""
    dev = _device_of(obs, poses, points4d)
    obs = _dev(obs, torch.float64, dev).contiguous()
    B, N, W = obs.shape[0], obs.shape[1], obs.shape[2]
    v8 = _dev(vis, None, dev).to(torch.uint8).contiguous()
    P = _dev(poses, torch.float64, dev).reshape(B, W, 16).clone()
    pts = _dev(points4d, torch.float64, dev).reshape(B, N, 4).clone()
    cnt = None if n_pts is None else _dev(n_pts, torch.int32, dev).contiguous()
    kinv = np.ascontiguousarray(np.linalg.inv(np.asarray(K, dtype=np.float64)))
    ctx = ctx or context(dev.index)
    cost = torch.zeros((B, 2), dtype=torch.float64, device=dev)
    acc = torch.zeros(B, dtype=torch.int32, device=dev)
    check(ctx.lib.tvg_bundle_adjust_window(ctx.handle, _p(obs), _p(v8), B, N, N, _p(cnt), int(W), int(n_fixed),
                                           kinv.ctypes.data_as(C.c_void_p), _p(P), _p(pts), int(iters), float(lambda0),
                                           _p(cost), _p(acc), _stream()), "tvg_bundle_adjust_window")
    return P.view(B, W, 4, 4), pts, cost, acc


def window_tracks(out, kps, b0, W):
    """Feature tracks over the W consecutive frames b0 .. b0+W-1 of ONE sequence, chained from the pair pipeline's
    inlier matches (outputs ``qidx, tidx, inl_idx, n_inl, points4d`` of :func:`process_pairs_device` with the default
    consecutive pairing): the inlier (t, q) of pair b links keypoint t of frame b to keypoint q of frame b+1; a track
    is extended while its keypoint keeps being matched (when several queries claim one train keypoint the lowest
    query wins).  Returns (obs f64[n,W,2] pixels, vis u8[n,W], points4d f64[n,4] = the triangulation of the pair that
    started the track) as CUDA tensors -- the inputs of :func:`bundle_adjust_window`.  Host-side bookkeeping helper
    (one small synchronisation per pair), not part of the hot path."""
    dev = kps.device
    N = kps.shape[1]
    cur = torch.full((N,), -1, dtype=torch.long, device=dev)      # track id of every keypoint of the current frame
    obs, vis, pts = [], [], []
    n_tracks = 0
    for c in range(W - 1):
        b = b0 + c
        n = int(out["n_inl"][b].item())
        inl = out["inl_idx"][b, :n].long()
        q, t = out["qidx"][b][inl].long(), out["tidx"][b][inl].long()
        ts, perm = torch.sort(t, stable=True)                     # one query per train keypoint: the lowest
        first = torch.ones_like(ts, dtype=torch.bool)
        first[1:] = ts[1:] != ts[:-1]
        keep = perm[first]
        q, t, row = q[keep], t[keep], keep
        ids = cur[t]
        new = ids < 0
        n_new = int(new.sum().item())
        ids = torch.where(new, n_tracks + torch.cumsum(new.long(), 0) - 1, ids)
        if n_new:
            obs.append(torch.zeros((n_new, W, 2), dtype=torch.float64, device=dev))
            vis.append(torch.zeros((n_new, W), dtype=torch.uint8, device=dev))
            pts.append(out["points4d"][b][row[new]])
        O, V = torch.cat(obs) if len(obs) > 1 else obs[0], torch.cat(vis) if len(vis) > 1 else vis[0]
        obs, vis = [O], [V]
        O[ids[new], c] = kps[b][t[new]]
        V[ids[new], c] = 1
        O[ids, c + 1] = kps[b + 1][q]
        V[ids, c + 1] = 1
        n_tracks += n_new
        nxt = torch.full((N,), -1, dtype=torch.long, device=dev)
        nxt[q] = ids
        cur = nxt
    return obs[0], vis[0], torch.cat(pts)


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


# ------------------------------------------------------------------------------- whole pairs
PAIR_OUTPUTS = ("idx", "dist", "qidx", "tidx", "m_count", "E", "E_best", "best", "counts", "mask", "n_inl",
                "inl_idx", "Rt", "poses", "points4d", "good", "status")


def _alloc_pair_outputs(B, N, H, nseq, want, factory):
    shapes = dict(idx=((B, N, 2), torch.int32), dist=((B, N, 2), torch.int32), qidx=((B, N), torch.int32),
                  tidx=((B, N), torch.int32), m_count=((B,), torch.int32), E=((B, 3, 3), torch.float64),
                  E_best=((B, 3, 3), torch.float64), best=((B,), torch.int32), counts=((B, H), torch.int32),
                  mask=((B, N), torch.uint8), n_inl=((B,), torch.int32), inl_idx=((B, N), torch.int32),
                  Rt=((B, 4, 4), torch.float64), poses=((B + nseq, 4, 4), torch.float64),
                  points4d=((B, N, 4), torch.float64), good=((B, N), torch.uint8), status=((B,), torch.int32),
                  pose_votes=((B, 4), torch.int32))
    return {k: factory(*shapes[k]) for k in want}


def process_pairs_device(des, kps, K, words=None, samples=None, H=None, frame_count=None, pair_q=None, pair_t=None,
                         seq_off=None, init_pose=None, ratio=RATIO, residual_threshold=RESIDUAL_THRESHOLD,
                         outputs=("m_count", "E", "n_inl", "Rt", "poses", "points4d", "good", "status"),
                         out=None, ctx=None, batch=None, pose_mode=0):
    """The ``else`` branch of process_frame (slam.py:249-266) for a batch of pairs, all on the device,
    no host synchronisation.  des u8[F,N,32], kps f64[F,N,2] frame pools (CUDA tensors).  Without
    pair_q/pair_t, pair b is (curr = frame b+1+s, last = frame b+s) inside sequence s (``seq_off``).
    ``words u32[H,8]`` / ``samples int32[H,8]`` is shared by all pairs, or per pair when 3-D.
    Returns a dict of CUDA tensors for the requested ``outputs``.
    """
    dev = _device_of(des)
    des = _dev(des, torch.uint8, dev)
    kps = _dev(kps, torch.float64, dev)
    F, N, _ = des.shape
    tab = samples if samples is not None else words
    tab = _dev(np.asarray(tab, dtype=np.int32 if samples is not None else np.uint32)
               if not isinstance(tab, torch.Tensor) else tab, torch.int32, dev)
    Hh = tab.shape[-2]
    so = None
    if pair_q is not None:
        pq = _dev(pair_q, torch.int32, dev)
        pt = _dev(pair_t, torch.int32, dev)
        B = pq.shape[0]
    else:
        pq = pt = None
        if batch is not None:
            B = int(batch)
        elif seq_off is not None:
            B = int(_np(seq_off)[-1])
        else:
            B = F - 1
    nseq = 1
    if seq_off is not None:
        so = seq_off if isinstance(seq_off, torch.Tensor) else np.asarray(seq_off, dtype=np.int32)
        so = _dev(so, torch.int32, dev)
        nseq = so.shape[0] - 1
    fc = _dev(frame_count, torch.int32, dev)
    ip = _dev(init_pose, torch.float64, dev)
    ctx = ctx or context(dev.index)
    if out is None:
        out = _alloc_pair_outputs(B, N, Hh, nseq, outputs,
                                  lambda shape, dt: torch.empty(shape, dtype=dt, device=dev))
    io = PairsIO()
    io.des, io.kps = des.data_ptr(), kps.data_ptr()
    io.frame_count = fc.data_ptr() if fc is not None else None
    io.pair_q = pq.data_ptr() if pq is not None else None
    io.pair_t = pt.data_ptr() if pt is not None else None
    io.batch, io.n_max, io.H = B, N, Hh
    io.seq_off = so.data_ptr() if so is not None else None
    io.nseq = nseq
    io.init_pose = ip.data_ptr() if ip is not None else None
    if samples is not None:
        io.samples = tab.data_ptr()
    else:
        io.words = tab.data_ptr()
    io.sample_stride = Hh * 8 if tab.dim() == 3 else 0
    kinv, _ = _kinv(K)
    for i in range(9):
        io.kinv[i] = float(kinv.flat[i])
    io.ratio, io.thr = float(ratio), float(residual_threshold)
    io.pose_mode = int(pose_mode)
    for k, v in out.items():
        setattr(io, k, v.data_ptr())
    check(ctx.lib.tvg_pairs_device(ctx.handle, C.byref(io), _stream()), "tvg_pairs_device")
    return out


def process_pairs_host(des, kps, K, words=None, samples=None, frame_count=None, pair_q=None, pair_t=None,
                       seq_off=None, init_pose=None, ratio=RATIO, residual_threshold=RESIDUAL_THRESHOLD,
                       outputs=("m_count", "E", "n_inl", "Rt", "poses", "status"), out=None, ctx=None, device=None,
                       wait=True):
    """Same pipeline behind HOST buffers (numpy arrays or pinned CPU tensors): host->device copies,
    kernels, device->host copies of the requested outputs and one synchronise happen inside the call
    (``tvg_pairs_host``).  Returns (dict of CPU tensors, h2d_bytes, d2h_bytes).

    ``wait=False`` only enqueues (``tvg_pairs_host_submit``): call :func:`wait_pairs_host` on the same context
    before reading the outputs or touching the inputs.  Alternating two contexts overlaps the copies of one
    batch with the kernels of the previous one."""
    def host(x, dt):
        if x is None:
            return None
        if isinstance(x, np.ndarray):
            if x.dtype == np.uint32:
                x = x.view(np.int32)
            x = torch.from_numpy(np.ascontiguousarray(x))
        assert not x.is_cuda
        return x.to(dt).contiguous() if x.dtype != dt else x.contiguous()
    des = host(des, torch.uint8)
    kps = host(kps, torch.float64)
    F, N, _ = des.shape
    tab = host(samples if samples is not None else
               (np.asarray(words, dtype=np.uint32) if not isinstance(words, torch.Tensor) else words), torch.int32)
    Hh = tab.shape[-2]
    pq, pt = host(pair_q, torch.int32), host(pair_t, torch.int32)
    so = host(np.asarray(seq_off, dtype=np.int32) if seq_off is not None and not isinstance(seq_off, torch.Tensor)
              else seq_off, torch.int32)
    nseq = so.shape[0] - 1 if so is not None else 1
    B = pq.shape[0] if pq is not None else (int(so[-1]) if so is not None else F - 1)
    fc, ip = host(frame_count, torch.int32), host(init_pose, torch.float64)
    ctx = ctx or context(device)
    if out is None:
        out = _alloc_pair_outputs(B, N, Hh, nseq, outputs,
                                  lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory())
    io = PairsIO()
    io.des, io.kps = des.data_ptr(), kps.data_ptr()
    io.frame_count = fc.data_ptr() if fc is not None else None
    io.pair_q = pq.data_ptr() if pq is not None else None
    io.pair_t = pt.data_ptr() if pt is not None else None
    io.batch, io.n_max, io.H = B, N, Hh
    io.seq_off = so.data_ptr() if so is not None else None
    io.nseq = nseq
    io.init_pose = ip.data_ptr() if ip is not None else None
    if samples is not None:
        io.samples = tab.data_ptr()
    else:
        io.words = tab.data_ptr()
    io.sample_stride = Hh * 8 if tab.dim() == 3 else 0
    kinv, _ = _kinv(K)
    for i in range(9):
        io.kinv[i] = float(kinv.flat[i])
    io.ratio, io.thr = float(ratio), float(residual_threshold)
    for k, v in out.items():
        setattr(io, k, v.data_ptr())
    h2d, d2h = C.c_longlong(0), C.c_longlong(0)
    with torch.cuda.device(ctx.device):
        fn = ctx.lib.tvg_pairs_host if wait else ctx.lib.tvg_pairs_host_submit
        check(fn(ctx.handle, C.byref(io), F, C.byref(h2d), C.byref(d2h), _stream()), "tvg_pairs_host")
    if not wait:
        ctx._inflight = (des, kps, tab, pq, pt, so, fc, ip, out)      # keep the host buffers alive
    return out, int(h2d.value), int(d2h.value)


def wait_pairs_host(ctx):
    """Block until the batch submitted with ``process_pairs_host(..., wait=False)`` on ``ctx`` is complete."""
    check(ctx.lib.tvg_pairs_host_wait(ctx.handle), "tvg_pairs_host_wait")
    ctx._inflight = None


# ------------------------------------------------------------------------------- one huge pair (config C5)
HUGE_PAIR_OUTPUTS = ("idx", "dist", "qidx", "tidx", "m_count", "E", "E_best", "best", "best_sse", "counts", "n_cand",
                     "mask", "n_inl", "inl_idx", "Rt", "poses", "points4d", "good", "status")


def huge_pair_rows(nq, nt, world):
    """Rows of every per-row output of :func:`huge_pair` (query rows padded to a multiple of the world size)."""
    return int(_lib.load().tvg_huge_pair_rows(int(nq), int(nt), int(world)))


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n items for `rank` (tvg_shard_range: the first n % world ranks own one more)."""
    lo, hi = C.c_int(0), C.c_int(0)
    _lib.load().tvg_shard_range(int(n), int(rank), int(world), C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def huge_pair(curr_kps, curr_des, last_kps, last_des, K, words=None, samples=None, ratio=RATIO,
              residual_threshold=RESIDUAL_THRESHOLD,
              outputs=("qidx", "tidx", "m_count", "E", "E_best", "best", "best_sse", "mask", "n_inl", "Rt", "poses",
                       "points4d", "good", "status"), out=None, ctx=None):
    """The ``else`` branch of process_frame (slam.py:249-266) for ONE huge keyframe pair, split over the ranks of the
    context's communicator (``tvg_huge_pair``; see ``openwork_b200.dist.init_comm``): query rows for the matcher,
    hypotheses for RANSAC, two fixed-size all-gathers, no host synchronisation.  Inputs are replicated CUDA tensors
    (or numpy arrays); ``words`` / ``samples`` is the whole [H,8] table on every rank.  Returns a dict of CUDA tensors
    (identical on all ranks, except ``counts`` which holds this rank's hypothesis block); per-row outputs have
    :func:`huge_pair_rows` rows, valid up to ``m_count`` / ``n_inl``."""
    dev = _device_of(curr_des, last_des, curr_kps)
    cd = _dev(curr_des, torch.uint8, dev)
    ld = _dev(last_des, torch.uint8, dev)
    ck = _dev(curr_kps, torch.float64, dev)
    lk = _dev(last_kps, torch.float64, dev)
    tab = samples if samples is not None else words
    tab = _dev(np.asarray(tab, dtype=np.int32 if samples is not None else np.uint32)
               if not isinstance(tab, torch.Tensor) else tab, torch.int32, dev)
    H = tab.shape[0]
    nq, nt = cd.shape[0], ld.shape[0]
    ctx = ctx or context(dev.index)
    _, world = ctx.comm_info()
    rows = huge_pair_rows(nq, nt, world)
    if out is None:
        shapes = dict(idx=((rows, 2), torch.int32), dist=((rows, 2), torch.int32), qidx=((rows,), torch.int32),
                      tidx=((rows,), torch.int32), m_count=((1,), torch.int32), E=((3, 3), torch.float64),
                      E_best=((3, 3), torch.float64), best=((1,), torch.int32), best_sse=((1,), torch.float64),
                      counts=((H,), torch.int32), n_cand=((1,), torch.int32), mask=((rows,), torch.uint8),
                      n_inl=((1,), torch.int32), inl_idx=((rows,), torch.int32), Rt=((4, 4), torch.float64),
                      poses=((2, 4, 4), torch.float64), points4d=((rows, 4), torch.float64), good=((rows,), torch.uint8),
                      status=((1,), torch.int32))
        out = {k: torch.zeros(shapes[k][0], dtype=shapes[k][1], device=dev) for k in outputs}
    io = HugePairIO()
    io.curr_des, io.last_des, io.curr_kps, io.last_kps = cd.data_ptr(), ld.data_ptr(), ck.data_ptr(), lk.data_ptr()
    io.nq, io.nt, io.H = nq, nt, H
    if samples is not None:
        io.samples = tab.data_ptr()
    else:
        io.words = tab.data_ptr()
    kinv, _ = _kinv(K)
    for i in range(9):
        io.kinv[i] = float(kinv.flat[i])
    io.ratio, io.thr = float(ratio), float(residual_threshold)
    for k, v in out.items():
        setattr(io, k, v.data_ptr())
    check(ctx.lib.tvg_huge_pair(ctx.handle, C.byref(io), _stream()), "tvg_huge_pair")
    ctx._hp_keep = (cd, ld, ck, lk, tab)     # inputs stay alive until the next call (the call is asynchronous)
    return out
