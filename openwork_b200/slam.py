"""Drop-in for the two-view functions of OpenWork's ``MonocularSlam/slam.py``.

Same names, positional signatures, array layouts, global-variable conventions and error behaviour as
the reference (``/root/reference/MonocularSlam/slam.py``), with every body running on the B200
through ``openwork_b200.ops`` (C ABI -> sm_100a kernels).  numpy in / numpy out, like the reference.

=====================  =========================  ==========================================
reference               here                       what runs
=====================  =========================  ==========================================
normalize        :87    normalize(K, pts)          device elementwise
match_points     :125   match_points(frame)        hamming_knn2 + ratio/compact kernels
fit_essential..  :146   fit_essential_matrix(mk)   ransac fit/score/reduce/rescore/refit kernels
extract_Rt       :167   extract_Rt(E)              extract_rt kernel
triangulate      :202   triangulate(p1,p2,P1,P2)   triangulate kernel
check_points     :233   check_points(points4d)     z > 0
process_frame    :239   process_frame(frame)       orchestration, same order as the reference
Frame            :95    Frame                      same class-attribute hand-off
=====================  =========================  ==========================================

Differences, all forced by the north star (BASELINE.json):
* ORB extraction is outside the path: ``Frame`` is built from keypoints + descriptors (or set the
  module-level ``extract_points`` hook to your own detector, signature as slam.py:112).
* RANSAC hypotheses come from a seeded host-side sample table instead of skimage's unseeded
  ``rng.choice``: ``ransac_config = dict(max_trials=200, seed=0)`` (defaults: slam.py:157-159).
* The viewer (``Map``, ``draw_points``) is out of scope; set ``mapp`` to any object with
  ``add_observation(pose, points)`` to receive what slam.py:269 hands to the map.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops, synth

# module globals, as in the reference (slam.py:147 `global K`, :161 global `frame`, :269 `mapp`)
K = None
frame = None
mapp = None
extract_points = None
ransac_config = dict(max_trials=200, residual_threshold=0.005, seed=0)
RATIO = 0.75
# process_frame runs the frame pair through ONE pipeline call (tvg_pairs_host: one upload, ten kernels, one download)
# instead of the function-by-function sequence with a host round trip per step; same results, same exceptions.
FUSED_PROCESS_FRAME = True


class Frame(object):
    """slam.py:95-109 -- carries the previous frame's kps / des / pose through class attributes."""
    idx = 0
    last_kps, last_des, last_pose = None, None, None

    def __init__(self, image=None, kps=None, des=None):
        Frame.idx += 1
        self.image = image
        self.idx = Frame.idx
        self.last_kps = Frame.last_kps
        self.last_des = Frame.last_des
        self.last_pose = Frame.last_pose
        self.kps, self.des = kps, des

    @classmethod
    def reset(cls):
        cls.idx = 0
        cls.last_kps, cls.last_des, cls.last_pose = None, None, None


def gpu_extract_points(frame):
    """slam.py:112-122 on the GPU (GFTT corners + ORB descriptors): returns (kps f64[N,2], des u8[N,32])."""
    kps, des, n = ops.extract_points(np.ascontiguousarray(frame.image))
    n = int(n[0])
    return kps[0, :n].cpu().numpy(), des[0, :n].cpu().numpy()


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def normalize(K, pts):
    """slam.py:87-92: ``(inv(K) @ [x, y, 1]^T)^T[:, :2]``."""
    kinv = np.linalg.inv(np.asarray(K))
    p = torch.as_tensor(np.ascontiguousarray(pts, dtype=np.float64)).cuda()
    x, y = p[:, 0], p[:, 1]
    out = torch.stack([kinv[0, 0] * x + kinv[0, 1] * y + kinv[0, 2],
                       kinv[1, 0] * x + kinv[1, 1] * y + kinv[1, 2]], dim=1)
    return out.cpu().numpy()


def match_points(frame):
    """slam.py:125-143: knnMatch(k=2) + Lowe ratio; rebinds ``frame.curr_kps`` / ``frame.last_kps`` to the
    matched rows and returns ``match_kps`` as f64[M,2,2] ([:,0] = curr xy, [:,1] = last xy)."""
    if len(frame.last_des) < 2:
        # the reference's `for m, n in matches` cannot unpack a 1-tuple (slam.py:130)
        raise ValueError("not enough values to unpack (expected 2, got %d)" % len(frame.last_des))
    idx, dist, keep = ops.knn2_hamming256(frame.curr_des, frame.last_des, ratio=RATIO, return_keep=True)
    out = ops.ratio_compact(idx, keep)
    idx1 = out["qidx"].cpu().numpy()
    idx2 = out["tidx"].cpu().numpy()
    assert len(idx1) >= 8
    curr_kps, last_kps = np.asarray(frame.curr_kps), np.asarray(frame.last_kps)
    match_kps = np.stack([curr_kps[idx1], last_kps[idx2]], axis=1)
    frame.curr_kps = curr_kps[idx1]
    frame.last_kps = last_kps[idx2]
    return match_kps


def fit_essential_matrix(match_kps):
    """slam.py:146-164: K-normalise, RANSAC (src = last, dst = curr), mask the module-global frame's
    keypoints by the inliers, return the refit ``model.params``."""
    global K, frame
    match_kps = np.array(match_kps)
    norm_curr_kps = normalize(K, match_kps[:, 0])
    norm_last_kps = normalize(K, match_kps[:, 1])
    H = int(ransac_config["max_trials"])
    fidx = getattr(frame, "idx", 0) if frame is not None else 0
    words = synth.sample_words(H, int(ransac_config["seed"]) * 1000003 + int(fidx))
    r = ops.ransac_essential(norm_last_kps, norm_curr_kps, words=words,
                             residual_threshold=float(ransac_config["residual_threshold"]))
    inliers = r["mask"].cpu().numpy()
    if frame is not None:
        frame.curr_kps = frame.curr_kps[inliers]
        frame.last_kps = frame.last_kps[inliers]
        frame.ransac = r
    return r["E"].cpu().numpy()


def extract_Rt(E):
    """slam.py:167-186."""
    return ops.extract_rt(np.asarray(E, dtype=np.float64)).cpu().numpy()


def triangulate(pts1, pts2, pose1, pose2):
    """slam.py:202-221 (reads the module-global K)."""
    global K
    return ops.triangulate(pts1, pts2, pose1, pose2, K).cpu().numpy()


def check_points(points4d):
    """slam.py:233-237."""
    return np.asarray(points4d)[:, 2] > 0


_fused_buffers = {}     # (n, H) -> pinned output buffers + staging arrays of the fused path, reused frame after frame


def _process_pair_fused(frame):
    """The else-branch of process_frame (slam.py:249-266) as one call of the batched pipeline on a batch of one pair.
    Leaves the frame exactly as the function-by-function sequence does."""
    curr_kps, last_kps = np.asarray(frame.curr_kps, dtype=np.float64), np.asarray(frame.last_kps, dtype=np.float64)
    curr_des, last_des = np.asarray(frame.curr_des), np.asarray(frame.last_des)
    nq, nt = len(curr_des), len(last_des)
    if nt < 2:
        raise ValueError("not enough values to unpack (expected 2, got %d)" % nt)      # slam.py:130
    n = max(nq, nt)
    H = int(ransac_config["max_trials"])
    buf = _fused_buffers.get((n, H))
    if buf is None:
        if len(_fused_buffers) > 8:
            _fused_buffers.clear()
        buf = _fused_buffers[(n, H)] = dict(des=torch.zeros((2, n, 32), dtype=torch.uint8).pin_memory(),
                                            kps=torch.zeros((2, n, 2), dtype=torch.float64).pin_memory(), out=None)
    des, kps = buf["des"].numpy(), buf["kps"].numpy()      # rows beyond the counts are never read
    des[0, :nt], des[1, :nq] = last_des, curr_des
    kps[0, :nt], kps[1, :nq] = last_kps, curr_kps
    words = synth.sample_words(H, int(ransac_config["seed"]) * 1000003 + int(frame.idx))
    out, _, _ = ops.process_pairs_host(
        buf["des"], buf["kps"], K, words=words, frame_count=np.array([nt, nq], dtype=np.int32), pair_q=np.array([1], dtype=np.int32),
        pair_t=np.array([0], dtype=np.int32), seq_off=np.array([0, 1], dtype=np.int32),
        init_pose=np.ascontiguousarray(frame.last_pose, dtype=np.float64)[None], ratio=RATIO,
        residual_threshold=float(ransac_config["residual_threshold"]),
        outputs=("qidx", "tidx", "m_count", "mask", "E", "best", "Rt", "poses", "points4d", "good", "n_inl", "status"),
        out=buf["out"])
    buf["out"] = out
    M = int(out["m_count"][0])
    assert M >= 8                                                                       # slam.py:138
    if M <= 8:
        raise ValueError("`min_samples` must be in range (0, <number-of-samples>)")     # skimage.measure.ransac
    idx1, idx2 = out["qidx"][0, :M].numpy(), out["tidx"][0, :M].numpy()
    inliers = out["mask"][0, :M].numpy().astype(bool)
    frame.curr_kps = curr_kps[idx1][inliers]
    frame.last_kps = last_kps[idx2][inliers]
    frame.ransac = dict(E=out["E"][0].clone(), mask=out["mask"][0, :M].clone(), best=int(out["best"][0]),
                        status=int(out["status"][0]))
    frame.essential_matrix = out["E"][0].numpy().copy()
    frame.curr_pose = out["poses"][1].numpy().copy()
    n_inl = int(out["n_inl"][0])
    points4d = out["points4d"][0, :n_inl].numpy()
    return points4d[out["good"][0, :n_inl].numpy().astype(bool)]                        # check_points filter


def process_frame(frame_):
    """slam.py:239-272, same order of operations."""
    global frame
    frame = frame_
    if extract_points is not None:
        frame.curr_kps, frame.curr_des = extract_points(frame)
    elif frame.kps is None and frame.image is not None:
        frame.curr_kps, frame.curr_des = gpu_extract_points(frame)
    else:
        frame.curr_kps, frame.curr_des = np.asarray(frame.kps), np.asarray(frame.des)
    Frame.last_kps, Frame.last_des = frame.curr_kps, frame.curr_des

    if frame.idx == 1:
        frame.curr_pose = np.eye(4)
        points4d = [[0, 0, 0, 1]]
    elif FUSED_PROCESS_FRAME:
        points4d = _process_pair_fused(frame)
    else:
        match_kps = match_points(frame)
        essential_matrix = fit_essential_matrix(match_kps)
        frame.essential_matrix = essential_matrix
        Rt = extract_Rt(essential_matrix)
        frame.curr_pose = np.dot(Rt, frame.last_pose)
        points4d = triangulate(frame.last_kps, frame.curr_kps, frame.last_pose, frame.curr_pose)
        good_pt4d = check_points(points4d)
        points4d = points4d[good_pt4d]
    frame.points4d = points4d
    if mapp is not None:
        mapp.add_observation(frame.curr_pose, points4d)
    Frame.last_pose = frame.curr_pose
    return frame
