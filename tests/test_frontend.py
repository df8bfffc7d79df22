"""Feature front-end: extract_points (slam.py:112-122).
CPU: the numpy restatement (oracle/frontend.py) is pinned against the REAL cv2 calls the reference makes, and
the committed ORB pattern against the installed cv2 binary.  GPU: the CUDA kernels against real cv2."""
import os
import sys

import numpy as np
import pytest

from oracle import frontend as fe

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pattern():
    return np.load(os.path.join(ROOT, "openwork_b200", "orb_pattern.npy"))


def test_committed_orb_pattern_matches_installed_cv2():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import extract_orb_pattern
    assert np.array_equal(extract_orb_pattern.extract(), _pattern())
    p = _pattern()
    assert p.shape == (512, 2) and p.dtype == np.int8 and p[:4].tolist() == [[8, -3], [9, 5], [4, 2], [7, -12]]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_numpy_restatement_equals_cv2(seed):
    import cv2
    img = fe.synthetic_image(seed)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(fe.bgr2gray(img), gray)
    e_ref = cv2.cornerMinEigenVal(gray, 3, ksize=3)
    e = fe.min_eig_map(gray)
    assert (e != e_ref).mean() < 1e-4 and np.abs(e - e_ref).max() < 1e-8      # running-sum last-bit effects only
    pts = fe.good_features_to_track(gray)
    ref = cv2.goodFeaturesToTrack(gray, 3000, qualityLevel=0.01, minDistance=3).reshape(-1, 2)
    assert pts.shape == ref.shape and np.array_equal(pts, ref)
    assert np.array_equal(fe.gaussian7_sigma2(), cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel())
    blur = cv2.GaussianBlur(gray.astype(np.float32), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(fe.blur_7x7_sigma2(gray), np.rint(blur).astype(np.uint8))
    kps_ref, des_ref = fe.extract_points_cv2(img)
    kps, des = fe.orb_compute(gray, pts, _pattern())
    assert np.array_equal(kps, kps_ref) and np.array_equal(des, des_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,channels,shape", [(3, 3, (480, 640)), (4, 1, (480, 640)), (5, 3, (540, 960)),
                                                  (6, 3, (120, 200))])
def test_extract_points_matches_cv2(seed, channels, shape):
    from openwork_b200 import ops
    img = fe.synthetic_image(seed, shape[0], shape[1], channels)
    kps_ref, des_ref = fe.extract_points_cv2(img)
    kps, des, n = ops.extract_points(img)
    n = int(n[0])
    assert n == len(kps_ref)
    assert np.array_equal(kps[0, :n].cpu().numpy(), kps_ref)
    assert np.array_equal(des[0, :n].cpu().numpy(), des_ref)


@pytest.mark.gpu
def test_extract_points_batched_and_fewer_corners():
    from openwork_b200 import ops
    imgs = np.stack([fe.synthetic_image(10 + i) for i in range(3)])
    kps, des, n = ops.extract_points(imgs, max_corners=500, quality_level=0.05, min_distance=5)
    import cv2
    for i in range(3):
        gray = cv2.cvtColor(imgs[i], cv2.COLOR_BGR2GRAY)
        pts = cv2.goodFeaturesToTrack(gray, 500, qualityLevel=0.05, minDistance=5)
        kp = [cv2.KeyPoint(x=float(p[0][0]), y=float(p[0][1]), size=20) for p in pts]
        kp, d = cv2.ORB_create().compute(gray, kp)
        m = int(n[i])
        assert m == len(kp)
        assert np.array_equal(kps[i, :m].cpu().numpy(), np.array([k.pt for k in kp]))
        assert np.array_equal(des[i, :m].cpu().numpy(), d)


@pytest.mark.gpu
def test_process_frame_from_images():
    """The reference's driver loop on images: extract_points -> match -> RANSAC -> pose -> triangulate, all GPU."""
    import cv2
    from openwork_b200 import slam, synth
    slam.K = synth.K_640x480
    slam.Frame.reset()
    slam.extract_points = None
    img0 = fe.synthetic_image(21)
    M = np.float32([[1.01, 0.002, 3.0], [-0.002, 1.01, -2.0]])          # a small similarity between the frames
    img1 = cv2.warpAffine(img0, M, (640, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    f0 = slam.process_frame(slam.Frame(image=img0))
    f1 = slam.process_frame(slam.Frame(image=img1))
    k0, d0 = fe.extract_points_cv2(img0)
    k1, d1 = fe.extract_points_cv2(img1)
    assert np.array_equal(slam.Frame.last_kps, k1) and np.array_equal(slam.Frame.last_des, d1)
    assert f1.curr_pose.shape == (4, 4) and np.isfinite(f1.curr_pose).all()
    assert len(f1.curr_kps) >= 8 and f1.points4d.shape[1] == 4


@pytest.mark.gpu
def test_run_sequences_from_images_matches_cv2_plus_oracle():
    """Raw images -> poses + map points entirely on the device == cv2 front-end + CPU oracle pair pipeline."""
    import cv2
    import torch
    from openwork_b200 import ops, synth
    from oracle import two_view as tv
    K = synth.K_640x480
    img0 = fe.synthetic_image(31)
    frames_img = [img0]
    for i in range(1, 4):
        M = np.float32([[1.0 + 0.008 * i, 0.002 * i, 2.5 * i], [-0.002 * i, 1.0 + 0.008 * i, -1.5 * i]])
        frames_img.append(cv2.warpAffine(img0, M, (640, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101))
    imgs = np.stack(frames_img)
    words = synth.sample_words(200, 12)
    res, n, status = ops.run_sequences_from_images(torch.from_numpy(imgs).cuda(), K, words)
    assert (status.cpu().numpy() == 0).all()
    feats = [fe.extract_points_cv2(im) for im in frames_img]
    assert n.cpu().numpy().tolist() == [len(f[0]) for f in feats]
    pose = np.eye(4)
    Xs, gs, poses_ref = [], [], [pose]
    for i in range(3):
        (lk, ld), (ck, cd) = feats[i], feats[i + 1]
        r = tv.process_pair(lk, ld, ck, cd, pose, K, words=words)
        Xs.append(r["points4d"]); gs.append(r["good"]); pose = r["curr_pose"]; poses_ref.append(pose)
    poses, pts = res[0][0].cpu().numpy(), res[0][1].cpu().numpy()
    np.testing.assert_allclose(poses, np.stack(poses_ref), atol=1e-5)
    ref_pts = tv.map_points(Xs, gs)
    assert pts.shape == ref_pts.shape
    rel = np.linalg.norm(pts[1:, :3] - ref_pts[1:, :3], axis=1) / np.maximum(np.linalg.norm(ref_pts[1:, :3], axis=1), 1e-12)
    assert np.median(rel) < 1e-4 and (rel < 1e-3).mean() > 0.98
