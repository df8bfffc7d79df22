"""GPU test of the drop-in call surface (openwork_b200/slam.py): the same driver loop a user of the
reference's slam.py writes (Frame -> process_frame), checked against the oracle pair by pair."""
import numpy as np
import pytest

from openwork_b200 import synth
from oracle import two_view as tv

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [True, False])
def test_process_frame_sequence_matches_oracle(fused):
    from openwork_b200 import slam
    K = synth.K_640x480
    slam.K = K
    slam.FUSED_PROCESS_FRAME = fused
    slam.ransac_config.update(max_trials=200, residual_threshold=0.005, seed=3)   # reference defaults slam.py:157-159
    slam.Frame.reset()
    observations = []

    class Mapp:
        def add_observation(self, pose, points):
            observations.append((np.array(pose), np.array(points)))
    slam.mapp = Mapp()
    frames, _ = synth.make_sequence(4, 900, seed=123)
    pose = None
    for i, (kps, des) in enumerate(frames):
        f = slam.Frame(kps=kps, des=des)
        f = slam.process_frame(f)
        if i == 0:
            assert np.array_equal(f.curr_pose, np.eye(4)) and f.points4d == [[0, 0, 0, 1]]
            pose = np.eye(4)
            continue
        (lk, ld), (ck, cd) = frames[i - 1], frames[i]
        words = synth.sample_words(200, 3 * 1000003 + f.idx)
        ref = tv.process_pair(lk, ld, ck, cd, pose, K, words=words)
        assert len(f.curr_kps) == int(ref["inliers"].sum())                     # rebinding at slam.py:161-162
        assert np.array_equal(f.curr_kps, ck[ref["idx1"]][ref["inliers"]])
        assert np.array_equal(f.last_kps, lk[ref["idx2"]][ref["inliers"]])
        E, Er = f.essential_matrix / np.linalg.norm(f.essential_matrix), ref["E"] / np.linalg.norm(ref["E"])
        assert min(np.linalg.norm(E - Er), np.linalg.norm(E + Er)) < 1e-4
        np.testing.assert_allclose(f.curr_pose, ref["curr_pose"], atol=1e-5)
        Xr = ref["points4d"][ref["good"]]
        assert f.points4d.shape == Xr.shape
        rel = np.linalg.norm(f.points4d[:, :3] - Xr[:, :3], axis=1) / np.linalg.norm(Xr[:, :3], axis=1)
        assert np.median(rel) < 1e-4
        pose = ref["curr_pose"]
    assert len(observations) == 4
    slam.mapp = None
    slam.FUSED_PROCESS_FRAME = True


def test_process_frame_fused_equals_stepwise():
    """One pipeline call per frame (default) leaves every frame attribute as the function-by-function sequence does,
    including ragged keypoint counts and the exceptions."""
    from openwork_b200 import slam
    slam.K = synth.K_640x480
    slam.mapp = None
    slam.ransac_config.update(max_trials=128, residual_threshold=0.005, seed=11)
    frames, _ = synth.make_sequence(5, 800, seed=77)
    frames = [(k[: 800 - 37 * i], d[: 800 - 37 * i]) for i, (k, d) in enumerate(frames)]     # different N per frame
    res = {}
    for fused in (True, False):
        slam.FUSED_PROCESS_FRAME = fused
        slam.Frame.reset()
        res[fused] = [slam.process_frame(slam.Frame(kps=k, des=d)) for k, d in frames]
    slam.FUSED_PROCESS_FRAME = True
    for a, b in zip(res[True][1:], res[False][1:]):
        assert np.array_equal(a.curr_kps, b.curr_kps) and np.array_equal(a.last_kps, b.last_kps)
        assert np.array_equal(a.essential_matrix, b.essential_matrix)          # the same kernels ran
        np.testing.assert_allclose(a.curr_pose, b.curr_pose, rtol=0, atol=1e-12)
        assert a.points4d.shape == b.points4d.shape
        np.testing.assert_allclose(a.points4d, b.points4d, rtol=1e-8, atol=1e-10)
    # exceptions of the fused path: too few train rows, too few matches
    slam.Frame.reset()
    lk, ld, ck, cd = synth.make_pair(64, 1)
    slam.process_frame(slam.Frame(kps=lk[:1], des=ld[:1]))
    with pytest.raises(ValueError):
        slam.process_frame(slam.Frame(kps=ck, des=cd))
    slam.Frame.reset()
    slam.process_frame(slam.Frame(kps=lk, des=synth.random_descriptors(64, 4)))
    with pytest.raises(AssertionError):
        slam.process_frame(slam.Frame(kps=ck, des=synth.random_descriptors(64, 3)))
    slam.Frame.reset()


def test_reference_error_conventions():
    from openwork_b200 import slam
    slam.K = synth.K_640x480

    class F:
        pass
    lk, ld, ck, cd = synth.make_pair(64, 1)
    f = F()
    f.curr_kps, f.curr_des, f.last_kps, f.last_des = ck, cd, lk[:1], ld[:1]
    with pytest.raises(ValueError):          # one train row: `for m, n in matches` fails (slam.py:130)
        slam.match_points(f)
    f.last_kps, f.last_des, f.curr_des = lk, synth.random_descriptors(64, 4), synth.random_descriptors(64, 3)
    with pytest.raises(AssertionError):      # fewer than 8 survivors (slam.py:138)
        slam.match_points(f)
    slam.frame = None
    with pytest.raises(ValueError):          # skimage: min_samples must be < number of samples
        slam.fit_essential_matrix(np.zeros((8, 2, 2)))
