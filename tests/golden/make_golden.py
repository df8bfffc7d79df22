"""Generate tests/golden/*.npz by running the REFERENCE's own functions in the build container.

    python tests/golden/make_golden.py

Imports /root/reference/MonocularSlam/slam.py through oracle/ref_loader.py (pangolin / OpenGL /
skimage stubbed, np.mat aliased) and records inputs + outputs of the functions that can run here:
match_points (real cv2.BFMatcher), normalize, extract_Rt, triangulate, check_points.  The RANSAC
stage cannot be generated from the reference (scikit-image is absent and not installable): its
fixture comes from the restated oracle and is marked `pinned_by_reference = False`.
The committed .npz files travel to the GPU box; /root/reference does not.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openwork_b200 import synth  # noqa: E402
from oracle import ref_loader, two_view as tv  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class _Frame:
    pass


def main():
    ref = ref_loader.load()
    K = synth.K_640x480
    ref.K = K

    # ---- pair fixture: match_points + normalize + (restated) ransac + extract_Rt + triangulate + check_points
    for name, n, seed, H, ties in (("pair_small", 320, 101, 96, 12), ("pair_medium", 900, 202, 160, 30)):
        lk, ld, ck, cd = synth.make_pair(n, seed, forced_ties=ties)
        f = _Frame()
        f.curr_kps, f.curr_des, f.last_kps, f.last_des = ck.copy(), cd, lk.copy(), ld
        raw = ref.cv2.BFMatcher(ref.cv2.NORM_HAMMING).knnMatch(cd, ld, k=2)
        knn_idx = np.array([[m.trainIdx, n_.trainIdx] for m, n_ in raw], dtype=np.int32)
        knn_dist = np.array([[m.distance, n_.distance] for m, n_ in raw], dtype=np.int32)
        match_kps = np.array(ref.match_points(f))          # REFERENCE slam.py:125-143
        matched_curr, matched_last = f.curr_kps, f.last_kps  # rebinding at slam.py:140-141
        norm_curr = ref.normalize(K, match_kps[:, 0])      # REFERENCE slam.py:87-92
        norm_last = ref.normalize(K, match_kps[:, 1])
        words = synth.sample_words(H, seed)
        samples = synth.samples_from_words(words, len(match_kps))
        r = tv.ransac_essential(norm_last, norm_curr, samples, 0.005, return_models=True)   # restated (unpinned)
        inl = r["mask"]
        E = r["E"]
        Rt = np.asarray(ref.extract_Rt(E))                 # REFERENCE slam.py:167-186
        last_pose = np.eye(4)
        last_pose[:3, 3] = [0.2, -0.1, 0.7]
        curr_pose = np.dot(Rt, last_pose)                  # slam.py:261
        pts = ref.triangulate(matched_last[inl], matched_curr[inl], last_pose, curr_pose)  # REFERENCE slam.py:202-221
        good = ref.check_points(pts)                       # REFERENCE slam.py:233-237
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"), K=K, last_kps=lk, last_des=ld, curr_kps=ck, curr_des=cd,
            knn_idx=knn_idx, knn_dist=knn_dist, match_kps=match_kps, norm_curr=norm_curr, norm_last=norm_last,
            words=words, samples=samples, ransac_counts=r["counts"], ransac_sse=r["sse"], ransac_best=r["best"],
            ransac_mask=inl, ransac_E_best=r["E_best"], ransac_models=r["models"], E=E, Rt=Rt, last_pose=last_pose,
            curr_pose=curr_pose, points4d=pts, good=good, pinned_by_reference=np.array(
                ["knn_idx", "knn_dist", "match_kps", "norm_curr", "norm_last", "Rt", "curr_pose", "points4d", "good"]))
        print(name, "M =", len(match_kps), "inliers =", int(inl.sum()), "good =", int(good.sum()))

    # ---- extract_Rt fixture: reference on closed-form E = [t]x R with small rotations
    rng = np.random.default_rng(7)
    Es, Rts = [], []
    for _ in range(32):
        R = synth._small_rotation(rng, 0.2)
        t = rng.normal(size=3)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        E = tx @ R * rng.uniform(0.2, 5.0) * rng.choice([-1.0, 1.0])
        Es.append(E)
        Rts.append(np.asarray(ref.extract_Rt(E)))
    np.savez_compressed(os.path.join(OUT, "extract_rt.npz"), E=np.stack(Es), Rt=np.stack(Rts))

    # ---- matcher tie-break fixture: real cv2 on inputs with many exact ties
    rng = np.random.default_rng(11)
    q = rng.integers(0, 256, (257, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (391, 32), dtype=np.uint8)
    t[rng.choice(391, 80, replace=False)] = t[rng.choice(391, 80, replace=False)]
    q[rng.choice(257, 40, replace=False)] = t[rng.integers(0, 391, 40)]
    raw = ref.cv2.BFMatcher(ref.cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    np.savez_compressed(os.path.join(OUT, "knn_ties.npz"), q=q, t=t,
                        idx=np.array([[m.trainIdx, n_.trainIdx] for m, n_ in raw], dtype=np.int32),
                        dist=np.array([[m.distance, n_.distance] for m, n_ in raw], dtype=np.int32))
    print("wrote fixtures to", OUT)


if __name__ == "__main__":
    main()
