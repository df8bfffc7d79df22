"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel, ragged sizes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import ops, synth  # noqa: E402

K = synth.K_640x480
ctx = ops.context(0)
frames, _ = synth.make_sequence(4, 300, seed=3)
des = np.stack([f[1] for f in frames])
kps = np.stack([f[0] for f in frames])
cnt = np.array([300, 257, 300, 129], dtype=np.int32)
words = synth.sample_words(70, 1)
for variant, r in ((0, 2), (1, 4), (2, 8), (3, 4), (5, 2), (7, 8)):
    ctx.set_matcher_config(r, variant, 6)
    out = ops.process_pairs_device(torch.from_numpy(des).cuda(), torch.from_numpy(kps).cuda(), K, words=words,
                                   frame_count=cnt, outputs=("m_count", "n_inl", "E", "Rt", "poses", "points4d",
                                                             "good", "status", "mask", "counts"), pose_mode=variant & 1)
    torch.cuda.synchronize()
    print(variant, r, out["m_count"].tolist(), out["n_inl"].tolist(), out["status"].tolist())
ctx.set_matcher_config(4, 11, 6)
# single-pair ops incl. split-Nt merge path and tiny / degenerate shapes
for nq, nt in ((1, 2), (3, 1), (130, 77), (700, 1500)):
    q, t = synth.random_descriptors(nq, 1), synth.random_descriptors(nt, 2)
    idx, dist, keep = ops.knn2_hamming256(q, t, return_keep=True)
    o = ops.ratio_compact(idx, keep, np.zeros((nq, 2)), np.zeros((nt, 2)), K)
lk, ld, ck, cd = synth.make_pair(400, 9)
idx, dist, keep = ops.knn2_hamming256(cd, ld, return_keep=True)
o = ops.ratio_compact(idx, keep, ck, lk, K)
r = ops.ransac_essential(o["src"], o["dst"], words=synth.sample_words(33, 2), return_models=True)
v = ops.cheirality(r["E"], o["src"], o["dst"], mask=r["mask"])
X = ops.triangulate(lk[:50], ck[:50], np.eye(4), r["Rt"].cpu().numpy(), K)
host, h2d, d2h = ops.process_pairs_host(des, kps, K, words=words, frame_count=cnt,
                                        outputs=("m_count", "n_inl", "E", "poses", "points4d", "status"))
ctx.set_host_chunk(2)
host, h2d, d2h = ops.process_pairs_host(des, kps, K, words=words, frame_count=cnt,
                                        outputs=("m_count", "n_inl", "E", "poses", "points4d", "status"))
torch.cuda.synchronize()
print("sanitize run ok", r["n_inl"], host["n_inl"].tolist())
