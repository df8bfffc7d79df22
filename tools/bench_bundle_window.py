"""Throughput of the sliding-window bundle adjustment behind the pair pipeline: S sequences of W frames at 2k keypoints
-> pair pipeline -> tracks chained from the inlier matches -> tvg_bundle_adjust_window (one CTA per window).
    python tools/bench_bundle_window.py [sequences=64] [W=6] [iters=8]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import ops, synth  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ITERS = int(sys.argv[3]) if len(sys.argv) > 3 else 8
K = synth.K_640x480
base = []
for d in range(4):
    fr, _ = synth.make_sequence(W, 2048, seed=700 + d)
    base.append((np.stack([f[1] for f in fr]), np.stack([f[0] for f in fr])))
des = torch.from_numpy(np.concatenate([base[s % 4][0] for s in range(S)])).cuda()
kps = torch.from_numpy(np.concatenate([base[s % 4][1] for s in range(S)])).cuda()
so = torch.arange(0, S + 1, dtype=torch.int32).cuda() * (W - 1)
out = ops.process_pairs_device(des, kps, K, words=synth.sample_words(1024, 1), seq_off=so, batch=S * (W - 1),
                               outputs=("n_inl", "inl_idx", "qidx", "tidx", "poses", "points4d", "good", "status"))
assert int((out["status"] != 0).sum()) == 0
tracks = []
for s in range(min(S, 4)):                      # tracks of the distinct sequences; the copies reuse them
    o = {k: v[s * (W - 1):(s + 1) * (W - 1)] for k, v in out.items() if k != "poses"}
    tracks.append(ops.window_tracks(o, kps[s * W:(s + 1) * W], 0, W))
n_max = max(t[0].shape[0] for t in tracks)
obs = torch.zeros((S, n_max, W, 2), dtype=torch.float64, device="cuda")
vis = torch.zeros((S, n_max, W), dtype=torch.uint8, device="cuda")
pts = torch.zeros((S, n_max, 4), dtype=torch.float64, device="cuda")
cnt = torch.zeros(S, dtype=torch.int32, device="cuda")
for s in range(S):
    o, v, p = tracks[s % 4]
    obs[s, :len(o)], vis[s, :len(o)], pts[s, :len(o)], cnt[s] = o, v, p, len(o)
poses = out["poses"].view(S, W, 4, 4)
for _ in range(2):
    r = ops.bundle_adjust_window(obs, vis, poses, pts, K, n_pts=cnt, n_fixed=1, iters=ITERS)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    r = ops.bundle_adjust_window(obs, vis, poses, pts, K, n_pts=cnt, n_fixed=1, iters=ITERS)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
cost = r[2].cpu().numpy()
print({"windows": S, "frames_per_window": W, "iters": ITERS, "tracks_per_window": float(cnt.float().mean()),
       "observations_per_window": float(vis.sum().item() / S), "ms": ms, "windows_per_s": S / ms * 1e3,
       "median_cost_ratio": float(np.median(cost[:, 1] / cost[:, 0])), "accepted_mean": float(r[3].float().mean())})
