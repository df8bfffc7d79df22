"""Throughput of the two-view bundle adjustment behind the pair pipeline: P C2-shaped pairs -> inliers -> tvg_bundle_adjust_pairs."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from openwork_b200 import ops, synth
P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
K = synth.K_640x480
des_h, kps_h = bench.make_batch(P)
des, kps = torch.from_numpy(des_h).cuda(), torch.from_numpy(kps_h).cuda()
pq = torch.arange(1, 2 * P, 2, dtype=torch.int32).cuda(); pt = pq - 1
so = torch.arange(0, P + 1, dtype=torch.int32).cuda()
out = ops.process_pairs_device(des, kps, K, words=synth.sample_words(1024, 1), pair_q=pq, pair_t=pt, seq_off=so,
                               outputs=("n_inl", "inl_idx", "qidx", "tidx", "poses", "points4d", "good", "status"))
N = out["points4d"].shape[1]
inl = out["inl_idx"].long().clamp(0, N - 1)
qi = torch.gather(out["qidx"].long().clamp(0, N - 1), 1, inl)
ti = torch.gather(out["tidx"].long().clamp(0, N - 1), 1, inl)
pix_curr = torch.stack([kps[2 * b + 1][qi[b]] for b in range(P)])
pix_last = torch.stack([kps[2 * b][ti[b]] for b in range(P)])
poses = out["poses"].view(P, 2, 4, 4)          # every pair is its own sequence: (first pose, second pose)
args = (pix_last, pix_curr, poses[:, 0].contiguous(), poses[:, 1].contiguous(), out["points4d"], K)
kw = dict(n_pts=out["n_inl"], use=out["good"], iters=ITERS)
for _ in range(2):
    r = ops.bundle_adjust_pairs(*args, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    r = ops.bundle_adjust_pairs(*args, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
cost = r[2].cpu().numpy()
print({"pairs": P, "iters": ITERS, "points_per_pair": float(out["n_inl"].float().mean()), "ms": ms, "pairs_per_s": P / ms * 1e3,
       "median_cost_ratio": float(np.median(cost[:, 1] / cost[:, 0])), "accepted_mean": float(r[3].float().mean())})
