"""Phase timing of dist.huge_pair (wall clock with syncs; diagnostic only)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import ops, synth
K = synth.K_640x480
n, H = 65536, 65536
lk, ld, ck, cd = synth.make_pair_large(n, 5005)
words = synth.sample_words(H, 5005)
ldd, cdd = torch.from_numpy(ld).cuda(), torch.from_numpy(cd).cuda()
lkd, ckd = torch.from_numpy(lk).cuda(), torch.from_numpy(ck).cuda()
def T(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    print("%-28s %8.3f ms" % (name, (time.perf_counter() - t) / reps * 1e3), flush=True)
    return r
for frac in (1, 8):
    print("--- query/hypothesis share 1/%d" % frac)
    q = n // frac
    idx, dist, keep = T("knn", lambda: ops.knn2_hamming256(cdd[:q], ldd, return_keep=True))
    loc = T("ratio_compact", lambda: ops.ratio_compact(idx, keep, ckd[:q], lkd, K))
idx, dist, keep = ops.knn2_hamming256(cdd, ldd, return_keep=True)
loc = ops.ratio_compact(idx, keep, ckd, lkd, K)
src, dst = loc["src"], loc["dst"]
for frac in (1, 8):
    r = T("ransac H/%d" % frac, lambda: ops.ransac_essential(src, dst, words=words[: H // frac]))
w = T("ransac winner (H=1)", lambda: ops.ransac_essential(src, dst, words=words[r["best"]:r["best"] + 1]))
m = w["mask"]
ckk, lkk = ckd[loc["qidx"].long()][m], lkd[loc["tidx"].long()][m]
T("triangulate all", lambda: ops.triangulate(lkk, ckk, np.eye(4), np.eye(4) @ w["Rt"].cpu().numpy(), K))
