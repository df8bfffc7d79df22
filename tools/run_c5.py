"""Config C5: one huge keyframe pair split over the GPUs of one node (torchrun, one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_c5.py --kpts 65536 --hyp 65536 [--check]

Matcher: query rows split; RANSAC: hypotheses split; exchange = one all-gather of match slices + ONE
all-reduce(MAX) of the packed key (openwork_b200/dist.py).  Prints one JSON line on rank 0.  With --check
the G-rank result is compared bit-for-bit with the same pair run on rank 0 alone.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import dist as tdist  # noqa: E402
from openwork_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kpts", type=int, default=65536)
    ap.add_argument("--hyp", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K = synth.K_640x480
    lk, ld, ck, cd = synth.make_pair_large(a.kpts, 5005)
    words = synth.sample_words(a.hyp, 5005)
    # replicated device copies (descriptor sets are <= 2 MiB)
    ldd, cdd = torch.from_numpy(ld).cuda(), torch.from_numpy(cd).cuda()
    lkd, ckd = torch.from_numpy(lk).cuda(), torch.from_numpy(ck).cuda()

    def run():
        return tdist.huge_pair(ckd, cdd, lkd, ldd, K, words)

    r = run()
    for _ in range(2):
        r = run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        r = run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ok = None
    if a.check and rank == 0:
        # single-rank reference of the same pair: plain ops calls, no process group involved
        idx, dist_, keep = ops.knn2_hamming256(cdd, ldd, return_keep=True)
        loc = ops.ratio_compact(idx, keep, ckd, lkd, K)
        one = ops.ransac_essential(loc["src"], loc["dst"], words=words)
        ok = bool(torch.equal(loc["qidx"], r["qidx"]) and torch.equal(loc["tidx"], r["tidx"]) and
                  one["best"] == r["best"] and one["n_inl"] == r["count"] and torch.equal(one["mask"], r["mask"]) and
                  torch.equal(one["E"], r["E"]))
    if rank == 0:
        print(json.dumps({"config": "C5 huge pair", "n": a.kpts, "hypotheses": a.hyp, "n_gpus": world, "ms_per_pair": ms,
                          "pairs_per_s": 1e3 / ms, "hamming_gcmp_s": a.kpts * a.kpts / ms / 1e6, "matches": int(r["M"]),
                          "best": int(r["best"]), "inliers": int(r["count"]), "points": int(r["points4d"].shape[0]),
                          "equals_single_rank": ok}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
