"""Matcher throughput sweep on the GPU box (CUDA events on the launching stream).
Writes gpurun_out/sweep_matcher.json.  Not part of the product."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import ops, synth  # noqa: E402


def time_it(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ctx = ops.context(0)
    res = []
    cases = [("8k", 8192, 1), ("2k", 2048, 1), ("2k_x256", 2048, 256), ("64k", 65536, 1)]
    for name, n, batch in cases:
        rng = np.random.default_rng(0)
        des = torch.from_numpy(rng.integers(0, 256, (batch + 1, n, 32), dtype=np.uint8)).cuda()
        pq = torch.arange(1, batch + 1, dtype=torch.int32).cuda()
        pt = torch.arange(0, batch, dtype=torch.int32).cuda()
        for r in (2, 4, 8):
            for variant in (0, 1, 2):
                for ctas in ((4, 8) if batch == 1 else (6,)):
                    if name == "64k" and (r == 2 or ctas != 4):
                        continue
                    ctx.set_matcher_config(r, variant, ctas)
                    fn = lambda: ops.knn2_hamming256_batched(des, pq, pt)
                    med, best = time_it(fn, iters=5 if name == "64k" else 20)
                    cmp = batch * n * n
                    rec = dict(case=name, r=r, variant=variant, ctas=ctas, ms_med=med, ms_min=best,
                               gcmp_s=cmp / med / 1e6)
                    print(rec, flush=True)
                    res.append(rec)
    ctx.set_matcher_config(4, 0, 6)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/sweep_matcher.json", "w"), indent=1)


if __name__ == "__main__":
    main()
