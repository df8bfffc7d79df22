"""Matcher throughput sweep on the GPU box (CUDA events on the launching stream, kernel launched
straight through the C ABI with preallocated outputs).  Writes gpurun_out/sweep_matcher.json."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import ops  # noqa: E402


def time_it(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ctx = ops.context(0)
    res = []
    cases = [("2k_x1024", 2048, 1024), ("8k", 8192, 1), ("2k", 2048, 1), ("64k", 65536, 1)]
    variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,3,5,7".split(","))]
    for name, n, batch in cases:
        rng = np.random.default_rng(0)
        des = torch.from_numpy(rng.integers(0, 256, (batch + 1, n, 32), dtype=np.uint8)).cuda()
        pq = torch.arange(1, batch + 1, dtype=torch.int32).cuda()
        pt = torch.arange(0, batch, dtype=torch.int32).cuda()
        idx = torch.empty((batch, n, 2), dtype=torch.int32, device="cuda")
        dist = torch.empty((batch, n, 2), dtype=torch.int32, device="cuda")
        keep = torch.empty((batch, n), dtype=torch.uint8, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = lambda t: C.c_void_p(t.data_ptr())

        def fn():
            ops.check(ctx.lib.tvg_knn2_hamming256(ctx.handle, P(des), P(des), batch, n, n, n, P(pq), P(pt), None, None,
                                                  0.75, P(idx), P(dist), P(keep), st), "knn2")
        ref = None
        for r in (2, 4, 8):
            for variant in variants:
                for ctas in ((4, 6, 8) if batch == 1 else (6,)):
                    if name == "64k" and (r == 2 or ctas != 6):
                        continue
                    try:
                        ctx.set_matcher_config(r, variant, ctas)
                    except Exception:
                        continue
                    ms = time_it(fn, iters=5 if name in ("64k", "2k_x1024") else 50)
                    cur = (idx.clone(), dist.clone(), keep.clone())
                    if ref is None:
                        ref = cur
                    same = all(torch.equal(a, b) for a, b in zip(ref, cur))
                    rec = dict(case=name, r=r, variant=variant, ctas=ctas, ms=ms, gcmp_s=batch * n * n / ms / 1e6,
                               identical=same)
                    print(rec, flush=True)
                    res.append(rec)
    ctx.set_matcher_config(4, 11, 6)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/sweep_matcher.json", "w"), indent=1)


if __name__ == "__main__":
    main()
