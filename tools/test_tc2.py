"""Bit-exactness and timing of the CTA-pair tensor-core matcher (variant 17) against the integer-pipe kernel (11) and
the single-CTA tensor-core kernel (16).   python tools/test_tc2.py [--time]"""
import argparse
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import ops, synth  # noqa: E402


def run(ctx, variant, des, pq, pt, fc=None):
    ctx.set_matcher_config(4, variant, 6)
    try:
        idx, dist, keep = ops.knn2_hamming256_batched(des, pq, pt, frame_count=fc, ctx=ctx)
        torch.cuda.synchronize()
    finally:
        ctx.set_matcher_config()
    return idx, dist, keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--only", type=int, default=None, help="time this variant only, skip the checks (for ncu)")
    a = ap.parse_args()
    ctx = ops.context(0)
    rng = np.random.default_rng(0)
    ok = True
    for (F, N, B, ragged) in [] if a.only is not None else [(2, 256, 1, False), (2, 2048, 1, False), (8, 2048, 7, False), (6, 700, 5, True), (4, 1000, 3, True),
                              (3, 4099, 2, False), (2, 8192, 1, False), (64, 2048, 63, False)]:
        q, t = synth.correlated_descriptors(N, N, seed=N + B, clutter=0.2)
        des = np.stack([(q if f % 2 else t) ^ rng.integers(0, 2, (N, 32), dtype=np.uint8) * (f // 2 % 2) for f in range(F)])
        if N <= 1000:
            des[1, 5:15] = des[1, 4]            # forced ties
        pq = np.array([(b + 1) % F for b in range(B)], dtype=np.int32)
        pt = np.array([b % F for b in range(B)], dtype=np.int32)
        fc = None
        if ragged:
            fc = rng.integers(1, N + 1, F).astype(np.int32)
            fc[0] = N
        d = torch.from_numpy(des).cuda()
        r11 = run(ctx, 11, d, pq, pt, fc)
        r17 = run(ctx, 17, d, pq, pt, fc)
        same = True
        for b in range(B):
            nq = N if fc is None else int(fc[pq[b]])
            for x, y in zip(r11, r17):
                same &= bool(torch.equal(x[b, :nq], y[b, :nq]))
        print("F=%d N=%d B=%d ragged=%s identical=%s" % (F, N, B, ragged, same), flush=True)
        ok &= same
    if a.time:
        import bench
        P = 1024
        des_h, _ = bench.make_batch(P, bench.SEED)
        d = torch.from_numpy(des_h).cuda()
        pq = torch.arange(1, 2 * P, 2, dtype=torch.int32).numpy()
        pt = torch.arange(0, 2 * P, 2, dtype=torch.int32).numpy()
        ref = None
        for v in ((a.only,) if a.only is not None else (16, 17)):
            for _ in range(2):
                r = run(ctx, v, d, pq, pt)
            ctx.set_matcher_config(4, v, 6)
            pqd, ptd = torch.from_numpy(pq).cuda(), torch.from_numpy(pt).cuda()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ops.knn2_hamming256_batched(d, pqd, ptd, ctx=ctx)
            e1.record()
            torch.cuda.synchronize()
            ctx.set_matcher_config()
            ms = e0.elapsed_time(e1) / 10
            print("variant %d: %.3f ms per 1024 pairs of 2048x2048 = %.2f Tcmp/s = %.0f TFLOP/s" %
                  (v, ms, P * 2048 * 2048 / ms / 1e9, P * 2048 * 2048 * 512 / ms / 1e9), flush=True)
            if ref is None:
                ref = r
            else:
                print("identical to variant 16:", all(bool(torch.equal(x, y)) for x, y in zip(ref, r)))
    print("ALL IDENTICAL" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
