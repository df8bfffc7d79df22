"""Times the RANSAC scoring stage of the C2 batch (bench.py's step) for every scoring configuration
(hypotheses per thread x resident CTAs per SM) with the stage events of tvg_pairs_device.

    python tools/score_sweep.py [--pairs 1024]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from openwork_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1024)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = ops.context(0)
    P = a.pairs
    des_h, kps_h = bench.make_batch(P, bench.SEED)
    des, kps = torch.from_numpy(des_h).to(dev), torch.from_numpy(kps_h).to(dev)
    words = torch.from_numpy(synth.sample_words(bench.N_HYP, bench.SEED).view(np.int32)).to(dev)
    pq = torch.arange(1, 2 * P, 2, dtype=torch.int32, device=dev)
    pt = torch.arange(0, 2 * P, 2, dtype=torch.int32, device=dev)
    so = torch.arange(0, P + 1, dtype=torch.int32, device=dev)
    outs = ("m_count", "n_inl", "status", "best", "counts")
    out = None
    ref = None
    rows = []
    for hpt, ctas in [(0, 0), (2, 0), (4, 0)] + [(2, c) for c in range(4, 11)] + [(4, c) for c in range(3, 8)]:
        ctx.set_score_config(hpt, ctas)
        ctx.set_profile(True)
        acc = None
        for i in range(6):
            out = ops.process_pairs_device(des, kps, synth.K_640x480, words=words, pair_q=pq, pair_t=pt, seq_off=so,
                                           outputs=outs, out=out, ctx=ctx)
            sm = ctx.stage_ms()
            if i >= 2:
                acc = sm if acc is None else {k: acc[k] + v for k, v in sm.items()}
        sm = {k: v / 4 for k, v in acc.items()}
        evals = bench.N_HYP * int(out["m_count"].sum().item())
        tf = evals * 36 / (sm["score"] / 1e3) / 1e12
        sig = (out["best"].clone(), out["n_inl"].clone(), out["counts"].clone())
        if ref is None:
            ref = sig
        same = all(torch.equal(x, y) for x, y in zip(ref, sig))
        rows.append({"hpt": hpt, "ctas_per_sm": ctas, "score_ms": sm["score"], "tflops": tf, "frac_fp32_peak": tf / 74.45,
                     "total_ms": sum(sm.values()), "identical_results": same})
        print(json.dumps(rows[-1]), flush=True)
    ctx.set_score_config(0, 0)
    ctx.set_profile(False)


if __name__ == "__main__":
    main()
