"""Short driver for ncu: a few steps of bench.py's C2 batch (1024 pairs) through tvg_pairs_device.

    python tools/prof_step.py [--pairs 1024] [--steps 4] [--score-hpt 0] [--score-ctas 0]
"""
import argparse
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
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--score-hpt", type=int, default=0)
    ap.add_argument("--score-ctas", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ctx = ops.context(0)
    ctx.set_score_config(a.score_hpt, a.score_ctas)
    P = a.pairs
    des_h, kps_h = bench.make_batch(P, bench.SEED)
    des, kps = torch.from_numpy(des_h).to(dev), torch.from_numpy(kps_h).to(dev)
    words = torch.from_numpy(synth.sample_words(bench.N_HYP, bench.SEED).view(np.int32)).to(dev)
    pq = torch.arange(1, 2 * P, 2, dtype=torch.int32, device=dev)
    pt = torch.arange(0, 2 * P, 2, dtype=torch.int32, device=dev)
    so = torch.arange(0, P + 1, dtype=torch.int32, device=dev)
    out = None
    for _ in range(a.steps):
        out = ops.process_pairs_device(des, kps, synth.K_640x480, words=words, pair_q=pq, pair_t=pt, seq_off=so,
                                       outputs=("m_count", "n_inl", "status", "points4d", "good", "poses"), out=out, ctx=ctx)
    torch.cuda.synchronize()
    assert int((out["status"] != 0).sum()) == 0
    print("ok", int(out["m_count"].sum()), int(out["n_inl"].sum()))


if __name__ == "__main__":
    main()
