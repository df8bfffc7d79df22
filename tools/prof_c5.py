"""Short driver for ncu: config C5 (one 64k x 64k pair, 65536 hypotheses) through tvg_huge_pair on ONE GPU.
    python tools/prof_c5.py [--reps 3]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openwork_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--kpts", type=int, default=65536)
    ap.add_argument("--hyp", type=int, default=65536)
    a = ap.parse_args()
    lk, ld, ck, cd = synth.make_pair_large(a.kpts, 5005)
    words = torch.from_numpy(synth.sample_words(a.hyp, 5005).view(np.int32)).cuda()
    t = [torch.from_numpy(x).cuda() for x in (ck, cd, lk, ld)]
    out = None
    for _ in range(a.reps):
        out = ops.huge_pair(t[0], t[1], t[2], t[3], synth.K_640x480, words=words, out=out)
    torch.cuda.synchronize()
    print("ok", int(out["m_count"]), int(out["n_inl"]), int(out["best"]), int(out["status"]))


if __name__ == "__main__":
    main()
