"""Experiment: do two half-batches on two streams (two contexts) overlap matcher (ALU/XU pipes) with
scoring (FMA pipe)?  Prints ms per 1024 pairs for 1, 2 and 4 streams."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from openwork_b200 import _lib, ops, synth  # noqa: E402

P = 1024
K = synth.K_640x480
des_h, kps_h = bench.make_batch(P)
des = torch.from_numpy(des_h).cuda()
kps = torch.from_numpy(kps_h).cuda()
words = torch.from_numpy(synth.sample_words(1024, 1).view(np.int32)).cuda()
outs = ("m_count", "E", "n_inl", "Rt", "poses", "points4d", "good", "status")
for ns in (1, 2, 4):
    ctxs = [_lib.Context(0) for _ in range(ns)]
    streams = [torch.cuda.Stream() for _ in range(ns)]
    per = P // ns
    args = []
    for i in range(ns):
        pq = torch.arange(2 * i * per + 1, 2 * (i + 1) * per, 2, dtype=torch.int32).cuda()
        pt = pq - 1
        so = torch.arange(0, per + 1, dtype=torch.int32).cuda()
        args.append([pq, pt, so, None])

    def step():
        for i in range(ns):
            with torch.cuda.stream(streams[i]):
                a = args[i]
                a[3] = ops.process_pairs_device(des, kps, K, words=words, pair_q=a[0], pair_t=a[1], seq_off=a[2],
                                                outputs=outs, out=a[3], ctx=ctxs[i])
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(10):
        step()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    print("streams", ns, "ms per 1024 pairs", e0.elapsed_time(e1) / 10, flush=True)
