"""Config C4: 64 synthetic sequences x 500 frame pairs at 2k keypoints, sequences sharded over the GPUs of
one node (sequence s -> rank s mod G), NO data-path collective.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29521 \
        tools/run_c4.py --seqs 64 --pairs 500 [--distinct 2] [--check]

Each rank synthesises `--distinct` different 501-frame sequences and places copies of them at distinct HBM
addresses for the rest of its share (timing needs the bytes, not new scenes).  Whole sequences are processed
per call (pose chain = prefix product per sequence on the device).  Rank 0 prints one JSON line:
frame-pairs/s over all ranks, device-timed, max over ranks.
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
    ap.add_argument("--seqs", type=int, default=64)
    ap.add_argument("--pairs", type=int, default=500)
    ap.add_argument("--kpts", type=int, default=2048)
    ap.add_argument("--hyp", type=int, default=1024)
    ap.add_argument("--distinct", type=int, default=2)
    ap.add_argument("--seqs-per-call", type=int, default=2)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K = synth.K_640x480
    mine = tdist.shard_sequences(a.seqs, rank, world)
    F = a.pairs + 1
    base = []
    for d in range(min(a.distinct, len(mine))):
        frames, _ = synth.make_sequence(F, a.kpts, seed=4000 + mine[d])      # seed = 1000*config + sequence_id
        base.append((np.stack([f[1] for f in frames]), np.stack([f[0] for f in frames])))
    des = torch.empty((len(mine), F, a.kpts, 32), dtype=torch.uint8, device="cuda")
    kps = torch.empty((len(mine), F, a.kpts, 2), dtype=torch.float64, device="cuda")
    for i in range(len(mine)):
        d, k = base[i % len(base)]
        des[i].copy_(torch.from_numpy(d))
        kps[i].copy_(torch.from_numpy(k))
    words = torch.from_numpy(synth.sample_words(a.hyp, 4004).view(np.int32)).cuda()
    spc = a.seqs_per_call
    outs = ("m_count", "n_inl", "E", "Rt", "poses", "points4d", "good", "status")
    bufs = {}

    def sweep():
        n_ok = 0
        for s0 in range(0, len(mine), spc):
            ns = min(spc, len(mine) - s0)
            seq_off = torch.arange(0, ns + 1, dtype=torch.int32, device="cuda") * a.pairs
            key = ns
            bufs[key] = ops.process_pairs_device(des[s0:s0 + ns].reshape(ns * F, a.kpts, 32),
                                                 kps[s0:s0 + ns].reshape(ns * F, a.kpts, 2), K, words=words,
                                                 seq_off=seq_off, outputs=outs, out=bufs.get(key), batch=ns * a.pairs)
        return n_ok

    # consecutive pairing inside a call needs B = ns*pairs; tell ops via pair lists-free path: seq_off[-1] on host
    sweep()
    torch.cuda.synchronize()
    status_bad = int((bufs[min(spc, len(mine))]["status"] != 0).sum().item())
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        sweep()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    check = None
    if a.check and rank == 0:
        from oracle import two_view as tv
        d, k = base[0]
        o = {kk: v.cpu().numpy() for kk, v in bufs[min(spc, len(mine))].items()}
        wn = words.cpu().numpy().view(np.uint32)
        pose = np.eye(4)
        check = True
        for b in range(3):
            ref = tv.process_pair(k[b], d[b], k[b + 1], d[b + 1], pose, K, words=wn)
            check &= int(o["n_inl"][b]) == int(ref["inliers"].sum()) and np.allclose(o["poses"][b + 1], ref["curr_pose"], atol=1e-5)
            pose = ref["curr_pose"]
    if rank == 0:
        total_pairs = a.seqs * a.pairs
        print(json.dumps({"config": "C4 sequences sharded", "seqs": a.seqs, "pairs_per_seq": a.pairs, "kpts": a.kpts,
                          "hypotheses": a.hyp, "n_gpus": world, "ms_per_sweep": ms, "pairs_per_s": total_pairs / ms * 1e3,
                          "hamming_gcmp_s": total_pairs * a.kpts * a.kpts / ms / 1e6, "status_nonzero_rank0": status_bad,
                          "oracle_check_first_pairs": check}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
