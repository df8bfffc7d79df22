"""Config C5 alone: one huge keyframe pair split over the GPUs of one node (torchrun, one rank per GPU) through
tvg_huge_pair over the context's NCCL communicator -- what `bench.py --gpus G` reports as its "c5" extra.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_c5.py [--reps 5]
Prints one JSON line on rank 0 (ms per pair = max over ranks, stage times of rank 0, equals_single_rank).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from openwork_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms, stage, same, info = bench.c5_huge_pair(ops, synth.K_640x480, dev, rank, world, local, barrier, reps=a.reps)
    t = torch.tensor([ms, 0.0 if same in (None, True) else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps(dict(info, n_gpus=world, ms_per_pair=float(t[0]), stage_ms_rank0=stage,
                              equals_single_rank=None if same is None else bool(t[1] == 0))))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
