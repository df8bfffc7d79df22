#!/usr/bin/env python
"""bench.py -- headline benchmark of the two-view geometry hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (our arm, sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path on host cores)

Workload (config C2 of BASELINE.json, "single frame pair 2k x 2k 256-bit descriptors, 1024 seeded
RANSAC hypotheses, triangulate inliers, 1 B200"): one STEP = one pass of the whole path
(match -> ratio/compact -> RANSAC fit/score/select/refit -> extract_Rt -> pose -> triangulate -> z>0)
over a batch of P independent synthetic frame pairs of that shape.  P is chosen so that the inputs
(descriptors + keypoints of 2P frames) exceed the 126 MB L2, hence no L2 flush between iterations.
`value` = frame pairs / s with inputs resident in HBM; `e2e` = the same through the host-buffer C-ABI
entry (tvg_pairs_host) with pinned host inputs, H2D + D2H inside the timed region.
Multi-GPU: every rank processes its own P pairs (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_KPTS = 2048
N_HYP = 1024
DISTINCT_PAIRS = 64      # synthetic pairs generated; the batch tiles them (distinct HBM addresses)
THR = 0.005
RATIO = 0.75
SEED = 2002              # 1000*config + id, SURVEY.md §8d


def make_batch(pairs, seed=SEED):
    """2*pairs frames: pair b = (last = frame 2b, curr = frame 2b+1). Returns des u8[2P,N,32], kps f64[2P,N,2]."""
    from openwork_b200 import synth
    base = []
    for i in range(min(pairs, DISTINCT_PAIRS)):
        lk, ld, ck, cd = synth.make_pair(N_KPTS, seed + i)
        base.append((lk, ld, ck, cd))
    des = np.empty((2 * pairs, N_KPTS, 32), dtype=np.uint8)
    kps = np.empty((2 * pairs, N_KPTS, 2), dtype=np.float64)
    for b in range(pairs):
        lk, ld, ck, cd = base[b % len(base)]
        des[2 * b], des[2 * b + 1] = ld, cd
        kps[2 * b], kps[2 * b + 1] = lk, ck
    return des, kps


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons during the timed region (NVML, 10 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_evt = index, threading.Event()
        self.sm, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self.stop_evt.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self.stop_evt.wait(0.01)

    def summary(self):
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "power_w_max": float(max(self.power)) if self.power else None,
                "samples": len(self.sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ----------------------------------------------------------------------------------------- reference arm
def _cpu_pair_worker(args):
    """One frame pair through the CPU oracle (cv2 BFMatcher + restated skimage RANSAC + per-point LAPACK)."""
    import cv2
    from oracle import two_view as tv
    cv2.setNumThreads(1)
    lk, ld, ck, cd, K, words = args
    r = tv.process_pair(lk, ld, ck, cd, np.eye(4), K, words=words, residual_threshold=THR)
    return int(r["inliers"].sum())


def cpu_reference_rate(n_pairs, cores, seed=SEED):
    """frame pairs/s of the reference's CPU path on `cores` host processes (bounded sample)."""
    import multiprocessing as mp
    from openwork_b200 import synth
    K = synth.K_640x480
    words = synth.sample_words(N_HYP, seed)
    base = [synth.make_pair(N_KPTS, seed + i) for i in range(min(n_pairs, 8))]
    jobs = [(*base[i % len(base)], K, words) for i in range(n_pairs)]
    if cores <= 1:
        _cpu_pair_worker(jobs[0])
        t0 = time.perf_counter()
        for j in jobs:
            _cpu_pair_worker(j)
        return n_pairs / (time.perf_counter() - t0)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_pair_worker, jobs[:cores])          # warm-up (imports, page-in)
        t0 = time.perf_counter()
        pool.map(_cpu_pair_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return n_pairs / dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0))
    per_step = 2 * cores
    import multiprocessing as mp
    from openwork_b200 import synth
    K = synth.K_640x480
    words = synth.sample_words(N_HYP, SEED)
    base = [synth.make_pair(N_KPTS, SEED + i) for i in range(8)]
    jobs = [(*base[i % len(base)], K, words) for i in range(per_step)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_pair_worker, jobs, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_pair_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = "%d pairs/step on a %d-process pool (cv2 BFMatcher + restated skimage ransac + reference triangulate loop)" % (
        per_step, cores)
    line = {
        "impl": "reference", "metric": "frame pairs/s (2k kpts, match+RANSAC+triangulate)", "value": value,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 popcount + f64", "data": "synthetic",
        "config": {"workload": "C2: frame pairs 2048x2048 256-bit descriptors, 1024 seeded RANSAC hypotheses, "
                               "triangulate inliers", "pairs_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0



def single_pair_latency(ops, ctx, K, n, H, dev, reps=30):
    """One frame pair (config C2: n=2048,H=1024 / C3: n=8192,H=4096) through the sync-free device pipeline:
    microseconds per pair launched eagerly and replayed from a CUDA graph, plus the matcher alone."""
    import ctypes as C
    import torch
    from openwork_b200 import synth
    lk, ld, ck, cd = synth.make_pair_large(n, SEED + n)
    des = torch.from_numpy(np.stack([ld, cd])).to(dev)
    kps = torch.from_numpy(np.stack([lk, ck])).to(dev)
    words = torch.from_numpy(synth.sample_words(H, SEED).view(np.int32)).to(dev)
    pq = torch.tensor([1], dtype=torch.int32, device=dev)
    pt = torch.tensor([0], dtype=torch.int32, device=dev)
    so = torch.tensor([0, 1], dtype=torch.int32, device=dev)
    outs = ("m_count", "E", "n_inl", "Rt", "poses", "points4d", "good", "status")
    out = None

    def call():
        nonlocal out
        out = ops.process_pairs_device(des, kps, K, words=words, pair_q=pq, pair_t=pt, seq_off=so, ratio=RATIO,
                                       residual_threshold=THR, outputs=outs, out=out, ctx=ctx)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    eager_us = timed(call, reps)
    graph_us = None
    try:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            call()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            call()
        graph_us = timed(g.replay, reps)
    except Exception as e:  # pragma: no cover
        graph_us = "capture failed: %r" % (e,)
    idx = torch.empty((1, n, 2), dtype=torch.int32, device=dev)
    dst = torch.empty((1, n, 2), dtype=torch.int32, device=dev)
    keep = torch.empty((1, n), dtype=torch.uint8, device=dev)
    P = lambda t: C.c_void_p(t.data_ptr())

    def knn():
        ops.check(ctx.lib.tvg_knn2_hamming256(ctx.handle, P(des), P(des), 1, n, n, n, P(pq), P(pt), None, None, RATIO,
                                              P(idx), P(dst), P(keep),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "knn2")
    knn_us = timed(knn, reps)
    ctx.set_matcher_config(4, 11, 6)
    knn_int_us = timed(knn, reps)
    ctx.set_matcher_config(4, 17, 6)
    knn_tc_us = timed(knn, reps)
    ctx.set_matcher_config()
    return {"matcher_us_int_pipe": knn_int_us, "matcher_us_tensor": knn_tc_us, "kpts": n, "hypotheses": H, "matches": int(out["m_count"][0]), "inliers": int(out["n_inl"][0]),
            "eager_us_per_pair": eager_us, "cuda_graph_us_per_pair": graph_us, "matcher_us": knn_us,
            "matcher_gcmp_s": n * n / knn_us / 1e3}


def process_frame_latency(K, n=2048, H=1024, reps=7):
    """Wall-clock microseconds of openwork_b200.slam.process_frame for one new frame (numpy in, numpy out, the call a
    user of the reference's slam.py makes per video frame): the default single-pipeline-call path and the
    function-by-function path."""
    import torch
    from openwork_b200 import slam, synth
    slam.K, slam.mapp = K, None
    slam.ransac_config.update(max_trials=H, residual_threshold=THR, seed=0)
    lk, ld, ck, cd = synth.make_pair_large(n, SEED + n)
    res = {"kpts": n, "hypotheses": H}
    for fused in (True, False):
        slam.FUSED_PROCESS_FRAME = fused
        times = []
        for _ in range(reps):
            slam.Frame.reset()
            slam.process_frame(slam.Frame(kps=lk, des=ld))
            torch.cuda.synchronize()
            t = time.perf_counter()
            slam.process_frame(slam.Frame(kps=ck, des=cd))
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t)
        res["fused_us" if fused else "stepwise_us"] = sorted(times[2:])[len(times[2:]) // 2] * 1e6
    slam.FUSED_PROCESS_FRAME = True
    slam.Frame.reset()
    return res


def frontend_rate(ops, ctx, dev, frames=256):
    """Not part of the metric (the north star feeds synthetic keypoints): frames/s of the GPU feature front-end
    (BGR2GRAY + goodFeaturesToTrack + ORB.compute, bit-exact vs cv2) on procedurally textured 640x480 images."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(7)
    low = torch.randint(0, 256, (8, 3, 60, 80), generator=g, dtype=torch.uint8).float()
    img = torch.nn.functional.interpolate(low, size=(480, 640), mode="bicubic", align_corners=False)
    img = (img + 12.0 * torch.randn(img.shape, generator=g)).clamp(0, 255).to(torch.uint8)
    img = img.permute(0, 2, 3, 1).contiguous().repeat(frames // 8, 1, 1, 1).to(dev)
    for _ in range(2):
        kps, des, n = ops.extract_points(img, ctx=ctx)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        kps, des, n = ops.extract_points(img, ctx=ctx)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / 3
    return {"frames_per_call": frames, "image": "640x480 BGR", "frames_per_s": frames / ms * 1e3,
            "keypoints_per_frame": float(n.float().mean().item())}


# ----------------------------------------------------------------------------------------- C4 / C5

C4_SEQS, C4_PAIRS = 64, 500       # BASELINE.json config C4: 64 sequences x 500 frame pairs at 2k keypoints
C5_N, C5_H = 65536, 65536         # config C5: one 64k x 64k pair, 65536 hypotheses


def c4_sequences(ops, ctx, K, dev, rank, world, barrier, reps=2, seqs_per_call=2):
    """Config C4: sequence s -> rank s mod G, NO data-path collective.  Frames are shared between consecutive pairs
    (slam.py:243, 271: pair (t-1, t) of a sequence), whole sequences per call, pose chain on the device.  Each rank
    synthesises two distinct 501-frame sequences and places copies at distinct HBM addresses for the rest of its
    share (timing needs the bytes, not new scenes).  Returns (local ms per sweep, info dict)."""
    import torch
    from openwork_b200 import dist as tdist, synth
    mine = tdist.shard_sequences(C4_SEQS, rank, world)
    F = C4_PAIRS + 1
    base = []
    for d in range(min(2, len(mine))):
        frames, _ = synth.make_sequence(F, N_KPTS, seed=4000 + mine[d])       # seed = 1000*config + sequence_id
        base.append((np.stack([f[1] for f in frames]), np.stack([f[0] for f in frames])))
    des = torch.empty((len(mine), F, N_KPTS, 32), dtype=torch.uint8, device=dev)
    kps = torch.empty((len(mine), F, N_KPTS, 2), dtype=torch.float64, device=dev)
    for i in range(len(mine)):
        d, k = base[i % len(base)]
        des[i].copy_(torch.from_numpy(d))
        kps[i].copy_(torch.from_numpy(k))
    words = torch.from_numpy(synth.sample_words(N_HYP, 4004).view(np.int32)).to(dev)
    outs = ("m_count", "n_inl", "E", "Rt", "poses", "points4d", "good", "status")
    bufs = {}

    def sweep():
        for s0 in range(0, len(mine), seqs_per_call):
            ns = min(seqs_per_call, len(mine) - s0)
            seq_off = torch.arange(0, ns + 1, dtype=torch.int32, device=dev) * C4_PAIRS
            bufs[ns] = ops.process_pairs_device(des[s0:s0 + ns].reshape(ns * F, N_KPTS, 32),
                                                kps[s0:s0 + ns].reshape(ns * F, N_KPTS, 2), K, words=words,
                                                seq_off=seq_off, ratio=RATIO, residual_threshold=THR, outputs=outs,
                                                out=bufs.get(ns), batch=ns * C4_PAIRS, ctx=ctx)
    sweep()
    torch.cuda.synchronize()
    bad = sum(int((b["status"] != 0).sum().item()) for b in bufs.values())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sweep()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    info = {"sequences": C4_SEQS, "pairs_per_sequence": C4_PAIRS, "sequences_this_rank": len(mine),
            "frames_shared_between_pairs": True, "pairs_per_call": seqs_per_call * C4_PAIRS,
            "status_nonzero_rank0": bad, "collectives": 0}
    del des, kps
    return ms, info


def c5_huge_pair(ops, K, dev, rank, world, local, barrier, reps=5):
    """Config C5: ONE 64k x 64k pair, 65536 hypotheses, query rows + hypotheses split over the ranks (tvg_huge_pair
    over the context's NCCL communicator: two fixed-size all-gathers).  Returns (local ms per pair, stage ms,
    equals_single_rank, info): equals_single_rank compares, on every rank, the G-rank outputs bit for bit with the
    same call on a context without a communicator."""
    import torch
    import torch.distributed as dist
    from openwork_b200 import _lib, dist as tdist, synth
    lk, ld, ck, cd = synth.make_pair_large(C5_N, 5005)
    words = torch.from_numpy(synth.sample_words(C5_H, 5005).view(np.int32)).to(dev)
    ldd, cdd = torch.from_numpy(ld).to(dev), torch.from_numpy(cd).to(dev)
    lkd, ckd = torch.from_numpy(lk).to(dev), torch.from_numpy(ck).to(dev)
    ctx = tdist.init_comm(ops.context(local))
    outs = ("idx", "dist", "qidx", "tidx", "m_count", "E", "E_best", "best", "best_sse", "mask", "n_inl", "inl_idx", "Rt",
            "poses", "points4d", "good", "status")
    out = None

    def run(c, o):
        return ops.huge_pair(ckd, cdd, lkd, ldd, K, words=words, ratio=RATIO, residual_threshold=THR, outputs=outs,
                             out=o, ctx=c)
    for _ in range(3):
        out = run(ctx, out)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launches
    e0.record()
    for _ in range(reps):
        out = run(ctx, out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / reps
    launches = (ctx.launches - l0) // reps
    ctx.set_profile(True)
    acc = None
    xacc = [0.0, 0.0]
    for _ in range(3):
        barrier()                                        # ranks enter together: what is left in the exchanges is theirs
        out = run(ctx, out)
        sm = ctx.stage_ms()
        acc = sm if acc is None else {k: acc[k] + v for k, v in sm.items()}
        x1, x2 = ctx.exchange_ms()
        xacc = [xacc[0] + x1, xacc[1] + x2]
    ctx.set_profile(False)
    stage = {k: v / 3 for k, v in acc.items()}
    # the two exchange steps on this rank (each includes waiting for the slowest rank): min / max over the ranks
    xt = torch.tensor([xacc[0] / 3, xacc[1] / 3, -xacc[0] / 3, -xacc[1] / 3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(xt, op=dist.ReduceOp.MAX)
    xt = [float(v) for v in xt.cpu()]
    exchange_ms = {"matcher_outputs_allgather": {"max_over_ranks": xt[0], "min_over_ranks": -xt[2]},
                   "best_records_allgather": {"max_over_ranks": xt[1], "min_over_ranks": -xt[3]},
                   "note": "CUDA events around the NCCL calls on every rank, ranks released by a barrier; a rank's time "
                           "includes waiting for the slowest rank, so the minimum is the transfer itself"}
    M, n_in = int(out["m_count"].item()), int(out["n_inl"].item())
    same = None
    if world > 1:
        solo = _lib.Context(local)                       # no communicator: the one-rank run of the same call
        ref = run(solo, None)
        torch.cuda.synchronize()
        same = int(ref["m_count"].item()) == M and int(ref["n_inl"].item()) == n_in
        for k in ("best", "best_sse", "E", "E_best", "Rt", "poses", "status"):
            same = same and bool(torch.equal(ref[k], out[k]))
        for k, n in (("idx", C5_N), ("dist", C5_N), ("qidx", M), ("tidx", M), ("mask", M), ("inl_idx", n_in),
                     ("points4d", n_in), ("good", n_in)):
            same = same and bool(torch.equal(ref[k][:n], out[k][:n]))
        solo.close()
    info = {"kpts": C5_N, "hypotheses": C5_H, "matches": M, "inliers": n_in, "best": int(out["best"].item()),
            "status": int(out["status"].item()), "launches_per_pair": int(launches),
            "collectives_per_pair": 2 if world > 1 else 0, "exchange_ms": exchange_ms,
            "exchange": "all-gather of idx/dist/keep blocks (17 B per query row) + all-gather of one 96-byte record per rank"}
    return ms, stage, same, info


# ----------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from openwork_b200 import ops, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = ops.context(local)
    if args.knn_variant is not None:
        ctx.set_matcher_config(args.knn_r or 4, args.knn_variant, args.knn_ctas)
    ctx.set_host_chunk(args.host_chunk)
    K = synth.K_640x480
    P = args.pairs
    des_h, kps_h = make_batch(P, SEED + 100000 * rank)
    words_h = synth.sample_words(N_HYP, SEED)
    des_pin = torch.from_numpy(des_h).pin_memory()
    kps_pin = torch.from_numpy(kps_h).pin_memory()
    words_pin = torch.from_numpy(words_h.view(np.int32)).pin_memory()
    des = des_pin.to(dev)
    kps = kps_pin.to(dev)
    words = words_pin.to(dev)
    pair_q = torch.arange(1, 2 * P, 2, dtype=torch.int32, device=dev)
    pair_t = torch.arange(0, 2 * P, 2, dtype=torch.int32, device=dev)
    seq_off = torch.arange(0, P + 1, dtype=torch.int32, device=dev)        # every pair is its own sequence
    outs = ("m_count", "E", "n_inl", "Rt", "poses", "points4d", "good", "status", "best")
    out = None

    def step():
        nonlocal out
        out = ops.process_pairs_device(des, kps, K, words=words, pair_q=pair_q, pair_t=pair_t, seq_off=seq_off,
                                       ratio=RATIO, residual_threshold=THR, outputs=outs, out=out, ctx=ctx)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    st = out["status"].cpu().numpy()
    if (st != 0).any():
        raise SystemExit("bench.py: pipeline reported status != 0")

    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0

    # ---- the same step held for seconds (thermal / power steady state), timed the same way; the headline `value`
    # stays the K-step figure the contract asks for
    sus_ms = 0.0
    sus_steps = 0
    if args.sustain_seconds > 0:
        sus_steps = max(args.steps, int(args.sustain_seconds * 1e3 / (ms / args.steps)))
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(sus_steps):
            step()
        s1.record()
        barrier()
        sus_ms = s0.elapsed_time(s1)

    # ---- per-stage times of the same step (CUDA events at the stage boundaries inside tvg_pairs_device)
    ctx.set_profile(True)
    stage_acc = None
    for _ in range(max(3, min(args.steps, 10))):
        step()
        sm = ctx.stage_ms()
        stage_acc = sm if stage_acc is None else {k: stage_acc[k] + v for k, v in sm.items()}
    n_prof = max(3, min(args.steps, 10))
    stage_ms = {k: v / n_prof for k, v in stage_acc.items()}
    ctx.set_profile(False)
    sum_m = int(out["m_count"].sum().item())
    sum_inl = int(out["n_inl"].sum().item())

    # ---- dominant kernel: the Hamming matcher over the same batch, timed alone on its stream
    idx = torch.empty((P, N_KPTS, 2), dtype=torch.int32, device=dev)
    dst = torch.empty((P, N_KPTS, 2), dtype=torch.int32, device=dev)
    keep = torch.empty((P, N_KPTS), dtype=torch.uint8, device=dev)
    import ctypes as C

    def knn():
        ops.check(ctx.lib.tvg_knn2_hamming256(ctx.handle, C.c_void_p(des.data_ptr()), C.c_void_p(des.data_ptr()), P,
                                              N_KPTS, N_KPTS, N_KPTS, C.c_void_p(pair_q.data_ptr()),
                                              C.c_void_p(pair_t.data_ptr()), None, None, RATIO,
                                              C.c_void_p(idx.data_ptr()), C.c_void_p(dst.data_ptr()),
                                              C.c_void_p(keep.data_ptr()),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "knn2")
    for _ in range(3):
        knn()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(args.steps, 5)
    k0.record()
    for _ in range(reps):
        knn()
    k1.record()
    k1.synchronize()
    knn_ms = k0.elapsed_time(k1) / reps
    # the tensor-core matcher's L2-resident mode (sub-batches whose e4m3 expansion never leaves L2)
    knn_l2_ms = None
    if args.knn_variant == 16:
        ctx.set_matcher_l2_budget(40 << 20)
        for _ in range(2):
            knn()
        k0.record()
        for _ in range(reps):
            knn()
        k1.record()
        k1.synchronize()
        knn_l2_ms = k0.elapsed_time(k1) / reps
        ctx.set_matcher_l2_budget(0)
    # the integer-pipe kernel (the north star's LOP3+POPC design) over the same batch, for its INT32/POPC fraction
    ctx.set_matcher_config(4, 11, args.knn_ctas)
    for _ in range(2):
        knn()
    k0.record()
    for _ in range(reps):
        knn()
    k1.record()
    k1.synchronize()
    knn_int_ms = k0.elapsed_time(k1) / reps
    if args.knn_variant is not None:
        ctx.set_matcher_config(args.knn_r or 4, args.knn_variant, args.knn_ctas)
    else:
        ctx.set_matcher_config()
    sampler.stop_evt.set()
    sampler.join()

    # ---- e2e: host buffers in, results out, through the C-ABI host entry (streaming form: two contexts
    # alternate, so the H2D copies of step i+1 overlap the kernels of step i; every step still moves its own
    # inputs host->device and its own results device->host inside the timed region)
    from openwork_b200 import _lib
    # everything process_frame hands on (slam.py:259-269): the pose, the matches, the inlier mask, the triangulated
    # points and their z > 0 flags -- padded per-pair arrays, copied device -> host inside the timed region
    e2e_outs = ("m_count", "E", "n_inl", "Rt", "poses", "status", "qidx", "tidx", "mask", "points4d", "good")
    NC = max(2, args.e2e_contexts)
    ctxs = [ctx] + [_lib.Context(local) for _ in range(NC - 1)]
    for c in ctxs:
        if args.knn_variant is not None:
            c.set_matcher_config(args.knn_r or 4, args.knn_variant, args.knn_ctas)
        c.set_host_chunk(args.host_chunk)
    host_out = [None] * NC
    h2d = d2h = 0
    pq_h, pt_h, so_h = pair_q.cpu(), pair_t.cpu(), seq_off.cpu()

    def e2e_submit(i):
        nonlocal h2d, d2h
        host_out[i % NC], h2d, d2h = ops.process_pairs_host(des_pin, kps_pin, K, words=words_pin, pair_q=pq_h,
                                                            pair_t=pt_h, seq_off=so_h, ratio=RATIO,
                                                            residual_threshold=THR, outputs=e2e_outs,
                                                            out=host_out[i % NC], ctx=ctxs[i % NC], wait=False)

    def e2e_loop(submit, steps):
        """NC batches in flight: the copies of step i overlap the kernels of step i-1 and the result copies of step
        i-2; the results of step i-NC+1 are on the host before step i+1 is submitted."""
        t0 = time.perf_counter()
        for i in range(steps):
            submit(i)
            if i >= NC - 1:
                ops.wait_pairs_host(ctxs[(i - NC + 1) % NC])
        for i in range(max(steps - NC + 1, 0), steps):
            ops.wait_pairs_host(ctxs[i % NC])
        barrier()
        return time.perf_counter() - t0
    for i in range(2 * NC):
        e2e_submit(i)
        ops.wait_pairs_host(ctxs[i % NC])
    barrier()
    e2e_steps = max(4, min(2 * args.steps, 20))
    e2e_s = e2e_loop(e2e_submit, e2e_steps)
    if any((o["status"].numpy() != 0).any() for o in host_out):
        raise SystemExit("bench.py: e2e pipeline reported status != 0")

    # ---- e2e on C4-shaped input: sequences whose consecutive pairs share a frame (slam.py:243, 271), so a pair
    # costs one frame of host->device traffic instead of two.  Same streaming entry, same outputs.
    n_seq_e2e = 8
    ppseq = max(P // n_seq_e2e, 1)
    from openwork_b200 import synth as _synth
    seq_frames = []
    for sidx in range(2):
        fr, _ = _synth.make_sequence(ppseq + 1, N_KPTS, seed=4000 + 100 * rank + sidx)
        seq_frames.append((np.stack([f[1] for f in fr]), np.stack([f[0] for f in fr])))
    des4 = torch.from_numpy(np.concatenate([seq_frames[i % 2][0] for i in range(n_seq_e2e)])).pin_memory()
    kps4 = torch.from_numpy(np.concatenate([seq_frames[i % 2][1] for i in range(n_seq_e2e)])).pin_memory()
    so4 = torch.arange(0, n_seq_e2e + 1, dtype=torch.int32) * ppseq
    host_out4 = [None] * NC
    h2d4 = d2h4 = 0

    def e2e4_submit(i):
        nonlocal h2d4, d2h4
        host_out4[i % NC], h2d4, d2h4 = ops.process_pairs_host(des4, kps4, K, words=words_pin, seq_off=so4, ratio=RATIO,
                                                               residual_threshold=THR, outputs=e2e_outs,
                                                               out=host_out4[i % NC], ctx=ctxs[i % NC], wait=False)
    for c in ctxs:
        c.set_host_chunk(args.host_chunk_c4)
    for i in range(NC):
        e2e4_submit(i)
        ops.wait_pairs_host(ctxs[i % NC])
    barrier()
    e2e4_s = e2e_loop(e2e4_submit, e2e_steps)
    ctx.set_host_chunk(args.host_chunk)
    if any((o["status"].numpy() != 0).any() for o in host_out4):
        raise SystemExit("bench.py: C4-shaped e2e pipeline reported status != 0")
    del ctxs[1:]

    # ---- bare host->device bandwidth with every rank copying at once (what bounds e2e when ranks share the host)
    probe = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    probe_d = torch.empty_like(probe, device=dev)
    probe_d.copy_(probe, non_blocking=True)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(4):
        probe_d.copy_(probe, non_blocking=True)
    p1.record()
    barrier()
    h2d_gbs = 4 * probe.numel() / (p0.elapsed_time(p1) / 1e3) / 1e9
    del probe, probe_d

    # ---- the sharded configs of BASELINE.json: C4 (sequences, no collective) and C5 (one huge pair over NCCL)
    c4_ms = c5_ms = 0.0
    c4_info = c5_info = c5_stage = c5_same = None
    if not args.no_c4:
        c4_ms, c4_info = c4_sequences(ops, ctx, K, dev, rank, world, barrier)
    if not args.no_c5:
        c5_ms, c5_stage, c5_same, c5_info = c5_huge_pair(ops, K, dev, rank, world, local, barrier)

    single = None
    if rank == 0 and not args.no_single_pair:
        single = {"C2": single_pair_latency(ops, ctx, K, 2048, 1024, dev), "C3": single_pair_latency(ops, ctx, K, 8192, 4096, dev),
                  "process_frame_C2": process_frame_latency(K)}

    frontend = None
    if rank == 0 and not args.no_frontend:
        frontend = frontend_rate(ops, ctx, dev)

    t = torch.tensor([ms, knn_ms, e2e_s, e2e4_s, c4_ms, c5_ms, -h2d_gbs, 0.0 if c5_same in (None, True) else 1.0,
                      sus_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, knn_ms, e2e_s, e2e4_s, c4_ms, c5_ms, h2d_min, c5_diff, sus_ms = [float(x) for x in t.cpu()]
    h2d_min = -h2d_min
    if c5_same is not None:
        c5_same = c5_diff == 0.0                      # every rank compared its own copy
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_kind = measured_peaks()
    clocks = sampler.summary()
    value = world * P * args.steps / (ms / 1e3)
    e2e_value = world * P * e2e_steps / e2e_s
    cmp_per_launch = P * N_KPTS * N_KPTS
    gcmp_s = cmp_per_launch / (knn_ms / 1e3) / 1e9
    f_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count
    popc_peak = 2.0 * n_sm * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e9      # 8 POPC/cmp at 16 lanes/clk/SM
    alg_bytes = P * (2 * N_KPTS * 32 + N_KPTS * 17)                           # descriptors in, idx+dist+keep out
    hbm_gbs = alg_bytes / (knn_ms / 1e3) / 1e9
    tensor_path = args.knn_variant in (None, -1, 16, 17)
    gcmp_int = cmp_per_launch / (knn_int_ms / 1e3) / 1e9
    matcher_int = {
        "kernel": "knn2_kernel (LOP3+POPC, variant 11), timed alone on the same batch", "gcmp_s": gcmp_int,
        "ms_per_launch": knn_int_ms, "popc_peak_gcmp_s": popc_peak, "frac_popc_peak": gcmp_int / popc_peak,
        "peak_note": "POPC-pipe bound of the plain 8xPOPC form: 2 cmp/clk/SM x %d SMs x sm_max_mhz (%s); the "
                     "Harley-Seal variant needs 4 POPC/cmp so it may exceed 1.0" % (n_sm, peak_kind),
        # dram__bytes_read.sum + dram__bytes_write.sum at P=1024 from profiles/r1_kernels_matcher_score.md
        "traffic": 154.74e6 * P / 1024,
    }
    if tensor_path:
        # one comparison = a 256-term dot product of +-1 values = 512 flop on the e4m3 tensor path; the peak is
        # measured on this GPU, now: back-to-back tcgen05.mma kind::f8f6f4 with resident operands (tvg_measure_fp8_peak)
        tflops = gcmp_s * 512 / 1e3
        bf16 = peaks.get("bf16_tflops", 2250.0)
        fp8_peak, fp8_ms = ctx.measure_fp8_peak()
        roofline = {
            "kernel": "knn2_tc2_kernel (CTA pairs: tcgen05.mma.cta_group::2 kind::f8f6f4, e4m3 operands expanded inside "
                      "the kernel, f16 accumulate, TMEM top-2 epilogue)" if args.knn_variant != 16 else
                      "knn2_tc_kernel (single-CTA tcgen05.mma) + unpack_train_kernel, timed together",
            "bound": "tensor", "achieved": tflops, "peak": fp8_peak, "unit": "TFLOP/s", "frac": tflops / fp8_peak,
            "peak_note": "measured live on this GPU: %.0f TF/s of e4m3 tcgen05.mma (M=128, N=256, K=32, f16 accumulate, "
                         "operands resident in shared memory, %.1f ms burst; tools/microbench/fp8_mma_peak.cu) -- nominal "
                         "4500; MEASURED_PEAKS.json holds bf16 only (%s %.1f TF/s, this kernel = %.2f x that)" %
                         (fp8_peak, fp8_ms, peak_kind, bf16, tflops / bf16),
            "ms_per_launch": knn_ms, "cmp_per_launch": cmp_per_launch, "gcmp_s": gcmp_s,
            "x_int32_popc_peak": gcmp_s / popc_peak, "x_measured_bf16_peak": tflops / bf16, "sm_mhz_during_run": f_mhz,
            # dram bytes of knn2_tc_kernel + unpack_train_kernel at P=1024 (profiles/r1_tensor_matcher_dram_traffic.md)
            # dram__bytes_read.sum + dram__bytes_write.sum at P=1024 from one ncu --set full capture
            "traffic": (TC2_TRAFFIC_BYTES_P1024 if args.knn_variant != 16 else TC_TRAFFIC_BYTES_P1024) * P / 1024,
            "traffic_note": ("0.93 x the algorithmic bytes (profiles/r2_kernels_matcher_tc2.md): the train rows are expanded "
                             "inside the kernel; part of the re-read descriptors hit in L2") if args.knn_variant != 16 else
                            ("7x the algorithmic bytes: the e4m3 expansion of the train side (512 KB per pair) is written "
                             "and read back once"),
            "l2_resident_mode": {"ms_per_launch": knn_l2_ms, "gcmp_s": cmp_per_launch / (knn_l2_ms / 1e3) / 1e9,
                                 "traffic": TC_TRAFFIC_L2MODE_BYTES_P1024 * P / 1024,
                                 "note": "tvg_set_matcher_l2_budget(40 MiB): sub-batches whose expansion stays in L2; DRAM "
                                         "traffic = the algorithmic bytes, more launches"} if knn_l2_ms else None,
            "algorithmic_flop_per_launch": cmp_per_launch * 512,
            "hbm": {"achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                    "frac": hbm_gbs / peaks.get("hbm_gbs", 6650.0), "algorithmic_bytes_per_launch": alg_bytes},
        }
    else:
        roofline = {
            "kernel": "knn2_kernel (Hamming kNN-2, LOP3+POPC)", "bound": "int_pipe", "achieved": gcmp_s,
            "peak": popc_peak, "unit": "Gcmp/s", "frac": gcmp_s / popc_peak, "peak_note": matcher_int["peak_note"],
            "ms_per_launch": knn_ms, "cmp_per_launch": cmp_per_launch, "sm_mhz_during_run": f_mhz,
            "traffic": 154.74e6 * P / 1024,
            "hbm": {"achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                    "frac": hbm_gbs / peaks.get("hbm_gbs", 6650.0), "algorithmic_bytes_per_launch": alg_bytes,
                    "note": "matcher is integer-pipe bound, far below the HBM roofline"},
        }
    fp32_peak = n_sm * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    evals = N_HYP * sum_m
    score_tflops = evals * 36 / (stage_ms["score"] / 1e3) / 1e12
    tri_bytes = sum_inl * 65
    tri_gbs = tri_bytes / (stage_ms["triangulate"] / 1e3) / 1e9
    kernels = {
        "stage_ms": stage_ms,
        "matcher_in_step": {"gcmp_s": cmp_per_launch / (stage_ms["knn2"] / 1e3) / 1e9,
                            "x_int32_popc_peak": cmp_per_launch / (stage_ms["knn2"] / 1e3) / 1e9 / popc_peak},
        "matcher_int_pipe": matcher_int,
        "ransac_score": {"bound": "fp32", "evaluations": evals, "flop_per_evaluation": 36,
                         "achieved_tflops": score_tflops, "peak_tflops": fp32_peak, "frac": score_tflops / fp32_peak,
                         "note": "19 FFMA per (hypothesis, correspondence) = 36 useful flop (two of the FFMA addends are the "
                                 "folded band-floor constants); peak = SMs x 128 lanes x 2 x sm_max_mhz"},
        "triangulate": {"bound": "hbm (algorithmic) / fp64 pipe (actual)", "points": sum_inl,
                        "algorithmic_bytes": tri_bytes, "achieved_gbs": tri_gbs, "peak_gbs": peaks.get("hbm_gbs"),
                        "frac": tri_gbs / peaks.get("hbm_gbs", 6650.0),
                        "note": "64 B/point (2 x (x,y) f64 in, (x,y,z,1) f64 out) + 1 B flag; per point ~450 fp64 "
                                "instructions (adjugate + repeated squaring; Jacobi SVD only as a fallback)"},
    }
    # `roofline` names the dominant kernel of the step (by the stage events of the timed configuration); the other
    # one stays beside it
    score_roofline = {
        "kernel": "ransac_score_kernel<2,false> (fp32 Sampson scoring of H x M (hypothesis, correspondence) pairs; "
                  "persistent CTAs on a device-side work queue, correspondences staged by TMA bulk copies)",
        "bound": "fp32", "achieved": score_tflops, "peak": fp32_peak, "unit": "TFLOP/s", "frac": score_tflops / fp32_peak,
        "peak_note": "FP32 FMA pipe: %d SMs x 128 lanes x 2 flop x sm_max_mhz (%s clock; MEASURED_PEAKS.json holds no "
                     "fp32 figure); tools/microbench/ffma_operands.cu measures 0.93 of it for FFMAs with two fresh "
                     "register operands and 0.46 with three (register-file read bandwidth)" % (n_sm, peak_kind),
        "ms_per_launch": stage_ms["score"], "evaluations_per_launch": evals, "flop_per_evaluation": 36,
        "algorithmic_flop_per_launch": evals * 36, "sm_mhz_during_run": f_mhz,
        # dram__bytes_read.sum + dram__bytes_write.sum at P=1024 (profiles/r2_kernels_score_fit_tri.md)
        "traffic": (62.87e6 + 7.18e6) * P / 1024,
        "traffic_note": "algorithmic bytes: fp32 correspondences (16 B x sum M) + E32 (36 B x H) per pair = %.1f MB" %
                        ((16.0 * sum_m + 36.0 * N_HYP * P) / 1e6),
    }
    dominant = "score" if stage_ms["score"] >= stage_ms["knn2"] else "knn2"
    roofline_other = roofline if dominant == "score" else score_roofline
    roofline = score_roofline if dominant == "score" else roofline
    cores = len(os.sched_getaffinity(0))
    if args.no_cpu_baseline or world > 1:
        cpu = None
    else:
        n_cpu = max(2 * cores, 16)
        cpu_rate = cpu_reference_rate(n_cpu, cores)
        cpu = {"value": cpu_rate, "unit": "pairs/s", "cores": cores, "kind": "port",
               "sample": "%d pairs of the same C2 workload on a %d-process pool (cv2 BFMatcher + restated skimage "
                         "ransac + reference triangulate loop)" % (n_cpu, cores)}
    line = {
        "metric": "frame pairs/s (2k kpts, match+RANSAC+triangulate)", "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": ("e4m3 (+-1, exact) tensor-core Hamming" if tensor_path else "u8 popcount") + " + f32/f64",
        "data": "synthetic",
        "config": {"workload": "C2: frame pairs 2048x2048 256-bit descriptors, 1024 seeded RANSAC hypotheses, "
                               "triangulate inliers", "pairs_per_step_per_gpu": P, "kpts": N_KPTS, "hypotheses": N_HYP,
                   "l2_policy": "inputs larger than L2 (%.0f MB of descriptors+keypoints per step)" %
                                ((des_h.nbytes + kps_h.nbytes) / 1e6),
                   "matcher": {"variant": "auto -> 17 (tensor cores, CTA pairs) for this batch" if args.knn_variant is None
                               else args.knn_variant, "r": args.knn_r, "ctas_per_sm": args.knn_ctas}},
        "hamming_gcmp_s": gcmp_s * world, "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "mode": "tvg_pairs_host_submit/_wait, %d contexts in rotation, host chunk %d pairs" % (NC, args.host_chunk)},
        "e2e_c4_shared_frames": {"value": world * n_seq_e2e * ppseq * e2e_steps / e2e4_s, "unit": "pairs/s",
                                 "h2d_bytes_per_step": int(h2d4), "d2h_bytes_per_step": int(d2h4), "steps": e2e_steps,
                                 "sequences_per_step": n_seq_e2e, "pairs_per_sequence": ppseq,
                                 "host_chunk_pairs": args.host_chunk_c4,
                                 "note": "same entry and outputs; consecutive pairs of a sequence share a frame, so a "
                                         "pair costs one frame (98 KB) of host->device traffic instead of two"},
        "h2d_gbs_per_rank_all_ranks_copying": h2d_min,
        "sustained": None if sus_steps == 0 else {
            "value": world * P * sus_steps / (sus_ms / 1e3), "unit": "pairs/s", "steps": sus_steps,
            "seconds": sus_ms / 1e3, "ms_per_step": sus_ms / sus_steps,
            "note": "the same device-resident step repeated back to back for seconds (max over ranks); the clock "
                    "sampler covers this region too"},
        "c4": None if c4_info is None else dict(c4_info, ms_per_sweep=c4_ms, n_gpus=world,
                                                value=C4_SEQS * C4_PAIRS / (c4_ms / 1e3), unit="pairs/s"),
        "c5": None if c5_info is None else dict(c5_info, ms_per_pair=c5_ms, n_gpus=world, stage_ms_rank0=c5_stage,
                                                equals_single_rank=c5_same,
                                                hamming_gcmp_s=C5_N * C5_N / (c5_ms / 1e3) / 1e9),
        "roofline": roofline, "roofline_second_kernel": roofline_other, "kernels": kernels, "cpu_baseline": cpu,
        "single_pair": single,
        "frontend_outside_metric": frontend,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# dram__bytes_read.sum + dram__bytes_write.sum of knn2_tc_kernel + unpack_train_kernel at P = 1024 (ncu --set full)
TC_TRAFFIC_BYTES_P1024 = 671.3e6 + 571.7e6        # default mode, ncu --cache-control none (profiles/r1_tensor_matcher_dram_traffic.md)
TC_TRAFFIC_L2MODE_BYTES_P1024 = 133.9e6 + 39.4e6   # L2-resident mode, 40 MiB budget
TC2_TRAFFIC_BYTES_P1024 = 134.35e6 + 23.09e6       # CTA-pair kernel (variant 17), profiles/r2_kernels_matcher_tc2.md


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1024, help="frame pairs per step per GPU")
    ap.add_argument("--knn-r", type=int, default=None, help="matcher queries/thread (default: library auto)")
    ap.add_argument("--knn-variant", type=int, default=None,
                    help="matcher variant (default: library auto; 16 = tensor cores, 0..11 = integer-pipe variants)")
    ap.add_argument("--knn-ctas", type=int, default=6)
    ap.add_argument("--host-chunk", type=int, default=1024,
                    help="pairs per chunk of the e2e host pipeline, independent pairs (measured best of 128..4096)")
    ap.add_argument("--host-chunk-c4", type=int, default=4096,
                    help="pairs per chunk of the e2e host pipeline, C4-shaped input (shared frames)")
    ap.add_argument("--e2e-contexts", type=int, default=3, help="batches in flight in the e2e measurement")
    ap.add_argument("--sustain-seconds", type=float, default=3.0,
                    help="also hold the step for this long and report it as `sustained` (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single-pair", action="store_true")
    ap.add_argument("--no-frontend", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 extra (64 sequences x 500 pairs, sharded)")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 extra (one 64k x 64k pair over the ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
