"""Multi-GPU partitioning of the two-view path (one process per GPU, torch.distributed).

Two levels shard naturally:

* independent sequences / frame pairs (config C4): sequence s -> rank s mod G, NO data-path collective
  (``shard_sequences``); results are gathered on the host if wanted.
* one huge keyframe pair (config C5): query rows are split into G contiguous blocks for the matcher
  (train set replicated), hypotheses are split into G contiguous blocks for RANSAC.  The only exchange
  steps are (1) an all-gather of the per-rank match slices -- concatenation in rank order preserves the
  reference's ascending-query order -- and (2) ONE all-reduce(MAX) of a 63-bit packed key
  ``count | ~sse | ~hypothesis`` that implements the reference's selection rule (most inliers, then
  smallest sum of squared residuals, then earliest trial).  The rank that owns
  the winning hypothesis then broadcasts its inlier mask, refit E and Rt (one small message).

The compute calls go through a ``backend`` (default: ``openwork_b200.ops``, i.e. the CUDA kernels).
Tests inject a CPU stand-in to exercise this host logic over gloo without a GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

COUNT_BITS, SSE_BITS, HYP_BITS = 17, 30, 16      # 63 bits: fits a signed int64 for NCCL MAX


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n items for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sequences(nseq, rank, world):
    """Config C4: sequence s is processed by rank s mod world; no communication."""
    return list(range(rank, nseq, world))


def pack_best_key(count, sse, hyp):
    """int64 key whose MAX is the reference's best hypothesis: most inliers, then smallest sse
    (compared at fp32 resolution minus one mantissa bit), then smallest hypothesis index."""
    count = torch.as_tensor(count).to(torch.int64)
    hyp = torch.as_tensor(hyp).to(torch.int64)
    sse32 = torch.as_tensor(sse).to(torch.float32).clamp_min(0)
    sse32 = torch.nan_to_num(sse32, nan=float("inf"))
    bits = sse32.view(torch.int32).to(torch.int64) >> 1              # monotone for non-negative floats
    inv_sse = ((1 << SSE_BITS) - 1) - bits.clamp(0, (1 << SSE_BITS) - 1)
    inv_hyp = ((1 << HYP_BITS) - 1) - hyp
    return (count << (SSE_BITS + HYP_BITS)) | (inv_sse << HYP_BITS) | inv_hyp


def unpack_best_key(key):
    key = int(key)
    count = key >> (SSE_BITS + HYP_BITS)
    hyp = ((1 << HYP_BITS) - 1) - (key & ((1 << HYP_BITS) - 1))
    return count, hyp


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_varlen(x, group=None):
    """Concatenate per-rank tensors of different first-dim length in RANK ORDER (one size exchange,
    one padded all_gather)."""
    rank, world = _world(group)
    if world == 1:
        return x
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def allreduce_best(key, group=None):
    """The single all-reduce(MAX) that picks the global best hypothesis."""
    rank, world = _world(group)
    if world > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)
    return key


def huge_pair(curr_kps, curr_des, last_kps, last_des, K, words, ratio=0.75, residual_threshold=0.005,
              backend=None, group=None):
    """Config C5: one keyframe pair split over the ranks of `group`.

    Inputs are replicated on every rank (descriptor sets are <= 2 MiB).  Returns the same dict on
    every rank: qidx, tidx (all matches, ascending query order), best (global hypothesis index),
    count, E (refit), mask, Rt, points4d (all inliers, match order), good.
    """
    if backend is None:
        from . import ops as backend
    rank, world = _world(group)
    nq = curr_des.shape[0]
    H = words.shape[0]
    if H > (1 << HYP_BITS):
        raise ValueError("at most 65536 hypotheses fit the packed key")

    # 1. matcher: this rank's block of query rows against the whole train set
    q_lo, q_hi = shard_range(nq, rank, world)
    idx, dist_, keep = backend.knn2_hamming256(curr_des[q_lo:q_hi], last_des, ratio=ratio, return_keep=True)
    loc = backend.ratio_compact(idx, keep, curr_kps[q_lo:q_hi], last_kps, K)
    # 2. exchange step 1: all-gather the match slices (rank order == ascending query order), packed into
    #    one buffer per rank: [qidx+q_lo, tidx] as two f64-exact columns next to src / dst  -> f64[M_loc, 6]
    packed = torch.cat([(loc["qidx"] + q_lo).to(torch.float64)[:, None], loc["tidx"].to(torch.float64)[:, None],
                        loc["src"], loc["dst"]], dim=1)
    packed = all_gather_varlen(packed, group)
    qidx = packed[:, 0].to(torch.int32)
    tidx = packed[:, 1].to(torch.int32)
    src = packed[:, 2:4].contiguous()
    dst = packed[:, 4:6].contiguous()
    M = src.shape[0]
    if M <= 8:
        raise ValueError("`min_samples` must be in range (0, <number-of-samples>)")

    # 3. RANSAC over this rank's block of hypotheses (rows of the same global word table everywhere;
    #    the word -> index mapping depends only on the row and on M, so slices map identically)
    words = np.ascontiguousarray(np.asarray(words, dtype=np.uint32))
    h_lo, h_hi = shard_range(H, rank, world)
    key = torch.zeros(1, dtype=torch.int64, device=src.device)
    if h_hi > h_lo:
        r = backend.ransac_essential(src, dst, words=words[h_lo:h_hi], residual_threshold=residual_threshold)
        if r["best"] >= 0 and r["n_inl"] >= 0:
            key = pack_best_key(torch.tensor([r["n_inl"]]), torch.tensor([r["best_sse"]]),
                                torch.tensor([r["best"] + h_lo])).to(src.device)
    # 4. exchange step 2: ONE all-reduce(MAX) of the packed key
    key = allreduce_best(key, group)
    count, best = unpack_best_key(key.item())

    # 5. the rank that owns the winning hypothesis already holds its mask / refit E / Rt (its local best IS
    #    the global best); it broadcasts them in one message.  (world == 1: nothing to do.)
    owner = next(rk for rk in range(world) if shard_range(H, rk, world)[0] <= best < shard_range(H, rk, world)[1])
    if world > 1:
        msg = torch.empty(M + 8 * (9 + 9 + 16), dtype=torch.uint8, device=src.device)
        if rank == owner:
            msg[:M] = r["mask"].to(torch.uint8)
            msg[M:] = torch.cat([r["E"].reshape(-1), r["E_best"].reshape(-1), r["Rt"].reshape(-1)]).view(torch.uint8)
        dist.broadcast(msg, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
        mask = msg[:M].bool()
        tail = msg[M:].clone().view(torch.float64)
        w = dict(E=tail[:9].view(3, 3), E_best=tail[9:18].view(3, 3), Rt=tail[18:34].view(4, 4))
    else:
        w, mask = r, r["mask"]
    # 6. triangulation of the inliers, split by points, gathered back in match order
    def _t(x):
        return (x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(src.device)
    ck = _t(curr_kps)[qidx.long()][mask]
    lk = _t(last_kps)[tidx.long()][mask]
    p_lo, p_hi = shard_range(ck.shape[0], rank, world)
    Rt = w["Rt"]
    pose0 = np.eye(4)
    pose1 = np.asarray(Rt.cpu()) @ pose0
    X = backend.triangulate(lk[p_lo:p_hi].contiguous(), ck[p_lo:p_hi].contiguous(), pose0, pose1, K)
    X = all_gather_varlen(X.to(src.device), group)
    return dict(qidx=qidx, tidx=tidx, M=M, best=best, count=count, E=w["E"], E_best=w["E_best"], mask=mask,
                Rt=Rt, points4d=X, good=X[:, 2] > 0)
