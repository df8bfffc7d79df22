"""Multi-GPU partitioning of the two-view path (one process per GPU, torch.distributed for the plumbing).

Two levels shard naturally:

* independent sequences / frame pairs (config C4): sequence s -> rank s mod G, NO data-path collective
  (``shard_sequences``); results are gathered on the host if wanted.
* one huge keyframe pair (config C5): query rows are split into G contiguous blocks for the matcher (train set
  replicated), hypotheses into G contiguous blocks for RANSAC.  The exchange steps are two FIXED-SIZE all-gathers:
  (1) the matcher outputs of the padded query blocks -- rank order is ascending query order, so every rank
  compacts the same match list -- and (2) one 96-byte record per rank ``{count, hypothesis, sum r^2 (fp64), E}``;
  every rank applies the reference's selection rule (most inliers, then smallest sum of squared residuals, then
  earliest trial; skimage ``ransac``) to the same G records with exact keys, so G ranks
  return bit for bit what one rank returns.  Mask, refit, pose and triangulation are then computed on every rank.

The product path is ``huge_pair``: one C call (``tvg_huge_pair``) over the NCCL communicator that ``init_comm``
binds to the context -- kernels and collectives on one stream, no host synchronisation.  ``huge_pair_protocol`` is
the same protocol written over ``torch.distributed`` with the compute calls injected through a ``backend``: the
CPU tests run it over gloo (world size 2) with an oracle stand-in to exercise the host logic without a GPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n items for `rank` (first n % world ranks get one extra); == tvg_shard_range."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sequences(nseq, rank, world):
    """Config C4: sequence s is processed by rank s mod world; no communication."""
    return list(range(rank, nseq, world))


def query_block(nq, rank, world):
    """Padded query block of the matcher: per = ceil(nq / world) rows per rank; returns (lo, hi, per)."""
    per = -(-nq // world)
    lo = min(nq, rank * per)
    return lo, min(nq, lo + per), per


def select_best_record(records):
    """The reference's selection rule over per-rank records ``(count, hypothesis, sse)`` (hypothesis < 0 = the rank
    has no model): most inliers, then smallest sum of squared residuals (NaN counts as +inf), then earliest trial.
    Returns the index of the winning record, or -1.  Mirrors ``ransac_merge_records_kernel``."""
    best, key = -1, None
    for i, (count, hyp, sse) in enumerate(records):
        if hyp < 0:
            continue
        s = float("inf") if sse != sse else float(sse)
        k = (-int(count), s, int(hyp))
        if key is None or k < key:
            best, key = i, k
    return best


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_fixed(x, group=None):
    """Concatenate equally-shaped per-rank tensors in RANK ORDER (one all_gather, no size exchange)."""
    rank, world = _world(group)
    if world == 1:
        return x
    bufs = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(bufs, x.contiguous(), group=group)
    return torch.cat(bufs, dim=0)


# ------------------------------------------------------------------------------------------ product path
def init_comm(ctx=None, group=None):
    """Bind an NCCL communicator spanning `group` (default: the world) to the context (``tvg_comm_init``): rank 0
    creates the 128-byte id, torch.distributed carries it to the other ranks, every rank joins.  Idempotent."""
    from . import _lib, ops
    ctx = ctx or ops.context()
    rank, world = _world(group)
    if ctx.comm_info()[1] == world and (world == 1 or getattr(ctx, "_comm_group", None) is group):
        return ctx
    box = [_lib.comm_unique_id() if rank == 0 else None]
    if world > 1:
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(box, src=src, group=group)
    ctx.comm_init(box[0], rank, world)
    ctx._comm_group = group
    return ctx


def huge_pair(curr_kps, curr_des, last_kps, last_des, K, words, ratio=0.75, residual_threshold=0.005, ctx=None,
              group=None, out=None, sync=True):
    """Config C5 on the GPUs of `group`: ``tvg_huge_pair`` over the context's communicator.  Inputs are replicated
    on every rank.  With ``sync`` the per-row outputs are trimmed to their valid lengths (one host synchronisation
    after the call) and the same dict is returned on every rank: qidx, tidx, M, best, count, best_sse, E, E_best,
    mask, Rt, points4d, good, status; without, the raw device tensors of ``ops.huge_pair``."""
    from . import ops
    ctx = init_comm(ctx, group)
    o = ops.huge_pair(curr_kps, curr_des, last_kps, last_des, K, words=words, ratio=ratio,
                      residual_threshold=residual_threshold, ctx=ctx, out=out)
    if not sync:
        return o
    M, n = int(o["m_count"].item()), int(o["n_inl"].item())
    if int(o["status"].item()) == 1:
        raise ValueError("`min_samples` must be in range (0, <number-of-samples>)")
    return dict(qidx=o["qidx"][:M], tidx=o["tidx"][:M], M=M, best=int(o["best"].item()), count=n,
                best_sse=float(o["best_sse"].item()), E=o["E"], E_best=o["E_best"], mask=o["mask"][:M].bool(),
                Rt=o["Rt"], points4d=o["points4d"][:n], good=o["good"][:n].bool(), status=int(o["status"].item()),
                raw=o)


# ------------------------------------------------------------------------------------------ protocol model
def huge_pair_protocol(curr_kps, curr_des, last_kps, last_des, K, words, backend, ratio=0.75, residual_threshold=0.005,
                       group=None):
    """The protocol of ``tvg_huge_pair`` over ``torch.distributed`` collectives with the compute calls injected:
    ``backend`` provides knn2_hamming256, ratio_compact, ransac_essential, refit_essential and triangulate (the
    CPU tests pass an oracle stand-in and run this over gloo).  Returns the same dict on every rank."""
    rank, world = _world(group)
    nq = curr_des.shape[0]
    H = words.shape[0]
    if H < world:
        raise ValueError("fewer hypotheses than ranks")
    # 1. matcher on the padded query block of this rank
    q_lo, q_hi, per = query_block(nq, rank, world)
    blk = torch.zeros((per, 3), dtype=torch.int32)           # idx0, idx1, keep per query row (fixed size)
    if q_hi > q_lo:
        idx, dist_, keep = backend.knn2_hamming256(curr_des[q_lo:q_hi], last_des, ratio=ratio, return_keep=True)
        blk[: q_hi - q_lo, 0] = torch.as_tensor(idx)[:, 0].to(torch.int32)
        blk[: q_hi - q_lo, 1] = torch.as_tensor(idx)[:, 1].to(torch.int32)
        blk[: q_hi - q_lo, 2] = torch.as_tensor(keep).to(torch.int32)
    # 2. exchange step 1: fixed-size all-gather, rank order == ascending query order; every rank compacts all rows
    full = all_gather_fixed(blk, group)[:nq]
    loc = backend.ratio_compact(full[:, :2].contiguous(), full[:, 2].bool(), curr_kps, last_kps, K)
    qidx, tidx, src, dst = loc["qidx"], loc["tidx"], loc["src"], loc["dst"]
    M = src.shape[0]
    if M <= 8:
        raise ValueError("`min_samples` must be in range (0, <number-of-samples>)")
    # 3. RANSAC over this rank's block of hypotheses (rows of the same global word table everywhere; the
    #    word -> index mapping depends only on the row and on M, so blocks map identically)
    words = np.ascontiguousarray(np.asarray(words, dtype=np.uint32))
    h_lo, h_hi = shard_range(H, rank, world)
    rec = torch.zeros((1, 12), dtype=torch.float64)
    rec[0, 1] = -1.0
    r = backend.ransac_essential(src, dst, words=words[h_lo:h_hi], residual_threshold=residual_threshold)
    if r["best"] >= 0 and r["n_inl"] >= 0:
        rec[0, 0], rec[0, 1], rec[0, 2] = float(r["n_inl"]), float(r["best"] + h_lo), float(r["best_sse"])
        rec[0, 3:] = torch.as_tensor(r["E_best"], dtype=torch.float64).reshape(-1)
    # 4. exchange step 2: ONE all-gather of the records; the same exact rule on every rank
    recs = all_gather_fixed(rec, group)
    win = select_best_record([(int(x[0]), int(x[1]), float(x[2])) for x in recs.tolist()])
    if win < 0:
        return dict(qidx=qidx, tidx=tidx, M=M, best=-1, count=0, status=2)
    best, E_best = int(recs[win, 1]), recs[win, 3:].reshape(3, 3).clone()
    # 5. inlier mask of the winner, refit, extract_Rt, triangulation: on every rank, identical
    w = backend.refit_essential(src, dst, E_best, residual_threshold)
    mask = torch.as_tensor(w["mask"]).bool()

    def _t(x):
        return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    ck = _t(curr_kps)[qidx.long()][mask]
    lk = _t(last_kps)[tidx.long()][mask]
    Rt = torch.as_tensor(w["Rt"])
    X = backend.triangulate(lk.contiguous(), ck.contiguous(), np.eye(4), np.asarray(Rt) @ np.eye(4), K)
    X = torch.as_tensor(X)
    return dict(qidx=qidx, tidx=tidx, M=M, best=best, count=int(mask.sum()), best_sse=float(recs[win, 2]),
                E=torch.as_tensor(w["E"]), E_best=E_best, mask=mask, Rt=Rt, points4d=X, good=X[:, 2] > 0, status=0)
