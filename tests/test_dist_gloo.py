"""CPU tests of the multi-GPU host logic (openwork_b200/dist.py) over gloo, world_size 2.

The product path is one C call (tvg_huge_pair over NCCL); dist.huge_pair_protocol is the same protocol over
torch.distributed with the compute calls injected through a backend.  Here a CPU stand-in built on the oracle
plays the kernels' role so the partitioning, the two fixed-size all-gathers and the exact record selection can be
exercised without a GPU.  The sharded result must equal the single-rank result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openwork_b200 import dist as tdist
from openwork_b200 import synth
from oracle import two_view as tv

K = synth.K_640x480


class OracleBackend:
    """Same call surface as openwork_b200.ops for the four ops dist.huge_pair uses (CPU tensors)."""

    @staticmethod
    def knn2_hamming256(q, t, ratio=0.75, return_keep=False):
        idx, d = tv.knn2_numpy(np.asarray(q), np.asarray(t))
        keep = (d[:, 0].astype(np.float64) < ratio * d[:, 1].astype(np.float64)) & (idx[:, 1] >= 0)
        return torch.from_numpy(idx), torch.from_numpy(d), torch.from_numpy(keep)

    @staticmethod
    def ratio_compact(idx, keep, curr_kps, last_kps, Kmat):
        idx, keep = idx.numpy(), keep.numpy()
        q = np.flatnonzero(keep).astype(np.int32)
        t = idx[keep, 0].astype(np.int32)
        src = tv.normalize(Kmat, np.asarray(last_kps)[t]) if len(q) else np.zeros((0, 2))
        dst = tv.normalize(Kmat, np.asarray(curr_kps)[q]) if len(q) else np.zeros((0, 2))
        return dict(qidx=torch.from_numpy(q), tidx=torch.from_numpy(t), M=len(q), src=torch.from_numpy(src),
                    dst=torch.from_numpy(dst))

    @staticmethod
    def ransac_essential(src, dst, samples=None, words=None, residual_threshold=0.005):
        if samples is None:
            samples = synth.samples_from_words(words, len(src))
        r = tv.ransac_essential(src.numpy(), dst.numpy(), samples, residual_threshold)
        n = int(r["mask"].sum()) if r["mask"] is not None else -1
        return dict(E=torch.from_numpy(r["E"]), mask=torch.from_numpy(r["mask"]), best=r["best"], n_inl=n,
                    best_sse=float(r["sse"][r["best"]]), E_best=torch.from_numpy(r["E_best"]),
                    Rt=torch.from_numpy(tv.extract_Rt(r["E"])))

    @staticmethod
    def refit_essential(src, dst, E_best, thr):
        src, dst, E_best = np.asarray(src), np.asarray(dst), np.asarray(E_best)
        mask = np.abs(tv.essential_residuals(E_best, src, dst)) < thr
        n_in = int(mask.sum())
        E = tv.essential_estimate(src[mask], dst[mask], full_matrices=(n_in < 9)) if n_in >= 8 else E_best
        return dict(E=torch.from_numpy(E), mask=torch.from_numpy(mask), Rt=torch.from_numpy(tv.extract_Rt(E)))

    @staticmethod
    def triangulate(p1, p2, pose1, pose2, Kmat):
        if len(p1) == 0:
            return torch.zeros((0, 4), dtype=torch.float64)
        return torch.from_numpy(tv.triangulate_batched(np.asarray(p1), np.asarray(p2), pose1, pose2, Kmat))


def _inputs():
    lk, ld, ck, cd = synth.make_pair(500, 77, forced_ties=20)
    words = synth.sample_words(96, 5)
    return lk, ld, ck, cd, words


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lk, ld, ck, cd, words = _inputs()
        r = tdist.huge_pair_protocol(ck, cd, lk, ld, K, words, backend=OracleBackend)
        np.savez(os.path.join(out_dir, "r%d.npz" % rank), qidx=r["qidx"].numpy(), tidx=r["tidx"].numpy(),
                 best=r["best"], count=r["count"], E=r["E"].numpy(), mask=r["mask"].numpy(), Rt=r["Rt"].numpy(),
                 X=r["points4d"].numpy())
    finally:
        dist.destroy_process_group()


def test_huge_pair_two_ranks_equals_single_rank(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    lk, ld, ck, cd, words = _inputs()
    one = tdist.huge_pair_protocol(ck, cd, lk, ld, K, words, backend=OracleBackend)   # world 1, no process group
    ref = tv.process_pair(lk, ld, ck, cd, np.eye(4), K, words=words)
    for rank in range(2):
        g = np.load(os.path.join(str(tmp_path), "r%d.npz" % rank))
        assert np.array_equal(g["qidx"], one["qidx"].numpy()) and np.array_equal(g["tidx"], one["tidx"].numpy())
        assert np.array_equal(g["qidx"], ref["idx1"]) and np.array_equal(g["tidx"], ref["idx2"])
        assert int(g["best"]) == one["best"] == ref["ransac"]["best"]
        assert int(g["count"]) == one["count"] == int(ref["inliers"].sum())
        assert np.array_equal(g["mask"], ref["inliers"])
        assert np.array_equal(g["E"], one["E"].numpy())
        np.testing.assert_allclose(g["E"], ref["E"], atol=1e-12)
        np.testing.assert_allclose(g["X"], ref["points4d"], rtol=1e-9, atol=1e-9)


def test_record_selection_is_the_reference_rule_with_exact_keys():
    """select_best_record == sequential skimage rule applied to the per-rank winners, including sums
    of squared residuals that differ only in the last fp64 bit (the old fp32-resolution key could not tell them
    apart) and ranks without a model."""
    rng = np.random.default_rng(0)
    for trial in range(300):
        G = int(rng.integers(1, 9))
        counts = rng.integers(0, 5, G)
        sse = rng.uniform(1.0, 2.0, G)
        if G > 1 and trial % 3 == 0:                       # plant equal counts with sse one ulp apart
            i, j = rng.choice(G, 2, replace=False)
            counts[j] = counts[i]
            sse[j] = np.nextafter(sse[i], 0.0 if trial % 2 else 4.0)
        if G > 1 and trial % 5 == 0:                       # plant an exact (count, sse) tie: earliest trial wins
            i, j = rng.choice(G, 2, replace=False)
            counts[j], sse[j] = counts[i], sse[i]
        hyp = rng.permutation(65536)[:G]
        hyp[rng.random(G) < 0.15] = -1                     # ranks whose block held no model
        recs = [(int(c), int(h), float(s)) for c, h, s in zip(counts, hyp, sse)]
        got = tdist.select_best_record(recs)
        # the reference scans trials in hypothesis order with strict improvements
        best, bn, bs = -1, 0, np.inf
        for i in sorted(range(G), key=lambda i: hyp[i]):
            if hyp[i] < 0:
                continue
            if counts[i] > bn or (counts[i] == bn and sse[i] < bs):
                best, bn, bs = i, counts[i], sse[i]
        assert got == best, (recs, got, best)
    assert tdist.select_best_record([(3, -1, 1.0)]) == -1
    assert tdist.select_best_record([(3, 7, float("nan")), (3, 9, 5.0)]) == 1     # NaN sum counts as +inf


def test_shard_range_matches_the_c_abi():
    from openwork_b200 import ops
    for n, w in ((10, 3), (7, 8), (65536, 8), (65537, 8), (0, 4), (5, 1)):
        for r in range(w):
            assert ops.shard_range(n, r, w) == tdist.shard_range(n, r, w)
    assert ops.huge_pair_rows(65536, 65536, 8) == 65536
    assert ops.huge_pair_rows(10, 4, 4) == 12 and ops.huge_pair_rows(10, 40, 4) == 40
    for nq, w in ((10, 4), (65536, 8), (3, 8)):
        blocks = [tdist.query_block(nq, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and max(b[1] for b in blocks) == nq
        assert all(b[2] * w >= nq for b in blocks)


def test_shard_helpers():
    for n, w in ((10, 3), (7, 8), (65536, 8), (0, 4)):
        blocks = [tdist.shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in blocks) - min(b - a for a, b in blocks) <= 1
    assert sorted(sum((tdist.shard_sequences(64, r, 8) for r in range(8)), [])) == list(range(64))
    assert tdist.shard_sequences(64, 3, 8) == [3, 11, 19, 27, 35, 43, 51, 59]
