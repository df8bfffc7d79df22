"""GPU parity tests of tvg_huge_pair (BASELINE.json config C5: one huge keyframe pair, query rows and hypotheses
split over the ranks of an NCCL communicator bound to the context).

* one rank: against the batched pair pipeline, the step-by-step operators and the oracle;
* C5 size (65536 x 65536, 65536 hypotheses): matcher against real cv2, RANSAC against the oracle on a hypothesis
  subsample, refit and points against the oracle on the GPU's inlier set;
* G ranks (needs G GPUs; skipped otherwise): bit for bit what one rank returns, including a planted tie -- equal
  counts and sums of squared residuals that differ in the last bits -- between hypotheses owned by different ranks.
"""
import os
import socket

import numpy as np
import pytest
import torch

from openwork_b200 import synth
from oracle import two_view as tv

pytestmark = pytest.mark.gpu
K = synth.K_640x480
THR = 0.005
BAND = 1e-5
E_TOL = 1e-4


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    from openwork_b200 import ops as _ops
    return _ops


def _fro_diff(A, B):
    A, B = np.asarray(A, dtype=float).reshape(3, 3), np.asarray(B, dtype=float).reshape(3, 3)
    A, B = A / np.linalg.norm(A), B / np.linalg.norm(B)
    return min(np.linalg.norm(A - B), np.linalg.norm(A + B))


ALL = ("idx", "dist", "qidx", "tidx", "m_count", "E", "E_best", "best", "best_sse", "counts", "n_cand", "mask", "n_inl",
       "inl_idx", "Rt", "poses", "points4d", "good", "status")


def _np(o):
    return {k: v.cpu().numpy() for k, v in o.items()}


@pytest.mark.parametrize("n,H,seed", [(700, 128, 5), (2048, 1024, 6), (5000, 600, 7)])
def test_one_rank_equals_pipeline_and_oracle(ops, n, H, seed):
    lk, ld, ck, cd = synth.make_pair_large(n, seed) if n > 2048 else synth.make_pair(n, seed, forced_ties=10)
    words = synth.sample_words(H, seed)
    o = _np(ops.huge_pair(ck, cd, lk, ld, K, words=words, residual_threshold=THR, outputs=ALL))
    assert int(o["status"][0]) == 0
    ref = tv.process_pair(lk, ld, ck, cd, np.eye(4), K, words=words, residual_threshold=THR)
    M, n_in = len(ref["idx1"]), int(ref["inliers"].sum())
    ri, rd = tv.knn2_cv2(cd, ld)
    assert np.array_equal(o["idx"][:n], ri) and np.array_equal(o["dist"][:n], rd)
    assert int(o["m_count"][0]) == M
    assert np.array_equal(o["qidx"][:M], ref["idx1"]) and np.array_equal(o["tidx"][:M], ref["idx2"])
    assert int(o["best"][0]) == ref["ransac"]["best"] and int(o["n_inl"][0]) == n_in
    assert np.array_equal(o["mask"][:M].astype(bool), ref["inliers"])
    assert np.isclose(o["best_sse"][0], ref["ransac"]["sse"][ref["ransac"]["best"]], rtol=1e-9)
    assert _fro_diff(o["E"], ref["E"]) < E_TOL and _fro_diff(o["E_best"], ref["ransac"]["E_best"]) < E_TOL
    np.testing.assert_allclose(o["Rt"], ref["Rt"], atol=1e-6)
    np.testing.assert_allclose(o["poses"][1], ref["curr_pose"], atol=1e-6)
    X, Xr = o["points4d"][:n_in], ref["points4d"]
    rel = np.linalg.norm(X[:, :3] - Xr[:, :3], axis=1) / np.maximum(np.linalg.norm(Xr[:, :3], axis=1), 1e-12)
    assert np.median(rel) < 1e-4 and (rel < 1e-3).mean() > 0.99
    assert np.array_equal(o["good"][:n_in].astype(bool), X[:, 2] > 0)
    # the batched pair pipeline on the same pair: same matches, same winner, same mask; E to rounding (its refit
    # sums in a different order below 4096 rows)
    des = torch.from_numpy(np.stack([ld, cd])).cuda()
    kps = torch.from_numpy(np.stack([lk, ck])).cuda()
    b = _np(ops.process_pairs_device(des, kps, K, words=words, residual_threshold=THR,
                                     outputs=("qidx", "tidx", "m_count", "best", "mask", "E", "n_inl", "points4d")))
    assert int(b["m_count"][0]) == M and int(b["best"][0]) == int(o["best"][0])
    assert np.array_equal(b["mask"][0, :M], o["mask"][:M])
    np.testing.assert_allclose(b["E"][0], o["E"], rtol=0, atol=1e-11 * np.abs(o["E"]).max())
    np.testing.assert_allclose(b["points4d"][0, :n_in], X, rtol=1e-8, atol=1e-8)


def test_too_few_matches_reports_status_1(ops):
    rng = np.random.default_rng(0)
    cd = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    ld = rng.integers(0, 256, (64, 32), dtype=np.uint8)       # unrelated descriptors: (almost) nothing passes the ratio test
    ck, lk = rng.uniform(0, 600, (64, 2)), rng.uniform(0, 600, (64, 2))
    o = _np(ops.huge_pair(ck, cd, lk, ld, K, words=synth.sample_words(32, 1), outputs=ALL))
    assert int(o["m_count"][0]) <= 8 and int(o["status"][0]) == 1 and int(o["best"][0]) == -1 and int(o["n_inl"][0]) == 0
    assert np.isnan(o["E"]).all()


def test_stage_and_exchange_timers_of_a_profiled_call(ops):
    """tvg_set_profile + tvg_get_stage_ms / tvg_get_exchange_ms after tvg_huge_pair: nine stage times and the two
    exchange times (near zero on one rank: nothing is exchanged), an error before any profiled call."""
    from openwork_b200 import _lib
    lk, ld, ck, cd = synth.make_pair(1500, 9)
    ctx = _lib.Context(0)
    try:
        with pytest.raises(_lib.TvgError):
            ctx.exchange_ms()
        ctx.set_profile(True)
        o = ops.huge_pair(ck, cd, lk, ld, K, words=synth.sample_words(256, 9), ctx=ctx)
        st = ctx.stage_ms()
        x1, x2 = ctx.exchange_ms()
        ctx.set_profile(False)
        assert int(o["status"][0]) == 0
        assert set(st) == set(_lib.Context.STAGES) and all(np.isfinite(v) and v >= 0 for v in st.values())
        assert st["score"] > 0 and st["knn2"] > 0
        assert 0 <= x1 < 1.0 and 0 <= x2 < 1.0
    finally:
        ctx.close()


def _check_c5_against_oracle(o, lk, ld, ck, cd, words, n_sub=48, seed=0):
    """Oracle checks that stay cheap at C5 size (used by the one-rank test and the multi-rank worker)."""
    n, H = len(cd), len(words)
    M, n_in, best = int(o["m_count"][0]), int(o["n_inl"][0]), int(o["best"][0])
    # matcher + ratio loop: real cv2 over all 64k x 64k comparisons, reference loop for the compaction
    ri, rd = tv.knn2_cv2(cd, ld)
    assert np.array_equal(o["idx"][:n], ri) and np.array_equal(o["dist"][:n], rd)
    keep = rd[:, 0].astype(np.float64) < 0.75 * rd[:, 1].astype(np.float64)
    assert M == int(keep.sum())
    assert np.array_equal(o["qidx"][:M], np.flatnonzero(keep)) and np.array_equal(o["tidx"][:M], ri[keep, 0])
    src, dst = tv.normalize(K, lk[o["tidx"][:M]]), tv.normalize(K, ck[o["qidx"][:M]])
    # RANSAC on a hypothesis subsample (+ the winner): oracle fit from the same 8 samples, counts inside the band
    rng = np.random.default_rng(seed)
    sub = np.unique(np.concatenate([rng.choice(H, n_sub, replace=False), [best]]))
    samples = synth.samples_from_words(words[sub], M)
    best_key = None
    for h, s in zip(sub, samples):
        E = tv.essential_estimate(src[s], dst[s])
        if E is None:
            continue
        r = np.abs(tv.essential_residuals(E, src, dst))
        cnt = int((r < THR).sum())
        band = np.abs(r - THR) < BAND
        lo, hi = cnt - int((band & (r < THR)).sum()), cnt + int((band & (r >= THR)).sum())
        if o.get("counts") is not None and o["counts"][h] != 0:            # this rank's block (others stay zero)
            assert lo <= o["counts"][h] <= hi, (h, o["counts"][h], cnt)
        key = (-cnt, float(np.sum(r ** 2)), int(h))
        if best_key is None or key < best_key:
            best_key = key
        if h == best:
            assert cnt == n_in and np.isclose(o["best_sse"][0], np.sum(r ** 2), rtol=1e-9)
            assert _fro_diff(o["E_best"], E) < E_TOL
            differs = o["mask"][:M].astype(bool) != (r < THR)
            assert not (differs & ~band).any() and differs.sum() <= 2
    assert best_key[2] == best                       # nothing in the subsample beats the GPU's winner
    # refit, pose, points on the GPU's inlier set
    mask = o["mask"][:M].astype(bool)
    E_ref = tv.essential_estimate(src[mask], dst[mask], full_matrices=False)
    assert _fro_diff(o["E"], E_ref) < E_TOL
    Rt = tv.extract_Rt(E_ref)
    np.testing.assert_allclose(o["Rt"], Rt, atol=1e-6)
    pick = rng.choice(n_in, 2000, replace=False)
    lki, cki = lk[o["tidx"][:M]][mask][pick], ck[o["qidx"][:M]][mask][pick]
    Xr = tv.triangulate_batched(lki, cki, np.eye(4), Rt @ np.eye(4), K)
    X = o["points4d"][:n_in][pick]
    rel = np.linalg.norm(X[:, :3] - Xr[:, :3], axis=1) / np.maximum(np.linalg.norm(Xr[:, :3], axis=1), 1e-12)
    assert np.median(rel) < 1e-4 and (rel < 1e-3).mean() > 0.99


def test_config_c5_one_rank_against_oracle_subsample(ops):
    """C5 on one GPU: 65536 x 65536 descriptors, 65536 hypotheses."""
    N = H = 65536
    lk, ld, ck, cd = synth.make_pair_large(N, 5005)
    words = synth.sample_words(H, 5005)
    o = _np(ops.huge_pair(ck, cd, lk, ld, K, words=words, residual_threshold=THR, outputs=ALL))
    assert int(o["status"][0]) == 0 and int(o["m_count"][0]) > 40000 and int(o["n_inl"][0]) > 20000
    _check_c5_against_oracle(o, lk, ld, ck, cd, words)
    # step by step through the single-pair operators: same matches, same winner, same mask, same E
    idx, dist, keep = ops.knn2_hamming256(cd, ld, return_keep=True)
    loc = ops.ratio_compact(idx, keep, ck, lk, K)
    one = ops.ransac_essential(loc["src"], loc["dst"], words=words, residual_threshold=THR)
    M = loc["M"]
    assert M == int(o["m_count"][0]) and one["best"] == int(o["best"][0]) and one["n_inl"] == int(o["n_inl"][0])
    assert np.array_equal(one["mask"].cpu().numpy(), o["mask"][:M].astype(bool))
    assert np.array_equal(one["E"].cpu().numpy(), o["E"])


# ------------------------------------------------------------------------------------------ several ranks
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tie_table(words, world, best, M):
    """The explicit sample table of `words` with, in every rank's block, one more copy of the winning row `best` whose
    8 indices come in a different order: the same model up to rounding -> equal counts, sums of squared residuals
    equal to ~1e-16 -- plus an exact duplicate at the very end (the earliest of exact ties must win)."""
    from openwork_b200.dist import shard_range
    tab = synth.samples_from_words(words, M).astype(np.int32)
    H = len(tab)
    row = tab[best].copy()
    for r in range(world):
        lo, hi = shard_range(H, r, world)
        at = lo + (hi - lo) // 2 + r
        if at != best:
            tab[at] = np.roll(row, r + 1)
    if best != H - 1:
        tab[H - 1] = row
    return tab


def _worker(rank, world, port, out_dir, n, H):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from openwork_b200 import _lib, dist as tdist, ops
        lk, ld, ck, cd = synth.make_pair_large(n, 4242)
        words = synth.sample_words(H, 4242)
        ctx = tdist.init_comm(ops.context(rank))
        assert ctx.comm_info() == (rank, world)
        solo = _lib.Context(rank)                         # no communicator: the one-rank run of the same call
        res = {}
        for name, kw in (("words", dict(words=words)), ("ties", None)):
            if kw is None:
                # second case: every rank's block holds a re-ordered copy of the first case's winning sample
                kw = dict(samples=_tie_table(words, world, int(res["words_g"]["best"][0]), int(res["words_g"]["m_count"][0])))
            g = _np(ops.huge_pair(ck, cd, lk, ld, K, residual_threshold=THR, outputs=ALL, ctx=ctx, **kw))
            s = _np(ops.huge_pair(ck, cd, lk, ld, K, residual_threshold=THR, outputs=ALL, ctx=solo, **kw))
            res[name + "_g"], res[name + "_s"] = g, s
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, "r%d.npz" % rank),
                 **{"%s__%s" % (case, k): v for case, o in res.items() for k, v in o.items()})
        solo.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_ranks_return_bit_for_bit_what_one_rank_returns(ops, world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    n, H = 6000, 4096
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), n, H), nprocs=world, join=True)
    first = None
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % rank))
        for case in ("words", "ties"):
            g = {k.split("__")[1]: z[k] for k in z.files if k.startswith(case + "_g__")}
            s = {k.split("__")[1]: z[k] for k in z.files if k.startswith(case + "_s__")}
            M, n_in = int(s["m_count"][0]), int(s["n_inl"][0])
            assert int(s["status"][0]) == 0 and M > 1000 and n_in > 500
            for k in ("m_count", "best", "best_sse", "n_inl", "E", "E_best", "Rt", "poses", "status"):
                assert np.array_equal(g[k], s[k]), (rank, case, k)
            for k in ("idx", "dist"):
                assert np.array_equal(g[k][:n], s[k][:n]), (rank, case, k)
            for k in ("qidx", "tidx", "mask"):
                assert np.array_equal(g[k][:M], s[k][:M]), (rank, case, k)
            for k in ("inl_idx", "points4d", "good"):
                assert np.array_equal(g[k][:n_in], s[k][:n_in]), (rank, case, k)
            from openwork_b200.dist import shard_range
            lo, hi = shard_range(H, rank, world)
            assert np.array_equal(g["counts"][lo:hi], s["counts"][lo:hi])
            if case == "ties":
                # the planted copies tie on the count; the winner is decided by the exact sums / the trial index
                assert int(s["n_cand"][0]) >= world + 1
            if first is None:
                first = {c: None for c in ("words", "ties")}
            if first[case] is None:
                first[case] = g
            else:                                            # and every rank holds the same answer
                for k in ("best", "E", "Rt", "n_inl"):
                    assert np.array_equal(first[case][k], g[k])
