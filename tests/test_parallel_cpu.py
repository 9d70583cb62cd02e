"""world_size-2 gloo tests (CPU) of the sharded path: shard planning, slicing, the id-broadcast plumbing, the
re-assembly of shard tables, and the PROTOCOL of the global steps as the kernels run it -- exact medians by all-reducing
2048-bin histogram counters round by round, the parametric trend by all-reducing 8 sums per pass with every rank taking
the same branch of glm.fit's control flow (tests/protocol_model.py restates both with the all-reduce passed in; here it
is torch.distributed over gloo), the small-df prior-variance rule by all-reducing the 40 bin counts of its histogram.  Nothing is gathered: a rank only ever sees its own regions plus those sums, and the
results must equal the unsharded oracle's."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _CommStub:
    """Stands in for engine.Engine's two communicator calls (cd_comm_unique_id / cd_comm_init need a GPU): the test is
    about what parallel.init_comm ships between the ranks."""

    def __init__(self):
        self.inited = None

    def comm_unique_id(self):
        return bytes(range(128))

    def comm_init(self, world, rank, uid):
        self.inited = (world, rank, uid)


def _allreduce(op):
    def f(a):
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint64:                     # gloo has no unsigned 64-bit reductions: order-preserving shift
            t = torch.from_numpy((a ^ np.uint64(1 << 63)).view(np.int64).copy())
            dist.all_reduce(t, op=op)
            return t.numpy().view(np.uint64) ^ np.uint64(1 << 63)
        t = torch.from_numpy(a.copy())
        dist.all_reduce(t, op=op)
        return t.numpy()
    return f


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import protocol_model as pm
    from chicdiff_b200 import parallel, synth
    from oracle import oracle as O
    ar_sum, ar_min = _allreduce(dist.ReduceOp.SUM), _allreduce(dist.ReduceOp.MIN)
    d = synth.generate("tiny")
    bounds = parallel.shard_slices(d.region_bait, d.row_off, world)
    off, (N, FMr), (lo, hi) = parallel.take_shard(d.row_off, [d.N_rows, d.FM_rows], bounds, rank)
    assert off[0] == 0 and off[-1] == N.shape[1] == FMr.shape[1]
    # id broadcast plumbing
    eng = _CommStub()
    w, r = parallel.init_comm(eng, dist)
    assert (w, r) == (world, rank) and eng.inited == (world, rank, bytes(range(128)))
    # stage 1 needs no exchange: shard-wise aggregation re-assembles to the full matrices
    K, FM = O.aggregate(off, N, FMr)
    full = parallel.gather_columns({"K": K, "FM": FM, "n_local": hi - lo}, dist)

    # the unsharded truth (every rank computes it for itself; the protocol below never looks at it except to compare)
    Kf, FMf = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
    ro = O.region_test(Kf, FMf, d.X)

    # global step 1 (estimateSizeFactors): medians of the local log-ratio columns, only histogram counters cross ranks
    with np.errstate(divide="ignore"):
        logK = np.log(K.astype(np.float64))
    lr = np.where((K > 0).all(axis=0), logK - logK.mean(axis=0), np.inf)
    sf = np.array([np.exp(pm.distributed_median(lr[s], ar_sum, ar_min)) for s in range(d.S)])
    # (the selection is exact; the log-ratios themselves are formed with NumPy's mean here and a sequential sum in the oracle)
    assert np.max(np.abs(sf - ro["sizeFactors"]) / ro["sizeFactors"]) < 1e-14, (sf, ro["sizeFactors"])
    one_rank = np.array([np.exp(np.median(x[np.isfinite(x)])) for x in np.where((Kf > 0).all(axis=0), np.log(np.maximum(Kf, 1).astype(np.float64)) - np.log(np.maximum(Kf, 1).astype(np.float64)).mean(axis=0), np.inf)])
    assert np.array_equal(sf, one_rank), "distributed selection must equal the median of the union exactly"

    # global step 2 (parametricDispersionFit): this rank's slice of (baseMean, dispGeneEst), 8 sums per pass cross ranks
    bm, ge = ro["baseMean"][lo:hi], ro["dispGeneEst"][lo:hi]
    a0, a1, status, outer, passes = pm.distributed_trend_fit(bm, ge, ar_sum)
    assert status == 0
    assert abs(a0 - ro["trend_a0"]) <= 1e-12 * ro["trend_a0"] and abs(a1 - ro["trend_a1"]) <= 1e-12 * ro["trend_a1"]
    # every rank followed the same control flow: same number of passes everywhere
    pp = ar_sum(np.array([passes, -passes * (rank == 0) * world], np.int64))
    assert pp[0] == passes * world

    # global step 3 (MAD of the log residuals): two more distributed medians
    with np.errstate(invalid="ignore", divide="ignore"):
        resid = np.where(ge >= 1e-6, np.log(ge) - np.log(a0 + a1 / bm), np.inf)
    med = pm.distributed_median(resid, ar_sum, ar_min)
    mad = pm.distributed_median(np.abs(resid - med), ar_sum, ar_min, scale=1.4826)
    assert abs(mad * mad - ro["varLogDispEsts"]) <= 1e-12 * ro["varLogDispEsts"]

    # global step 4, designs with S - p <= 3 only (here exercised on this design's residuals as if it were one): DESeq2's
    # Monte-Carlo prior-variance rule needs the 40-bin histogram of the residuals, and bin counts add up -- what
    # cd_region_test all-reduces in a sharded 2-vs-2 run (csrc/priorvar.cpp)
    import ctypes as C
    from chicdiff_b200 import engine
    Lp = engine.load_library()
    Lp.cd_prior_var_hist.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
    Lp.cd_prior_var_from_hist.restype = C.c_double
    Lp.cd_prior_var_from_hist.argtypes = [C.c_int, C.c_void_p]
    mine = np.ascontiguousarray(resid[np.isfinite(resid)])
    counts = np.zeros(40)
    Lp.cd_prior_var_hist(len(mine), mine.ctypes.data, counts.ctypes.data)
    total = np.ascontiguousarray(ar_sum(counts))
    with np.errstate(invalid="ignore", divide="ignore"):
        all_resid = np.where(ro["dispGeneEst"] >= 1e-6, np.log(ro["dispGeneEst"]) - np.log(ro["trend_a0"] + ro["trend_a1"] / ro["baseMean"]), np.inf)
    all_resid = np.ascontiguousarray(all_resid[np.isfinite(all_resid)])
    whole = np.zeros(40)
    Lp.cd_prior_var_hist(len(all_resid), all_resid.ctypes.data, whole.ctypes.data)
    assert np.array_equal(total, whole) and total.sum() > 0
    Lp.cd_prior_var_small_df.restype = C.c_double
    Lp.cd_prior_var_small_df.argtypes = [C.c_int, C.c_int64, C.c_void_p]
    assert Lp.cd_prior_var_from_hist(2, total.ctypes.data) == Lp.cd_prior_var_small_df(2, len(all_resid), all_resid.ctypes.data)

    if rank == 0:
        assert np.array_equal(full["K"], Kf)
        assert np.array_equal(np.isnan(full["FM"]), np.isnan(FMf))
        assert np.allclose(np.nan_to_num(full["FM"]), np.nan_to_num(FMf), rtol=0, atol=0)
        np.save(os.path.join(tmp, "ok.npy"), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_protocol(tmp_path, built):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok.npy"))


def test_protocol_model_alone_matches_numpy_and_the_oracle(built):
    """one rank (identity all-reduce): the radix selection is R's median for odd / even / empty / tied inputs, and the
    trend state machine reproduces the oracle's parametric fit"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import protocol_model as pm
    from chicdiff_b200 import synth
    from oracle import oracle as O
    ident = lambda a: a
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 5, 1000, 1001):
        v = rng.normal(size=n)
        v[rng.random(n) < 0.1] = np.inf
        fin = v[np.isfinite(v)]
        got = pm.distributed_median(v, ident, ident)
        assert (np.isnan(got) and len(fin) == 0) or got == np.median(fin)
    ties = np.repeat([-1.5, 0.0, 0.0, 2.0, 2.0, 2.0], 7)
    assert pm.distributed_median(ties, ident, ident) == np.median(ties)
    assert pm.distributed_median(-np.abs(rng.normal(size=64)), ident, ident) < 0
    for name in ("tiny", "c1"):
        d = synth.generate(name)
        K, FM = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
        ro = O.region_test(K, FM, d.X, theta=0.5, prior_var=0.5)
        a0, a1, st, _, _ = pm.distributed_trend_fit(ro["baseMean"], ro["dispGeneEst"], ident)
        assert st == 0 and abs(a0 - ro["trend_a0"]) <= 1e-12 * a0 and abs(a1 - ro["trend_a1"]) <= 1e-12 * a1


def test_take_shard_covers_everything_once(built):
    from chicdiff_b200 import parallel, synth
    d = synth.generate("c1")
    for world in (2, 3, 8):
        bounds = parallel.shard_slices(d.region_bait, d.row_off, world)
        seen_regions, seen_rows = 0, 0
        for r in range(world):
            off, (N,), (lo, hi) = parallel.take_shard(d.row_off, [d.N_rows], bounds, r)
            assert len(off) == hi - lo + 1 and off[-1] == N.shape[1]
            assert np.array_equal(np.diff(off), np.diff(d.row_off[lo:hi + 1]))
            seen_regions += hi - lo
            seen_rows += N.shape[1]
            if lo < hi and lo > 0:
                assert d.region_bait[lo] != d.region_bait[lo - 1]
        assert seen_regions == d.n and seen_rows == d.R
