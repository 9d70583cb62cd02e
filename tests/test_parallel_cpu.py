"""world_size-2 gloo tests (CPU) of the host side of the sharded path: shard planning, slicing, the
id-broadcast plumbing and the re-assembly of shard tables.  The kernels themselves cannot run here; the
per-shard compute is played by the oracle so that the protocol (what is exchanged, in which order, and that
shard-wise results re-assemble to the unsharded answer) is what gets checked."""
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


class _FakeEngine:
    """Stands in for engine.Engine's communicator calls (no GPU on this box)."""

    def __init__(self):
        self.inited = None

    def comm_unique_id(self):
        return bytes(range(128))

    def comm_init(self, world, rank, uid):
        self.inited = (world, rank, uid)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chicdiff_b200 import parallel, synth
    from oracle import oracle as O
    d = synth.generate("tiny")
    bounds = parallel.shard_slices(d.region_bait, d.row_off, world)
    off, (N, FMr), (lo, hi) = parallel.take_shard(d.row_off, [d.N_rows, d.FM_rows], bounds, rank)
    assert off[0] == 0 and off[-1] == N.shape[1] == FMr.shape[1]
    # id broadcast plumbing
    eng = _FakeEngine()
    w, r = parallel.init_comm(eng, dist)
    assert (w, r) == (world, rank) and eng.inited == (world, rank, bytes(range(128)))
    # stage 1 needs no exchange: shard-wise aggregation re-assembles to the full matrices
    K, FM = O.aggregate(off, N, FMr)
    full = parallel.gather_columns({"K": K, "FM": FM, "n_local": hi - lo}, dist)
    # global step 1: size factors need every region -> all-gather of the counts (what cd_region_test does)
    parts = [None] * world
    dist.all_gather_object(parts, K)
    K_all = np.concatenate(parts, axis=1)
    sf = O.size_factors(K_all)
    # global step 2: gene-wise estimates are local, the trend is fitted on the gathered (baseMean, dispGeneEst)
    FM_parts = [None] * world
    dist.all_gather_object(FM_parts, FM)
    if rank == 0:
        Kf, FMf = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
        assert np.array_equal(full["K"], Kf)
        assert np.array_equal(np.isnan(full["FM"]), np.isnan(FMf))
        assert np.allclose(np.nan_to_num(full["FM"]), np.nan_to_num(FMf), rtol=0, atol=0)
        assert np.array_equal(K_all, Kf)
        assert np.allclose(sf, O.size_factors(Kf), rtol=0, atol=0)
        np.save(os.path.join(tmp, "ok.npy"), np.array([1]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_protocol(tmp_path, built):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok.npy"))


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
