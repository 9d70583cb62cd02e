"""Launched by torchrun (one rank per GPU): the sharded run must reproduce the single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/multi_gpu_check.py [workload] [regions]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chicdiff_b200 import engine, parallel, synth  # noqa: E402


def run(eng, X, row_off, N, FMr, **kw):
    eng.set_design(X)
    eng.set_regions(row_off)
    for s in range(X.shape[0]):
        eng.set_sample_rows(s, N[s], FMr[s])
    K, FM = eng.aggregate()
    r = eng.region_test(**kw)
    r["K"], r["FMagg"] = K, FM
    return r


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
    nreg = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = synth.generate(workload, n_regions=nreg)
    bounds = parallel.shard_slices(d.region_bait, d.row_off, world)
    off, (N, FMr), (lo, hi) = parallel.take_shard(d.row_off, [d.N_rows, d.FM_rows], bounds, rank)
    eng = engine.Engine(local)
    parallel.init_comm(eng, dist)
    r = run(eng, d.X, off, N, FMr)
    cols = {k: v for k, v in r.items() if isinstance(v, np.ndarray) and v.ndim >= 1 and k not in ("sizeFactors", "deviances")}
    for k in ("sizeFactors", "deviances", "theta", "trend_a0", "trend_a1", "varLogDispEsts", "dispPriorVar"):
        cols[k] = r[k]
    full = parallel.gather_columns(cols, dist)
    ok = True
    if rank == 0:
        ref = run(engine.Engine(local), d.X, d.row_off, d.N_rows, d.FM_rows)
        assert np.array_equal(full["K"], ref["K"]), "aggregated counts differ"
        print("shards", bounds.tolist(), "theta", full["theta"], ref["theta"])
        assert full["theta"] == ref["theta"]
        worst = 0.0
        for k in ("sizeFactors", "deviances"):
            e = float(np.max(np.abs(full[k] - ref[k]) / np.abs(ref[k])))
            print("%-14s max rel %.3e" % (k, e)); worst = max(worst, e if k == "sizeFactors" else 0)
        for k in ("trend_a0", "trend_a1", "varLogDispEsts", "dispPriorVar"):
            e = abs(full[k] - ref[k]) / abs(ref[k])
            print("%-14s rel %.3e" % (k, e)); worst = max(worst, e)
        for k in ("baseMean", "normFactors", "dispGeneEst", "dispFit", "dispMAP", "dispersion", "log2FoldChange", "lfcSE", "stat", "pvalue", "deviance", "maxCooks"):
            a, b = full[k], ref[k]
            assert np.array_equal(np.isnan(a), np.isnan(b)), k
            with np.errstate(invalid="ignore", divide="ignore"):
                e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
            e[np.isnan(e) | (a == b)] = 0
            print("%-14s max rel %.3e  #>1e-9: %d" % (k, e.max(), (e > 1e-9).sum()))
            if k not in ("pvalue", "stat", "log2FoldChange", "dispGeneEst"):
                worst = max(worst, float(np.quantile(e, 0.999)))
        print("iters equal:", np.array_equal(full["dispIter"], ref["dispIter"]), np.array_equal(full["betaIter"], ref["betaIter"]))
        ok = worst < 1e-6
        print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", "world", world, "worst %.3e" % worst)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    try:
        main()
    except SystemExit:
        raise
    except BaseException:
        import traceback
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(1)          # do not leave the other ranks waiting in a collective
