"""Launched by torchrun (one rank per GPU): a sharded 2-vs-2 run without a given dispersion prior variance.  Every fit has
S - p <= 3, so the library's restatement of DESeq2's Monte-Carlo rule supplies the prior variance; in a sharded run the
ranks all-reduce the 40-bin histogram of their residuals and evaluate the rule on the sum.  Must give the single-GPU
run's value.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/multi_gpu_small_df.py
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
    eng.aggregate(fetch=False)
    return eng.region_test(**kw)


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = synth.generate("c1")
    bounds = parallel.shard_slices(d.region_bait, d.row_off, world)
    off, (N, FMr), (lo, hi) = parallel.take_shard(d.row_off, [d.N_rows, d.FM_rows], bounds, rank)
    eng = engine.Engine(local)
    parallel.init_comm(eng, dist)
    ok = True
    for kw in (dict(theta=0.5), dict()):                                  # one fit (df = 2); theta grid (5 x df = 3, then df = 2)
        r = run(eng, d.X, off, N, FMr, **kw)
        cols = parallel.gather_columns({"pvalue": r["pvalue"], "dispPriorVar": r["dispPriorVar"], "theta": r["theta"]}, dist)
        if rank == 0:
            ref = run(engine.Engine(local), d.X, d.row_off, d.N_rows, d.FM_rows, **kw)
            with np.errstate(invalid="ignore"):
                rel = np.abs(cols["pvalue"] - ref["pvalue"]) / np.maximum(ref["pvalue"], 1e-300)
            same = (cols["dispPriorVar"] == ref["dispPriorVar"]) and (cols["theta"] == ref["theta"])
            print("options", kw, "dispPriorVar", cols["dispPriorVar"], ref["dispPriorVar"], "theta", cols["theta"], ref["theta"],
                  "p-values beyond 1e-4:", int(np.nansum(rel > 1e-4)), "of", d.n)
            ok = ok and same and int(np.nansum(rel > 1e-4)) <= 5
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    if rank == 0:
        print("MULTI_GPU_SMALL_DF", "OK" if ok else "FAILED", "world", world)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
