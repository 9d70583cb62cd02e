"""Runs exactly one step of the bench workload (cd_aggregate + cd_region_test) and writes the line-search evaluation
counts of that step (cd_last_search_counts) as JSON: the run ncu profiles to count the FP64 instructions of the
line-search kernels, see scripts/flop_per_eval.py.

    ncu --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum --clock-control none -k regex:fit_disp --csv \
        --log-file gpurun_out/flop_launches.csv python scripts/flop_probe.py c3 full gpurun_out/flop_counts.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chicdiff_b200 import engine  # noqa: E402
from _cache import cached_generate  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
nreg = None if len(sys.argv) < 3 or sys.argv[2] == "full" else int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/flop_counts.json"
d = cached_generate(workload, nreg)
e = engine.Engine(0)
e.set_design(d.X); e.set_regions(d.row_off)
for s in range(d.S):
    e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
e.aggregate(fetch=False)
e.region_test(fetch="none")
calls = e.last_search_counts()
with open(out, "w") as fh:
    json.dump({"workload": workload, "regions": d.n, "S": d.S, "calls": [{"evaluations": ev, "p": p, "regions": rg} for ev, p, rg in calls]}, fh)
print("SEARCH_COUNTS", calls)
e.close()
