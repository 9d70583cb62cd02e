"""Development aid (needs a GPU): times one step of the hot path (cd_aggregate + cd_region_test, rows resident) under the
build-time variants that can be switched at run time, on the bench workload, and prints the stage split.

    python scripts/fit_variants.py [workload] [n_regions|full] [steps]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chicdiff_b200 import engine  # noqa: E402
from _cache import cached_generate  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
nreg = None if len(sys.argv) < 3 or sys.argv[2] == "full" else int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
d = cached_generate(workload, nreg)
e = engine.Engine(0)
e.set_design(d.X); e.set_regions(d.row_off)
for s in range(d.S):
    e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
names = ["aggregate", "region_test", "fit_disp", "wald", "grid_refits", "trend_mad", "size_factors"]
ref = None
variants = [("table log", "1")] if os.environ.get("CHICDIFF_B200_LIB") else [("table log", "1"), ("fdlibm log", "0"), ("table log", "1")]
if os.environ.get("CHICDIFF_B200_LIB"):
    print("library:", os.environ["CHICDIFF_B200_LIB"])
for label, env in variants:
    os.environ["CHICDIFF_B200_TABLE_LOG"] = env
    for _ in range(2):
        e.aggregate(fetch=False); r = e.region_test(fetch="none")
    l0 = e.launch_count()
    e.timer_start()
    tm = np.zeros(8)
    for _ in range(steps):
        e.aggregate(fetch=False); r = e.region_test(fetch="none")
        tm += e.last_timings()
    ms = e.timer_stop() / steps
    tm /= steps
    print("%-11s n=%d: %.2f ms/step, %d launches/step | " % (label, d.n, ms, (e.launch_count() - l0) // steps) +
          "  ".join("%s %.2f" % (k, v) for k, v in zip(names, tm)), "| theta", r["theta"], flush=True)
    full = e.region_test(fetch="table")
    import zlib
    print("   checksums (bit-level, to compare two builds): pvalue %08x  lfcSE %08x  sum(p) %r" % (
        zlib.crc32(np.ascontiguousarray(full["pvalue"]).tobytes()), zlib.crc32(np.ascontiguousarray(full["lfcSE"]).tobytes()),
        float(np.nansum(full["pvalue"]))), flush=True)
    if ref is None:
        ref = full
    else:
        with np.errstate(invalid="ignore"):
            rel = np.abs(full["pvalue"] - ref["pvalue"]) / np.maximum(ref["pvalue"], 1e-300)
        print("   p-values vs first variant: max rel %.2e, beyond 1e-6: %d" % (np.nanmax(rel), int((rel > 1e-6).sum())))
e.close()
