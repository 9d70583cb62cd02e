"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv > profiles/rNN_launches.txt
    python scripts/summarize_ncu.py raw gpurun_out/prof_raw.csv      > profiles/rNN_kernels.txt
(the raw csv comes from `ncu -i prof.ncu-rep --page raw --csv`)
"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        nm = r["Kernel Name"].split("(")[0][:80]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
        agg[nm][0] += 1
        agg[nm][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised:")
    print("# compare SHARES, not absolutes).  total %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))
    print("%-82s %6s %11s %7s" % ("kernel", "n", "ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-82s %6d %11.3f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))


def raw(path):
    rd = list(csv.reader(open(path)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        print("== %s" % r[ix["Kernel Name"]][:110])
        for k in KEYS:
            if k in ix:
                print("   %-78s %s %s" % (k, r[ix[k]], units[ix[k]]))
        for h in hdr:
            if "issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                v = float(r[ix[h]] or 0)
                if v > 5:
                    print("   stall %-72s %.1f %%" % (h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
