"""FP64 flop per line-search evaluation, from an ncu launch list of scripts/flop_probe.py and the evaluation counts the
same run wrote (convention of SURVEY.md section 8d: DFMA = 2 flop, DMUL = DADD = 1; thread-level, predicated-on
instruction counts).  Every dispersion line search is two launches (first pass, then the parked regions), in the order
cd_last_search_counts reports the searches.  Output: profiles/<name>.json, read by bench.py for the roofline entry of
the line-search kernel -- the bench multiplies it with the evaluations it counts live.

    python scripts/flop_per_eval.py gpurun_out/flop_launches.csv gpurun_out/flop_counts.json profiles/r02_fit_disp_flop_per_eval.json
"""
import csv
import json
import sys


def read_launches(path):
    """-> [(kernel name, {metric: value})] in launch order"""
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    rd = csv.DictReader(lines)
    by_id = {}
    order = []
    for r in rd:
        k = r["ID"]
        if k not in by_id:
            by_id[k] = (r["Kernel Name"], {})
            order.append(k)
        v = r["Metric Value"].replace(",", "")
        try:
            val = float(v)
        except ValueError:
            continue
        if r["Metric Unit"] in ("usecond", "us"):
            val *= 1e3
        elif r["Metric Unit"] in ("msecond", "ms"):
            val *= 1e6
        by_id[k][1][r["Metric Name"]] = val
    return [by_id[k] for k in order]


def main():
    launches = [(nm, m) for nm, m in read_launches(sys.argv[1]) if "fit_disp" in nm and "grid" not in nm]
    counts = json.load(open(sys.argv[2]))
    calls = counts["calls"]
    assert len(launches) == 2 * len(calls), (len(launches), len(calls))
    out = {"source": "ncu thread-level FP64 instruction counts of the line-search launches of one step of %s (%d regions, S = %d); "
                     "DFMA = 2 flop, DMUL = DADD = 1" % (counts["workload"], counts["regions"], counts["S"]),
           "S": counts["S"], "per_design_columns": {}, "calls": []}
    agg = {}
    for k, c in enumerate(calls):
        fl = ns = inst = 0.0
        for nm, m in launches[2 * k: 2 * k + 2]:
            fma = m["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
            other = m["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] + m["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
            fl += 2 * fma + other
            inst += fma + other
            ns += m.get("gpu__time_duration.sum", 0.0)
        out["calls"].append({"p": c["p"], "regions": c["regions"], "evaluations": c["evaluations"], "fp64_flop": fl,
                             "fp64_instructions": inst, "flop_per_evaluation": fl / c["evaluations"], "ncu_ns": ns})
        a = agg.setdefault(str(c["p"]), [0.0, 0, 0.0])
        a[0] += fl; a[1] += c["evaluations"]; a[2] += inst
    # fp64_instructions: DFMA + DMUL + DADD thread instructions; every one of them takes one slot of the FP64 pipe, so
    # instructions / (peak flop/s / 2) is the pipe's utilisation, which the flop fraction understates when the mix is not all FMA
    for p, (fl, ev, inst) in agg.items():
        out["per_design_columns"][p] = {"flop_per_evaluation": fl / ev, "flop_per_replicate_evaluation": fl / ev / counts["S"],
                                        "fp64_instructions_per_evaluation": inst / ev}
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out["per_design_columns"]))


if __name__ == "__main__":
    main()
