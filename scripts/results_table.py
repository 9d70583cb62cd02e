"""Prints the results tables of DESIGN.md section 4 from the bench records under profiles/ (every number in those tables
is a field of a JSON line bench.py printed on the GPU box; nothing is typed in by hand).

    python scripts/results_table.py [profiles/r02_final_bench_1gpu.json profiles/r02_final_bench_2gpu.json ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def last_json_line(path):
    with open(path) as fh:
        lines = [ln for ln in fh.read().splitlines() if ln.startswith("{")]
    return json.loads(lines[-1])


def main():
    paths = sys.argv[1:] or [os.path.join(ROOT, "profiles", f) for f in
                             ("r02_final_bench_1gpu.json", "r02_final_bench_2gpu.json", "r02_final_bench_4gpu.json", "r02_final_bench_8gpu.json")]
    recs = [(p, last_json_line(p)) for p in paths if os.path.exists(p)]
    print("| GPUs | ms/step | regions/s (resident) | e2e ms/step | e2e regions/s | launches/step | fit_disp ms | FP64 flop frac | FP64 pipe frac | strong: ms, speed-up | file |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    base = None
    for p, r in recs:
        n = r["n_gpus"]
        if n == 1:
            base = r
        st = r.get("strong") or {}
        rf = r.get("roofline", {})
        strong = "-"
        if st:
            strong = "%.2f, %.2fx" % (st.get("ms_per_step", float("nan")), st.get("speedup_vs_one_gpu", float("nan")))
        print("| %d | %.2f | %.1f M | %.1f | %.1f M | %.0f | %.2f | %.3f | %s | %s | `%s` |" % (
            n, r["ms_per_step"], r["value"] / 1e6, r["e2e"]["ms_per_step"], r["e2e"]["value"] / 1e6,
            r.get("gpu_launches_per_step", 0), r["stage_ms"]["fit_disp_kernels"], rf.get("frac") or float("nan"),
            ("%.3f" % rf["fp64_pipe_frac"]) if rf.get("fp64_pipe_frac") else "-", strong, os.path.basename(p)))
    if base:
        print()
        print("weak-scaling efficiency against the 1-GPU line (driver computes its own): " + ", ".join(
            "N=%d: %.3f (resident), %.3f (e2e)" % (r["n_gpus"], r["value"] / r["n_gpus"] / base["value"],
                                                   r["e2e"]["value"] / r["n_gpus"] / base["e2e"]["value"]) for _, r in recs if r["n_gpus"] > 1))
        cb = base.get("cpu_baseline")
        if cb:
            print("CPU baseline (%s, %d cores): %.0f regions/s on %s" % (cb["kind"], cb["cores"], cb["value"], cb["sample"]))
    for f in ("r02_sweep_1gpu.json", "r02_sweep_8gpu.json"):
        p = os.path.join(ROOT, "profiles", f)
        if not os.path.exists(p):
            continue
        sw = last_json_line(p)
        print()
        print("sweep, %d GPU(s) (`%s`):" % (sw["n_gpus"], f))
        print("| regions total | per GPU | ms/step | regions/s | fit_disp FP64 flop frac | aggregate HBM frac |")
        print("|---|---|---|---|---|---|")
        for e in sw["sweep"]:
            print("| %d | %d | %.2f | %.1f M | %.3f | %.3f |" % (e["regions_total"], e["regions_per_gpu"], e["ms_per_step"],
                                                                 e["value"] / 1e6, e["fit_disp_fp64_frac"], e["aggregate_hbm_frac"]))


if __name__ == "__main__":
    main()
