#!/usr/bin/env python
"""bench.py -- regions tested per second through the Chicdiff hot path (aggregation + offsets + NB GLM +
dispersion + Wald, with the default norm="combined" theta grid) on synthetic genome-wide 3-vs-3 PCHi-C data
(BASELINE.json configs[2], the configuration the metric's target is quoted on; it fits one GPU).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # CPU restatement of the reference (oracle port)

One step = one pass of DESeq2Wrap's numeric core over one synthetic batch: cd_aggregate + cd_region_test
(size factors, 5 intercept-only theta-grid fits, final dispersion fit, IRLS, Cook's, Wald).
`value` times steps whose inputs are already resident in HBM; `e2e` times the same step through the host-
buffer C-ABI calls, with the pinned-host -> device copy of every input column and the device -> host read of
the output-table columns inside the timed region.  Under torchrun each rank owns a genome-wide shard of its
own (weak scaling: regions are partitioned by bait; the collectives are the NCCL all-gathers / all-reduces of
the global steps).  Inputs (1.7 GB per rank) are far larger than L2 (126 MB), so no explicit L2 flush is needed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class stdout_to_stderr:
    """NCCL prints its version banner on stdout when the first communicator comes up; stdout must carry only
    the one JSON line, so file descriptor 1 points at stderr while the communicators are created."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)

METRIC = "regions tested/sec (agg+NB GLM+dispersion+Wald)"
UNIT = "regions/s"


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_data(workload, n_regions, rank):
    from chicdiff_b200 import synth
    return synth.generate(workload, n_regions=n_regions, seed_offset=1000 * rank)


def cpu_reference_rate(d, sample_regions, threads):
    """Times the oracle port (CPU restatement of the reference) on the first `sample_regions` regions."""
    from oracle import oracle as O
    m = min(sample_regions, d.n)
    r_hi = int(d.row_off[m])
    row_off = d.row_off[: m + 1]
    N = np.ascontiguousarray(d.N_rows[:, :r_hi])
    FMr = np.ascontiguousarray(d.FM_rows[:, :r_hi])
    t0 = time.perf_counter()
    K, FM = O.aggregate(row_off, N, FMr)
    O.region_test(K, FM, d.X, nthreads=threads)
    dt = time.perf_counter() - t0
    return m / dt, dt, m


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    threads = os.cpu_count() or 1
    d = make_data(args.workload, args.regions, 0)
    # bounded sample: the whole --steps K run should end within a few minutes whatever K is.  A small probe gives the
    # rate of this box; the per-step sample is what fits ~150 s / K, between 20 000 regions and --cpu-sample
    probe_rate, _, _ = cpu_reference_rate(d, min(args.cpu_sample, 20000), threads)
    sample = int(max(20000, min(args.cpu_sample, probe_rate * 150.0 / max(args.steps, 1))))
    sample = min(sample, d.n)
    times = []
    m = 0
    for _ in range(args.steps):
        rate, dt, m = cpu_reference_rate(d, sample, threads)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = m / (ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "synthetic genome-wide 3-vs-3 PCHi-C (BASELINE configs[2])", "regions_per_step": m,
                       "samples": d.S, "design_columns": int(d.X.shape[1]), "norm": "combined", "theta_grid": 5},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "first %d of %d regions of the same synthetic set, full DESeq2Wrap numerics "
                                       "(aggregation, size factors, 5 theta-grid fits, final fit), OpenMP over regions" % (m, d.n)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--regions", type=int, default=None, help="override the number of regions per rank")
    ap.add_argument("--cpu-sample", type=int, default=600000, help="regions in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): every rank owns a genome-wide set of its own; strong: ONE set, regions partitioned by "
                         "bait across the ranks (cd_plan_shards)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from chicdiff_b200 import engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; chicdiff_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    e = engine.Engine(local_rank)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            uid = [e.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            e.comm_init(world, rank, uid[0])
            dist.barrier()
            torch.cuda.synchronize()

    d = make_data(args.workload, args.regions, rank if args.scaling == "weak" else 0)
    if args.scaling == "strong" and world > 1:
        # one genome-wide set, regions partitioned by bait (every rank generates the same set and keeps its shard)
        from chicdiff_b200 import parallel
        bounds = parallel.shard_slices(d.region_bait, d.row_off, world)
        off, (Nl, FMl, rb, ro), (lo, hi) = parallel.take_shard(d.row_off, [d.N_rows, d.FM_rows, d.row_bait, d.row_oe], bounds, rank)
        d.row_off, d.N_rows, d.FM_rows, d.row_bait, d.row_oe = off, Nl, FMl, rb, ro
        d.region_bait, d.region_seed, d.true_lfc = d.region_bait[lo:hi], d.region_seed[lo:hi], d.true_lfc[lo:hi]
    S, p, n, R = d.S, int(d.X.shape[1]), d.n, d.R
    e.set_design(d.X)
    e.set_regions(d.row_off)

    # pinned host copies of the inputs (what the R glue would hand over) and device-resident copies
    N_host = torch.from_numpy(d.N_rows).pin_memory()
    FM_host = torch.from_numpy(d.FM_rows).pin_memory()
    N_dev = N_host.cuda(non_blocking=True)
    FM_dev = FM_host.cuda(non_blocking=True)
    torch.cuda.synchronize()
    fp64_peak = e.measure_fp64_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        e.aggregate(fetch=False)
        return e.region_test(fetch="none")

    def step_e2e():
        for s in range(S):
            e.set_sample_rows_ptr(s, R, N_host[s].data_ptr(), FM_host[s].data_ptr())
        e.aggregate(fetch=False)
        return e.region_test(fetch="table")

    def timed(fn, steps):
        barrier()
        e.timer_start()
        t0 = time.perf_counter()
        tm = np.zeros(8)
        for _ in range(steps):
            fn()
            tm += e.last_timings()
        dev_ms = e.timer_stop()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / steps, float(t[1]) / steps, tm / steps

    # device-resident steps
    e.set_rows_device(R, N_dev.data_ptr(), FM_dev.data_ptr())
    for _ in range(max(args.warmup, 3)):
        step_resident()
    launches0 = e.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    dev_ms, wall_ms, tm = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    launches = (e.launch_count() - launches0)

    # end-to-end steps (host buffers in, table columns out)
    e.set_regions(d.row_off)
    for _ in range(2):
        step_e2e()
    e2e_dev_ms, e2e_wall_ms, _ = timed(step_e2e, max(2, args.steps // 2))

    # the same step started one stage earlier: per-replicate CHiCAGO tables -> fused assembly + aggregation
    # (cd_assemble) -> region test.  Reported beside the main line, not instead of it.
    asm = None
    if "tables" in d.extra:
        pin = lambda a: torch.from_numpy(a).pin_memory()
        packed = [engine.Engine.pack_sample_tables(t, pin=pin) for t in d.extra["tables"]]
        asm_h2d = sum(int(k.numel() * k.element_size()) for _, keep in packed for k in keep)
        e.set_rmap(d.frag_chr, d.frag_start, d.frag_end, 1)
        e.set_regions(d.row_off)
        e.set_region_rows(d.row_bait, d.row_oe)
        for si in range(S):
            e.set_sample_tables(si, packed[si])

        def step_asm_resident():
            e.assemble(fetch=False)
            return e.region_test(fetch="none")

        def step_asm_e2e():
            for si in range(S):
                e.set_sample_tables(si, packed[si])
            e.assemble(fetch=False)
            return e.region_test(fetch="table")

        for _ in range(2):
            step_asm_resident()
        a_dev_ms, _, a_tm = timed(step_asm_resident, max(2, args.steps // 2))
        for _ in range(2):
            step_asm_e2e()
        a_e2e_ms, _, _ = timed(step_asm_e2e, max(2, args.steps // 2))
        asm = {"what": "per-replicate CHiCAGO tables (s_j, s_i, tblb/tlb, Tmean table, distance function, sparse counts) -> cd_assemble "
                       "(joins + Bmean/Tmean + count merge + region sums in one kernel) -> cd_region_test",
               "ms_per_step": a_dev_ms, "assemble_kernel_ms": a_tm[0], "e2e_ms_per_step": a_e2e_ms, "h2d_bytes_per_step": asm_h2d}

    # the "next" step after the Wald test, outside the metric: results() (Cook's cutoff, independent filtering, BH) on the
    # arrays still in device memory, beside the host routine on the same columns
    res_step = None
    if world == 1:
        r_tab = e.region_test(fetch="table")
        e.results_resident()
        t0 = time.perf_counter()
        for _ in range(3):
            adj_dev = e.results_resident()
        res_dev_ms = (time.perf_counter() - t0) / 3 * 1e3
        t0 = time.perf_counter()
        adj_host = engine.results_adjust(r_tab["baseMean"], r_tab["maxCooks"], r_tab["flags"], r_tab["pvalue"], S, p)
        res_host_ms = (time.perf_counter() - t0) * 1e3
        res_step = {"what": "results(): cd_results_resident (device sorts + prefix counts, padj and filtered p-values copied to the "
                            "host) vs cd_results_adjust (host) on the same columns",
                    "device_wall_ms": res_dev_ms, "host_wall_ms": res_host_ms,
                    "identical": bool(np.array_equal(adj_dev["padj"], adj_host["padj"], equal_nan=True)),
                    "d2h_bytes": int(n * 16)}

    n_tot = torch.tensor([n], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(n_tot)
    n_tot = int(n_tot)

    if rank == 0:
        peaks, peak_src = read_peaks()
        W = R / n
        agg_bytes = n * (W * S * 12 + 8 + S * 12)          # rows read once + CSR offsets + outputs written
        agg_ms = tm[0]
        achieved = agg_bytes / (agg_ms * 1e-3) / 1e9
        value = n_tot / (dev_ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "synthetic genome-wide 3-vs-3 PCHi-C (BASELINE configs[2]): %d regions, %d region rows, "
                                       "%d samples per GPU" % (n, R, S),
                           "regions_total": n_tot, "design_columns": p, "norm": "combined", "theta_grid": 5,
                           "fits_per_step": 6, "l2": "inputs (%.2f GB per GPU) larger than L2; no flush" % ((N_host.numel() * 4 + FM_host.numel() * 8) / 1e9),
                           "parallelism": "regions sharded by bait, %d rank(s); global steps by all-reduce only (trend sums: %s; median "
                                          "histograms: %s; offsets sums, deviance: NCCL); nothing is gathered"
                                          % (world, "in-kernel over NVLink peer memory" if e.comm_info()["peer_memory_allreduce"] else "NCCL",
                                             "in-kernel over NVLink peer memory" if e.comm_info()["peer_memory_medians"] else "NCCL")},
                "wall_ms_per_step": wall_ms,
                "stage_ms": {"aggregate": tm[0], "region_test": tm[1], "fit_disp_kernels": tm[2], "wald_kernels": tm[3],
                             "grid_refits": tm[4], "trend_and_mad": tm[5], "size_factors": tm[6]},
                "roofline": {"bound": "hbm", "kernel": "aggregate_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                             # dram__bytes_read.sum + dram__bytes_write.sum of one launch on this workload, from the
                             # committed ncu --set full capture (profiles/r01_final_kernels_aggregate_assemble_irls.txt)
                             "traffic": 1841515000.0 if (n == 2135814 and S == 6) else None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": agg_bytes},
                "fp64": {"peak_tflops_measured_dfma": fp64_peak, "fit_disp_ms": tm[2], "wald_ms": tm[3],
                         # from the committed ncu capture of the line-search kernel on this workload (not re-measured here):
                         # (2 DFMA + DMUL + DADD) per cycle x SM clock, and the FP64 pipe's busy cycles
                         "line_search_tflops_ncu": 15.0, "line_search_frac_of_dfma_peak_ncu": 0.41,
                         "line_search_fp64_pipe_busy_ncu": 0.59,
                         "ncu_source": "profiles/r01_final_kernels_fit_disp_trend.txt"},
                "e2e": {"value": n_tot / (e2e_dev_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_dev_ms,
                        "h2d_bytes_per_step": int(N_host.numel() * 4 + FM_host.numel() * 8),
                        "d2h_bytes_per_step": int(n * (6 * 8 + 1))},
                "gpu_launches": int(launches), "clocks": clocks}
        if asm is not None:
            asm["value"] = n_tot / (asm["ms_per_step"] * 1e-3)
            asm["e2e_value"] = n_tot / (asm["e2e_ms_per_step"] * 1e-3)
            line["assembly_path"] = asm
        if res_step is not None:
            line["results_step"] = res_step
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            rate, dt, m = cpu_reference_rate(d, args.cpu_sample, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "first %d of %d regions of the same synthetic set (%.1f s), full DESeq2Wrap numerics, "
                                              "OpenMP over regions on all host cores" % (m, n, dt)}
            # the reference's DESeq2 loops are single-threaded and Chicdiff never enables BiocParallel: one core as well
            rate1, dt1, m1 = cpu_reference_rate(d, max(2000, args.cpu_sample // 16), 1)
            line["cpu_baseline_single_thread"] = {"value": rate1, "unit": UNIT, "cores": 1, "kind": "port",
                                                  "sample": "first %d regions (%.1f s)" % (m1, dt1)}
        print(json.dumps(line), flush=True)
    e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
