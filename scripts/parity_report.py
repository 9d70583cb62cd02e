"""Per-region parity report of the CUDA path against the CPU oracle on the parity-test configurations
(BASELINE.json configs[0..3] shapes).  Written to stdout; kept under profiles/."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chicdiff_b200 import engine, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [("c1", None, 0.5, "chr19-shaped 2-vs-2 (configs[0] shape), dispPriorVar given (S-p=2)"),
         ("c2", None, 0.6, "1 chromosome 2-vs-2 (configs[1]), dispPriorVar given"),
         ("c3", 400000, None, "genome-wide 3-vs-3 (configs[2]) first 400k-region universe"),
         ("c4", 60000, None, "8-vs-8 + batch, 3-column GLM (configs[3]) 60k-region universe")]


def rel(a, b, scale=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    ref = np.abs(b) if scale is None else np.maximum(np.abs(b), scale)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.abs(a - b) / np.maximum(ref, 1e-300)
    e[np.isnan(e) | (a == b)] = 0
    return e


for name, nreg, prior, what in CASES:
    d = synth.generate(name, n_regions=nreg)
    e = engine.Engine(0)
    e.set_design(d.X); e.set_regions(d.row_off)
    for s in range(d.S):
        e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    K, FM = e.aggregate()
    t0 = time.time()
    r = e.region_test(disp_prior_var=prior, disp_prior_var_grid=prior)
    t_gpu = time.time() - t0
    Ko, FMo = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
    nan = float("nan")
    t0 = time.time()
    ro = O.region_test(Ko, FMo, d.X, prior_var=nan if prior is None else prior, prior_var_grid=nan if prior is None else prior)
    t_cpu = time.time() - t0
    p = d.X.shape[1]
    print("== %s: %s" % (name, what))
    print("   n = %d regions, R = %d rows, S = %d, p = %d ; GPU region_test %.3f s, oracle (%d threads) %.2f s" % (
        d.n, d.R, d.S, p, t_gpu, O.lib().orc_num_threads(), t_cpu))
    print("   aggregated counts bit-exact: %s ; FullMean sums max rel %.2e ; theta %s / %s" % (
        np.array_equal(K, Ko), np.nanmax(rel(FM, FMo)), r["theta"], ro["theta"]))
    coup = max(abs(r["trend_a0"] - ro["trend_a0"]) / ro["trend_a0"], abs(r["trend_a1"] - ro["trend_a1"]) / ro["trend_a1"])
    print("   trend coefficients rel diff %.2e ; varLogDispEsts rel diff %.2e ; dispPriorVar %.6g / %.6g" % (
        coup, abs(r["varLogDispEsts"] - ro["varLogDispEsts"]) / ro["varLogDispEsts"], r["dispPriorVar"], ro["dispPriorVar"]))
    floor = ro["dispGeneEst"] < 1e-6
    cols = [("dispGeneEst (>= 1e-6)", np.where(floor, ro["dispGeneEst"], r["dispGeneEst"]), ro["dispGeneEst"], None),
            ("dispFit", r["dispFit"], ro["dispFit"], None), ("dispMAP", r["dispMAP"], ro["dispMAP"], None),
            ("dispersion", r["dispersion"], ro["dispersion"], None),
            ("log2FoldChange (scale: lfcSE)", r["log2FoldChange"], ro["beta"][p - 1], ro["betaSE"][p - 1]),
            ("lfcSE", r["lfcSE"], ro["betaSE"][p - 1], None), ("stat (scale: 1)", r["stat"], ro["stat"], 1.0),
            ("pvalue", r["pvalue"], ro["pvalue"], None),
            ("pvalue / max(1, z^2)", r["pvalue"], ro["pvalue"], np.abs(ro["pvalue"]) * np.maximum(1.0, ro["stat"] ** 2))]
    print("   %-32s %10s %10s %10s %10s" % ("quantity", "max rel", "#>1e-6", "#>1e-5", "frac<=1e-6"))
    for nm, a, b, sc in cols:
        ee = rel(a, b, sc)
        print("   %-32s %10.2e %10d %10d %10.6f" % (nm, ee.max(), (ee > 1e-6).sum(), (ee > 1e-5).sum(), (ee <= 1e-6).mean()))
    print("   gene-wise estimates at the floor (<1e-6): oracle %d, of which GPU also at the floor %d" % (floor.sum(), (r["dispGeneEst"][floor] < 1e-6 * (1 + 1e-9)).sum()))
    print("   iteration counts equal: dispGeneIter %.4f  dispIter %.4f  betaIter %.4f ; flags equal %.6f" % (
        (r["dispGeneIter"] == ro["dispGeneIter"]).mean(), (r["dispIter"] == ro["dispIter"]).mean(),
        (r["betaIter"] == ro["betaIter"]).mean(), ((r["flags"] & 63) == ro["flags"]).mean()))
    adj = engine.results_adjust(r["baseMean"], r["maxCooks"], r["flags"], r["pvalue"], d.S, p)
    res_o = O.results(ro, Ko, d.X)
    with np.errstate(invalid="ignore"):
        sg, so = adj["padj"] < 0.05, res_o["padj"] < 0.05
    print("   results(): filter index %d / %d ; padj NA pattern equal %s ; significant calls (padj<0.05): %d / %d, differing %d" % (
        adj["filterIndex"], res_o["filterIndex"] + 1, np.array_equal(np.isnan(adj["padj"]), np.isnan(res_o["padj"])),
        sg.sum(), so.sum(), (sg != so).sum()))
    e.close()
