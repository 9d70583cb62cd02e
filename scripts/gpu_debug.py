"""Development aid: runs the CUDA path and the oracle on one synthetic set and prints per-column errors."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chicdiff_b200 import synth, engine
from oracle import oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
nreg = int(sys.argv[2]) if len(sys.argv) > 2 else None
d = synth.generate(name, n_regions=nreg)
print("config", name, "n", d.n, "R", d.R, "S", d.S, flush=True)
e = engine.Engine(0)
e.set_design(d.X)
e.set_regions(d.row_off)
for s in range(d.S):
    e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
K, FM = e.aggregate()
Ko, FMo = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
print("K exact:", np.array_equal(K, Ko), " FM max rel:", np.nanmax(np.abs(FM - FMo) / np.abs(FMo)),
      " NaN pattern:", np.array_equal(np.isnan(FM), np.isnan(FMo)))
pv = None if d.S - d.X.shape[1] > 3 else 0.5
t = time.time()
r = e.region_test(disp_prior_var=pv, disp_prior_var_grid=None if d.S - 1 > 3 else 0.5)
print("gpu region_test %.3f s; timings" % (time.time() - t), e.last_timings()[:2], "launches", e.launch_count())
t = time.time()
ro = O.region_test(Ko, FMo, d.X, prior_var=float("nan") if pv is None else pv,
                   prior_var_grid=float("nan") if d.S - 1 > 3 else 0.5)
print("oracle %.3f s" % (time.time() - t))
print("sf", np.max(np.abs(r["sizeFactors"] - ro["sizeFactors"]) / ro["sizeFactors"]))
print("deviances gpu", r["deviances"], "\n          orc", ro["deviances"], "theta", r["theta"], ro["theta"])
for k in ["trend_a0", "trend_a1", "varLogDispEsts", "dispPriorVar"]:
    print(k, r[k], ro[k], abs(r[k] - ro[k]) / abs(ro[k]))


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    bad_nan = np.isnan(a) != np.isnan(b)
    with np.errstate(invalid="ignore", divide="ignore"):
        e_ = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    e_[np.isnan(e_)] = 0
    e_[(a == b)] = 0
    return e_, bad_nan.sum()


for k, ko in [("baseMean", "baseMean"), ("normFactors", "nf"), ("mu", "mu"), ("dispGeneEst", "dispGeneEst"), ("dispFit", "dispFit"),
              ("dispMAP", "dispMAP"), ("dispersion", "dispersion"), ("beta", "beta"), ("betaSE", "betaSE"),
              ("stat", "stat"), ("pvalue", "pvalue"), ("deviance", "deviance"), ("maxCooks", "maxCooks")]:
    e_, nb = rel(r[k], ro[ko])
    print("%-12s max rel %.3e  #>1e-6: %d  nan mismatch %d" % (k, e_.max(), (e_ > 1e-6).sum(), nb))
for k in ["dispGeneIter", "dispIter", "betaIter"]:
    print(k, "equal:", np.array_equal(r[k], ro[k]), " ndiff", (r[k] != ro[k]).sum())
fo = ro["flags"].astype(int); fg = r["flags"].astype(int) & 63
print("flags equal", np.array_equal(fo, fg), "ndiff", (fo != fg).sum(), np.bincount(fg))
# results()
adj = engine.results_adjust(r["baseMean"], r["maxCooks"], r["flags"], r["pvalue"], d.S, d.X.shape[1])
resO = O.results(ro, Ko, d.X)
e_, nb = rel(adj["padj"], resO["padj"])
print("padj max rel %.3e nan mismatch %d ; sig set equal %s (%d)" % (e_.max(), nb,
      np.array_equal(adj["padj"] < 0.05, resO["padj"] < 0.05), np.nansum(resO["padj"] < 0.05)))
print("filter idx", adj["filterIndex"], resO["filterIndex"] + 1, "cooks outliers", int(np.isnan(adj["pvalue"]).sum() - np.isnan(r["pvalue"]).sum()), int(resO["cooksOutlier"].sum()))
# mismatching gene-wise estimates in detail
e_, _ = rel(r["dispGeneEst"], ro["dispGeneEst"])
bad = np.flatnonzero(e_ > 1e-6)
print("gene-est mismatches: %d ; of which oracle est > 1e-6: %d ; gpu est > 1e-6: %d" % (
    len(bad), (ro["dispGeneEst"][bad] > 1e-6).sum(), (r["dispGeneEst"][bad] > 1e-6).sum()))
order = bad[np.argsort(-np.maximum(ro["dispGeneEst"][bad], r["dispGeneEst"][bad]))]
for i in order[:25]:
    print(i, "K", Ko[:, i].tolist(), "gpu est %.6e it %d fl %d | orc est %.6e it %d fl %d | fit %.3e" % (
        r["dispGeneEst"][i], r["dispGeneIter"][i], r["flags"][i], ro["dispGeneEst"][i], ro["dispGeneIter"][i], ro["flags"][i], ro["dispFit"][i]))
np.savez_compressed("gpurun_out/debug_%s.npz" % name, bad=bad, K=Ko[:, bad], mu=ro["mu"][:, bad], gpu_est=r["dispGeneEst"][bad],
                    orc_est=ro["dispGeneEst"][bad], gpu_it=r["dispGeneIter"][bad], orc_it=ro["dispGeneIter"][bad], flags=ro["flags"][bad])
