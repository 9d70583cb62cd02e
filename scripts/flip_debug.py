"""Development aid (needs a GPU): lists the regions of a 400 k-region C3 subset whose gene-wise dispersion differs between
the CUDA path and the oracle, with the oracle's posterior at both estimates -- the tool behind the "decision flip"
analysis in DESIGN.md section 5 (all such rows sit on noise-level Armijo / stop decisions)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chicdiff_b200 import engine, synth
from oracle import oracle as O
d = synth.generate("c3", n_regions=400000)
e = engine.Engine(0)
e.set_design(d.X); e.set_regions(d.row_off)
for s in range(d.S): e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
K, FM = e.aggregate()
r = e.region_test(theta=0.0)
Ko, FMo = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
ro = O.region_test(Ko, FMo, d.X, theta=0.0)
ge, geo = r["dispGeneEst"], ro["dispGeneEst"]
rel = np.abs(ge - geo) / geo
bad = np.flatnonzero((rel > 1e-6) & (geo >= 1e-6) | ((ge >= 1e-6) & (geo < 1e-6)))
print("bad rows", len(bad))
for i in bad:
    print(i, "K", Ko[:, i].tolist(), "mu", np.round(ro["mu"][:, i], 3).tolist())
    print("   gpu est %.9e it %d fl %d | orc est %.9e it %d fl %d | fit %.4e  baseMean %.3f" % (ge[i], r["dispGeneIter"][i], r["flags"][i], geo[i], ro["dispGeneIter"][i], ro["flags"][i], ro["dispFit"][i], ro["baseMean"][i]))
    # oracle lp along a grid to see the landscape
    y = Ko[:, i].astype(float); mu = np.ascontiguousarray(ro["mu"][:, i]); X = np.ascontiguousarray(d.X)
    L = O.lib()
    for la in [np.log(ge[i]), np.log(geo[i])]:
        print("   lp(%.6f) = %.12f  dlp = %.6e" % (la, L.orc_log_posterior(la, 6, 2, X.ctypes.data, y.ctypes.data, mu.ctypes.data, 0., 1., 0, 1), L.orc_dlog_posterior(la, 6, 2, X.ctypes.data, y.ctypes.data, mu.ctypes.data, 0., 1., 0, 1)))
