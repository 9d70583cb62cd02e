"""Two independent CPU restatements of the DESeq2 numerics must agree: oracle/chicdiff_oracle.c (own special functions,
Rmath-style dnbinom, packed Cholesky) against oracle/crosscheck.py (SciPy special functions and NB density, QR of the
ridge-augmented design, lstsq trend fit), both written from SURVEY.md Appendix A.  No R session exists to pin either
("parity unpinned"); a transcription error in one of them shows up here as disagreement."""
import numpy as np
import pytest

from chicdiff_b200 import synth
from oracle import crosscheck as X2
from oracle import oracle as O

TOL = 1e-9


def _agree(name, a, b, skip=None):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert np.array_equal(np.isnan(a), np.isnan(b)), name
    ok = ~np.isnan(a)
    if skip is not None:
        ok &= ~skip
    rel = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(a[ok]), 1e-300)
    assert rel.max() < TOL, (name, float(rel.max()))


def _both(d, theta, m):
    K, FM = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
    nf = O.norm_factors(FM, O.size_factors(K), "combined", theta)
    return np.ascontiguousarray(K[:, :m]), np.ascontiguousarray(nf[:, :m])


@pytest.mark.parametrize("case", ["3v3 ~condition", "3v3 ~1", "8v8 ~batch+condition"])
def test_c_and_numpy_restatements_agree(case):
    if case.startswith("3v3"):
        d = synth.generate("tiny")
        K, nf = _both(d, 0.5, 300)
        X = d.X if case.endswith("~condition") else np.ones((d.S, 1))
    else:
        d = synth.generate("c4", n_regions=150)       # 4 design cells, p = 3: mu comes from an NB GLM, not the hat matrix
        K, nf = _both(d, 0.25, 150)
        X = d.X
        assert len(np.unique(X, axis=0)) == 4 and X.shape[1] == 3
    p = X.shape[1]
    a = O.deseq(K, nf, X)
    b = X2.deseq(K, nf, X)
    # At the 1e-8 floor the gene-wise posterior is rounding noise of lgamma(1e8) in any implementation (DESeq2 excludes
    # those rows from the trend for that reason): both sides must be at the floor, iteration counts may differ there.
    with np.errstate(invalid="ignore"):
        floor = (a["dispGeneEst"] < 1e-6) | (b["dispGeneEst"] < 1e-6)
        assert np.array_equal(a["dispGeneEst"] < 1e-6, b["dispGeneEst"] < 1e-6)
    _agree("baseMean", a["baseMean"], b["baseMean"])
    _agree("dispGeneEst", a["dispGeneEst"], b["dispGeneEst"], skip=floor)
    assert abs(a["trend_a0"] - b["trend"][0]) < TOL * a["trend_a0"] and abs(a["trend_a1"] - b["trend"][1]) < TOL * a["trend_a1"]
    assert abs(a["varLogDispEsts"] - b["varLogDispEsts"]) < TOL and abs(a["dispPriorVar"] - b["dispPriorVar"]) < TOL
    for k, kb in (("dispFit", "dispFit"), ("dispMAP", "dispMAP"), ("dispersion", "dispersion"), ("stat", "stat"),
                  ("pvalue", "pvalue"), ("deviance", "deviance")):
        _agree(k, a[k], b[kb])
    _agree("log2FoldChange", a["beta"][p - 1], b["lfc"])
    _agree("lfcSE", a["betaSE"][p - 1], b["lfcSE"])
    if p > 1:
        _agree("maxCooks", a["maxCooks"], b["maxCooks"])
    assert np.array_equal(a["dispGeneIter"][~floor], b["dispGeneIter"][~floor])
    assert np.array_equal(a["dispIter"], b["dispIter"])
    assert np.array_equal(a["betaIter"], b["betaIter"])
