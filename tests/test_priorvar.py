"""DESeq2's dispersion prior variance for 1 <= S - p <= 3 residual degrees of freedom (csrc/priorvar.cpp): R's random
streams against values R is known to print, the samplers and the smoother against an independent restatement
(oracle/priorvar.py) and against their definitions, the whole rule against that restatement and against the variance it
is meant to recover.  No GPU involved: the rule is host code."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

from chicdiff_b200 import engine
from oracle import priorvar as P


@pytest.fixture(scope="module")
def L():
    lib = engine.load_library()
    lib.cd_prior_var_debug_stream.argtypes = [C.c_uint, C.c_int, C.c_double, C.c_int, C.c_void_p]
    lib.cd_prior_var_small_df.restype = C.c_double
    lib.cd_prior_var_small_df.argtypes = [C.c_int, C.c_int64, C.c_void_p]
    lib.cd_prior_var_from_hist.restype = C.c_double
    lib.cd_prior_var_from_hist.argtypes = [C.c_int, C.c_void_p]
    lib.cd_prior_var_hist.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
    lib.cd_prior_var_debug_curve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def stream(L, seed, what, n, shape=1.0):
    out = np.empty(n)
    assert L.cd_prior_var_debug_stream(seed, what, shape, n, out.ctypes.data) == 0
    return out


def test_streams_reproduce_what_r_prints(L):
    """set.seed(s); runif(3) / rnorm(3) / rexp(3) as printed by R (7 significant digits): pins the seed scrambling, the
    Mersenne-Twister, the fix-up, the inversion with two uniforms, Wichura's quantile function and Ahrens-Dieter's exp_rand"""
    known = {(42, 0): [0.9148060, 0.9370754, 0.2861395], (1, 0): [0.2655087, 0.3721239, 0.5728534],
             (123, 0): [0.2875775, 0.7883051, 0.4089769],
             (42, 1): [1.37095845, -0.56469817, 0.36312841], (1, 1): [-0.6264538, 0.1836433, -0.8356286],
             (123, 1): [-0.56047565, -0.23017749, 1.55870831],
             (1, 2): [0.7551818, 1.1816428, 0.1457067], (42, 2): [0.1983368, 0.6608953]}
    for (seed, what), vals in known.items():
        got = stream(L, seed, what, len(vals))
        assert np.allclose(got, vals, rtol=0, atol=6e-8 if what != 1 or seed != 1 else 6e-8), (seed, what, got)


def test_streams_agree_with_the_independent_restatement(L):
    """NumPy's MT19937 seeded with R's scrambled state, SciPy's normal quantile, the samplers transcribed a second time"""
    for seed in (2, 7):
        g = P.RStream(seed)
        assert np.array_equal(stream(L, seed, 0, 5000), np.array([g.unif() for _ in range(5000)]))
        g = P.RStream(seed)
        assert np.allclose(stream(L, seed, 1, 3000), [P.norm_rand(g) for _ in range(3000)], rtol=0, atol=2e-15)
        g = P.RStream(seed)
        assert np.array_equal(stream(L, seed, 2, 3000), np.array([P.exp_rand(g) for _ in range(3000)]))
        for shape in (0.5, 1.0, 1.5):
            g = P.RStream(seed)
            ref = np.array([P.rgamma(g, shape) for _ in range(4000)])
            assert np.allclose(stream(L, seed, 3, 4000, shape), ref, rtol=1e-13, atol=0), shape


def test_samplers_have_the_right_distributions(L):
    n = 200000
    assert stats.kstest(stream(L, 11, 1, n), "norm").pvalue > 1e-3
    assert stats.kstest(stream(L, 12, 2, n), "expon").pvalue > 1e-3
    for shape in (0.5, 1.0, 1.5):
        x = stream(L, 13, 3, n, shape)
        assert stats.kstest(x, "gamma", args=(shape,)).pvalue > 1e-3, shape
        assert abs(x.mean() - shape) < 4 * np.sqrt(shape / n) and abs(x.var() - shape) < 0.02 * shape + 0.01


def test_smoother_reproduces_quadratics_and_follows_the_restatement(L):
    """local quadratic fits reproduce a quadratic exactly, and so does the cubic blending of exact values and slopes; on a
    noisy curve the C++ smoother and the NumPy / SciPy one agree"""
    grid = np.linspace(0, 8, 200)
    fine = np.linspace(0, 8, 1000)
    f = P.loess_interpolate(grid, 3.0 - 1.5 * grid + 0.4 * grid ** 2)
    assert np.max(np.abs(f(fine) - (3.0 - 1.5 * fine + 0.4 * fine ** 2))) < 1e-9
    rng = np.random.default_rng(5)
    resid = np.log(rng.chisquare(2, 30000) / 2) + rng.normal(0, 1.0, 30000)
    counts = np.zeros(40)
    L.cd_prior_var_hist(len(resid), resid.ctypes.data, counts.ctypes.data)
    assert np.array_equal(counts, P.hist_density(resid)[1])
    kl, fitted = np.empty(200), np.empty(1000)
    assert L.cd_prior_var_debug_curve(2, counts.ctypes.data, kl.ctypes.data, fitted.ctypes.data) == 0
    ref = P.loess_interpolate(grid, kl)(fine)
    assert np.max(np.abs(fitted - ref)) < 1e-9 * max(1.0, np.max(np.abs(kl)))
    # 33 vertices: 31 median cuts of the 200 grid points down to cells of 6 or 7, and the two ends of the box
    assert len(P.loess_interpolate(grid, kl).x) == 33
    # the blended surface stays close to the direct one (a local quadratic fit at every point of the fine grid)
    direct = np.empty(len(fine))
    for k, z in enumerate(fine):
        d = np.abs(grid - z)
        h = np.sort(d)[39]
        keep = d < h
        w = np.sqrt((1 - (d[keep] / h) ** 3) ** 3)
        A = np.vander(grid[keep] - z, 3, increasing=True) * w[:, None]
        direct[k] = np.linalg.lstsq(A, kl[keep] * w, rcond=None)[0][0]
    assert np.max(np.abs(fitted - direct)) < 0.02 * (np.max(kl) - np.min(kl))
    assert abs(fine[np.argmin(fitted)] - fine[np.argmin(direct)]) <= 0.1


@pytest.mark.parametrize("df", [2, 3])
def test_rule_follows_the_restatement_and_recovers_the_variance(L, df):
    """df = 2: the final fit of a 2-vs-2 run (S = 4, p = 2); df = 3: its intercept-only theta-grid fits (p = 1)"""
    rng = np.random.default_rng(100 + df)
    for v in (0.5, 1.2, 2.5):
        resid = np.log(rng.chisquare(df, 40000) / df) + rng.normal(0, np.sqrt(v), 40000)
        resid[:50] = 25.0                                                # outside (-10, 10): ignored by the rule
        got = L.cd_prior_var_small_df(df, len(resid), resid.ctypes.data)
        pv, kl, fitted, counts = P.prior_var_small_df(df, resid, return_curves=True)
        kl_c, fit_c = np.empty(200), np.empty(1000)
        L.cd_prior_var_debug_curve(df, counts.ctypes.data, kl_c.ctypes.data, fit_c.ctypes.data)
        assert np.max(np.abs(kl_c - kl)) < 1e-12 * max(1.0, np.max(np.abs(kl)))        # same simulated histograms, bin for bin
        assert got == pv
        assert abs(got - v) < 0.25 + 0.1 * v, (df, v, got)               # the rule estimates the variance it was built to find
    # the floor, and the additivity of the histogram over shards
    tight = np.log(rng.chisquare(df, 20000) / df)
    assert L.cd_prior_var_small_df(df, len(tight), tight.ctypes.data) == 0.25
    a, b = np.zeros(40), np.zeros(40)
    L.cd_prior_var_hist(8000, tight[:8000].ctypes.data, a.ctypes.data)
    L.cd_prior_var_hist(12000, tight[8000:].ctypes.data, b.ctypes.data)
    tot = a + b
    assert L.cd_prior_var_from_hist(df, tot.ctypes.data) == 0.25
    assert np.isnan(L.cd_prior_var_from_hist(5, tot.ctypes.data)) and np.isnan(L.cd_prior_var_from_hist(df, np.zeros(40).ctypes.data))
