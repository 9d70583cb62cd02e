"""Accuracy of the device math of chicdiff_b200/csrc/common.cuh, run on the CPU.

The line-search and GLM kernels do not call libdevice's lgamma/digamma/log: they use branch-free versions (shift-10
gamma rationals, an fdlibm-style log, a Newton reciprocal) so that the lanes of a warp never diverge.  The header is
compiled here for the host (tests/device_math provides a stand-in for <cuda_runtime.h>; the MUFU reciprocal seed is the
only line that differs) and checked against SciPy and the oracle over argument ranges far wider than any synthetic data
set produces: dispersions from the 1e-8 floor to 1e2, counts up to 1e5, means up to 1e6."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy import special, stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("device_math") / "libdm.so")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
           "-I" + os.path.join(ROOT, "tests", "device_math"), "-I" + os.path.join(ROOT, "chicdiff_b200", "csrc"),
           "-o", out, os.path.join(ROOT, "tests", "device_math", "shim.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = C.CDLL(out)
    L.dm_vec.argtypes = [C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dm_dnbinom_vec.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.dm_chol_logdet2.restype = C.c_double
    L.dm_chol_logdet2.argtypes = [C.c_double] * 3
    return L


def _vec(L, what, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    o, o2 = np.empty_like(x), np.empty_like(x)
    L.dm_vec(what, len(x), x.ctypes.data, o.ctypes.data, o2.ctypes.data)
    return o, o2


def _loguniform(rng, lo, hi, n):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), n))


def test_log_and_reciprocal_are_within_one_ulp(dm):
    rng = np.random.default_rng(0)
    x = np.concatenate([_loguniform(rng, 1e-300, 1e300, 200000), 1 + rng.uniform(-1e-3, 1e-3, 50000), rng.uniform(0.5, 2, 50000)])
    a, _ = _vec(dm, 0, x)
    b = np.log(x)
    assert np.max(np.abs(a - b) / np.spacing(np.abs(b) + 1e-300)) <= 1.5
    a, _ = _vec(dm, 1, x)
    assert np.max(np.abs(a - 1 / x) / np.spacing(1 / x)) <= 1.0


def test_shift10_lgamma_and_digamma(dm):
    rng = np.random.default_rng(1)
    x = _loguniform(rng, 1e-6, 1e8, 300000)
    lg, dg = _vec(dm, 2, x)
    ref = special.gammaln(x) - 0.5 * np.log(2 * np.pi)
    assert np.max(np.abs(lg - ref) / np.maximum(1, np.abs(ref))) < 5e-14
    refd = special.digamma(x)
    assert np.max(np.abs(dg - refd) / np.maximum(1, np.abs(refd))) < 1e-14
    lc, _ = _vec(dm, 3, x)
    assert np.array_equal(lc, lg)                       # the lgamma-only variant is the same arithmetic
    big = _loguniform(rng, 1e8, 1e15, 100000)
    lg, dg = _vec(dm, 2, big)
    assert np.max(np.abs(lg - (special.gammaln(big) - 0.5 * np.log(2 * np.pi))) / special.gammaln(big)) < 2e-15
    assert np.max(np.abs(dg - special.digamma(big)) / special.digamma(big)) < 2e-15
    tg, _ = _vec(dm, 4, x)
    assert np.max(np.abs(tg - special.polygamma(1, x)) / special.polygamma(1, x)) < 1e-14


def test_lgamma_difference_of_the_nb_likelihood(dm):
    """lgamma(y + 1/alpha) - lgamma(1/alpha), the term the Cox-Reid posterior is built from: as accurate as a
    difference of two doubles of that size can be (the reference forms the same difference in double)."""
    rng = np.random.default_rng(2)
    y = rng.integers(0, 2000, 300000).astype(float)
    r = _loguniform(rng, 0.05, 1e8, 300000)
    a1, _ = _vec(dm, 3, y + r)
    a2, _ = _vec(dm, 3, r)
    ref = special.gammaln(y + r) - special.gammaln(r)
    size_of_terms = np.maximum(1.0, np.abs(special.gammaln(y + r)))
    assert np.max(np.abs((a1 - a2) - ref) / size_of_terms) < 5e-14        # two values of ~1e-14 absolute accuracy each


def test_dnbinom_mu_log_follows_the_oracle_and_scipy(dm):
    from oracle import oracle as O
    Lo = O.lib()
    Lo.orc_dnbinom_mu_log.restype = C.c_double
    Lo.orc_dnbinom_mu_log.argtypes = [C.c_double] * 3
    rng = np.random.default_rng(3)
    n = 100000
    y = np.concatenate([rng.integers(0, 50, n // 2), rng.integers(0, 100000, n // 2)]).astype(float)
    size = _loguniform(rng, 1e-2, 1e8, n)
    mu = _loguniform(rng, 0.5, 1e6, n)
    out = np.empty(n)
    dm.dm_dnbinom_vec(n, y.ctypes.data, size.ctypes.data, mu.ctypes.data, out.ctypes.data)
    m = 20000
    ro = np.array([Lo.orc_dnbinom_mu_log(a, b, c) for a, b, c in zip(y[:m], size[:m], mu[:m])])
    assert np.max(np.abs(out[:m] - ro) / np.maximum(1, np.abs(ro))) < 1e-12
    sp = stats.nbinom.logpmf(y, size, size / (size + mu))
    ok = np.isfinite(sp) & (size < 1e5)                  # SciPy's own lgamma differences degrade beyond that
    assert np.max(np.abs(out[ok] - sp[ok]) / np.maximum(1, np.abs(sp[ok]))) < 1e-9


def test_packed_cholesky_logdet(dm):
    rng = np.random.default_rng(4)
    for _ in range(200):
        A = rng.normal(size=(2, 2))
        A = A @ A.T + 1e-3 * np.eye(2)
        got = dm.dm_chol_logdet2(A[0, 0], A[1, 0], A[1, 1])
        assert abs(got - np.linalg.slogdet(A)[1]) < 1e-10 * max(1.0, abs(got))
    assert np.isnan(dm.dm_chol_logdet2(1.0, 2.0, 1.0))   # not positive definite


def test_fused_posterior_and_derivative_follow_the_oracle(dm):
    """eval_post (posterior.cuh), the function every trip of the dispersion line search evaluates on the GPU, against
    the oracle's log_posterior / dlog_posterior (DESeq2.cpp) and a central difference, over dispersions from the 1e-8
    floor to 10, for the intercept-only, two-level and batch designs, with and without the MAP prior."""
    from oracle import oracle as O
    Lo = O.lib()
    dm.dm_eval_post.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                                C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    designs = {
        1: np.ones((6, 1)),
        2: np.column_stack([np.ones(6), [0, 0, 0, 1, 1, 1]]).astype(float),
        3: np.column_stack([np.ones(8), [0, 1, 0, 1, 0, 1, 0, 1], [0, 0, 0, 0, 1, 1, 1, 1]]).astype(float),
    }
    worst_lp = worst_dlp = 0.0
    for p, X in designs.items():
        S = X.shape[0]
        X = np.ascontiguousarray(X)
        for _ in range(400):
            mu = np.maximum(np.exp(rng.uniform(np.log(0.5), np.log(3000.0), S)), 0.5)
            alpha_true = np.exp(rng.uniform(np.log(1e-3), np.log(2.0)))
            y = rng.negative_binomial(1.0 / alpha_true, 1.0 / (1.0 + mu * alpha_true)).astype(float)
            a = rng.uniform(np.log(1e-8), np.log(10.0))
            use_prior = int(rng.random() < 0.5)
            pm, ps = rng.normal(-2.0, 1.0), rng.uniform(0.25, 2.0)
            lp, dlp = C.c_double(), C.c_double()
            dm.dm_eval_post(S, p, X.ctypes.data, y.ctypes.data, mu.ctypes.data, a, pm, ps, use_prior, C.byref(lp), C.byref(dlp))
            lo = Lo.orc_log_posterior(a, S, p, X.ctypes.data, y.ctypes.data, mu.ctypes.data, pm, ps, use_prior, 1)
            dlo = Lo.orc_dlog_posterior(a, S, p, X.ctypes.data, y.ctypes.data, mu.ctypes.data, pm, ps, use_prior, 1)
            # The posterior is a sum of terms of size ~ lgamma(y + 1/alpha): that size sets the attainable accuracy
            scale = 1.0 + np.sum(np.abs(special.gammaln(y + np.exp(-a)))) * 1e-3
            worst_lp = max(worst_lp, abs(lp.value - lo) / max(abs(lo), scale))
            worst_dlp = max(worst_dlp, abs(dlp.value - dlo) / max(abs(dlo), 1e-3 * scale, 1.0))
    assert worst_lp < 1e-11, worst_lp
    assert worst_dlp < 1e-8, worst_dlp
    # the derivative really is the derivative of the posterior (moderate dispersion, where differences are meaningful)
    X = designs[2]
    mu = np.array([20.0, 35.0, 18.0, 60.0, 44.0, 52.0]); y = np.array([25.0, 30.0, 11.0, 71.0, 38.0, 60.0])
    def f(a):
        lp, dlp = C.c_double(), C.c_double()
        dm.dm_eval_post(6, 2, X.ctypes.data, y.ctypes.data, mu.ctypes.data, a, 0.0, 1.0, 0, C.byref(lp), C.byref(dlp))
        return lp.value, dlp.value
    for a in (-4.0, -2.5, -1.0, 0.5):
        h = 1e-5
        num = (f(a + h)[0] - f(a - h)[0]) / (2 * h)
        assert abs(num - f(a)[1]) < 1e-6 * max(1.0, abs(num))


@pytest.mark.parametrize("variant", [1, 2])
def test_posterior_formulations_are_the_same_function(dm, variant):
    """eval_post (posterior.cuh: one logarithm per pair of replicates for the gamma rationals, closed-form 1x1 / 2x2
    Cox-Reid term; variant 2 = with the table-assisted logarithm, which is what the kernel runs) against the first
    formulation eval_post_ref on the host: same value and derivative up to the rounding of the lgamma-sized terms."""
    f = dm.dm_eval_post_variant
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int,
                  C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(7)
    designs = [np.ones((6, 1)), np.ones((5, 1)), np.column_stack([np.ones(6), [0, 0, 0, 1, 1, 1]]).astype(float),
               np.column_stack([np.ones(8), [0, 1, 0, 1, 0, 1, 0, 1], [0, 0, 0, 0, 1, 1, 1, 1]]).astype(float)]
    for X in designs:
        S, p = X.shape
        X = np.ascontiguousarray(X)
        worst_lp = worst_dlp = 0.0
        for _ in range(1500):
            mu = np.exp(rng.uniform(np.log(0.5), np.log(1e5), S))
            alpha_true = np.exp(rng.uniform(np.log(1e-3), np.log(2.0)))
            y = rng.negative_binomial(1.0 / alpha_true, 1.0 / (1.0 + mu * alpha_true)).astype(float)
            if rng.random() < 0.2:
                y[rng.integers(0, S)] = 0.0
            a = rng.uniform(np.log(1e-8), np.log(10.0))
            use_prior = int(rng.random() < 0.5)
            out = []
            for v in (0, variant):
                lp, dlp = C.c_double(), C.c_double()
                f(v, S, p, X.ctypes.data, y.ctypes.data, mu.ctypes.data, a, -2.0, 0.7, use_prior, C.byref(lp), C.byref(dlp))
                out.append((lp.value, dlp.value))
            scale = 1.0 + np.sum(np.abs(special.gammaln(y + np.exp(-a))))
            worst_lp = max(worst_lp, abs(out[0][0] - out[1][0]) / scale)
            # the derivative multiplies a sum of S digamma-sized terms by 1/alpha: their last-bit differences (the two
            # logarithms differ by an ulp) come back magnified by that factor
            noise = 4e-15 * S * np.exp(-a) * (abs(special.digamma(np.exp(-a))) + 1.0)
            worst_dlp = max(worst_dlp, max(0.0, abs(out[0][1] - out[1][1]) - noise) / max(abs(out[0][1]), 1.0))
        assert worst_lp < 5e-15, (S, p, worst_lp)
        assert worst_dlp < 1e-12, (S, p, worst_dlp)


def test_posterior_gamma_rational_overflow_path(dm):
    """counts in the hundreds of millions at alpha near its cap: the product of the replicates' gamma rationals leaves the
    double range and eval_post takes one logarithm per replicate instead (posterior.cuh); same function as before"""
    f = dm.dm_eval_post_variant
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int,
                  C.c_void_p, C.c_void_p]
    X = np.ascontiguousarray(np.column_stack([np.ones(6), [0, 0, 0, 1, 1, 1]]).astype(float))
    y = np.array([2.0e9, 1.5e9, 1.9e9, 2.1e9, 1.0e9, 7.0e8])
    mu = y * np.array([1.1, 0.9, 1.0, 1.2, 0.8, 1.0])
    for a in (np.log(10.0), np.log(3.0), np.log(0.5)):
        with np.errstate(over="ignore"):
            assert np.prod(((y + np.exp(-a)) / np.exp(-a)) ** 10.0) == np.inf   # the single product does overflow
        out = []
        for v in (0, 2):
            lp, dlp = C.c_double(), C.c_double()
            f(v, 6, 2, X.ctypes.data, y.ctypes.data, mu.ctypes.data, a, -2.0, 0.7, 1, C.byref(lp), C.byref(dlp))
            out.append((lp.value, dlp.value))
        assert np.isfinite(out[1][0]) and np.isfinite(out[1][1])
        assert abs(out[0][0] - out[1][0]) <= 5e-15 * np.sum(np.abs(special.gammaln(y + np.exp(-a))))
        assert abs(out[0][1] - out[1][1]) <= 1e-9 * max(abs(out[0][1]), 1.0)
    # 16 replicates, three design columns, alpha up to 16: products from far inside the double range to far outside it,
    # including those that end within a few binades of the largest double
    rng = np.random.default_rng(11)
    S = 16
    X = np.ascontiguousarray(np.column_stack([np.ones(S), np.arange(S) % 2, np.arange(S) >= 8]).astype(float))
    n_over = 0
    for _ in range(3000):
        mu = np.exp(rng.uniform(np.log(0.5), np.log(1e5), S))
        at = np.exp(rng.uniform(np.log(1e-3), np.log(2.0)))
        y = rng.negative_binomial(1.0 / at, 1.0 / (1.0 + mu * at)).astype(float)
        a = rng.uniform(np.log(1e-2), np.log(16.0))
        out = []
        for v in (0, 2):
            lp, dlp = C.c_double(), C.c_double()
            f(v, S, 3, X.ctypes.data, y.ctypes.data, mu.ctypes.data, a, -2.0, 0.7, 0, C.byref(lp), C.byref(dlp))
            out.append((lp.value, dlp.value))
        with np.errstate(over="ignore"):
            n_over += int(np.prod(((y + np.exp(-a)) / np.exp(-a)) ** 10.0) > 2.0 ** 1000)
        assert abs(out[0][0] - out[1][0]) <= 5e-15 * (1.0 + np.sum(np.abs(special.gammaln(y + np.exp(-a)))))
    assert n_over > 100


def test_table_log(dm):
    """log_pos_v2 (common.cuh): within 1.5 ulp for every argument >= 1 (all the kernel's logarithms except the determinant's)
    and within 3e-16 of max(1, |log x|) below 1, where the table entry and k ln2 cancel near x = 1."""
    rng = np.random.default_rng(8)
    for x in (np.exp(rng.uniform(0, np.log(1e300), 200000)), rng.uniform(1, 4, 200000),
              1 + np.exp(rng.uniform(np.log(1e-16), np.log(1e-1), 200000)),
              np.array([1.0, 2.0, 4.0, 1 + 2.0 ** -7, 2 - 2.0 ** -7, 2 - 2.0 ** -6, np.nextafter(2, 1), np.nextafter(1, 2)])):
        a, _ = _vec(dm, 10, x)
        ref = np.log(x.astype(np.longdouble))
        err = np.abs(a.astype(np.longdouble) - ref)
        assert float(np.max(err / np.spacing(np.abs(ref.astype(np.float64)) + 1e-300))) <= 1.5
    below = np.concatenate([rng.uniform(0.5, 1, 200000), np.exp(rng.uniform(np.log(1e-300), np.log(0.5), 100000))])
    a, _ = _vec(dm, 10, below)
    ref = np.log(below.astype(np.longdouble))
    err = np.abs(a.astype(np.longdouble) - ref)
    assert float(np.max(err / np.maximum(np.abs(ref), 1.0))) < 3e-16
    assert _vec(dm, 10, np.array([1.0]))[0][0] == 0.0


def test_exp_mid(dm):
    """exp_mid (common.cuh), the exponential of the line search: within 2 ulp over the range the search keeps log alpha in
    and well beyond it"""
    rng = np.random.default_rng(9)
    x = np.concatenate([rng.uniform(-35, 15, 400000), rng.uniform(-700, 700, 100000), rng.uniform(-1e-3, 1e-3, 50000),
                        np.array([0.0, -30.0, 10.0, np.log(2) / 2, -np.log(2) / 2, 1e-300])])
    a, _ = _vec(dm, 11, x)
    ref = np.exp(x.astype(np.longdouble))
    err = np.abs(a.astype(np.longdouble) - ref) / np.spacing(ref.astype(np.float64))
    assert float(np.max(err)) <= 2.0, float(np.max(err))
    assert _vec(dm, 11, np.array([0.0]))[0][0] == 1.0
