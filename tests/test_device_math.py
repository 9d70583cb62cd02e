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
