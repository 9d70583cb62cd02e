"""Host-side model of what the ranks of a sharded run exchange in the global steps (test infrastructure).

The product runs these protocols inside CUDA kernels (select.cu: sel_fused_kernel, offsets.cu: trend_fit_kernel), with
the all-reduces going through NVLink peer-memory mailboxes.  This module restates the same protocols in NumPy with the
all-reduce passed in as a function, so that tests/test_parallel_cpu.py can play them over gloo with world_size 2 on a
box without GPUs and check them against the unsharded oracle:

  * exact medians: six rounds of 2048-bin histograms of the order-preserving 64-bit image of the doubles; only the
    counters are summed across ranks (plus one count, one count <= v1 and one minimum);
  * parametricDispersionFit: every pass sums 8 numbers over the local regions, the 8 sums are all-reduced, and every
    rank then takes the same branch of glm.fit's control flow.
"""
import numpy as np

BITS = (11, 11, 11, 11, 11, 9)
SHIFTS = (53, 42, 31, 20, 9, 0)


def key_of(x):
    b = np.asarray(x, np.float64).view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    return np.where(neg, ~b, b | np.uint64(1 << 63))


def value_of(k):
    k = np.uint64(k)
    b = (k & np.uint64((1 << 63) - 1)) if (k >> np.uint64(63)) else ~k
    return np.array([b], np.uint64).view(np.float64)[0]


def distributed_median(local, allreduce_sum, allreduce_min, scale=1.0):
    """R median() of the finite entries of the union of every rank's `local`; NaN if there are none."""
    v = np.asarray(local, np.float64)
    keys = key_of(v[np.isfinite(v)])
    m = int(allreduce_sum(np.array([len(keys)], np.int64))[0])
    if m == 0:
        return float("nan")
    k = (m - 1) // 2
    prefix = 0
    for bits, shift in zip(BITS, SHIFTS):
        hi_shift = shift + bits
        match = keys if hi_shift >= 64 else keys[(keys >> np.uint64(hi_shift)) == np.uint64(prefix)]
        digit = ((match >> np.uint64(shift)) & np.uint64((1 << bits) - 1)).astype(np.int64)
        hist = allreduce_sum(np.bincount(digit, minlength=2048).astype(np.int64))          # the only exchange of the round
        cum = 0
        b = 0
        while b < 2047 and cum + hist[b] <= k:
            cum += int(hist[b]); b += 1
        prefix = (prefix << bits) | b
        k -= cum
    v1 = np.uint64(prefix)
    le = int(allreduce_sum(np.array([int((keys <= v1).sum())], np.int64))[0])
    gt = keys[keys > v1]
    mg = int(allreduce_min(np.array([gt.min() if len(gt) else np.uint64((1 << 64) - 1)], np.uint64))[0])
    x1 = value_of(v1)
    if m & 1:
        med = x1
    else:
        x2 = x1 if le > m // 2 else value_of(np.uint64(mg))
        med = 0.5 * (x1 + x2)
    return scale * med


def trend_pass(baseMean, disp, keep_state, refresh, c, b):
    """the 8 local sums of one pass (offsets.cu trend_row): s00 s01 s11 t0 t1 deviance #invalid #rows"""
    usable = disp > 100 * 1e-8
    x = 1.0 / baseMean
    if refresh:
        with np.errstate(invalid="ignore", divide="ignore"):
            r = disp / (c[0] + c[1] * x)
        keep_state[:] = usable & (r > 1e-4) & (r < 15.0)
    keep = keep_state
    mu = b[0] + b[1] * x[keep]
    d = disp[keep]
    v = np.zeros(8)
    v[7] = keep.sum()
    bad = ~(mu > 0) | ~np.isfinite(mu)
    v[6] = bad.sum()
    mu, d, xx = mu[~bad], d[~bad], x[keep][~bad]
    w = 1.0 / (mu * mu)
    v[0], v[1], v[2] = w.sum(), (w * xx).sum(), (w * xx * xx).sum()
    v[3], v[4] = (w * d).sum(), (w * xx * d).sum()
    t = d / mu
    v[5] = (-2.0 * (np.log(t) - (t - 1.0))).sum()
    return v


def distributed_trend_fit(baseMean, disp, allreduce_sum):
    """parametricDispersionFit over the union of every rank's regions -> (a0, a1, status, outer iterations, passes);
    the control flow of offsets.cu TrendFit (= R's glm.fit with family Gamma(link = "identity"))."""
    baseMean = np.asarray(baseMean, np.float64)
    disp = np.asarray(disp, np.float64)
    keep = np.zeros(len(disp), bool)
    c = [0.1, 1.0]
    it_outer = 0
    passes = 0

    def do_pass(refresh, b):
        nonlocal passes
        passes += 1
        return allreduce_sum(trend_pass(baseMean, disp, keep, refresh, c, b))

    while True:
        b = list(c)
        ob = list(c)
        v = do_pass(True, b)
        if v[7] < 2:
            return c[0], c[1], 1, it_outer + 1, passes
        if v[6] > 0:
            return c[0], c[1], 2, it_outer + 1, passes
        devold = v[5]
        conv = False
        for _ in range(25):
            det = v[0] * v[2] - v[1] * v[1]
            nb = [(v[2] * v[3] - v[1] * v[4]) / det, (v[0] * v[4] - v[1] * v[3]) / det]
            halv = 0
            while True:
                w = do_pass(False, nb)
                if w[6] == 0 and np.isfinite(w[5]):
                    break
                halv += 1
                if halv > 25:
                    return c[0], c[1], 3, it_outer + 1, passes
                nb = [0.5 * (nb[0] + ob[0]), 0.5 * (nb[1] + ob[1])]
            b = nb
            v = w
            dev = w[5]
            if abs(dev - devold) / (abs(dev) + 0.1) < 1e-8:
                conv = True
                break
            devold = dev
            ob = list(b)
        oc = list(c)
        c = list(b)
        if not (c[0] > 0 and c[1] > 0):
            return c[0], c[1], 4, it_outer + 1, passes
        if (np.log(c[0] / oc[0]) ** 2 + np.log(c[1] / oc[1]) ** 2 < 1e-6) and conv:
            return c[0], c[1], 0, it_outer + 1, passes
        it_outer += 1
        if it_outer > 10:
            return c[0], c[1], 5, it_outer + 1, passes
