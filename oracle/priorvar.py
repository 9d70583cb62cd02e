"""Test infrastructure: a second, independent restatement of DESeq2's small-df dispersion prior variance
(estimateDispersionsPriorVar, the Monte-Carlo branch for 1 <= S - p <= 3) -- the checker of
chicdiff_b200/csrc/priorvar.cpp.  Nothing here is shared with that file: the Mersenne-Twister stream comes from NumPy's
own MT19937 bit generator (seeded with R's scrambled state), the normal quantile from SciPy, the local regressions of the
smoother from numpy.linalg.lstsq and the blending from scipy.interpolate.CubicHermiteSpline.  The samplers (exp_rand,
rgamma) are transcribed a second time from Ahrens & Dieter (1972, 1974, 1982) as R's nmath states them.

"Parity unpinned" like the C++ side: the uniform / normal / exponential streams reproduce the values R prints
(tests/test_priorvar.py); the gamma sampler and loess() could not be run against R here.
"""
import numpy as np
from scipy import special
from scipy.interpolate import CubicHermiteSpline

BREAKS = np.arange(-20, 21) / 2.0


class RStream:
    """unif_rand() after set.seed(seed), kind = Mersenne-Twister"""

    def __init__(self, seed):
        s = np.uint64(seed)
        m = np.uint64(0xFFFFFFFF)
        for _ in range(50):
            s = (np.uint64(69069) * s + np.uint64(1)) & m
        key = np.empty(625, np.uint64)
        for j in range(625):
            s = (np.uint64(69069) * s + np.uint64(1)) & m
            key[j] = s
        bg = np.random.MT19937()
        st = bg.state
        st["state"]["key"] = key[1:].astype(np.uint32)                   # i_seed[0] is the position word, set to 624
        st["state"]["pos"] = 624
        bg.state = st
        self.bg = bg
        self.buf = np.empty(0)
        self.i = 0

    def unif(self):
        if self.i >= len(self.buf):
            raw = self.bg.random_raw(1 << 16).astype(np.float64) * 2.3283064365386963e-10
            half = 0.5 * 2.328306437080797e-10
            raw[raw <= 0.0] = half
            raw[(1.0 - raw) <= 0.0] = 1.0 - half
            self.buf, self.i = raw, 0
        v = self.buf[self.i]
        self.i += 1
        return v


def norm_rand(g):
    big = 134217728.0
    u = g.unif()
    u = float(int(big * u)) + g.unif()
    return float(special.ndtri(u / big))


_Q = np.cumsum([np.log(2.0) ** k / special.factorial(k) for k in range(1, 17)])
_Q[15] = 1.0


def exp_rand(g):
    a = 0.0
    u = g.unif()
    while u <= 0.0 or u >= 1.0:
        u = g.unif()
    while True:
        u += u
        if u > 1.0:
            break
        a += _Q[0]
    u -= 1.0
    if u <= _Q[0]:
        return a + u
    i = 0
    ustar = g.unif()
    umin = ustar
    while True:
        ustar = g.unif()
        if umin > ustar:
            umin = ustar
        i += 1
        if not (u > _Q[i]):
            break
    return a + umin * _Q[0]


_QC = (0.04166669, 0.02083148, 0.00801191, 0.00144121, -7.388e-5, 2.4511e-4, 2.424e-4)
_AC = (0.3333333, -0.250003, 0.2000062, -0.1662921, 0.1423657, -0.1367177, 0.1233795)


def _horner_desc(c, x):
    """((c[6] x + c[5]) x + ... + c[0]) x"""
    r = c[6]
    for k in (5, 4, 3, 2, 1, 0):
        r = r * x + c[k]
    return r * x


def rgamma(g, a, scale=1.0):
    """one draw (no caching: the GD set-up quantities are recomputed; they are pure functions of a)"""
    if a < 1.0:
        e = 1.0 + 0.36787944117144233 * a
        while True:
            p = e * g.unif()
            if p >= 1.0:
                x = -np.log((e - p) / a)
                if exp_rand(g) >= (1.0 - a) * np.log(x):
                    break
            else:
                x = np.exp(np.log(p) / a)
                if exp_rand(g) >= x:
                    break
        return scale * x
    s2 = a - 0.5
    s = np.sqrt(s2)
    d = 5.656854 - s * 12.0
    t = norm_rand(g)
    x = s + 0.5 * t
    ret = x * x
    if t >= 0.0:
        return scale * ret
    u = g.unif()
    if d * u <= t * t * t:
        return scale * ret
    q0 = _horner_desc(_QC, 1.0 / a)
    if a <= 3.686:
        b, si, c = 0.463 + s + 0.178 * s2, 1.235, 0.195 / s - 0.079 + 0.16 * s
    elif a <= 13.022:
        b, si, c = 1.654 + 0.0076 * s2, 1.68 / s + 0.275, 0.062 / s + 0.024
    else:
        b, si, c = 1.77, 0.75, 0.1515 / s

    def quotient(t):
        v = t / (s + s)
        if abs(v) <= 0.25:
            return q0 + 0.5 * t * t * _horner_desc(_AC, v)
        return q0 - s * t + 0.25 * t * t + (s2 + s2) * np.log(1.0 + v)
    if x > 0.0 and np.log(1.0 - u) <= quotient(t):
        return scale * ret
    while True:
        e = exp_rand(g)
        u = g.unif()
        u = u + u - 1.0
        t = b - si * e if u < 0.0 else b + si * e
        if t >= -0.71874483771719:
            q = quotient(t)
            if q > 0.0:
                w = np.expm1(q)
                if c * abs(u) <= w * np.exp(e - 0.5 * t * t):
                    break
    x = s + 0.5 * t
    return scale * x * x


def hist_density(x):
    """hist(x[x > -10 & x < 10], breaks = -20:20/2)$density and the counts"""
    x = np.asarray(x, dtype=np.float64)
    x = x[(x > BREAKS[0]) & (x < BREAKS[-1])]
    fuzz = np.full(len(BREAKS), 1e-7 * np.median(np.diff(BREAKS)))
    fuzz[0] = -fuzz[0]
    fb = BREAKS + fuzz
    idx = np.searchsorted(fb, x, side="left") - 1                        # right-closed bins (fb[k], fb[k+1]]
    counts = np.bincount(idx, minlength=len(BREAKS) - 1).astype(np.float64)[:len(BREAKS) - 1]
    return counts / (counts.sum() * np.diff(BREAKS)), counts


_SIM = {}


def sim_densities(df, n_draws=10000, n_grid=200):
    if df in _SIM:
        return _SIM[df]
    g = RStream(2)
    out = np.empty((n_grid, len(BREAKS) - 1))
    grid = np.linspace(0.0, 8.0, n_grid)
    for k, v in enumerate(grid):
        chi = np.log(np.array([rgamma(g, df / 2.0, 2.0) for _ in range(n_draws)]))
        sd = np.sqrt(v)
        z = np.zeros(n_draws) if sd == 0.0 else sd * np.array([norm_rand(g) for _ in range(n_draws)])
        out[k] = hist_density(chi + z - np.log(df))[0]
    _SIM[df] = out
    return out


def loess_interpolate(x, y, span=0.2, cell=0.2, degree=2):
    """loess(y ~ x, span, degree, surface = "interpolate"): returns a callable"""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(x)
    q = min(n, int(np.floor(n * span + 1e-5)))
    fc = int(np.floor(n * span * cell))
    cuts = []

    def split(lo, hi):                      # 1-based inclusive
        if hi - lo + 1 <= fc:
            return
        m = (lo + hi) // 2
        while m > 1 and x[m - 2] == x[m - 1]:
            m -= 1
        cuts.append((x[m - 1] + x[m]) / 2.0)
        split(lo, m)
        split(m + 1, hi)
    split(1, n)
    mu = 0.005 * max(x[-1] - x[0], 1e-10 * max(abs(x[0]), abs(x[-1])) + 1e-30)
    vx = np.array(sorted([x[0] - mu, x[-1] + mu] + cuts))
    val, slope = np.empty(len(vx)), np.empty(len(vx))
    for k, v in enumerate(vx):
        d = np.abs(x - v)
        h = np.sort(d)[q - 1]
        keep = d < h
        w = (1.0 - (d[keep] / h) ** 3) ** 3
        u = x[keep] - v
        A = np.vander(u, degree + 1, increasing=True) * np.sqrt(w)[:, None]
        coef = np.linalg.lstsq(A, y[keep] * np.sqrt(w), rcond=None)[0]
        val[k], slope[k] = coef[0], coef[1]
    return CubicHermiteSpline(vx, val, slope)


def prior_var_small_df(df, resid, return_curves=False):
    obs, counts = hist_density(resid)
    sim = sim_densities(df)
    grid = np.linspace(0.0, 8.0, 200)
    kl = np.empty(200)
    for k in range(200):
        z = np.concatenate([obs, sim[k]])
        small = z[z > 0].min()
        kl[k] = np.sum(obs * (np.log(obs + small) - np.log(sim[k] + small)))
    fit = loess_interpolate(grid, kl)
    fine = np.linspace(0.0, 8.0, 1000)
    fitted = fit(fine)
    pv = max(fine[int(np.argmin(fitted))], 0.25)
    return (pv, kl, fitted, counts) if return_curves else pv
