"""Second, independent restatement of the DESeq2 numerics that DESeq2Wrap runs (chicdiff.R:1573-1574, 1602-1603,
1643-1644, 1673-1674), written from SURVEY.md Appendix A with NumPy / SciPy building blocks only.

TEST INFRASTRUCTURE (like everything under oracle/): nothing in chicdiff_b200/ may import it.  Its purpose is to make
transcription errors of the C restatement (chicdiff_oracle.c) visible as disagreement, since no R session exists to
pin either of them ("parity unpinned").  It shares no code with the C file and deliberately takes different routes
to the same numbers:

    C restatement                         this file
    -----------------------------------   ---------------------------------------------------
    own lgamma / digamma / trigamma       scipy.special
    saddle-point dnbinom_mu (Rmath port)  scipy.stats.nbinom.logpmf
    packed Cholesky normal equations      numpy.linalg.qr on the ridge-augmented design (as DESeq2's fitBeta does)
    closed-form 2x2 weighted LS (trend)   numpy.linalg.lstsq on the weighted design
    hand-written medians                  numpy.median

Plain Python loops over regions: meant for a few hundred regions."""
import math

import numpy as np
from scipy import special, stats

MIN_DISP = 1e-8
MINMU = 0.5
KAPPA0 = 1.0
DISP_TOL = 1e-6
MAXIT = 100
BETA_TOL = 1e-8
EPSILON = 1e-4
LARGE = 30.0


# ---- A.4: Cox-Reid adjusted profile log-posterior of log(alpha) and its derivative ----------------------------
def log_posterior(a, y, mu, X, prior_mean, prior_sigmasq, use_prior):
    alpha = math.exp(a)
    w = 1.0 / (1.0 / mu + alpha)
    B = X.T @ (w[:, None] * X)
    cr = -0.5 * np.linalg.slogdet(B)[1]
    r = 1.0 / alpha
    ll = np.sum(special.gammaln(y + r) - special.gammaln(r) - y * np.log(mu + r) - r * np.log1p(mu * alpha))
    pr = -0.5 * (a - prior_mean) ** 2 / prior_sigmasq if use_prior else 0.0
    return ll + pr + cr


def dlog_posterior(a, y, mu, X, prior_mean, prior_sigmasq, use_prior):
    alpha = math.exp(a)
    w = 1.0 / (1.0 / mu + alpha)
    dw = -w * w
    B = X.T @ (w[:, None] * X)
    dB = X.T @ (dw[:, None] * X)
    dcr = -0.5 * np.trace(np.linalg.solve(B, dB))
    r = 1.0 / alpha
    dll = r * r * np.sum(special.digamma(r) + np.log1p(mu * alpha) - mu * alpha / (1.0 + mu * alpha)
                         - special.digamma(y + r) + y / (mu + r))
    dpr = -(a - prior_mean) / prior_sigmasq if use_prior else 0.0
    return (dll + dcr) * alpha + dpr


def fit_disp(y, mu, X, log_alpha, prior_mean, prior_sigmasq, use_prior, min_log_alpha):
    """DESeq2.cpp fitDisp for one row: Armijo backtracking line search on log(alpha)."""
    f = lambda v: log_posterior(v, y, mu, X, prior_mean, prior_sigmasq, use_prior)
    g = lambda v: dlog_posterior(v, y, mu, X, prior_mean, prior_sigmasq, use_prior)
    a = log_alpha
    lp = f(a)
    initial_lp = lp
    dlp = g(a)
    kappa = KAPPA0
    it = it_accept = 0
    for _ in range(MAXIT):
        it += 1
        a_prop = a + kappa * dlp
        if a_prop < -30.0:
            kappa = (-30.0 - a) / dlp
        if a_prop > 10.0:
            kappa = (10.0 - a) / dlp
        theta_kappa = -f(a + kappa * dlp)
        theta_hat_kappa = -lp - kappa * EPSILON * dlp * dlp
        if theta_kappa <= theta_hat_kappa:
            it_accept += 1
            a = a + kappa * dlp
            lpnew = f(a)
            change = lpnew - lp
            if change < DISP_TOL:
                lp = lpnew
                break
            if a < min_log_alpha:
                break
            lp = lpnew
            dlp = g(a)
            kappa = min(kappa * 1.1, KAPPA0)
            if it_accept % 5 == 0:
                kappa = kappa / 2.0
        else:
            kappa = kappa / 2.0
    return a, it, initial_lp, lp


def fit_disp_grid(y, mu, X, max_disp, prior_mean, prior_sigmasq, use_prior, n_grid=20):
    """A.5: coarse grid over log(alpha), then a fine grid one coarse step either side of the best point."""
    f = lambda v: log_posterior(v, y, mu, X, prior_mean, prior_sigmasq, use_prior)
    grid = np.linspace(math.log(MIN_DISP), math.log(max_disp), n_grid)
    vals = np.array([f(v) for v in grid])
    a_hat = grid[int(np.argmax(vals))]
    delta = grid[1] - grid[0]
    fine = np.linspace(a_hat - delta, a_hat + delta, n_grid)
    vals = np.array([f(v) for v in fine])
    return math.exp(fine[int(np.argmax(vals))])


# ---- A.6: parametric trend ---------------------------------------------------------------------------------------
def gamma_identity_glm(x, y, start):
    """R glm.fit(family = Gamma(link = "identity"), start = start) for the model y ~ 1 + x."""
    D = np.column_stack([np.ones_like(x), x])
    beta = np.array(start, float)
    mu = D @ beta
    if not (np.all(np.isfinite(mu)) and np.all(mu > 0)):
        raise FloatingPointError("invalid starting values")
    dev_of = lambda m: -2.0 * np.sum(np.log(y / m) - (y - m) / m)
    devold = dev_of(mu)
    beta_old = beta.copy()
    converged = False
    for _ in range(25):
        sw = 1.0 / mu                                     # sqrt of the working weights 1/mu^2
        beta_new = np.linalg.lstsq(D * sw[:, None], y * sw, rcond=None)[0]
        mu_new = D @ beta_new
        halv = 0
        while not (np.all(mu_new > 0) and np.isfinite(dev_of(mu_new))):
            halv += 1
            if halv > 25:
                raise FloatingPointError("no valid step")
            beta_new = (beta_new + beta_old) / 2.0
            mu_new = D @ beta_new
        beta, mu = beta_new, mu_new
        dev = dev_of(mu)
        if abs(dev - devold) / (abs(dev) + 0.1) < 1e-8:
            converged = True
            break
        devold = dev
        beta_old = beta.copy()
    return beta, converged


def parametric_trend(base_mean, disp):
    use = disp > 100.0 * MIN_DISP
    m, d = base_mean[use], disp[use]
    coefs = np.array([0.1, 1.0])
    it = 0
    while True:
        resid = d / (coefs[0] + coefs[1] / m)
        good = (resid > 1e-4) & (resid < 15.0)
        new, conv = gamma_identity_glm(1.0 / m[good], d[good], coefs)
        old, coefs = coefs, new
        if not np.all(coefs > 0):
            raise FloatingPointError("parametric dispersion fit failed")
        if np.sum(np.log(coefs / old) ** 2) < 1e-6 and conv:
            break
        it += 1
        if it > 10:
            raise FloatingPointError("dispersion fit did not converge")
    return coefs


# ---- A.8: NB GLM --------------------------------------------------------------------------------------------------
def nb_loglik(y, mu, alpha):
    size = 1.0 / alpha
    return np.sum(stats.nbinom.logpmf(y, size, size / (size + mu)))


def fit_beta(y, nf, X, alpha, beta0, lam):
    """DESeq2.cpp fitBeta for one row (natural-log scale), QR of the ridge-augmented weighted design."""
    S, p = X.shape
    beta = beta0.copy()
    mu = np.maximum(nf * np.exp(X @ beta), MINMU)
    ridge = np.diag(np.sqrt(lam))
    dev_old = 0.0
    it = 0
    for t in range(MAXIT):
        it += 1
        w = mu / (1.0 + alpha * mu)
        z = np.log(mu / nf) + (y - mu) / mu
        A = np.vstack([np.sqrt(w)[:, None] * X, ridge])
        rhs = np.concatenate([np.sqrt(w) * z, np.zeros(p)])
        Q, R = np.linalg.qr(A)
        beta = np.linalg.solve(R, Q.T @ rhs)
        if np.any(np.abs(beta) > LARGE):
            it = MAXIT
            break
        mu = np.maximum(nf * np.exp(X @ beta), MINMU)
        dev = -2.0 * nb_loglik(y, mu, alpha)
        conv = abs(dev - dev_old) / (abs(dev) + 0.1)
        if math.isnan(conv):
            it = MAXIT
            break
        if t > 0 and conv < BETA_TOL:
            break
        dev_old = dev
    w = mu / (1.0 + alpha * mu)
    A = X.T @ (w[:, None] * X)
    Ar = np.linalg.inv(A + np.diag(lam))
    sigma = Ar @ A @ Ar
    hat = w * np.einsum("ju,uv,jv->j", X, Ar, X)
    return beta, np.sqrt(np.maximum(np.diag(sigma), 0.0)), it, hat


def trimmed_mean(x, trim):
    """R mean(x, trim = t): sort, drop floor(n t) values from each end."""
    x = np.sort(x)
    lo = int(math.floor(len(x) * trim))
    return np.mean(x[lo:len(x) - lo])


def trimmed_cell_variance(q, cells):
    """robustMethodOfMomentsDisp's variance when some design cell has >= 3 samples: per cell the scaled trimmed mean of
    squared errors around the cell's trimmed mean, maximum over those cells."""
    best = -np.inf
    for c in np.unique(cells):
        idx = np.flatnonzero(cells == c)
        n = len(idx)
        if n < 3:
            continue
        ratio, scale = (1 / 3, 2.04) if n <= 3 else ((1 / 4, 1.86) if n <= 23 else (1 / 8, 1.51))
        cm = trimmed_mean(q[idx], ratio)
        best = max(best, scale * trimmed_mean((q[idx] - cm) ** 2, ratio))
    return best


# ---- the whole estimateDispersions + nbinomWaldTest for a small matrix ----------------------------------------------
def deseq(K, nf, X, prior_var=None):
    """K int[S, n], nf float[S, n], X float[S, p] -> dict of per-region arrays (NaN rows for all-zero regions)."""
    K = np.asarray(K, float)
    S, n = K.shape
    p = X.shape[1]
    max_disp = max(10.0, float(S))
    q = K / nf
    base_mean = q.mean(axis=0)
    base_var = q.var(axis=0, ddof=1)
    nz = K.sum(axis=0) > 0
    H = X @ np.linalg.solve(X.T @ X, X.T)                    # hat matrix: linearModelMu(y, X) = y H
    lin_mu = (H @ q)                                         # [S, n]
    mu_t = np.maximum(1.0, lin_mu)
    rough = np.maximum(0.0, np.sum(((q - mu_t) ** 2 - mu_t) / mu_t ** 2, axis=0) / (S - p))
    xim = np.mean(1.0 / nf[:, nz].mean(axis=1))
    with np.errstate(divide="ignore", invalid="ignore"):
        moments = (base_var - xim * base_mean) / base_mean ** 2
    alpha_init = np.clip(np.minimum(rough, moments), MIN_DISP, max_disp)
    distinct_rows = len(np.unique(X, axis=0))
    out = {k: np.full(n, np.nan) for k in ("dispGeneEst", "dispFit", "dispMAP", "dispersion", "stat", "pvalue", "deviance",
                                           "maxCooks", "lfc", "lfcSE")}
    out["baseMean"], out["baseVar"] = base_mean, base_var
    out["dispGeneIter"] = np.zeros(n, int); out["dispIter"] = np.zeros(n, int); out["betaIter"] = np.zeros(n, int)
    out["alpha_init"] = alpha_init
    lam = np.full(p, 1e-6 / math.log(2.0) ** 2)
    Qx, Rx = np.linalg.qr(X)
    mu_all = np.full((S, n), np.nan)
    # gene-wise estimates
    for i in np.flatnonzero(nz):
        y = K[:, i]
        if distinct_rows == p:
            mu = lin_mu[:, i] * nf[:, i]
        else:
            beta0 = np.linalg.solve(Rx, Qx.T @ np.log(q[:, i] + 0.1))
            b, _, _, _ = fit_beta(y, nf[:, i], X, alpha_init[i], beta0, lam)
            mu = nf[:, i] * np.exp(X @ b)
        mu = np.maximum(mu, MINMU)
        mu_all[:, i] = mu
        a, it, lp0, lp1 = fit_disp(y, mu, X, math.log(alpha_init[i]), 0.0, 1.0, False, math.log(MIN_DISP / 10.0))
        d = min(math.exp(a), max_disp)
        if lp1 < lp0 + abs(lp0) / 1e6:
            d = alpha_init[i]
        conv = it < MAXIT and it != 1
        if (not conv) and d > 10.0 * MIN_DISP:
            d = fit_disp_grid(y, mu, X, max_disp, 0.0, 1.0, False)
        out["dispGeneEst"][i] = min(max(d, MIN_DISP), max_disp)
        out["dispGeneIter"][i] = it
    # trend, prior
    ge = out["dispGeneEst"]
    a0, a1 = parametric_trend(base_mean[nz], ge[nz])
    out["trend"] = (a0, a1)
    out["dispFit"][nz] = a0 + a1 / base_mean[nz]
    above = nz & (ge >= 100.0 * MIN_DISP)
    resid = np.log(ge[above]) - np.log(out["dispFit"][above])
    var_log = (1.4826 * np.median(np.abs(resid - np.median(resid)))) ** 2
    out["varLogDispEsts"] = var_log
    df = S - p
    if prior_var is None:
        if df <= 3:
            raise NotImplementedError("Monte-Carlo prior variance (S - p <= 3)")
        prior_var = max(var_log - float(special.polygamma(1, df / 2.0)), 0.25)
    out["dispPriorVar"] = prior_var
    # MAP + GLM + Wald
    cells = np.unique(X, axis=0, return_inverse=True)[1].ravel()
    cell_sizes = np.bincount(cells)
    any3 = np.any(cell_sizes >= 3)
    for i in np.flatnonzero(nz):
        y, mu = K[:, i], mu_all[:, i]
        fit_i = out["dispFit"][i]
        init = ge[i] if ge[i] > 0.1 * fit_i else fit_i
        a, it, _, _ = fit_disp(y, mu, X, math.log(init), math.log(fit_i), prior_var, True, math.log(MIN_DISP / 10.0))
        d = math.exp(a)
        if it >= MAXIT:
            d = fit_disp_grid(y, mu, X, max_disp, math.log(fit_i), prior_var, True)
        d = min(max(d, MIN_DISP), max_disp)
        out["dispMAP"][i] = d
        out["dispIter"][i] = it
        outlier = math.log(ge[i]) > math.log(fit_i) + 2.0 * math.sqrt(var_log)
        alpha = ge[i] if outlier else d
        out["dispersion"][i] = alpha
        if p == 1:
            beta = np.array([math.log(np.mean(q[:, i]))])
            mu_w = nf[:, i] * np.exp(beta[0])
            w = 1.0 / (1.0 / mu_w + alpha)
            se = np.array([1.0 / math.sqrt(np.sum(w))])
            it_b = 1
            hat = w / np.sum(w)
        else:
            beta0 = np.linalg.solve(Rx, Qx.T @ np.log(q[:, i] + 0.1))
            beta, se, it_b, hat = fit_beta(y, nf[:, i], X, alpha, beta0, lam)
            mu_w = nf[:, i] * np.exp(X @ beta)
        out["betaIter"][i] = it_b
        out["deviance"][i] = -2.0 * nb_loglik(y, mu_w, alpha)
        out["lfc"][i] = beta[-1] / math.log(2.0)
        out["lfcSE"][i] = se[-1] / math.log(2.0)
        out["stat"][i] = beta[-1] / se[-1]
        out["pvalue"][i] = 2.0 * stats.norm.sf(abs(out["stat"][i]))
        if any3 and S > p:
            qi = q[:, i]
            v = trimmed_cell_variance(qi, cells)
            alpha_r = max((v - np.mean(qi)) / np.mean(qi) ** 2, 0.04)
            V = mu_w + alpha_r * mu_w ** 2
            cooks = (y - mu_w) ** 2 / V / p * hat / (1.0 - hat) ** 2
            out["maxCooks"][i] = np.max(cooks[cell_sizes[cells] >= 3])
    return out
