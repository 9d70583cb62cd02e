// priorvar.cpp -- DESeq2's dispersion prior variance for designs with 1 <= S - p <= 3 residual degrees of freedom
// (estimateDispersionsPriorVar, the Monte-Carlo branch: SURVEY.md Appendix A.7; reached from chicdiff.R:1573, 1603, 1643,
// 1673 for every 2-vs-2 run, the reference's bundled example included).
//
// With so few degrees of freedom the sampling variance of log(dispGeneEst) is not trigamma((S-p)/2) any more; DESeq2
// matches the histogram of the observed residuals log(dispGeneEst) - log(dispFit) against simulated ones: for each of 200
// candidate prior variances x in [0, 8] it draws, after set.seed(2), 1e4 values log(rchisq(df)) + rnorm(0, sqrt(x)) -
// log(df), bins them like hist(breaks = -20:20/2), takes the Kullback-Leibler divergence of the observed histogram from
// the simulated one, smooths the 200 divergences with loess(span = 0.2) and takes the minimiser on a 1000-point grid,
// floored at 0.25.
//
// Host code (it is 2e6 sequential random draws and a 200-point smoother, once per residual degrees of freedom).  What is
// restated here, from the published algorithms R implements (R's sources are not in this image; nothing could be run
// against R):
//   * set.seed(): the 50-step LCG scrambling of the seed and the 625-word fill of the Mersenne-Twister state (RNG.c);
//     unif_rand() = MT19937 output x 2^-32 with R's fix-up away from 0 and 1;
//   * norm_rand() by inversion: (int)(2^27 u1) + u2 over 2^27 through qnorm (Wichura's AS 241, PPND16);
//   * exp_rand(): Ahrens & Dieter 1972 (algorithm SA); rgamma(): Ahrens & Dieter 1974 (GS, a < 1) and 1982 (GD, a >= 1);
//     rchisq(df) = rgamma(df / 2, scale 2); rnorm(mu, 0) returns mu without touching the stream;
//   * hist(): right-closed bins with the 1e-7 x median(width) fuzz on the breaks, density = count / (n x width);
//   * loess(): degree 2, tricube weights over the floor(0.2 n) nearest points, surface = "interpolate": local fits (value
//     and slope) at the vertices of the k-d tree (cells split at the median until <= floor(n span cell) = 8 points; box
//     expanded by 0.5 %), cubic Hermite blending inside a cell (Cleveland & Grosse's dloess, ehg126 / ehg124 / ehg127 /
//     ehg128).
// The uniform and normal streams are pinned by values R is known to print (tests/test_priorvar.py: set.seed(42);
// runif(3), rnorm(3) ...); the gamma sampler and the smoother are checked against an independent restatement (the
// checker's priorvar.py: NumPy's Mersenne-Twister, SciPy's quantile function and splines) and against their mathematical
// definitions, not against R: "parity unpinned" for this function.
// The simulated histograms depend only on df (the seed is fixed), so they are built once per df and process.
#include "../../include/chicdiff_b200.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <mutex>
#include <vector>

namespace {

// ---- R's Mersenne-Twister ------------------------------------------------------------------
struct RRng {
    uint32_t mt[624];
    int mti;

    explicit RRng(uint32_t seed)
    {
        for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;               // initial scrambling (RNG.c, RNG_Init)
        uint32_t dummy0;
        seed = 69069u * seed + 1u; dummy0 = seed; (void)dummy0;               // i_seed[0] is the position word ...
        for (int j = 0; j < 624; j++) { seed = 69069u * seed + 1u; mt[j] = seed; }
        mti = 624;                                                            // ... which FixupSeeds sets to N: regenerate first
    }
    uint32_t next32()
    {
        static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
        if (mti >= 624) {
            int kk;
            for (kk = 0; kk < 624 - 397; kk++) {
                const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
            }
            for (; kk < 623; kk++) {
                const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
            }
            const uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
            mti = 0;
        }
        uint32_t y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double unif()
    {
        const double i2_32m1 = 2.328306437080797e-10;
        const double x = (double)next32() * 2.3283064365386963e-10;           // [0, 1)
        if (x <= 0.0) return 0.5 * i2_32m1;
        if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
        return x;
    }
};

// qnorm(p) for 0 < p < 1: Wichura (1988), algorithm AS 241, PPND16
double qnorm_std(double p)
{
    const double q = p - 0.5;
    double r, val;
    if (std::fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        val = q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                       45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                     133.14166789178437745) * r + 3.387132872796366608) /
              (((((((r * 5226.495278852854561 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                   21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
                42.313330701600911252) * r + 1.0);
        return val;
    }
    r = (q < 0) ? p : 1.0 - p;
    r = std::sqrt(-std::log(r));
    if (r <= 5.0) {
        r -= 1.6;
        val = (((((((r * 7.7454501427834140764e-4 + 0.0227238449892691845833) * r + 0.24178072517745061177) * r +
                   1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
                4.6303378461565452959) * r + 1.42343711074968357734) /
              (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r +
                   0.14810397642748007459) * r + 0.68976733498510000455) * r + 1.6763848301838038494) * r +
                2.05319162663775882187) * r + 1.0);
    } else {
        r -= 5.0;
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r +
                   0.026532189526576123093) * r + 0.29656057182850489123) * r + 1.7848265399172913358) * r +
                5.4637849111641143699) * r + 6.6579046435011037772) /
              (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
                   7.868691311456132591e-4) * r + 0.0148753612908506148525) * r + 0.13692988092273580531) * r +
                0.59983220655588793769) * r + 1.0);
    }
    return (q < 0.0) ? -val : val;
}

double norm_rand(RRng& g)
{
    const double BIG = 134217728.0;                                           // 2^27: one uniform is not precise enough
    double u = g.unif();
    u = (double)(int)(BIG * u) + g.unif();
    return qnorm_std(u / BIG);
}

double exp_rand(RRng& g)
{
    // q[k-1] = sum_{i=1..k} ln(2)^i / i!
    static double q[16];
    static bool have = false;
    if (!have) {
        double term = 1.0, sum = 0.0;
        for (int k = 1; k <= 16; k++) { term *= M_LN2 / k; sum += term; q[k - 1] = sum; }
        q[15] = 1.0;
        have = true;
    }
    double a = 0.0;
    double u = g.unif();
    while (u <= 0.0 || u >= 1.0) u = g.unif();
    for (;;) {
        u += u;
        if (u > 1.0) break;
        a += q[0];
    }
    u -= 1.0;
    if (u <= q[0]) return a + u;
    int i = 0;
    double ustar = g.unif(), umin = ustar;
    do {
        ustar = g.unif();
        if (umin > ustar) umin = ustar;
        i++;
    } while (u > q[i]);
    return a + umin * q[0];
}

// rgamma(a, scale) as in R's nmath/rgamma.c; the cached quantities of the GD algorithm live in the struct
struct RGamma {
    double aa = 0.0, aaa = 0.0, s = 0.0, s2 = 0.0, d = 0.0, q0 = 0.0, b = 0.0, si = 0.0, c = 0.0;

    double draw(RRng& g, double a, double scale)
    {
        const double sqrt32 = 5.656854, exp_m1 = 0.36787944117144233;
        const double q1 = 0.04166669, q2 = 0.02083148, q3 = 0.00801191, q4 = 0.00144121, q5 = -7.388e-5, q6 = 2.4511e-4, q7 = 2.424e-4;
        const double a1 = 0.3333333, a2 = -0.250003, a3 = 0.2000062, a4 = -0.1662921, a5 = 0.1423657, a6 = -0.1367177, a7 = 0.1233795;
        double e, p, q, r, t, u, v, w, x, ret_val;
        if (a < 1.0) {                                                        // GS algorithm for 0 < a < 1
            e = 1.0 + exp_m1 * a;
            for (;;) {
                p = e * g.unif();
                if (p >= 1.0) {
                    x = -std::log((e - p) / a);
                    if (exp_rand(g) >= (1.0 - a) * std::log(x)) break;
                } else {
                    x = std::exp(std::log(p) / a);
                    if (exp_rand(g) >= x) break;
                }
            }
            return scale * x;
        }
        // GD algorithm.  Step 1: recalculations of s2, s, d if a has changed
        if (a != aa) { aa = a; s2 = a - 0.5; s = std::sqrt(s2); d = sqrt32 - s * 12.0; }
        // Step 2: t = standard normal deviate, x = (s, 1/2)-normal deviate; immediate acceptance
        t = norm_rand(g);
        x = s + 0.5 * t;
        ret_val = x * x;
        if (t >= 0.0) return scale * ret_val;
        // Step 3: u = uniform sample; squeeze acceptance
        u = g.unif();
        if (d * u <= t * t * t) return scale * ret_val;
        // Step 4: recalculations of q0, b, si, c if necessary
        if (a != aaa) {
            aaa = a;
            r = 1.0 / a;
            q0 = ((((((q7 * r + q6) * r + q5) * r + q4) * r + q3) * r + q2) * r + q1) * r;
            if (a <= 3.686) { b = 0.463 + s + 0.178 * s2; si = 1.235; c = 0.195 / s - 0.079 + 0.16 * s; }
            else if (a <= 13.022) { b = 1.654 + 0.0076 * s2; si = 1.68 / s + 0.275; c = 0.062 / s + 0.024; }
            else { b = 1.77; si = 0.75; c = 0.1515 / s; }
        }
        // Step 5: no quotient test if x not positive
        if (x > 0.0) {
            // Step 6: calculation of v and quotient q
            v = t / (s + s);
            if (std::fabs(v) <= 0.25)
                q = q0 + 0.5 * t * t * ((((((a7 * v + a6) * v + a5) * v + a4) * v + a3) * v + a2) * v + a1) * v;
            else
                q = q0 - s * t + 0.25 * t * t + (s2 + s2) * std::log(1.0 + v);
            // Step 7: quotient acceptance
            if (std::log(1.0 - u) <= q) return scale * ret_val;
        }
        for (;;) {
            // Step 8: e = standard exponential deviate, u = uniform deviate, t = (b, si)-double exponential sample
            e = exp_rand(g);
            u = g.unif();
            u = u + u - 1.0;
            t = (u < 0.0) ? b - si * e : b + si * e;
            // Step 9: rejection if t < tau(1) = -0.71874483771719
            if (t >= -0.71874483771719) {
                // Step 10: calculation of v and quotient q
                v = t / (s + s);
                if (std::fabs(v) <= 0.25)
                    q = q0 + 0.5 * t * t * ((((((a7 * v + a6) * v + a5) * v + a4) * v + a3) * v + a2) * v + a1) * v;
                else
                    q = q0 - s * t + 0.25 * t * t + (s2 + s2) * std::log(1.0 + v);
                // Step 11: hat acceptance
                if (q > 0.0) {
                    w = std::expm1(q);
                    if (c * std::fabs(u) <= w * std::exp(e - 0.5 * t * t)) break;
                }
            }
        }
        x = s + 0.5 * t;
        return scale * x * x;
    }
};

// ---- hist(breaks = -20:20/2) -----------------------------------------------------------------
constexpr int kBins = 40;

struct Breaks {
    double b[kBins + 1], fuzzy[kBins + 1];
    Breaks()
    {
        for (int k = 0; k <= kBins; k++) b[k] = (double)(k - 20) / 2.0;
        const double diddle = 1e-7 * 0.5;                                     // 1e-7 * median(diff(breaks))
        for (int k = 0; k <= kBins; k++) fuzzy[k] = b[k] + (k == 0 ? -diddle : diddle);
    }
};

// counts of the values strictly inside (min(breaks), max(breaks)), binned like C_BinCount(right = TRUE, include.lowest = TRUE)
void hist_add(const Breaks& B, double x, double counts[kBins])
{
    if (!(x > B.b[0] && x < B.b[kBins])) return;                              // the rule filters before hist() sees the value
    int lo = 0, hi = kBins;
    if (B.fuzzy[lo] <= x && (x < B.fuzzy[hi] || x == B.fuzzy[hi])) {
        while (hi - lo >= 2) {
            const int mid = (hi + lo) / 2;
            if (x > B.fuzzy[mid]) lo = mid; else hi = mid;
        }
        counts[lo] += 1.0;
    }
}

void density_of(const double counts[kBins], double dens[kBins])
{
    double n = 0.0;
    for (int k = 0; k < kBins; k++) n += counts[k];
    for (int k = 0; k < kBins; k++) dens[k] = counts[k] / (n * 0.5);
}

// ---- the simulated histograms, once per df ------------------------------------------------------
constexpr int kGrid = 200, kFine = 1000, kDraws = 10000;

struct SimTable { bool have = false; double dens[kGrid][kBins]; };
SimTable g_sim[4];
std::mutex g_sim_mutex;

const SimTable& sim_table(int df)
{
    std::lock_guard<std::mutex> lock(g_sim_mutex);
    SimTable& T = g_sim[df];
    if (T.have) return T;
    const Breaks B;
    RRng g(2u);                                                               // set.seed(2)
    RGamma gam;
    std::vector<double> chi(kDraws);
    const double ldf = std::log((double)df);
    for (int k = 0; k < kGrid; k++) {
        const double x = (double)k * (8.0 / (double)(kGrid - 1));             // seq(0, 8, length = 200): from + k * by
        const double sd = std::sqrt(x);
        for (int i = 0; i < kDraws; i++) chi[i] = std::log(gam.draw(g, (double)df / 2.0, 2.0));   // log(rchisq(1e4, df))
        double counts[kBins] = {0.0};
        for (int i = 0; i < kDraws; i++) {
            const double z = (sd == 0.0) ? 0.0 : 0.0 + sd * norm_rand(g);    // rnorm(1e4, 0, sd): no draw when sd == 0
            hist_add(B, chi[i] + z - ldf, counts);
        }
        density_of(counts, T.dens[k]);
    }
    T.have = true;
    return T;
}

// ---- loess(y ~ x, span = 0.2), degree 2, surface = "interpolate", evaluated on a grid ---------------
struct Loess1D {
    std::vector<double> vx, vval, vslope;                                     // vertices of the k-d tree, ascending

    static void local_fit(const std::vector<double>& x, const std::vector<double>& y, double at, int q, double& val, double& slope)
    {
        const int n = (int)x.size();
        std::vector<double> d(n);
        for (int i = 0; i < n; i++) d[i] = std::fabs(x[i] - at);
        std::vector<double> sorted = d;
        std::nth_element(sorted.begin(), sorted.begin() + (q - 1), sorted.end());
        const double h = sorted[q - 1];
        // weighted least squares of y on (1, u, u^2), u = x - at, tricube weights inside h: normal equations in long double
        long double S[5] = {0, 0, 0, 0, 0}, T[3] = {0, 0, 0};
        for (int i = 0; i < n; i++) {
            if (!(d[i] < h)) continue;
            const double r = d[i] / h;
            const double t3 = 1.0 - r * r * r;
            const long double w = (long double)(t3 * t3 * t3);
            const long double u = (long double)(x[i] - at);
            long double p = w;
            for (int k = 0; k < 5; k++) { S[k] += p; if (k < 3) T[k] += p * (long double)y[i]; p *= u; }
        }
        // solve [[S0 S1 S2][S1 S2 S3][S2 S3 S4]] b = T by Cramer's rule
        const long double a11 = S[0], a12 = S[1], a13 = S[2], a22 = S[2], a23 = S[3], a33 = S[4];
        const long double det = a11 * (a22 * a33 - a23 * a23) - a12 * (a12 * a33 - a23 * a13) + a13 * (a12 * a23 - a22 * a13);
        const long double b0 = (T[0] * (a22 * a33 - a23 * a23) - a12 * (T[1] * a33 - a23 * T[2]) + a13 * (T[1] * a23 - a22 * T[2])) / det;
        const long double b1 = (a11 * (T[1] * a33 - a23 * T[2]) - T[0] * (a12 * a33 - a23 * a13) + a13 * (a12 * T[2] - T[1] * a13)) / det;
        val = (double)b0;
        slope = (double)b1;
    }

    // cells of the k-d tree over the (sorted) x: split l..u (1-based, inclusive) at m = (l + u) / 2 while more than fc points
    static void cuts(const std::vector<double>& x, int l, int u, int fc, std::vector<double>& out)
    {
        if (u - l + 1 <= fc) return;
        int m = (l + u) / 2;
        while (m > l && x[m - 2] == x[m - 1]) m--;                            // ties go with the upper son
        out.push_back((x[m - 1] + x[m]) / 2.0);
        cuts(x, l, m, fc, out);
        cuts(x, m + 1, u, fc, out);
    }

    Loess1D(const std::vector<double>& x, const std::vector<double>& y, double span, double cell)
    {
        const int n = (int)x.size();
        const int q = std::min(n, (int)std::floor((double)n * span + 1e-5));
        const int fc = (int)std::floor((double)n * span * cell);
        const double lo = x.front(), hi = x.back();
        const double mu = 0.005 * std::max(hi - lo, 1e-10 * std::max(std::fabs(lo), std::fabs(hi)) + 1e-30);   // "expand the box a little"
        vx.push_back(lo - mu);
        vx.push_back(hi + mu);
        cuts(x, 1, n, fc, vx);
        std::sort(vx.begin(), vx.end());
        vval.resize(vx.size()); vslope.resize(vx.size());
        for (size_t k = 0; k < vx.size(); k++) local_fit(x, y, vx[k], q, vval[k], vslope[k]);
    }

    double at(double z) const
    {
        size_t k = (size_t)(std::upper_bound(vx.begin(), vx.end(), z) - vx.begin());
        if (k == 0) k = 1;
        if (k >= vx.size()) k = vx.size() - 1;
        const double v0 = vx[k - 1], v1 = vx[k], h = v1 - v0, u = (z - v0) / h;
        const double phi0 = (1 - u) * (1 - u) * (1 + 2 * u), phi1 = u * u * (3 - 2 * u);
        const double psi0 = u * (1 - u) * (1 - u), psi1 = -u * u * (1 - u);
        return phi0 * vval[k - 1] + phi1 * vval[k] + (psi0 * vslope[k - 1] + psi1 * vslope[k]) * h;
    }
};

double prior_var_from_counts(int df, const double obs_counts[kBins], double* kl_out /*200 or null*/)
{
    double obs[kBins];
    density_of(obs_counts, obs);
    const SimTable& T = sim_table(df);
    std::vector<double> grid(kGrid), kl(kGrid);
    for (int k = 0; k < kGrid; k++) {
        grid[k] = (double)k * (8.0 / (double)(kGrid - 1));
        double small = INFINITY;
        for (int b = 0; b < kBins; b++) {
            if (obs[b] > 0.0 && obs[b] < small) small = obs[b];
            if (T.dens[k][b] > 0.0 && T.dens[k][b] < small) small = T.dens[k][b];
        }
        double s = 0.0;
        for (int b = 0; b < kBins; b++) s += obs[b] * (std::log(obs[b] + small) - std::log(T.dens[k][b] + small));
        kl[k] = s;
        if (kl_out) kl_out[k] = s;
    }
    const Loess1D fit(grid, kl, 0.2, 0.2);
    double best = INFINITY, arg = 0.0;
    for (int k = 0; k < kFine; k++) {
        const double z = (double)k * (8.0 / (double)(kFine - 1));
        const double f = fit.at(z);
        if (f < best) { best = f; arg = z; }                                  // which.min: the first minimum
    }
    return std::max(arg, 0.25);
}

}  // namespace

extern "C" {

int cd_prior_var_hist(int64_t n_resid, const double* resid, double counts[40])
{
    if (n_resid < 0 || (n_resid > 0 && !resid) || !counts) return CD_EINVAL;
    const Breaks B;
    for (int k = 0; k < kBins; k++) counts[k] = 0.0;
    for (int64_t i = 0; i < n_resid; i++) hist_add(B, resid[i], counts);
    return CD_OK;
}

double cd_prior_var_from_hist(int df, const double counts[40])
{
    if (df < 1 || df > 3 || !counts) return NAN;
    double n = 0.0;
    for (int k = 0; k < kBins; k++) n += counts[k];
    if (!(n > 0.0)) return NAN;
    return prior_var_from_counts(df, counts, nullptr);
}

double cd_prior_var_small_df(int df, int64_t n_resid, const double* resid)
{
    double counts[kBins];
    if (cd_prior_var_hist(n_resid, resid, counts) != CD_OK) return NAN;
    return cd_prior_var_from_hist(df, counts);
}

// test hooks: the first n values of the streams after set.seed(seed); what = 0 unif_rand, 1 norm_rand, 2 exp_rand,
// 3 rgamma(shape, 1); and the Kullback-Leibler curve of a histogram
int cd_prior_var_debug_stream(unsigned int seed, int what, double shape, int n, double* out)
{
    if (n < 0 || !out) return CD_EINVAL;
    RRng g(seed);
    RGamma gam;
    for (int i = 0; i < n; i++)
        out[i] = what == 0 ? g.unif() : what == 1 ? norm_rand(g) : what == 2 ? exp_rand(g) : gam.draw(g, shape, 1.0);
    return CD_OK;
}

int cd_prior_var_debug_curve(int df, const double counts[40], double kl_out[200], double fitted_out[1000])
{
    if (df < 1 || df > 3 || !counts || !kl_out) return CD_EINVAL;
    prior_var_from_counts(df, counts, kl_out);
    if (fitted_out) {
        std::vector<double> grid(kGrid), kl(kl_out, kl_out + kGrid);
        for (int k = 0; k < kGrid; k++) grid[k] = (double)k * (8.0 / (double)(kGrid - 1));
        const Loess1D fit(grid, kl, 0.2, 0.2);
        for (int k = 0; k < kFine; k++) fitted_out[k] = fit.at((double)k * (8.0 / (double)(kFine - 1)));
    }
    return CD_OK;
}

}  // extern "C"
