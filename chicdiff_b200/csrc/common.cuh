// common.cuh -- shared definitions for the chicdiff_b200 CUDA path (sm_100a).
//
// Matrix layout everywhere: "sample-major" = R's column-major n x S matrix, element
// (region i, sample s) at [s*n + i], so a warp of consecutive regions reads every
// per-sample column fully coalesced.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define CD_MAXS 32          // samples (replicates over both conditions)
#define CD_MAXP 4           // design columns

// per-region status bits (cd_results.flags)
#define CD_FLAG_ALLZERO 1          // every count is zero: all outputs NA (DESeq2 allZero)
#define CD_FLAG_GENE_GRID 2        // gene-wise estimate refitted on the grid (fitDispGrid)
#define CD_FLAG_MAP_GRID 4         // MAP estimate refitted on the grid
#define CD_FLAG_BETA_NOCONV 8      // IRLS did not converge: the reference would switch to optim()
#define CD_FLAG_OUTLIER 16         // dispersion outlier: final dispersion = gene-wise estimate
#define CD_FLAG_GENE_NOINCREASE 32 // line search did not raise the posterior: kept the start value
#define CD_FLAG_COOKS_KEEP 64      // >= 3 counts above the max-Cook's sample (results() heuristic)

struct CdDesign {
    int S, p;
    int linear_mu;                   // #distinct design rows == p: mu by hat-matrix projection
    int ncell, any3;
    double X[CD_MAXS * CD_MAXP];     // S x p, row-major
    double hat[CD_MAXS * CD_MAXS];   // X (X'X)^-1 X'
    double ls[CD_MAXP * CD_MAXS];    // (X'X)^-1 X'  (p x S): least-squares start for IRLS
    int cell[CD_MAXS];               // design cell of each sample
    int cell_size[CD_MAXS];
};

namespace cd {

constexpr double kMinDisp = 1e-8;
constexpr double kMinMu = 0.5;
constexpr double kLnSqrt2Pi = 0.918938533204672741780329736406;
constexpr double kLn2Pi = 1.837877066409345483560659472811;
constexpr double kLn2 = 0.693147180559945309417232121458;
constexpr double kLog2e = 1.442695040888963407359924681002;

// ---------------------------------------------------------------------------------------
// special functions (FP64)
// ---------------------------------------------------------------------------------------

// ---- building blocks of the line-search inner loop ----------------------------------------------
// Polynomial coefficients live in __constant__ memory so that the DFMAs take them as constant-bank
// operands; as literals every 64-bit coefficient costs two UMOV issue slots per use (measured: 20 %
// of all issued instructions in the first version of the kernel).
static __constant__ double kLogC[9] = {
    6.93147180369123816490e-01,   // ln2_hi
    1.90821492927058770002e-10,   // ln2_lo
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
    1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01};
// Stirling series of log Gamma in 1/x (odd powers): B_2k / (2k (2k-1)), k = 1..7
static __constant__ double kLgamC[7] = {1.0 / 12.0, -1.0 / 360.0, 1.0 / 1260.0, -1.0 / 1680.0, 1.0 / 1188.0,
                                        -691.0 / 360360.0, 1.0 / 156.0};
// asymptotic series of digamma in 1/x^2: -B_2k / (2k), k = 1..7
static __constant__ double kDigamC[7] = {-1.0 / 12.0, 1.0 / 120.0, -1.0 / 252.0, 1.0 / 240.0, -1.0 / 132.0,
                                         691.0 / 32760.0, -1.0 / 12.0};

// 1/b for a positive normal b: MUFU.RCP64H seed + two Newton steps (no slow-path branch; the
// arguments here are never zero, denormal, infinite or NaN)
__device__ __forceinline__ double rcp_pos(double b)
{
    double r;
#ifdef __CUDA_ARCH__
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
#else
    r = (1.0 / b) * (1.0 + 0x1p-21);   // host build of the accuracy tests (tests/device_math): a seed of similar quality
#endif
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    return fma(r, e, r);
}

// exp(x) for |x| <= 700 (the line search keeps log alpha in [-30, 10]): x = k ln2 + r, |r| <= ln2 / 2, degree-13 Taylor
// polynomial (the next term is 4e-18 of the result), k by the round-to-nearest trick of adding 1.5 * 2^52, 2^k by an
// integer add on the exponent.  18 FP64 + 2 integer instructions and no 64-bit literal below 1/10! (the four highest
// coefficients are rounded to 21 bits: <= 3e-18); libdevice's exp() spends half of its ~90 instructions moving 64-bit
// literals into registers (profiles/r02_d_fit_disp_source_lines.txt: 6.3 % of the line-search kernel).  <= 2 ulp.
static __constant__ double kExpC[8] = {1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0,
                                       1.0 / 362880.0};
__device__ __forceinline__ double exp_mid(double x)
{
    const double magic = 6755399441055744.0;                 // 1.5 * 2^52: the low word of x log2(e) + magic is rint(.)
    const double t = fma(x, kLog2e, magic);
    const int k = __double2loint(t);
    const double dk = t - magic;
    double r = fma(-dk, kLogC[0], x);                         // k ln2_hi is exact
    r = fma(-dk, kLogC[1], r);
    double p = fma(r, 0x1.61246p-33 /* 1/13! */, 0x1.1eed9p-29 /* 1/12! */);
    p = fma(r, p, 0x1.ae645p-26 /* 1/11! */);
    p = fma(r, p, 0x1.27e5p-22 /* 1/10! */);
#pragma unroll
    for (int j = 7; j >= 0; j--) p = fma(r, p, kExpC[j]);
    p = fma(r, p, 1.0);
    p = fma(r, p, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// natural log of a positive normal x (< 1 ulp): x = 2^k m with m in [sqrt(1/2), sqrt(2)),
// log m = f - f^2/2 + s (f^2/2 + R(s^2)), s = f / (2 + f), f = m - 1   (the classic fdlibm scheme)
__device__ __forceinline__ double log_pos(double x)
{
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int i = (hi + 0x95f64) & 0x100000;
    hi |= (i ^ 0x3ff00000);
    k += (i >> 20);
    const double f = __hiloint2double(hi, lo) - 1.0;
    const double s = f * rcp_pos(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, kLogC[7], kLogC[5]), kLogC[3]);
    const double t2 = z * fma(w, fma(w, fma(w, kLogC[8], kLogC[6]), kLogC[4]), kLogC[2]);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    return dk * kLogC[0] - ((hfsq - fma(s, hfsq + R, dk * kLogC[1])) - f);
}

// Table-assisted natural logarithm for the line-search kernel, where the logarithm runs 2 + 1/S times per replicate
// and evaluation.  The mantissa is reduced multiplicatively with a 512-entry table instead of the division
// s = f / (2 + f): x = 2^k m, m rc_i = 1 + r with |r| <= 2^-9, log x = k ln2 - log rc_i + log1p(r), log1p by a degree-6
// Taylor polynomial (the first dropped term, r^7 / 7, is below 1e-17 of the result even where the result is r itself;
// its two highest coefficients are rounded to 21 bits, which costs 2^-58 relative, so that they are instruction
// immediates).  No reciprocal, 11 FP64 + 6 integer instructions + the load (log_pos: ~25 FP64 + the MUFU seed).  rc_i has 20
// significant bits, so fma(m, rc_i, -1) is exact up to its single rounding.  The first interval uses rc = 1 (r = m - 1),
// the last one rc = 1/2 with lc = ln2 (r = m/2 - 1, exact), so the result keeps its relative accuracy at and above
// x = 1, which log(1 + mu alpha) at small alpha needs; just below 1 the absolute error is that of ln2 as a double
// (2.3e-17).  <= 1.5 ulp for x >= 1, <= 3e-16 max(1, |log x|) below (tests/test_device_math.py).
// tab = kLogTabN x {rc, -log rc}: kLogTab copied to shared memory by the kernel (the lanes of a warp index it with
// unrelated mantissas, which constant memory would serialise), 16-byte aligned.
constexpr int kLogTabN = 512;
alignas(16) static __constant__ double kLogTab[2 * kLogTabN] = {
#include "log_table.inc"
};

// the same table in global memory: what the kernels copy to shared memory (coalesced; a constant-memory read with a
// different index per lane is replayed 32 times)
#ifdef __CUDACC__
alignas(16) static __device__ const double kLogTabGlobal[2 * kLogTabN] = {
#include "log_table.inc"
};
// cooperative copy by the whole CTA; the caller synchronises
__device__ __forceinline__ void load_log_table(double* tab_shared)
{
    for (int k = threadIdx.x; k < 2 * kLogTabN; k += blockDim.x) tab_shared[k] = kLogTabGlobal[k];
}
#endif

struct alignas(16) LogTabEntry { double rc, lc; };      // one 16-byte shared-memory load per logarithm

// The table as the device code names it: its 32-bit shared-memory address (log_tab_handle), so that the look-up is a
// single LDS.128 [offset + base] -- through a generic pointer the compiler rebuilds the shared window's base on the way.
// The host build of the accuracy tests uses the pointer.
#ifdef __CUDA_ARCH__
typedef unsigned LogTab;
__device__ __forceinline__ LogTab log_tab_handle(const double* tab_in_shared) { return (unsigned)__cvta_generic_to_shared(tab_in_shared); }
#else
typedef const double* LogTab;
__device__ __forceinline__ LogTab log_tab_handle(const double* tab) { return tab; }
#endif

__device__ __forceinline__ double log_pos_v2(double x, LogTab tab)
{
    const int hi = __double2hiint(x);
    const int k = (hi >> 20) - 1023;
    LogTabEntry e;
#ifdef __CUDA_ARCH__
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(e.rc), "=d"(e.lc) : "r"(tab + ((unsigned)(hi & ((kLogTabN - 1) << 11)) >> 7)));
#else
    e = reinterpret_cast<const LogTabEntry*>(tab)[(hi >> 11) & (kLogTabN - 1)];      // top 9 bits of the mantissa
#endif
    // m rc = x (2^-k rc): the exponent of x is taken off rc's (two integer instructions on rc's high word; rc's low
    // word is zero) instead of building m.  Needs 2^-1020 < x < 2^1021, far beyond what the posterior produces.
    const double rcs = __hiloint2double(__double2hiint(e.rc) - (hi & 0x7ff00000) + 0x3ff00000, __double2loint(e.rc));
    const double lc = e.lc;
    const double r = fma(x, rcs, -1.0);
    double q = fma(r, -0x1.55555p-3 /* -1/6 */, 0x1.9999ap-3 /* 1/5 */);
    q = fma(r, q, -1.0 / 4.0);
    q = fma(r, q, 1.0 / 3.0);
    q = fma(r, q, -0.5);
    const double dk = (double)k;
    const double t = fma(dk, kLogC[0], lc);           // k ln2_hi is exact (ln2_hi has 32 trailing zero bits)
    const double u = fma(r * r, q, dk * kLogC[1]);
    return t + (r + u);
}

// 1/b for the posterior: the same seed and ONE cubic step, y0 (1 + e + e^2) = (1 - e^3) / b with e = 1 - b y0: three
// DFMAs instead of rcp_pos's four; |e| <= 2^-19 (the seed has ~20 bits) leaves 2^-57 plus the final rounding (<= 1 ulp,
// not always correctly rounded, which nothing in the posterior needs)
__device__ __forceinline__ double rcp_fast(double b)
{
    double r;
#ifdef __CUDA_ARCH__
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
#else
    r = (1.0 / b) * (1.0 + 0x1p-20);   // host build of the accuracy tests: a seed at the low end of the quality assumed
#endif
    const double e = fma(-b, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// log Gamma(x) - 0.5 log(2 pi) and digamma(x) for x > 0 in one go, same instruction sequence for every
// argument (libdevice's lgamma picks an argument-range code path; with one region per lane the lanes of a
// warp hold unrelated arguments y_j + 1/alpha, so those paths serialise).  Shift by 10 with the
// recurrence, the ten factors kept as one rational:  Gamma(x) = Gamma(x + 10) / prod (x + k),
// psi(x) = psi(x + 10) - sum 1/(x + k) = psi(x + 10) - num/den; asymptotic series at x + 10 >= 10
// (truncation < 1e-16).  log(x + 10) and 1/(x + 10) are shared by the two functions.
__device__ __forceinline__ void lgamma_digamma_pos(double x, double& lg, double& dg)
{
    double num = 1.0, den = x;
#pragma unroll
    for (int k = 1; k < 10; k++) {
        const double t = x + (double)k;
        num = fma(num, t, den);
        den *= t;
    }
    const double xs = x + 10.0;
    const double xi = rcp_pos(xs);
    const double f = xi * xi;
    const double lxs = log_pos(xs);
    double t = kLgamC[6];
#pragma unroll
    for (int k = 5; k >= 0; k--) t = fma(f, t, kLgamC[k]);
    lg = (((xs - 0.5) * lxs - xs) + xi * t) - log_pos(den);
    double u = kDigamC[6];
#pragma unroll
    for (int k = 5; k >= 0; k--) u = fma(f, u, kDigamC[k]);
    dg = ((lxs - 0.5 * xi) + f * u) - num * rcp_pos(den);
}

// log Gamma(x) - 0.5 log(2 pi) alone (same scheme as above)
__device__ __forceinline__ double lgamma_c_pos(double x)
{
    double den = x;
#pragma unroll
    for (int k = 1; k < 10; k++) den *= x + (double)k;
    const double xs = x + 10.0;
    const double xi = rcp_pos(xs);
    const double f = xi * xi;
    double t = kLgamC[6];
#pragma unroll
    for (int k = 5; k >= 0; k--) t = fma(f, t, kLgamC[k]);
    return (((xs - 0.5) * log_pos(xs) - xs) + xi * t) - log_pos(den);
}

__device__ __forceinline__ double trigamma_pos(double x)
{
    double acc = 0.0;
    while (x < 10.0) { acc += 1.0 / (x * x); x += 1.0; }
    const double xi = 1.0 / x;
    const double f = xi * xi;
    double t = 7.0 / 6.0;
    t = fma(f, t, -691.0 / 2730.0);
    t = fma(f, t, 5.0 / 66.0);
    t = fma(f, t, -1.0 / 30.0);
    t = fma(f, t, 1.0 / 42.0);
    t = fma(f, t, -1.0 / 30.0);
    t = fma(f, t, 1.0 / 6.0);
    return acc + xi + 0.5 * f + xi * f * t;
}

// Saddle-point pieces of the binomial/negative-binomial density (Loader 2000), in the
// branch structure R's nmath uses so that the deviance follows dnbinom_mu().
__device__ const double kSferrHalves[31] = {
    0.0, 0.1534264097200273452913848, 0.0810614667953272582196702,
    0.0548141210519176538961390, 0.0413406959554092940938221,
    0.03316287351993628748511048, 0.02767792568499833914878929,
    0.02374616365629749597132920, 0.02079067210376509311152277,
    0.01848845053267318523077934, 0.01664469118982119216319487,
    0.01513497322191737887351255, 0.01387612882307074799874573,
    0.01281046524292022692424986, 0.01189670994589177009505572,
    0.01110455975820691732662991, 0.010411265261972096497478567,
    0.009799416126158803298389475, 0.009255462182712732917728637,
    0.008768700134139385462952823, 0.008330563433362871256469318,
    0.007934114564314020547248100, 0.007573675487951840794972024,
    0.007244554301320383179543912, 0.006942840107209529865664152,
    0.006665247032707682442354394, 0.006408994188004207068439631,
    0.006171712263039457647532867, 0.005951370112758847735624416,
    0.005746216513010115682023589, 0.005554733551962801371038690};

__device__ __forceinline__ double stirlerr(double n)
{
    if (n <= 15.0) {
        const double nn = n + n;
        if (nn == (double)(int)nn) return kSferrHalves[(int)nn];
        return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - kLnSqrt2Pi;
    }
    const double nn = n * n;
    const double S0 = 1.0 / 12.0, S1 = 1.0 / 360.0, S2 = 1.0 / 1260.0, S3 = 1.0 / 1680.0, S4 = 1.0 / 1188.0;
    if (n > 500.0) return (S0 - S1 / nn) / n;
    if (n > 80.0) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35.0) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

__device__ __forceinline__ double bd0(double x, double np)
{
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < 2.2250738585072014e-308) return s;
        double ej = 2.0 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            const double s1 = s + ej / (double)((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

// log dnbinom(y; size, mu) split into the part that does not depend on mu (computed once
// per row and sample: nb_const) and the part that does (nb_var, evaluated in every IRLS
// iteration).  kind: 0 = y == 0, 1 = tiny y / size (Poisson-like expansion), 2 = general.
struct NbConst { double c; int kind; };

__device__ __forceinline__ NbConst nb_const(double y, double size)
{
    NbConst r;
    if (y == 0.0) { r.c = 0.0; r.kind = 0; return r; }
    if (y < 1e-10 * size) {
        r.c = -lgamma(y + 1.0) + log1p(y * (y - 1.0) / (2.0 * size));
        r.kind = 1;
        return r;
    }
    const double n = y + size;
    // log(size/(size+y)) + [stirlerr(n) - stirlerr(size) - stirlerr(y)] - 0.5*[ln 2pi + log(size) + log1p(-size/n)]
    r.c = log(size / n) + (stirlerr(n) - stirlerr(size) - stirlerr(y)) -
          0.5 * (kLn2Pi + log(size) + log1p(-size / n));
    r.kind = 2;
    return r;
}

__device__ __forceinline__ double nb_var(double y, double size, double mu, NbConst k)
{
    if (k.kind == 0)
        return size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu)));
    if (k.kind == 1) {
        const double p = (size < mu ? log(size / (1.0 + size / mu)) : log(mu / (1.0 + mu / size)));
        return y * p - mu + k.c;
    }
    const double n = y + size;
    const double pp = size / (size + mu), qq = mu / (size + mu);
    return k.c - bd0(size, n * pp) - bd0(y, n * qq);
}

__device__ __forceinline__ double dnbinom_mu_log(double y, double size, double mu)
{
    return nb_var(y, size, mu, nb_const(y, size));
}

// ---------------------------------------------------------------------------------------
// tiny SPD algebra, P <= CD_MAXP.  Symmetric matrices are stored packed lower:
// index(a, b) = a*(a+1)/2 + b for b <= a.
// ---------------------------------------------------------------------------------------
template <int P> struct Sym { double v[P * (P + 1) / 2]; };

template <int P> __device__ __forceinline__ int sidx(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

// in-place Cholesky A = L L'; returns log det A (NaN if not positive definite)
template <int P> __device__ __forceinline__ double chol_logdet(Sym<P>& A)
{
    double det = 1.0;
#pragma unroll
    for (int j = 0; j < P; j++) {
        double d = A.v[sidx<P>(j, j)];
#pragma unroll
        for (int k = 0; k < j; k++) d -= A.v[sidx<P>(j, k)] * A.v[sidx<P>(j, k)];
        det *= d;
        const double l = sqrt(d);
        A.v[sidx<P>(j, j)] = l;
#pragma unroll
        for (int i = j + 1; i < P; i++) {
            double s = A.v[sidx<P>(i, j)];
#pragma unroll
            for (int k = 0; k < j; k++) s -= A.v[sidx<P>(i, k)] * A.v[sidx<P>(j, k)];
            A.v[sidx<P>(i, j)] = s / l;
        }
    }
    return log(det);
}

// solve L L' x = b given the Cholesky factor
template <int P> __device__ __forceinline__ void chol_solve(const Sym<P>& L, double* x)
{
#pragma unroll
    for (int i = 0; i < P; i++) {
        double s = x[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= L.v[sidx<P>(i, k)] * x[k];
        x[i] = s / L.v[sidx<P>(i, i)];
    }
#pragma unroll
    for (int i = P - 1; i >= 0; i--) {
        double s = x[i];
#pragma unroll
        for (int k = i + 1; k < P; k++) s -= L.v[sidx<P>(k, i)] * x[k];
        x[i] = s / L.v[sidx<P>(i, i)];
    }
}

// explicit inverse (packed) from the Cholesky factor
template <int P> __device__ __forceinline__ void chol_inverse(const Sym<P>& L, Sym<P>& inv)
{
#pragma unroll
    for (int c = 0; c < P; c++) {
        double e[P];
#pragma unroll
        for (int k = 0; k < P; k++) e[k] = (k == c) ? 1.0 : 0.0;
        chol_solve<P>(L, e);
#pragma unroll
        for (int r = c; r < P; r++) inv.v[sidx<P>(r, c)] = e[r];
    }
}

}  // namespace cd
