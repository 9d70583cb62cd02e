// posterior_v2.cuh -- EXPERIMENT, not part of the product build (nothing under chicdiff_b200/ includes it).
//
// A lighter formulation of eval_post (chicdiff_b200/csrc/posterior.cuh), prepared from the per-source-line instruction
// profile of the line-search kernel (profiles/r01_final_fit_disp_source_lines.txt: the three log_pos calls per sample
// are ~60 % of the executed instructions, the Cholesky epilogue with its sqrt / divisions / libdevice log another ~10 %):
//
//   1. log(den_j) of the shift-10 gamma rational is no longer taken per sample.  With q_j = den_j / den_r (>= 1, the
//      ratio against the rational of r = 1/alpha alone) the sum over samples of [log den_j - log den_r] is
//      log(q_a q_b) per PAIR of samples: 2.5 instead of 3 logarithms per sample, and a zero count still contributes
//      (almost) exactly zero.  q_a q_b stays far from overflow for counts up to ~1e15.
//   2. P = 1 and P = 2 (all fits of the default pipeline) use the closed-form determinant and inverse of X'WX with one
//      Newton reciprocal and log_pos instead of Cholesky + sqrt + three divisions + libdevice log.
//
// Numerically it is the same function: tests/test_device_math.py::test_experimental_posterior compares it with eval_post
// on the host.  Whether it is faster on the GPU has to be measured (next round); it changes sums at the 1e-16 level, so
// it also has to go through the GPU parity suite before it replaces eval_post.
#pragma once
#include "posterior.cuh"

namespace cd {

// the pieces of lgamma_digamma_pos, with the rational's denominator handed back instead of logged
struct GammaParts { double st, dgs, num, den; };      // st: Stirling part of lgamma; dgs: series part of digamma

__device__ __forceinline__ GammaParts gamma_parts(double x)
{
    GammaParts g;
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
    g.st = ((xs - 0.5) * lxs - xs) + xi * t;
    double u = kDigamC[6];
#pragma unroll
    for (int k = 5; k >= 0; k--) u = fma(f, u, kDigamC[k]);
    g.dgs = (lxs - 0.5 * xi) + f * u;
    g.num = num; g.den = den;
    return g;
}

template <int P, bool WANT_D>
__device__ __forceinline__ void eval_post_v2(double a, const double* ys, const double* mus, int stride, int S,
                                             double prior_mean, double prior_sigmasq, bool use_prior,
                                             double& lp_out, double& dlp_out)
{
    const double alpha = exp(a);
    const double r = rcp_pos(alpha);
    const double log_r = -a;
    const GammaParts gr = gamma_parts(r);
    const double inv_den_r = rcp_pos(gr.den);
    const double dgr = gr.dgs - gr.num * inv_den_r;
    Sym<P> B, dB;
#pragma unroll
    for (int k = 0; k < P * (P + 1) / 2; k++) { B.v[k] = 0.0; dB.v[k] = 0.0; }
    double ll = 0.0, ds = 0.0, qprod = 1.0;
#pragma unroll 1
    for (int j = 0; j < S; j++) {
        const double yj = ys[j * stride], muj = mus[j * stride];
        const double ma = muj * alpha;
        const double ropm = rcp_pos(1.0 + ma);
        const double w = muj * ropm;
        const double dw = -w * w;
        if (P == 1) {
            B.v[0] += w;
            if (WANT_D) dB.v[0] += dw;
        } else {
#pragma unroll
            for (int u = 0; u < P; u++)
#pragma unroll
                for (int v = 0; v <= u; v++) {
                    const double xx = c_des.X[j * P + u] * c_des.X[j * P + v];
                    B.v[u * (u + 1) / 2 + v] += w * xx;
                    if (WANT_D) dB.v[u * (u + 1) / 2 + v] += dw * xx;
                }
        }
        const double l1 = log_pos(1.0 + ma);
        const GammaParts g = gamma_parts(yj + r);
        qprod *= g.den * inv_den_r;                       // den_j / den_r >= 1
        if (j & 1) { ll -= log_pos(qprod); qprod = 1.0; }  // one logarithm per pair of samples
        ll += ((g.st - gr.st) - yj * (log_r + l1)) - r * l1;
        if (WANT_D) {
            const double dg = g.dgs - g.num * rcp_pos(g.den);
            ds += ((dgr - dg) + (l1 - ma * ropm)) + yj * (alpha * ropm);
        }
    }
    if (S & 1) ll -= log_pos(qprod);
    double cr, dcr = 0.0;
    if (P == 1) {
        const double b = B.v[0];
        cr = -0.5 * ((b > 0.0) ? log_pos(b) : NAN);
        if (WANT_D) dcr = -0.5 * (dB.v[0] * rcp_pos(b));
    } else if (P == 2) {
        const double det = B.v[0] * B.v[2] - B.v[1] * B.v[1];
        const bool ok = (B.v[0] > 0.0) && (det > 0.0);
        cr = -0.5 * (ok ? log_pos(det) : NAN);
        if (WANT_D) {
            // tr(B^-1 dB) = (B11 dB00 - 2 B10 dB10 + B00 dB11) / det
            const double tr = (B.v[2] * dB.v[0] - 2.0 * B.v[1] * dB.v[1] + B.v[0] * dB.v[2]) * rcp_pos(det);
            dcr = -0.5 * tr;
        }
    } else {
        Sym<P> L = B;
        cr = -0.5 * chol_logdet<P>(L);
        if (WANT_D) {
            Sym<P> Bi;
            chol_inverse<P>(L, Bi);
            double tr = 0.0;
#pragma unroll
            for (int u = 0; u < P; u++)
#pragma unroll
                for (int v = 0; v < P; v++) tr += Bi.v[sidx<P>(u, v)] * dB.v[sidx<P>(v, u)];
            dcr = -0.5 * tr;
        }
    }
    double pr = 0.0;
    if (use_prior) {
        const double d = a - prior_mean;
        pr = -0.5 * d * d / prior_sigmasq;
    }
    lp_out = ll + pr + cr;
    if (WANT_D) {
        const double dpr = use_prior ? -1.0 * (a - prior_mean) / prior_sigmasq : 0.0;
        dlp_out = ((r * r) * ds + dcr) * alpha + dpr;
    }
}

}  // namespace cd
