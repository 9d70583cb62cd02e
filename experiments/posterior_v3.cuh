// posterior_v3.cuh -- EXPERIMENT: posterior_v2 with every logarithm of the sample loop taken by log_pos_v2 (log_v2.cuh).
// `tab` is the 128 x {rc, -log rc} table, in shared memory on the device.
#pragma once
#include "posterior_v2.cuh"
#include "log_v2.cuh"

namespace cd {

__device__ __forceinline__ GammaParts gamma_parts_t(double x, const double* tab)
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
    const double lxs = log_pos_v2(xs, tab);
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

// P <= 2 only (the closed-form Cox-Reid term); larger designs stay on eval_post_v2
template <int P, bool WANT_D>
__device__ __forceinline__ void eval_post_v3(double a, const double* ys, const double* mus, int stride, int S,
                                             double prior_mean, double prior_sigmasq, bool use_prior, const double* tab,
                                             double& lp_out, double& dlp_out)
{
    static_assert(P <= 2, "closed forms for p = 1, 2");
    const double alpha = exp(a);
    const double r = rcp_pos(alpha);
    const double log_r = -a;
    const GammaParts gr = gamma_parts_t(r, tab);
    const double inv_den_r = rcp_pos(gr.den);
    const double dgr = gr.dgs - gr.num * inv_den_r;
    double b00 = 0.0, b10 = 0.0, b11 = 0.0, d00 = 0.0, d10 = 0.0, d11 = 0.0;
    double ll = 0.0, ds = 0.0, qprod = 1.0;
#pragma unroll 1
    for (int j = 0; j < S; j++) {
        const double yj = ys[j * stride], muj = mus[j * stride];
        const double ma = muj * alpha;
        const double ropm = rcp_pos(1.0 + ma);
        const double w = muj * ropm;
        const double dw = -w * w;
        b00 += w;
        if (WANT_D) d00 += dw;
        if (P == 2) {
            const double x1 = c_des.X[j * P + 1];           // column 0 is the intercept
            b10 += w * x1; b11 += w * x1 * x1;
            if (WANT_D) { d10 += dw * x1; d11 += dw * x1 * x1; }
        }
        const double l1 = log_pos_v2(1.0 + ma, tab);
        const GammaParts g = gamma_parts_t(yj + r, tab);
        qprod *= g.den * inv_den_r;
        if (j & 1) { ll -= log_pos_v2(qprod, tab); qprod = 1.0; }
        ll += ((g.st - gr.st) - yj * (log_r + l1)) - r * l1;
        if (WANT_D) {
            const double dg = g.dgs - g.num * rcp_pos(g.den);
            ds += ((dgr - dg) + (l1 - ma * ropm)) + yj * (alpha * ropm);
        }
    }
    if (S & 1) ll -= log_pos_v2(qprod, tab);
    double cr, dcr = 0.0;
    if (P == 1) {
        cr = -0.5 * ((b00 > 0.0) ? log_pos_v2(b00, tab) : NAN);
        if (WANT_D) dcr = -0.5 * (d00 * rcp_pos(b00));
    } else {
        const double det = b00 * b11 - b10 * b10;
        const bool ok = (b00 > 0.0) && (det > 0.0);
        cr = -0.5 * (ok ? log_pos_v2(det, tab) : NAN);
        if (WANT_D) dcr = -0.5 * ((b11 * d00 - 2.0 * b10 * d10 + b00 * d11) * rcp_pos(det));
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
