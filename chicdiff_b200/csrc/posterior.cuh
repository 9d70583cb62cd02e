// posterior.cuh -- the Cox-Reid adjusted profile log-posterior of log(alpha) and its derivative, the function every
// trip of the dispersion line search evaluates (dispersion.cu).  Kept in a header of its own so that the accuracy
// tests can compile it for the host (tests/device_math) and compare it with the oracle without a GPU.
#pragma once
#include "common.cuh"

namespace cd {

// c_des, the design of the running fit, must be visible here: dispersion.cu defines it in constant memory before
// including this header (set_design_dispersion fills it); the host build of the accuracy tests supplies a plain global.
#ifndef __CUDACC__
extern CdDesign c_des;
#endif

// ---------------------------------------------------------------------------------------
// Cox-Reid adjusted profile log-posterior of log(alpha) (DESeq2.cpp log_posterior) and its
// derivative (dlog_posterior), evaluated together at one point.  The line search needs the
// posterior at every proposal and the derivative at every accepted one; computed together they
// share exp(a), w_j = 1/(1/mu_j + alpha), log(1 + mu_j alpha), log(y_j + r + 10) and the shifted
// gamma-function rationals, so the pair costs ~1.3x the posterior alone, and the lanes of a warp
// never split into "needs the derivative" and "does not".
// ys / mus point at the region's replicates in shared memory, element j at [j * stride].
// WANT_D = false (grid refit) skips the derivative.
// ---------------------------------------------------------------------------------------
template <int P, bool WANT_D>
__device__ __forceinline__ void eval_post(double a, const double* ys, const double* mus, int stride, int S,
                                          double prior_mean, double prior_sigmasq, bool use_prior,
                                          double& lp_out, double& dlp_out)
{
    const double alpha = exp(a);
    const double r = rcp_pos(alpha);
    const double log_r = -a;                            // log(1/alpha)
    double lgr, dgr;
    lgamma_digamma_pos(r, lgr, dgr);
    Sym<P> B, dB;
#pragma unroll
    for (int k = 0; k < P * (P + 1) / 2; k++) { B.v[k] = 0.0; dB.v[k] = 0.0; }
    double ll = 0.0, ds = 0.0;
    // (measured: unrolling this loop by 2 doubles the registers to 188 and is 24 % slower)
#pragma unroll 1
    for (int j = 0; j < S; j++) {
        const double yj = ys[j * stride], muj = mus[j * stride];
        const double ma = muj * alpha;
        const double ropm = rcp_pos(1.0 + ma);
        const double w = muj * ropm;                    // = 1 / (1/mu + alpha)
        const double dw = -w * w;
#pragma unroll
        for (int u = 0; u < P; u++)
#pragma unroll
            for (int v = 0; v <= u; v++) {
                const double xx = c_des.X[j * P + u] * c_des.X[j * P + v];
                B.v[u * (u + 1) / 2 + v] += w * xx;
                if (WANT_D) dB.v[u * (u + 1) / 2 + v] += dw * xx;
            }
        const double l1 = log_pos(1.0 + ma);
        double lg, dg;
        lgamma_digamma_pos(yj + r, lg, dg);
        // mu + r = r (1 + mu alpha): log(mu + r) = log r + log(1 + mu alpha), 1/(mu + r) = alpha / (1 + mu alpha)
        // for a zero count lg - lgr and dgr - dg are exactly zero (same instruction sequence, same input)
        ll += ((lg - lgr) - yj * (log_r + l1)) - r * l1;
        if (WANT_D) ds += ((dgr - dg) + (l1 - ma * ropm)) + yj * (alpha * ropm);
    }
    const double cr = -0.5 * chol_logdet<P>(B);
    double pr = 0.0;
    if (use_prior) {
        const double d = a - prior_mean;
        pr = -0.5 * d * d / prior_sigmasq;
    }
    lp_out = ll + pr + cr;
    if (WANT_D) {
        Sym<P> Bi;
        chol_inverse<P>(B, Bi);
        double tr = 0.0;
#pragma unroll
        for (int u = 0; u < P; u++)
#pragma unroll
            for (int v = 0; v < P; v++) tr += Bi.v[sidx<P>(u, v)] * dB.v[sidx<P>(v, u)];
        const double dcr = -0.5 * tr;
        const double dpr = use_prior ? -1.0 * (a - prior_mean) / prior_sigmasq : 0.0;
        dlp_out = ((r * r) * ds + dcr) * alpha + dpr;
    }
}

}  // namespace cd
