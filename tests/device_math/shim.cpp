// extern "C" wrappers around the device math of chicdiff_b200/csrc/common.cuh, compiled for the host (see cuda_runtime.h here)
#include "common.cuh"

namespace cd { CdDesign c_des; }      // host stand-in for the constant-memory design of dispersion.cu
#include "posterior.cuh"
#include "../../experiments/posterior_v2.cuh"
#include "../../experiments/log_v2.cuh"
#include "../../experiments/posterior_v3.cuh"

extern "C" {
double dm_rcp_pos(double x) { return cd::rcp_pos(x); }
double dm_log_pos(double x) { return cd::log_pos(x); }
void dm_lgamma_digamma_pos(double x, double* lg, double* dg) { cd::lgamma_digamma_pos(x, *lg, *dg); }
double dm_lgamma_c_pos(double x) { return cd::lgamma_c_pos(x); }
double dm_trigamma_pos(double x) { return cd::trigamma_pos(x); }
double dm_dnbinom_mu_log(double y, double size, double mu) { return cd::dnbinom_mu_log(y, size, mu); }
double dm_chol_logdet2(double a00, double a10, double a11)
{
    cd::Sym<2> A;
    A.v[0] = a00; A.v[1] = a10; A.v[2] = a11;
    return cd::chol_logdet<2>(A);
}
void dm_vec(int what, long n, const double* x, double* out, double* out2)
{
    for (long i = 0; i < n; i++) {
        if (what == 10) { out[i] = cd::log_pos_v2(x[i], cd::kLogTab); continue; }      // experiments/log_v2.cuh
        if (what == 0) out[i] = cd::log_pos(x[i]);
        else if (what == 1) out[i] = cd::rcp_pos(x[i]);
        else if (what == 2) cd::lgamma_digamma_pos(x[i], out[i], out2[i]);
        else if (what == 3) out[i] = cd::lgamma_c_pos(x[i]);
        else if (what == 4) out[i] = cd::trigamma_pos(x[i]);
    }
}
// the fused Cox-Reid posterior and derivative of the line search, at log(alpha) = a, for one region
void dm_eval_post(int S, int p, const double* X, const double* y, const double* mu, double a, double prior_mean,
                  double prior_sigmasq, int use_prior, double* lp, double* dlp)
{
    cd::c_des.S = S; cd::c_des.p = p;
    for (int k = 0; k < S * p; k++) cd::c_des.X[k] = X[k];
    if (p == 1) cd::eval_post<1, true>(a, y, mu, 1, S, prior_mean, prior_sigmasq, use_prior != 0, *lp, *dlp);
    else if (p == 2) cd::eval_post<2, true>(a, y, mu, 1, S, prior_mean, prior_sigmasq, use_prior != 0, *lp, *dlp);
    else if (p == 3) cd::eval_post<3, true>(a, y, mu, 1, S, prior_mean, prior_sigmasq, use_prior != 0, *lp, *dlp);
    else cd::eval_post<4, true>(a, y, mu, 1, S, prior_mean, prior_sigmasq, use_prior != 0, *lp, *dlp);
}
// experiments/posterior_v2.cuh: the lighter formulation prepared for the next round, same contract as dm_eval_post
void dm_eval_post_v2(int S, int p, const double* X, const double* y, const double* mu, double a, double prior_mean,
                     double prior_sigmasq, int use_prior, double* lp, double* dlp)
{
    cd::c_des.S = S; cd::c_des.p = p;
    for (int k = 0; k < S * p; k++) cd::c_des.X[k] = X[k];
    if (p == 1) cd::eval_post_v2<1, true>(a, y, mu, 1, S, prior_mean, prior_sigmasq, use_prior != 0, *lp, *dlp);
    else if (p == 2) cd::eval_post_v2<2, true>(a, y, mu, 1, S, prior_mean, prior_sigmasq, use_prior != 0, *lp, *dlp);
    else if (p == 3) cd::eval_post_v2<3, true>(a, y, mu, 1, S, prior_mean, prior_sigmasq, use_prior != 0, *lp, *dlp);
    else cd::eval_post_v2<4, true>(a, y, mu, 1, S, prior_mean, prior_sigmasq, use_prior != 0, *lp, *dlp);
}
// experiments/posterior_v3.cuh (p <= 2): v2 with the table-assisted logarithm
void dm_eval_post_v3(int S, int p, const double* X, const double* y, const double* mu, double a, double prior_mean,
                     double prior_sigmasq, int use_prior, double* lp, double* dlp)
{
    cd::c_des.S = S; cd::c_des.p = p;
    for (int k = 0; k < S * p; k++) cd::c_des.X[k] = X[k];
    if (p == 1) cd::eval_post_v3<1, true>(a, y, mu, 1, S, prior_mean, prior_sigmasq, use_prior != 0, cd::kLogTab, *lp, *dlp);
    else cd::eval_post_v3<2, true>(a, y, mu, 1, S, prior_mean, prior_sigmasq, use_prior != 0, cd::kLogTab, *lp, *dlp);
}
void dm_dnbinom_vec(long n, const double* y, const double* size, const double* mu, double* out)
{
    for (long i = 0; i < n; i++) out[i] = cd::dnbinom_mu_log(y[i], size[i], mu[i]);
}
}
