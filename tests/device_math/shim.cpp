// extern "C" wrappers around the device math of chicdiff_b200/csrc/common.cuh, compiled for the host (see cuda_runtime.h here)
#include "common.cuh"

#include "posterior.cuh"

// the fused Cox-Reid posterior and derivative of the line search, at log(alpha) = a, for one region.
// variant 0: the round-1 formulation (eval_post_ref); 1: pairwise logarithms + closed-form Cox-Reid term with log_pos;
// 2: the same with the table-assisted logarithm (what the line-search kernel runs)
template <int P>
static void eval_variant(int variant, int S, const double* X, const double* y, const double* mu, double a, double prior_mean,
                         double prior_sigmasq, bool use_prior, double* lp, double* dlp)
{
    if (variant == 0) cd::eval_post_ref<P, true>(a, y, mu, 1, S, X, prior_mean, prior_sigmasq, use_prior, *lp, *dlp);
    else if (variant == 1) cd::eval_post<P, true, false>(a, y, mu, 1, S, X, cd::kLogTab, prior_mean, 1.0 / prior_sigmasq, use_prior, *lp, *dlp);
    else cd::eval_post<P, true, true>(a, y, mu, 1, S, X, cd::kLogTab, prior_mean, 1.0 / prior_sigmasq, use_prior, *lp, *dlp);
}
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
        if (what == 10) { out[i] = cd::log_pos_v2(x[i], cd::kLogTab); continue; }      // the table-assisted logarithm
        if (what == 11) { out[i] = cd::exp_mid(x[i]); continue; }                        // the line search's exponential
        if (what == 0) out[i] = cd::log_pos(x[i]);
        else if (what == 1) out[i] = cd::rcp_pos(x[i]);
        else if (what == 2) cd::lgamma_digamma_pos(x[i], out[i], out2[i]);
        else if (what == 3) out[i] = cd::lgamma_c_pos(x[i]);
        else if (what == 4) out[i] = cd::trigamma_pos(x[i]);
    }
}
void dm_eval_post_variant(int variant, int S, int p, const double* X, const double* y, const double* mu, double a,
                          double prior_mean, double prior_sigmasq, int use_prior, double* lp, double* dlp)
{
    if (p == 1) eval_variant<1>(variant, S, X, y, mu, a, prior_mean, prior_sigmasq, use_prior != 0, lp, dlp);
    else if (p == 2) eval_variant<2>(variant, S, X, y, mu, a, prior_mean, prior_sigmasq, use_prior != 0, lp, dlp);
    else if (p == 3) eval_variant<3>(variant, S, X, y, mu, a, prior_mean, prior_sigmasq, use_prior != 0, lp, dlp);
    else eval_variant<4>(variant, S, X, y, mu, a, prior_mean, prior_sigmasq, use_prior != 0, lp, dlp);
}
void dm_eval_post(int S, int p, const double* X, const double* y, const double* mu, double a, double prior_mean,
                  double prior_sigmasq, int use_prior, double* lp, double* dlp)
{
    dm_eval_post_variant(2, S, p, X, y, mu, a, prior_mean, prior_sigmasq, use_prior, lp, dlp);
}
void dm_dnbinom_vec(long n, const double* y, const double* size, const double* mu, double* out)
{
    for (long i = 0; i < n; i++) out[i] = cd::dnbinom_mu_log(y[i], size[i], mu[i]);
}
}
