// wald.cu -- stages 3 and 5: batched FP64 negative-binomial GLM (IRLS), hat diagonals,
// Cook's distances, Wald statistic and p-value.
//
// Restates what nbinomWaldTest() does when Chicdiff calls it (chicdiff.R:1574,1603,1644,1674):
// fitNbinomGLMs -> DESeq2.cpp fitBeta (ridge lambda = 1e-6 / ln(2)^2, mu floored at 0.5,
// convergence on the relative deviance change, |beta| > 30 => failure), the intercept-only
// shortcut for design ~ 1 (the theta grid), calculateCooksDistance / recordMaxCooks, and
// p = 2 * pnorm(-|beta / SE|).
//
// Mapping: one thread per region; y_j, nf_j and the mu-independent part of the NB log density
// are staged per thread in conflict-free shared-memory columns.  The ridge normal equations
// (X'WX + lambda I) beta = X'Wz are solved by an unrolled Cholesky (p <= 4); DESeq2 solves the
// same system by QR of the row-augmented matrix.
#include "kernels.h"

namespace cd {

__constant__ CdDesign c_desw;

cudaError_t set_design_wald(const CdDesign& d, cudaStream_t st)
{
    return cudaMemcpyToSymbolAsync(c_desw, &d, sizeof(CdDesign), 0, cudaMemcpyHostToDevice, st);
}

constexpr int kWaldThreads = 128;

// R mean(x, trim): sort, drop floor(n * trim) from each end
__device__ __forceinline__ double trimmed_mean_dev(double* v, int n, double trim)
{
    for (int a = 1; a < n; a++) {
        const double x = v[a];
        int b = a - 1;
        while (b >= 0 && v[b] > x) { v[b + 1] = v[b]; b--; }
        v[b + 1] = x;
    }
    const int lo = (int)floor((double)n * trim);
    double s = 0.0;
    for (int a = lo; a < n - lo; a++) s += v[a];
    return s / (double)(n - 2 * lo);
}

__device__ __forceinline__ int trim_bin(int n) { return n <= 3 ? 0 : (n <= 23 ? 1 : 2); }

// one sweep over the samples at coefficients beta: X'WX (packed, no ridge), X'Wz and the deviance
template <int P>
__device__ __forceinline__ void irls_pass(const double* beta, double alpha, double size, int S, int stride,
                                          const double* ys, const double* nfs, const double* cs, const int* kinds,
                                          Sym<P>& A, double* b, double& dev)
{
#pragma unroll
    for (int k = 0; k < P * (P + 1) / 2; k++) A.v[k] = 0.0;
#pragma unroll
    for (int u = 0; u < P; u++) b[u] = 0.0;
    dev = 0.0;
#pragma unroll 1
    for (int j = 0; j < S; j++) {
        const double yj = ys[j * stride], nfj = nfs[j * stride];
        double eta = 0.0;
#pragma unroll
        for (int u = 0; u < P; u++) eta += c_desw.X[j * P + u] * beta[u];
        double mu = nfj * exp(eta);
        double lmn = eta;
        if (!(mu >= kMinMu)) { mu = kMinMu; lmn = log(kMinMu / nfj); }     // fmax(mu, minmu)
        NbConst kc; kc.c = cs[j * stride]; kc.kind = kinds[j * stride];
        dev += -2.0 * nb_var(yj, size, mu, kc);
        const double w = mu / (1.0 + alpha * mu);
        const double z = lmn + (yj - mu) / mu;
#pragma unroll
        for (int u = 0; u < P; u++) {
            const double xu = c_desw.X[j * P + u];
            b[u] += w * z * xu;
#pragma unroll
            for (int v = 0; v <= u; v++) A.v[u * (u + 1) / 2 + v] += w * xu * c_desw.X[j * P + v];
        }
    }
}

template <int P>
__global__ void __launch_bounds__(kWaldThreads)
wald_kernel(int64_t n, int S, const int32_t* __restrict__ K, const double* __restrict__ nf,
            const double* __restrict__ dispersion, uint8_t* __restrict__ flags,
            double* __restrict__ beta_out, double* __restrict__ se_out, double* __restrict__ stat_out,
            double* __restrict__ pvalue_out, double* __restrict__ deviance_out, double* __restrict__ maxCooks_out,
            int32_t* __restrict__ betaIter_out, double* __restrict__ mu_out)
{
    extern __shared__ double smem[];
    const int stride = kWaldThreads;
    double* ys = smem + threadIdx.x;
    double* nfs = smem + (size_t)S * stride + threadIdx.x;
    double* cs = smem + (size_t)2 * S * stride + threadIdx.x;
    int* kinds = reinterpret_cast<int*>(smem + (size_t)3 * S * stride) + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool mu_only = (mu_out != nullptr);
    if (flags[i] & CD_FLAG_ALLZERO) {
        if (mu_only) { for (int j = 0; j < S; j++) mu_out[(int64_t)j * n + i] = NAN; return; }
        for (int u = 0; u < P; u++) { beta_out[(int64_t)u * n + i] = NAN; se_out[(int64_t)u * n + i] = NAN; }
        stat_out[i] = NAN; pvalue_out[i] = NAN; deviance_out[i] = NAN; betaIter_out[i] = 0;
        if (maxCooks_out) maxCooks_out[i] = NAN;
        return;
    }
    const double alpha = dispersion[i];
    const double size = 1.0 / alpha;
    double qsum = 0.0;
    for (int j = 0; j < S; j++) {
        const double yj = (double)K[(int64_t)j * n + i];
        const double nfj = nf[(int64_t)j * n + i];
        ys[j * stride] = yj;
        nfs[j * stride] = nfj;
        const NbConst kc = nb_const(yj, size);
        cs[j * stride] = kc.c;
        kinds[j * stride] = kc.kind;
        qsum += yj / nfj;
    }
    double beta[P], se[P];
    Sym<P> A;
    double rhs[P];
    double dev = 0.0;
    int iter = 0;
    bool noconv = false;
    const double lambda = 1e-6 / (kLn2 * kLn2);
    if (P == 1) {
        // fitNbinomGLMs intercept-only shortcut: beta = log2(mean normalised count)
        beta[0] = log(qsum / S);
        iter = 1;
    } else {
        // start: least squares of log(q + 0.1) on X
#pragma unroll
        for (int u = 0; u < P; u++) beta[u] = 0.0;
        for (int j = 0; j < S; j++) {
            const double l = log(ys[j * stride] / nfs[j * stride] + 0.1);
#pragma unroll
            for (int u = 0; u < P; u++) beta[u] += c_desw.ls[u * S + j] * l;
        }
        double dev_old = 0.0, dev_new;
        irls_pass<P>(beta, alpha, size, S, stride, ys, nfs, cs, kinds, A, rhs, dev_new);
        for (int t = 0; t < 100; t++) {
            iter++;
            Sym<P> Ar = A;
#pragma unroll
            for (int u = 0; u < P; u++) Ar.v[u * (u + 1) / 2 + u] += lambda;
            chol_logdet<P>(Ar);
            double bn[P];
#pragma unroll
            for (int u = 0; u < P; u++) bn[u] = rhs[u];
            chol_solve<P>(Ar, bn);
            bool big = false;
#pragma unroll
            for (int u = 0; u < P; u++) { beta[u] = bn[u]; big = big || (fabs(bn[u]) > 30.0); }
            if (big) { iter = 100; break; }
            irls_pass<P>(beta, alpha, size, S, stride, ys, nfs, cs, kinds, A, rhs, dev_new);
            dev = dev_new;
            const double conv_test = fabs(dev - dev_old) / (fabs(dev) + 0.1);
            if (isnan(conv_test)) { iter = 100; break; }
            if (t > 0 && conv_test < 1e-8) break;
            dev_old = dev;
        }
        noconv = !(iter < 100);
    }
    if (mu_only) {
        for (int j = 0; j < S; j++) {
            double eta = 0.0;
#pragma unroll
            for (int u = 0; u < P; u++) eta += c_desw.X[j * P + u] * beta[u];
            mu_out[(int64_t)j * n + i] = fmax(nfs[j * stride] * exp(eta), kMinMu);
        }
        return;
    }
    // covariance, hat diagonals, log likelihood at the unclamped mu, Cook's distance
    Sym<P> Ari;
    if (P == 1) {
        double sw = 0.0;
        const double m0 = exp(beta[0]);                  // = 2^log2(mean q)
        for (int j = 0; j < S; j++) sw += 1.0 / (1.0 / (nfs[j * stride] * m0) + alpha);
        A.v[0] = sw;
        Ari.v[0] = 1.0 / sw;
        se[0] = kLog2e * sqrt(1.0 / sw);
    } else {
        Sym<P> Ar = A;
#pragma unroll
        for (int u = 0; u < P; u++) Ar.v[u * (u + 1) / 2 + u] += lambda;
        chol_logdet<P>(Ar);
        chol_inverse<P>(Ar, Ari);
        bool bad = false;
#pragma unroll
        for (int u = 0; u < P; u++) {
            // sigma_uu = sum_kl Ari[u][k] A[k][l] Ari[l][u]
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < P; k++)
#pragma unroll
                for (int l = 0; l < P; l++) s += Ari.v[sidx<P>(u, k)] * A.v[sidx<P>(k, l)] * Ari.v[sidx<P>(l, u)];
            se[u] = kLog2e * sqrt(fmax(s, 0.0));
            bad = bad || !(s > 0.0) || isnan(beta[u]);
        }
        noconv = noconv || bad;
    }
    double loglike = 0.0;
    // robust method-of-moments dispersion for Cook's distance
    const bool want_cooks = (maxCooks_out != nullptr);
    double ar = 0.0;
    if (want_cooks) {
        double v;
        double tmp[CD_MAXS];
        if (c_desw.any3) {
            v = -INFINITY;
            for (int c = 0; c < c_desw.ncell; c++) {
                const int nc = c_desw.cell_size[c];
                if (nc < 3) continue;
                const double trimr = (trim_bin(nc) == 0) ? 1.0 / 3.0 : (trim_bin(nc) == 1 ? 1.0 / 4.0 : 1.0 / 8.0);
                const double scalec = (trim_bin(nc) == 0) ? 2.04 : (trim_bin(nc) == 1 ? 1.86 : 1.51);
                int k = 0;
                for (int j = 0; j < S; j++) if (c_desw.cell[j] == c) tmp[k++] = ys[j * stride] / nfs[j * stride];
                const double cm = trimmed_mean_dev(tmp, nc, trimr);
                k = 0;
                for (int j = 0; j < S; j++) if (c_desw.cell[j] == c) {
                    const double d = ys[j * stride] / nfs[j * stride] - cm;
                    tmp[k++] = d * d;
                }
                const double ve = scalec * trimmed_mean_dev(tmp, nc, trimr);
                if (ve > v) v = ve;
            }
        } else {
            for (int j = 0; j < S; j++) tmp[j] = ys[j * stride] / nfs[j * stride];
            const double rm = trimmed_mean_dev(tmp, S, 1.0 / 8.0);
            for (int j = 0; j < S; j++) { const double d = ys[j * stride] / nfs[j * stride] - rm; tmp[j] = d * d; }
            v = 1.51 * trimmed_mean_dev(tmp, S, 1.0 / 8.0);
        }
        const double mq = qsum / S;
        ar = fmax((v - mq) / (mq * mq), 0.04);
    }
    double mc = -INFINITY, ck_best = -INFINITY;
    double y_best = 0.0;
    for (int j = 0; j < S; j++) {
        const double yj = ys[j * stride], nfj = nfs[j * stride];
        double eta = 0.0;
#pragma unroll
        for (int u = 0; u < P; u++) eta += c_desw.X[j * P + u] * beta[u];
        const double muw = nfj * exp(eta);                      // unclamped, as stored by nbinomWaldTest
        NbConst kc; kc.c = cs[j * stride]; kc.kind = kinds[j * stride];
        loglike += nb_var(yj, size, muw, kc);
        if (want_cooks) {
            // hat diagonal from the weights at the floored mu (fitBeta), or at mu itself (p = 1 shortcut)
            const double muc = (P == 1) ? muw : fmax(muw, kMinMu);
            const double w = (P == 1) ? 1.0 / (1.0 / muc + alpha) : muc / (1.0 + alpha * muc);
            double h = 0.0;
#pragma unroll
            for (int u = 0; u < P; u++)
#pragma unroll
                for (int v2 = 0; v2 < P; v2++)
                    h += c_desw.X[j * P + u] * Ari.v[sidx<P>(u, v2)] * c_desw.X[j * P + v2];
            h *= w;
            const double V = muw + ar * muw * muw;
            const double ck = (yj - muw) * (yj - muw) / V / (double)P * h / ((1.0 - h) * (1.0 - h));
            if (c_desw.cell_size[c_desw.cell[j]] >= 3 && ck > mc) mc = ck;
            if (ck > ck_best) { ck_best = ck; y_best = yj; }        // which.max: first maximum
        }
    }
    uint8_t f = flags[i];
    if (noconv) f |= CD_FLAG_BETA_NOCONV;
    if (want_cooks) {
        int greater = 0;
        for (int j = 0; j < S; j++) greater += (ys[j * stride] > y_best);
        if (greater >= 3) f |= CD_FLAG_COOKS_KEEP;
        maxCooks_out[i] = (S > P && c_desw.any3) ? mc : NAN;
    }
    flags[i] = f;
#pragma unroll
    for (int u = 0; u < P; u++) {
        beta_out[(int64_t)u * n + i] = kLog2e * beta[u];
        se_out[(int64_t)u * n + i] = se[u];
    }
    const double st = (kLog2e * beta[P - 1]) / se[P - 1];
    stat_out[i] = st;
    pvalue_out[i] = erfc(fabs(st) * 0.70710678118654752440);
    deviance_out[i] = -2.0 * loglike;
    betaIter_out[i] = iter;
}

cudaError_t launch_wald(int64_t n, int S, int p, const int32_t* K, const double* nf, const double* dispersion,
                        uint8_t* flags, double* beta, double* betaSE, double* stat, double* pvalue, double* deviance,
                        double* maxCooks, int32_t* betaIter, double* mu_out, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    const int threads = kWaldThreads;
    const int blocks = (int)((n + threads - 1) / threads);
    const size_t smem = (size_t)S * threads * (3 * sizeof(double) + sizeof(int));
    cudaError_t e;
#define CD_LAUNCH(P_)                                                                                               \
    e = cudaFuncSetAttribute(wald_kernel<P_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    if (e != cudaSuccess) return e;                                                                                 \
    wald_kernel<P_><<<blocks, threads, smem, st>>>(n, S, K, nf, dispersion, flags, beta, betaSE, stat, pvalue,      \
                                                   deviance, maxCooks, betaIter, mu_out)
    switch (p) {
        case 1: CD_LAUNCH(1); break;
        case 2: CD_LAUNCH(2); break;
        case 3: CD_LAUNCH(3); break;
        case 4: CD_LAUNCH(4); break;
        default: return cudaErrorInvalidValue;
    }
#undef CD_LAUNCH
    return cudaGetLastError();
}

}  // namespace cd
