/*
 * r_glue.c -- .Call entry points that an R build of Chicdiff links against libchicdiff_b200.so.
 *
 * SEXP <-> pointer marshalling only: no arithmetic lives here (all of it is behind include/chicdiff_b200.h).
 * This file needs R's headers (R.h, Rinternals.h), which are not present in the build image, so it is not
 * compiled by __graft_entry__.build(); INTEGRATION.md shows the R CMD SHLIB line.  The Python ctypes binding
 * (chicdiff_b200/engine.py) exercises exactly the same C entry points in the tests.
 *
 * Replaces, inside DESeq2Wrap (Chicdiff/R/chicdiff.R:1494-1777):
 *   :1540-1547  fragData[, list(N = sum(N), ..., FullMean = sum(FullMean)), by = ...]   -> cdR_aggregate
 *   :1551-1674  DESeqDataSetFromMatrix / estimateSizeFactors / theta grid /
 *               estimateDispersions / nbinomWaldTest                                    -> cdR_region_test
 *   :1721-1739  results()                                                              -> cdR_results_resident,
 *                                                                                         cdR_results_adjust
 * and, inside IHWcorrection (:2038-2049), the weight application                         -> cdR_ihw_apply
 */
#include <R.h>
#include <Rinternals.h>
#include <stdint.h>
#include <string.h>
#include "../include/chicdiff_b200.h"

static void ctx_finalizer(SEXP ptr)
{
    cd_ctx* ctx = (cd_ctx*)R_ExternalPtrAddr(ptr);
    if (ctx) { cd_destroy(ctx); R_ClearExternalPtr(ptr); }
}

static cd_ctx* get_ctx(SEXP ptr)
{
    cd_ctx* ctx = (cd_ctx*)R_ExternalPtrAddr(ptr);
    if (!ctx) error("chicdiff_b200: context was destroyed");
    return ctx;
}

#define CD_CHECK(ctx, call) do { int rc_ = (call); if (rc_ != CD_OK) error("chicdiff_b200: %s", cd_last_error(ctx)); } while (0)

SEXP cdR_create(SEXP device)
{
    cd_ctx* ctx = NULL;
    if (cd_create(&ctx, asInteger(device)) != CD_OK) error("chicdiff_b200: %s", cd_last_error(NULL));
    SEXP ptr = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* X: S x p numeric matrix (column-major in R) */
SEXP cdR_set_design(SEXP ptr, SEXP X)
{
    cd_ctx* ctx = get_ctx(ptr);
    int S = nrows(X), p = ncols(X);
    double* rowmajor = (double*)R_alloc((size_t)S * p, sizeof(double));
    for (int j = 0; j < S; j++) for (int u = 0; u < p; u++) rowmajor[j * p + u] = REAL(X)[u * S + j];
    CD_CHECK(ctx, cd_set_design(ctx, S, p, rowmajor));
    return R_NilValue;
}

/* row_off: numeric vector of length n + 1 (doubles hold the 64-bit offsets exactly up to 2^53) */
SEXP cdR_set_regions(SEXP ptr, SEXP row_off)
{
    cd_ctx* ctx = get_ctx(ptr);
    R_xlen_t m = XLENGTH(row_off);
    int64_t* off = (int64_t*)R_alloc((size_t)m, sizeof(int64_t));
    for (R_xlen_t i = 0; i < m; i++) off[i] = (int64_t)REAL(row_off)[i];
    CD_CHECK(ctx, cd_set_regions(ctx, (int64_t)m - 1, off));
    return R_NilValue;
}

/* N: integer vector, fullmean: numeric vector (NA_real_ is a NaN: passed through) */
SEXP cdR_set_sample_rows(SEXP ptr, SEXP s, SEXP N, SEXP fullmean)
{
    cd_ctx* ctx = get_ctx(ptr);
    CD_CHECK(ctx, cd_set_sample_rows(ctx, asInteger(s) - 1, (int64_t)XLENGTH(N), (const int32_t*)INTEGER(N), REAL(fullmean)));
    return R_NilValue;
}

/* returns list(K = integer matrix n x S, FullMean = numeric matrix n x S); R's column-major n x S is the
 * library's sample-major layout, so no transposition happens */
SEXP cdR_aggregate(SEXP ptr, SEXP n_, SEXP S_)
{
    cd_ctx* ctx = get_ctx(ptr);
    int n = asInteger(n_), S = asInteger(S_);
    SEXP K = PROTECT(allocMatrix(INTSXP, n, S));
    SEXP FM = PROTECT(allocMatrix(REALSXP, n, S));
    CD_CHECK(ctx, cd_aggregate(ctx, (int32_t*)INTEGER(K), REAL(FM)));
    SEXP out = PROTECT(allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, K); SET_VECTOR_ELT(out, 1, FM);
    SEXP nm = PROTECT(allocVector(STRSXP, 2));
    SET_STRING_ELT(nm, 0, mkChar("K")); SET_STRING_ELT(nm, 1, mkChar("FullMean"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(4);
    return out;
}

/* norm: 0/1/2; theta, priorVar, priorVarGrid: NA_real_ = let the library decide; grid: numeric vector */
SEXP cdR_region_test(SEXP ptr, SEXP n_, SEXP S_, SEXP p_, SEXP norm, SEXP theta, SEXP grid, SEXP priorVar, SEXP priorVarGrid)
{
    cd_ctx* ctx = get_ctx(ptr);
    int n = asInteger(n_), S = asInteger(S_);
    (void)p_;
    cd_options opt;
    memset(&opt, 0, sizeof(opt));
    opt.norm = asInteger(norm);
    opt.theta = asReal(theta);                 /* NA_real_ is a NaN */
    opt.theta_grid = REAL(grid); opt.n_theta_grid = LENGTH(grid);
    opt.disp_prior_var = asReal(priorVar); opt.disp_prior_var_grid = asReal(priorVarGrid);
    cd_results res;
    memset(&res, 0, sizeof(res));
    const char* names[] = {"baseMean", "log2FoldChange", "lfcSE", "stat", "pvalue", "maxCooks", "dispGeneEst", "dispFit",
                           "dispMAP", "dispersion", "deviance", "flags", "theta", "deviances", "sizeFactors"};
    SEXP out = PROTECT(allocVector(VECSXP, 15));
    SEXP col[11];
    for (int k = 0; k < 11; k++) { col[k] = PROTECT(allocVector(REALSXP, n)); SET_VECTOR_ELT(out, k, col[k]); }
    res.baseMean = REAL(col[0]); res.log2FoldChange = REAL(col[1]); res.lfcSE = REAL(col[2]); res.stat = REAL(col[3]);
    res.pvalue = REAL(col[4]); res.maxCooks = REAL(col[5]); res.dispGeneEst = REAL(col[6]); res.dispFit = REAL(col[7]);
    res.dispMAP = REAL(col[8]); res.dispersion = REAL(col[9]); res.deviance = REAL(col[10]);
    SEXP flags = PROTECT(allocVector(RAWSXP, n));
    res.flags = RAW(flags);
    SET_VECTOR_ELT(out, 11, flags);
    CD_CHECK(ctx, cd_region_test(ctx, &opt, &res));
    SET_VECTOR_ELT(out, 12, ScalarReal(res.theta));
    SEXP dv = PROTECT(allocVector(REALSXP, res.n_deviances));
    for (int k = 0; k < res.n_deviances; k++) REAL(dv)[k] = res.deviances[k];
    SET_VECTOR_ELT(out, 13, dv);
    SEXP sf = PROTECT(allocVector(REALSXP, S));
    for (int k = 0; k < S; k++) REAL(sf)[k] = res.sizeFactors[k];
    SET_VECTOR_ELT(out, 14, sf);
    SEXP nm = PROTECT(allocVector(STRSXP, 15));
    for (int k = 0; k < 15; k++) SET_STRING_ELT(nm, k, mkChar(names[k]));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(16);
    return out;
}

/* returns list(pvalue, padj) after Cook's cutoff + independent filtering + BH */
SEXP cdR_results_adjust(SEXP S_, SEXP p_, SEXP baseMean, SEXP maxCooks, SEXP flags, SEXP pvalue)
{
    R_xlen_t n = XLENGTH(baseMean);
    SEXP pv = PROTECT(duplicate(pvalue));
    SEXP padj = PROTECT(allocVector(REALSXP, n));
    double sc[4];
    if (cd_results_adjust((int64_t)n, asInteger(S_), asInteger(p_), REAL(baseMean), REAL(maxCooks), RAW(flags), REAL(pv), REAL(padj), sc) != CD_OK)
        error("chicdiff_b200: cd_results_adjust: bad arguments");
    SEXP out = PROTECT(allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, pv); SET_VECTOR_ELT(out, 1, padj);
    SEXP nm = PROTECT(allocVector(STRSXP, 2));
    SET_STRING_ELT(nm, 0, mkChar("pvalue")); SET_STRING_ELT(nm, 1, mkChar("padj"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(4);
    return out;
}

/* results() on the arrays the last cdR_region_test left on the device (no columns cross the bus except the two
 * returned): list(pvalue, padj, filterThreshold) */
SEXP cdR_results_resident(SEXP ptr, SEXP n_)
{
    cd_ctx* ctx = get_ctx(ptr);
    R_xlen_t n = (R_xlen_t)asReal(n_);
    SEXP pv = PROTECT(allocVector(REALSXP, n));
    SEXP padj = PROTECT(allocVector(REALSXP, n));
    SEXP thr = PROTECT(allocVector(REALSXP, 1));
    double sc[4];
    CD_CHECK(ctx, cd_results_resident(ctx, REAL(pv), REAL(padj), sc));
    REAL(thr)[0] = sc[1];
    SEXP out = PROTECT(allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, pv); SET_VECTOR_ELT(out, 1, padj); SET_VECTOR_ELT(out, 2, thr);
    SEXP nm = PROTECT(allocVector(STRSXP, 3));
    SET_STRING_ELT(nm, 0, mkChar("pvalue")); SET_STRING_ELT(nm, 1, mkChar("padj")); SET_STRING_ELT(nm, 2, mkChar("filterThreshold"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}

/* IHWcorrection(), chicdiff.R:2038-2049: list(group, weight, weighted_pvalue, weighted_padj) in input row order.
 * NaN outputs where R has NA_real_ are turned into NA by the adapter (is.nan -> NA), group INT32_MIN is NA_integer_. */
SEXP cdR_ihw_apply(SEXP avDist, SEXP pvalue, SEXP minLogDist, SEXP maxLogDist, SEXP avWeights)
{
    R_xlen_t n = XLENGTH(avDist);
    int G = (int)XLENGTH(avWeights);
    SEXP group = PROTECT(allocVector(INTSXP, n));
    SEXP weight = PROTECT(allocVector(REALSXP, n));
    SEXP wp = PROTECT(allocVector(REALSXP, n));
    SEXP wpadj = PROTECT(allocVector(REALSXP, n));
    if (cd_ihw_apply((int64_t)n, REAL(avDist), REAL(pvalue), G, REAL(minLogDist), REAL(maxLogDist), REAL(avWeights),
                     INTEGER(group), REAL(weight), REAL(wp), REAL(wpadj)) != CD_OK)
        error("chicdiff_b200: cd_ihw_apply: bad arguments or 'breaks' are not unique");
    SEXP out = PROTECT(allocVector(VECSXP, 4));
    SET_VECTOR_ELT(out, 0, group); SET_VECTOR_ELT(out, 1, weight); SET_VECTOR_ELT(out, 2, wp); SET_VECTOR_ELT(out, 3, wpadj);
    SEXP nm = PROTECT(allocVector(STRSXP, 4));
    SET_STRING_ELT(nm, 0, mkChar("group")); SET_STRING_ELT(nm, 1, mkChar("weight"));
    SET_STRING_ELT(nm, 2, mkChar("weighted_pvalue")); SET_STRING_ELT(nm, 3, mkChar("weighted_padj"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(6);
    return out;
}
