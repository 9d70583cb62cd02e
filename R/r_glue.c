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
 * and, inside IHWcorrection (:2038-2049), the weight application                         -> cdR_ihw_apply,
 *                                                                                         cdR_ihw_apply_device
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

/* the sizes the context holds: outputs are allocated from these, and a caller whose idea of n / S / R differs is told */
static void check_dims(cd_ctx* ctx, const char* who, long long n_given, int S_given, long long R_given)
{
    int64_t n = 0, R = 0;
    int S = 0, p = 0;
    if (cd_get_dims(ctx, &n, &S, &p, &R) != CD_OK) error("chicdiff_b200: %s: bad context", who);
    if (n_given >= 0 && n_given != (long long)n) error("chicdiff_b200: %s: n = %lld but the context holds %lld regions", who, n_given, (long long)n);
    if (S_given >= 0 && S_given != S) error("chicdiff_b200: %s: S = %d but the context's design has %d samples", who, S_given, S);
    if (R_given >= 0 && R_given != (long long)R) error("chicdiff_b200: %s: R = %lld but the context holds %lld region rows", who, R_given, (long long)R);
}

/* returns list(K = integer matrix n x S, FullMean = numeric matrix n x S); R's column-major n x S is the
 * library's sample-major layout, so no transposition happens */
SEXP cdR_aggregate(SEXP ptr, SEXP n_, SEXP S_)
{
    cd_ctx* ctx = get_ctx(ptr);
    int n = asInteger(n_), S = asInteger(S_);
    check_dims(ctx, "cdR_aggregate", n, S, -1);
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

/* cd_options.prior_var_fn -> an R closure function(df, resid): called on the R thread that is inside .Call, so
 * evaluating R code here is allowed; an R error inside it longjmps out through cd_region_test, which holds no locks
 * and whose device buffers belong to the context (the next call reuses them). */
static double prior_var_trampoline(void* user, int df, int64_t n_resid, const double* resid)
{
    SEXP fn = (SEXP)user;
    SEXP r = PROTECT(allocVector(REALSXP, (R_xlen_t)n_resid));
    memcpy(REAL(r), resid, sizeof(double) * (size_t)n_resid);
    SEXP d = PROTECT(ScalarInteger(df));
    SEXP call = PROTECT(lang3(fn, d, r));
    /* an R error must not longjmp through the C++ frames of cd_region_test (they own heap buffers): evaluate under
     * R_tryEvalSilent and report failure as NaN, which the library turns into CD_ENUMERIC */
    int failed = 0;
    SEXP val = R_tryEvalSilent(call, R_GlobalEnv, &failed);
    double v = NA_REAL;
    if (!failed) { PROTECT(val); v = asReal(val); UNPROTECT(1); }
    UNPROTECT(3);
    return v;
}


/* norm: 0/1/2; theta, priorVar, priorVarGrid: NA_real_ = let the library decide; grid: numeric vector;
 * priorVarFn: NULL or function(df, resid) returning dispPriorVar for designs with S - p <= 3.
 * run: cd_region_test on a context or cd_multi_region_test on a multi-GPU handle (same contract). */
typedef int (*region_test_fn)(void* handle, const cd_options* opt, cd_results* out);
static int run_single(void* h, const cd_options* opt, cd_results* out) { return cd_region_test((cd_ctx*)h, opt, out); }
static int run_multi(void* h, const cd_options* opt, cd_results* out) { return cd_multi_region_test((cd_multi*)h, opt, out); }

static SEXP region_test_common(region_test_fn run, void* handle, const char* (*last_error)(void*), int n, int S, SEXP norm, SEXP theta,
                               SEXP grid, SEXP priorVar, SEXP priorVarGrid, SEXP priorVarFn)
{
    cd_options opt;
    memset(&opt, 0, sizeof(opt));              /* trend_a0 / trend_a1 / var_log_disp = 0: estimate */
    opt.norm = asInteger(norm);
    opt.theta = asReal(theta);                 /* NA_real_ is a NaN */
    opt.theta_grid = REAL(grid); opt.n_theta_grid = LENGTH(grid);
    opt.disp_prior_var = asReal(priorVar); opt.disp_prior_var_grid = asReal(priorVarGrid);
    if (!isNull(priorVarFn)) { opt.prior_var_fn = prior_var_trampoline; opt.prior_var_user = (void*)priorVarFn; }
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
    if (run(handle, &opt, &res) != CD_OK) error("chicdiff_b200: %s", last_error(handle));
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

static const char* last_error_single(void* h) { return cd_last_error((cd_ctx*)h); }
static const char* last_error_multi(void* h) { return cd_multi_last_error((cd_multi*)h); }

SEXP cdR_region_test(SEXP ptr, SEXP n_, SEXP S_, SEXP p_, SEXP norm, SEXP theta, SEXP grid, SEXP priorVar, SEXP priorVarGrid,
                     SEXP priorVarFn)
{
    cd_ctx* ctx = get_ctx(ptr);
    int n = asInteger(n_), S = asInteger(S_);
    (void)p_;
    check_dims(ctx, "cdR_region_test", n, S, -1);
    return region_test_common(run_single, ctx, last_error_single, n, S, norm, theta, grid, priorVar, priorVarGrid, priorVarFn);
}

/* ---- several GPUs from this one R session (cd_multi_*): same calls on the whole problem ---- */
static void multi_finalizer(SEXP ptr)
{
    cd_multi* m = (cd_multi*)R_ExternalPtrAddr(ptr);
    if (m) { cd_multi_destroy(m); R_ClearExternalPtr(ptr); }
}

static cd_multi* get_multi(SEXP ptr)
{
    cd_multi* m = (cd_multi*)R_ExternalPtrAddr(ptr);
    if (!m) error("chicdiff_b200: multi-GPU handle was destroyed");
    return m;
}

#define CD_MCHECK(m, call) do { int rc_ = (call); if (rc_ != CD_OK) error("chicdiff_b200: %s", cd_multi_last_error(m)); } while (0)

SEXP cdR_multi_create(SEXP n_gpus)
{
    cd_multi* m = NULL;
    if (cd_multi_create(&m, asInteger(n_gpus), NULL) != CD_OK) error("chicdiff_b200: %s", cd_multi_last_error(NULL));
    SEXP ptr = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, multi_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

SEXP cdR_multi_set_design(SEXP ptr, SEXP X)
{
    cd_multi* m = get_multi(ptr);
    int S = nrows(X), p = ncols(X);
    double* rowmajor = (double*)R_alloc((size_t)S * p, sizeof(double));
    for (int j = 0; j < S; j++) for (int u = 0; u < p; u++) rowmajor[j * p + u] = REAL(X)[u * S + j];
    CD_MCHECK(m, cd_multi_set_design(m, S, p, rowmajor));
    return R_NilValue;
}

/* row_off: numeric n + 1; region_bait: integer n (baitID of every region, regions in regionID order) */
SEXP cdR_multi_set_regions(SEXP ptr, SEXP row_off, SEXP region_bait)
{
    cd_multi* m = get_multi(ptr);
    R_xlen_t len = XLENGTH(row_off);
    if (XLENGTH(region_bait) != len - 1) error("chicdiff_b200: cdR_multi_set_regions: one baitID per region");
    int64_t* off = (int64_t*)R_alloc((size_t)len, sizeof(int64_t));
    for (R_xlen_t i = 0; i < len; i++) off[i] = (int64_t)REAL(row_off)[i];
    CD_MCHECK(m, cd_multi_set_regions(m, (int64_t)len - 1, off, (const int32_t*)INTEGER(region_bait)));
    return R_NilValue;
}

SEXP cdR_multi_set_sample_rows(SEXP ptr, SEXP s, SEXP N, SEXP fullmean)
{
    cd_multi* m = get_multi(ptr);
    CD_MCHECK(m, cd_multi_set_sample_rows(m, asInteger(s) - 1, (int64_t)XLENGTH(N), (const int32_t*)INTEGER(N), REAL(fullmean)));
    return R_NilValue;
}

SEXP cdR_multi_aggregate(SEXP ptr, SEXP n_, SEXP S_)
{
    cd_multi* m = get_multi(ptr);
    int n = asInteger(n_), S = asInteger(S_);
    SEXP K = PROTECT(allocMatrix(INTSXP, n, S));
    SEXP FM = PROTECT(allocMatrix(REALSXP, n, S));
    CD_MCHECK(m, cd_multi_aggregate(m, (int32_t*)INTEGER(K), REAL(FM)));
    SEXP out = PROTECT(allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, K); SET_VECTOR_ELT(out, 1, FM);
    SEXP nm = PROTECT(allocVector(STRSXP, 2));
    SET_STRING_ELT(nm, 0, mkChar("K")); SET_STRING_ELT(nm, 1, mkChar("FullMean"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(4);
    return out;
}

SEXP cdR_multi_region_test(SEXP ptr, SEXP n_, SEXP S_, SEXP norm, SEXP theta, SEXP grid, SEXP priorVar, SEXP priorVarGrid)
{
    cd_multi* m = get_multi(ptr);
    return region_test_common(run_multi, m, last_error_multi, asInteger(n_), asInteger(S_), norm, theta, grid, priorVar, priorVarGrid,
                              R_NilValue);
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
    check_dims(ctx, "cdR_results_resident", (long long)n, -1, -1);
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

/* The same on the context's GPU (cd_ihw_apply_device).  avDist = NULL: the avDist column cd_assemble left on the device. */
SEXP cdR_ihw_apply_device(SEXP ptr, SEXP avDist, SEXP pvalue, SEXP minLogDist, SEXP maxLogDist, SEXP avWeights)
{
    cd_ctx* ctx = get_ctx(ptr);
    R_xlen_t n = XLENGTH(pvalue);
    int G = (int)XLENGTH(avWeights);
    if (!isNull(avDist) && XLENGTH(avDist) != n) error("chicdiff_b200: avDist and pvalue differ in length");
    if (XLENGTH(minLogDist) != G || XLENGTH(maxLogDist) != G) error("chicdiff_b200: the distance lookup must have one row per weight");
    SEXP group = PROTECT(allocVector(INTSXP, n));
    SEXP weight = PROTECT(allocVector(REALSXP, n));
    SEXP wp = PROTECT(allocVector(REALSXP, n));
    SEXP wpadj = PROTECT(allocVector(REALSXP, n));
    CD_CHECK(ctx, cd_ihw_apply_device(ctx, (int64_t)n, isNull(avDist) ? NULL : REAL(avDist), REAL(pvalue), G, REAL(minLogDist),
                                      REAL(maxLogDist), REAL(avWeights), INTEGER(group), REAL(weight), REAL(wp), REAL(wpadj)));
    SEXP out = PROTECT(allocVector(VECSXP, 4));
    SET_VECTOR_ELT(out, 0, group); SET_VECTOR_ELT(out, 1, weight); SET_VECTOR_ELT(out, 2, wp); SET_VECTOR_ELT(out, 3, wpadj);
    SEXP nm = PROTECT(allocVector(STRSXP, 4));
    SET_STRING_ELT(nm, 0, mkChar("group")); SET_STRING_ELT(nm, 1, mkChar("weight"));
    SET_STRING_ELT(nm, 2, mkChar("weighted_pvalue")); SET_STRING_ELT(nm, 3, mkChar("weighted_padj"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(6);
    return out;
}

/* ---- per-replicate assembly (getFullRegionData1, chicdiff.R:609-702, 820-910) and its neighbours ---- */

static int64_t* offsets_from_real(SEXP v)
{
    R_xlen_t m = XLENGTH(v);
    int64_t* off = (int64_t*)R_alloc((size_t)m, sizeof(int64_t));
    for (R_xlen_t i = 0; i < m; i++) off[i] = (int64_t)REAL(v)[i];
    return off;
}

/* chr: integer codes of the rmap's chromosome column; start/end integer; id0 = first fragment ID (IDs contiguous) */
SEXP cdR_set_rmap(SEXP ptr, SEXP chr, SEXP start, SEXP end, SEXP id0)
{
    cd_ctx* ctx = get_ctx(ptr);
    CD_CHECK(ctx, cd_set_rmap(ctx, (int64_t)XLENGTH(chr), asInteger(id0), INTEGER(chr), INTEGER(start), INTEGER(end)));
    return R_NilValue;
}

/* getRegionUniverse(): peaks (baitID, otherEndID) -> list(row_off (numeric, m + 1), baitID, otherEndID) */
SEXP cdR_region_universe(SEXP ptr, SEXP peak_bait, SEXP peak_oe, SEXP ru_expand)
{
    cd_ctx* ctx = get_ctx(ptr);
    R_xlen_t m = XLENGTH(peak_bait);
    int64_t R = 0;
    CD_CHECK(ctx, cd_region_universe(ctx, (int64_t)m, INTEGER(peak_bait), INTEGER(peak_oe), asInteger(ru_expand), &R));
    int64_t* off = (int64_t*)R_alloc((size_t)m + 1, sizeof(int64_t));
    SEXP rb = PROTECT(allocVector(INTSXP, (R_xlen_t)R));
    SEXP ro = PROTECT(allocVector(INTSXP, (R_xlen_t)R));
    CD_CHECK(ctx, cd_get_region_universe(ctx, off, INTEGER(rb), INTEGER(ro)));
    SEXP roff = PROTECT(allocVector(REALSXP, m + 1));
    for (R_xlen_t i = 0; i <= m; i++) REAL(roff)[i] = (double)off[i];
    SEXP out = PROTECT(allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, roff); SET_VECTOR_ELT(out, 1, rb); SET_VECTOR_ELT(out, 2, ro);
    SEXP nm = PROTECT(allocVector(STRSXP, 3));
    SET_STRING_ELT(nm, 0, mkChar("row_off")); SET_STRING_ELT(nm, 1, mkChar("baitID")); SET_STRING_ELT(nm, 2, mkChar("otherEndID"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}

SEXP cdR_set_region_rows(SEXP ptr, SEXP row_bait, SEXP row_oe)
{
    cd_ctx* ctx = get_ctx(ptr);
    CD_CHECK(ctx, cd_set_region_rows(ctx, (int64_t)XLENGTH(row_bait), INTEGER(row_bait), INTEGER(row_oe)));
    return R_NilValue;
}

/* s: 1-based replicate; s_j / s_i numeric (NA -> NaN is what REAL() already holds), tblb / tlb 0-based bin codes with
 * -1 = NA; tmean: numeric matrix passed TRANSPOSED (t(tmean)) so that R's column-major storage is n_tblb x n_tlb
 * row-major; distfun: the 10 numbers of .chicEstimateDistFun; cnt_off numeric (F + 1), cnt_oe / cnt_N integer */
SEXP cdR_set_sample_tables(SEXP ptr, SEXP s, SEXP s_j, SEXP tblb, SEXP s_i, SEXP tlb, SEXP tmean_t, SEXP distfun,
                           SEXP cnt_off, SEXP cnt_oe, SEXP cnt_N)
{
    cd_ctx* ctx = get_ctx(ptr);
    cd_sample_tables t;
    memset(&t, 0, sizeof(t));
    t.s_j = REAL(s_j); t.tblb = INTEGER(tblb); t.s_i = REAL(s_i); t.tlb = INTEGER(tlb);
    t.n_tlb = nrows(tmean_t); t.n_tblb = ncols(tmean_t);
    t.tmean = REAL(tmean_t);
    if (XLENGTH(distfun) != 10) error("chicdiff_b200: distfun must hold 10 numbers");
    memcpy(t.distfun, REAL(distfun), sizeof(t.distfun));
    t.cnt_off = offsets_from_real(cnt_off); t.cnt_oe = INTEGER(cnt_oe); t.cnt_N = INTEGER(cnt_N);
    CD_CHECK(ctx, cd_set_sample_tables(ctx, asInteger(s) - 1, &t));
    return R_NilValue;
}

/* One replicate's raw CHiCAGO columns -> the per-fragment tables, built on the device (cd_build_sample_tables): no setkey,
 * no first-per-group pass in R.  tblb / tlb: integer codes (factor codes - 1, NA -> -1); cnt_*: the .chinput rows, or
 * NULL to take the table's own N column. */
SEXP cdR_build_sample_tables(SEXP ptr, SEXP s, SEXP baitID, SEXP otherEndID, SEXP s_j, SEXP s_i, SEXP tblb, SEXP tlb, SEXP Tmean,
                             SEXP N, SEXP n_tblb, SEXP n_tlb, SEXP distfun, SEXP cnt_bait, SEXP cnt_oe, SEXP cnt_N)
{
    cd_ctx* ctx = get_ctx(ptr);
    cd_chicago_table t;
    memset(&t, 0, sizeof(t));
    t.rows = (int64_t)XLENGTH(baitID);
    t.baitID = INTEGER(baitID); t.otherEndID = INTEGER(otherEndID); t.s_j = REAL(s_j); t.s_i = REAL(s_i);
    t.tblb = INTEGER(tblb); t.tlb = INTEGER(tlb); t.Tmean = REAL(Tmean);
    t.N = isNull(N) ? NULL : INTEGER(N);
    t.n_tblb = asInteger(n_tblb); t.n_tlb = asInteger(n_tlb);
    if (XLENGTH(distfun) != 10) error("chicdiff_b200: distfun must hold 10 numbers");
    memcpy(t.distfun, REAL(distfun), sizeof(t.distfun));
    if (!isNull(cnt_bait)) {
        t.cnt_rows = (int64_t)XLENGTH(cnt_bait);
        t.cnt_baitID = INTEGER(cnt_bait); t.cnt_otherEndID = INTEGER(cnt_oe); t.cnt_N = INTEGER(cnt_N);
    }
    CD_CHECK(ctx, cd_build_sample_tables(ctx, asInteger(s) - 1, &t));
    return R_NilValue;
}

/* list(K = n x S integer matrix, FullMean = n x S numeric matrix, avDist = numeric n) */
SEXP cdR_assemble(SEXP ptr, SEXP n_, SEXP S_, SEXP keep_rows)
{
    cd_ctx* ctx = get_ctx(ptr);
    R_xlen_t n = (R_xlen_t)asReal(n_);
    int S = asInteger(S_);
    check_dims(ctx, "cdR_assemble", (long long)n, S, -1);
    SEXP K = PROTECT(allocMatrix(INTSXP, (int)n, S));
    SEXP FM = PROTECT(allocMatrix(REALSXP, (int)n, S));
    SEXP av = PROTECT(allocVector(REALSXP, n));
    CD_CHECK(ctx, cd_assemble(ctx, asLogical(keep_rows) ? 1 : 0, INTEGER(K), REAL(FM), REAL(av)));
    SEXP out = PROTECT(allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, K); SET_VECTOR_ELT(out, 1, FM); SET_VECTOR_ELT(out, 2, av);
    SEXP nm = PROTECT(allocVector(STRSXP, 3));
    SET_STRING_ELT(nm, 0, mkChar("K")); SET_STRING_ELT(nm, 1, mkChar("FullMean")); SET_STRING_ELT(nm, 2, mkChar("avDist"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}

/* per-row columns of replicate s (1-based) after cdR_assemble(keep_rows = TRUE): list(N, FullMean, Bmean) */
SEXP cdR_get_sample_rows(SEXP ptr, SEXP s, SEXP R_)
{
    cd_ctx* ctx = get_ctx(ptr);
    R_xlen_t R = (R_xlen_t)asReal(R_);
    check_dims(ctx, "cdR_get_sample_rows", -1, -1, (long long)R);
    SEXP N = PROTECT(allocVector(INTSXP, R));
    SEXP FM = PROTECT(allocVector(REALSXP, R));
    SEXP BM = PROTECT(allocVector(REALSXP, R));
    CD_CHECK(ctx, cd_get_sample_rows(ctx, asInteger(s) - 1, INTEGER(N), REAL(FM)));
    CD_CHECK(ctx, cd_get_sample_bmean(ctx, asInteger(s) - 1, REAL(BM)));
    SEXP out = PROTECT(allocVector(VECSXP, 3));
    SET_VECTOR_ELT(out, 0, N); SET_VECTOR_ELT(out, 1, FM); SET_VECTOR_ELT(out, 2, BM);
    SEXP nm = PROTECT(allocVector(STRSXP, 3));
    SET_STRING_ELT(nm, 0, mkChar("N")); SET_STRING_ELT(nm, 1, mkChar("FullMean")); SET_STRING_ELT(nm, 2, mkChar("Bmean"));
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(5);
    return out;
}

/* countput of one condition (chicdiff.R:755-770).  reps: list of data.frames / lists with columns baitID, otherEndID, N
 * (integer), Bmean, score (numeric) in that order -- the CHiCAGO rows that have a distance (:715).
 * Returns list(baitID, otherEndID, Nav, Bav, score, oeID_mid). */
SEXP cdR_countput(SEXP ptr, SEXP reps)
{
    cd_ctx* ctx = get_ctx(ptr);
    int nr = (int)XLENGTH(reps);
    cd_chicago_rows* rows = (cd_chicago_rows*)R_alloc((size_t)nr, sizeof(cd_chicago_rows));
    for (int k = 0; k < nr; k++) {
        SEXP t = VECTOR_ELT(reps, k);
        rows[k].rows = (int64_t)XLENGTH(VECTOR_ELT(t, 0));
        rows[k].baitID = INTEGER(VECTOR_ELT(t, 0)); rows[k].otherEndID = INTEGER(VECTOR_ELT(t, 1));
        rows[k].N = INTEGER(VECTOR_ELT(t, 2));
        rows[k].Bmean = REAL(VECTOR_ELT(t, 3)); rows[k].score = REAL(VECTOR_ELT(t, 4));
    }
    int64_t G = 0;
    CD_CHECK(ctx, cd_countput(ctx, nr, rows, &G));
    SEXP b = PROTECT(allocVector(INTSXP, (R_xlen_t)G)), o = PROTECT(allocVector(INTSXP, (R_xlen_t)G));
    SEXP nav = PROTECT(allocVector(REALSXP, (R_xlen_t)G)), bav = PROTECT(allocVector(REALSXP, (R_xlen_t)G));
    SEXP sc = PROTECT(allocVector(REALSXP, (R_xlen_t)G)), mid = PROTECT(allocVector(REALSXP, (R_xlen_t)G));
    CD_CHECK(ctx, cd_get_countput(ctx, INTEGER(b), INTEGER(o), REAL(nav), REAL(bav), REAL(sc), REAL(mid)));
    static const char* names[6] = {"baitID", "otherEndID", "Nav", "Bav", "score", "oeID_mid"};
    SEXP cols[6] = {b, o, nav, bav, sc, mid};
    SEXP out = PROTECT(allocVector(VECSXP, 6));
    SEXP nm = PROTECT(allocVector(STRSXP, 6));
    for (int k = 0; k < 6; k++) { SET_VECTOR_ELT(out, k, cols[k]); SET_STRING_ELT(nm, k, mkChar(names[k])); }
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(8);
    return out;
}

/* fread() of a .chinput file that is already in memory as a raw vector -> list of its five columns */
SEXP cdR_parse_chinput(SEXP ptr, SEXP raw)
{
    cd_ctx* ctx = get_ctx(ptr);
    int64_t rows = 0;
    CD_CHECK(ctx, cd_parse_chinput(ctx, (const char*)RAW(raw), (int64_t)XLENGTH(raw), &rows));
    SEXP b = PROTECT(allocVector(INTSXP, (R_xlen_t)rows)), o = PROTECT(allocVector(INTSXP, (R_xlen_t)rows));
    SEXP N = PROTECT(allocVector(INTSXP, (R_xlen_t)rows)), len = PROTECT(allocVector(INTSXP, (R_xlen_t)rows));
    SEXP dist = PROTECT(allocVector(REALSXP, (R_xlen_t)rows));
    CD_CHECK(ctx, cd_get_chinput(ctx, INTEGER(b), INTEGER(o), INTEGER(N), INTEGER(len), REAL(dist)));
    static const char* names[5] = {"baitID", "otherEndID", "N", "otherEndLen", "distSign"};
    SEXP cols[5] = {b, o, N, len, dist};
    SEXP out = PROTECT(allocVector(VECSXP, 5));
    SEXP nm = PROTECT(allocVector(STRSXP, 5));
    for (int k = 0; k < 5; k++) { SET_VECTOR_ELT(out, k, cols[k]); SET_STRING_ELT(nm, k, mkChar(names[k])); }
    setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(7);
    return out;
}
