/* Declarations-only stand-in for R's Rinternals.h, used by tests/test_abi.py to SYNTAX-check R/r_glue.c against
 * include/chicdiff_b200.h in an image without R (gcc -fsyntax-only).  Test tooling: nothing links against it. */
#ifndef CD_TEST_RINTERNALS_MOCK_H
#define CD_TEST_RINTERNALS_MOCK_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned char Rbyte;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef unsigned int SEXPTYPE;
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define RAWSXP 24
extern SEXP R_NilValue, R_NamesSymbol;
typedef void (*R_CFinalizer_t)(SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
void Rf_error(const char*, ...);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_asLogical(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
char* R_alloc(size_t, int);
double* REAL(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
Rbyte* RAW(SEXP);
R_xlen_t XLENGTH(SEXP);
int LENGTH(SEXP);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_mkChar(const char*);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_duplicate(SEXP);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP Rf_lang3(SEXP, SEXP, SEXP);
SEXP Rf_eval(SEXP, SEXP);
SEXP R_tryEvalSilent(SEXP, SEXP, int*);
extern double R_NaReal;
#define NA_REAL R_NaReal
Rboolean Rf_isNull(SEXP);
extern SEXP R_GlobalEnv;
#define error Rf_error
#define asInteger Rf_asInteger
#define asReal Rf_asReal
#define asLogical Rf_asLogical
#define nrows Rf_nrows
#define ncols Rf_ncols
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define mkChar Rf_mkChar
#define setAttrib Rf_setAttrib
#define duplicate Rf_duplicate
#define ScalarReal Rf_ScalarReal
#define ScalarInteger Rf_ScalarInteger
#define lang3 Rf_lang3
#define eval Rf_eval
#define isNull Rf_isNull
#endif
