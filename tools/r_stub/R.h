/* Minimal stand-in for R's headers: ONLY for a syntax / prototype check of r/src/r_glue.c in an image without R
 * (tests/test_abi.py::test_r_glue_compiles_against_stub).  Not R, not linked, not shipped. */
#ifndef R_STUB_RINTERNALS_H
#define R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
double* REAL(SEXP);
int* INTEGER(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_asLogical(SEXP);
int Rf_length(SEXP);
int Rf_nrows(SEXP);
R_xlen_t XLENGTH(SEXP);
SEXP Rf_duplicate(SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_alloc3DArray(unsigned int, int, int, int);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char*, ...);
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo*, const void*, const R_CallMethodDef*, const void*, const void*);
int R_useDynamicSymbols(DllInfo*, int);
#define FALSE 0
#define TRUE 1
#endif
