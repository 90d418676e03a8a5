/* Minimal stand-in for Rinternals.h (declarations used by r_shim.c): ONLY for the syntax check in an image without R. */
#include <stddef.h>
typedef struct SEXPREC *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;
#define VECSXP 19
#define STRSXP 16
#define REALSXP 14
#define INTSXP 13
void *R_ExternalPtrAddr(SEXP); void R_ClearExternalPtr(SEXP); SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void R_RegisterCFinalizerEx(SEXP, void (*)(SEXP), Rboolean);
SEXP Rf_protect(SEXP); void Rf_unprotect(int);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)
int Rf_asInteger(SEXP); int Rf_asLogical(SEXP); double Rf_asReal(SEXP);
int Rf_isReal(SEXP); int Rf_isMatrix(SEXP); int Rf_isNull(SEXP); int Rf_isInteger(SEXP);
R_xlen_t XLENGTH(SEXP); int Rf_nrows(SEXP); int Rf_ncols(SEXP);
double *REAL(SEXP); int *INTEGER(SEXP);
SEXP Rf_allocVector(int, R_xlen_t); SEXP Rf_allocMatrix(int, int, int); SEXP Rf_mkChar(const char*);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP); void SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP); SEXP Rf_ScalarReal(double);
int Rf_isString(SEXP); SEXP STRING_ELT(SEXP, R_xlen_t); const char *CHAR(SEXP);
