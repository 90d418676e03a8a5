/*
 * rstub_runtime.c -- a miniature stand-in for the part of the R runtime that revealer_b200/r/r_shim.c uses.
 *
 * TEST INFRASTRUCTURE.  R is not installed in the build image or on the GPU box, so the .Call shim cannot be loaded
 * into R there.  This file implements, in ~250 lines of plain C, the handful of R API entry points the shim calls
 * (vectors, matrices, lists with names / dim attributes, external pointers with finalizers, R_alloc, Rf_error as a
 * longjmp) so that the shim's object code can be LINKED and EXECUTED against librevealer_b200.so from a test
 * (tests/test_r_shim_gpu.py): argument checking, layouts (column-major matrices, 1-based indices), result lists and
 * error conversion are exercised for real.  It is not R: no garbage collector (objects live until rstub_reset),
 * PROTECT is bookkeeping only, Rf_error unwinds to the rstub_call that made the call.
 */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R.h"
#include "Rinternals.h"

#define RSTUB_EXTPTR 22
#define RSTUB_CHAR 9
#define RSTUB_LGL 10
#define RSTUB_NIL 0

struct SEXPREC {
    int type;
    R_xlen_t len;
    int nrow, ncol, is_matrix;
    void *data;           /* double[], int[], SEXP[] or char[] */
    SEXP names, dim;
    void *ext;            /* external pointer address */
    void (*fin)(SEXP);
    struct SEXPREC *next; /* all objects, for rstub_reset */
};

static struct SEXPREC nil_obj = { RSTUB_NIL, 0, 0, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL };
static struct SEXPREC names_sym, dim_sym;
SEXP R_NilValue = &nil_obj, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;

static SEXP all_objects = NULL;
static void *ralloc_list = NULL; /* R_alloc blocks: [next][payload] */
static jmp_buf *err_jmp = NULL;
static char err_msg[1024];
static int protect_depth = 0, protect_max = 0, protect_underflow = 0;

static SEXP new_obj(int type, R_xlen_t len, size_t elt)
{
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = type;
    s->len = len;
    s->names = s->dim = R_NilValue;
    s->data = calloc((size_t)(len > 0 ? len : 1), elt ? elt : 1);
    s->next = all_objects;
    all_objects = s;
    return s;
}

/* ---- the R API subset ---------------------------------------------------------------------- */

char *R_alloc(size_t n, int size)
{
    void **blk = (void **)malloc(sizeof(void *) + n * (size_t)size + 16);
    blk[0] = ralloc_list;
    ralloc_list = blk;
    return (char *)(blk + 1);
}

void Rf_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_msg, sizeof err_msg, fmt, ap);
    va_end(ap);
    if (err_jmp) longjmp(*err_jmp, 1);
    fprintf(stderr, "rstub: Rf_error outside rstub_call: %s\n", err_msg);
    abort();
}

SEXP Rf_protect(SEXP s) { if (++protect_depth > protect_max) protect_max = protect_depth; return s; }
void Rf_unprotect(int n) { protect_depth -= n; if (protect_depth < 0) { protect_underflow = 1; protect_depth = 0; } }

SEXP Rf_allocVector(int type, R_xlen_t n)
{
    size_t elt = type == REALSXP ? sizeof(double) : (type == INTSXP || type == RSTUB_LGL) ? sizeof(int) : sizeof(SEXP);
    SEXP s = new_obj(type, n, elt);
    if (type == VECSXP || type == STRSXP)
        for (R_xlen_t i = 0; i < n; i++) ((SEXP *)s->data)[i] = R_NilValue;
    return s;
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol)
{
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    s->is_matrix = 1; s->nrow = nrow; s->ncol = ncol;
    return s;
}
SEXP Rf_mkChar(const char *c)
{
    SEXP s = new_obj(RSTUB_CHAR, (R_xlen_t)strlen(c), 1);
    free(s->data);
    s->data = strdup(c);
    return s;
}
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); ((double *)s->data)[0] = v; return s; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; }
void SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP *)x->data)[i]; }
const char *CHAR(SEXP x) { return (const char *)x->data; }
SEXP Rf_setAttrib(SEXP x, SEXP sym, SEXP v)
{
    if (sym == R_NamesSymbol) x->names = v;
    else if (sym == R_DimSymbol) {
        x->dim = v;
        if (v->len == 2) { x->is_matrix = 1; x->nrow = ((int *)v->data)[0]; x->ncol = ((int *)v->data)[1]; }
    }
    return v;
}
double *REAL(SEXP x) { if (x->type != REALSXP) Rf_error("rstub: REAL() on a non-double object"); return (double *)x->data; }
int *INTEGER(SEXP x) { if (x->type != INTSXP && x->type != RSTUB_LGL) Rf_error("rstub: INTEGER() on a non-integer object"); return (int *)x->data; }
R_xlen_t XLENGTH(SEXP x) { return x->len; }
int Rf_nrows(SEXP x) { return x->is_matrix ? x->nrow : (int)x->len; }
int Rf_ncols(SEXP x) { return x->is_matrix ? x->ncol : 1; }
int Rf_isReal(SEXP x) { return x->type == REALSXP; }
int Rf_isInteger(SEXP x) { return x->type == INTSXP; }
int Rf_isString(SEXP x) { return x->type == STRSXP; }
int Rf_isMatrix(SEXP x) { return x->is_matrix; }
int Rf_isNull(SEXP x) { return x == R_NilValue; }
int Rf_asInteger(SEXP x)
{
    if (x->len < 1) Rf_error("rstub: asInteger of length 0");
    if (x->type == REALSXP) return (int)((double *)x->data)[0];
    return ((int *)x->data)[0];
}
int Rf_asLogical(SEXP x) { return Rf_asInteger(x) != 0; }
double Rf_asReal(SEXP x)
{
    if (x->len < 1) Rf_error("rstub: asReal of length 0");
    if (x->type == REALSXP) return ((double *)x->data)[0];
    return (double)((int *)x->data)[0];
}
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot)
{
    (void)tag; (void)prot;
    SEXP s = new_obj(RSTUB_EXTPTR, 0, 1);
    s->ext = p;
    return s;
}
void *R_ExternalPtrAddr(SEXP s) { return s->ext; }
void R_ClearExternalPtr(SEXP s) { s->ext = NULL; }
void R_RegisterCFinalizerEx(SEXP s, void (*fin)(SEXP), Rboolean onexit) { (void)onexit; s->fin = fin; }

/* ---- harness (called from Python through ctypes) --------------------------------------------- */

SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_real(const double *v, long n) { SEXP s = Rf_allocVector(REALSXP, n); memcpy(s->data, v, sizeof(double) * (size_t)n); return s; }
/* column-major nrow x ncol, like an R matrix */
SEXP rstub_real_matrix(const double *v, int nrow, int ncol)
{
    SEXP s = Rf_allocMatrix(REALSXP, nrow, ncol);
    memcpy(s->data, v, sizeof(double) * (size_t)nrow * ncol);
    return s;
}
SEXP rstub_int(const int *v, long n) { SEXP s = Rf_allocVector(INTSXP, n); memcpy(s->data, v, sizeof(int) * (size_t)n); return s; }
SEXP rstub_logical(int v) { SEXP s = Rf_allocVector(RSTUB_LGL, 1); ((int *)s->data)[0] = v; return s; }
SEXP rstub_string(const char *c) { SEXP s = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(s, 0, Rf_mkChar(c)); return s; }

int rstub_type(SEXP s) { return s->type; }
long rstub_length(SEXP s) { return (long)s->len; }
int rstub_nrow(SEXP s) { return s->nrow; }
int rstub_ncol(SEXP s) { return s->ncol; }
long rstub_dim(SEXP s, int i) { return (s->dim == R_NilValue || i >= s->dim->len) ? -1 : ((int *)s->dim->data)[i]; }
const double *rstub_real_ptr(SEXP s) { return (const double *)s->data; }
const int *rstub_int_ptr(SEXP s) { return (const int *)s->data; }
SEXP rstub_elt(SEXP s, long i) { return ((SEXP *)s->data)[i]; }
const char *rstub_chars(SEXP s, long i) { return s->type == STRSXP ? CHAR(STRING_ELT(s, i)) : ""; }
const char *rstub_name(SEXP s, long i) { return s->names == R_NilValue ? "" : CHAR(STRING_ELT(s->names, i)); }
const char *rstub_last_error(void) { return err_msg; }
int rstub_protect_balance(void) { return protect_underflow ? -1 : protect_depth; }

/* .Call(fn, a0, ..., a(n-1)) with Rf_error caught: returns NULL on error (message in rstub_last_error). */
typedef SEXP (*fn0)(void);
SEXP rstub_call(void *fn, int nargs, SEXP *a)
{
    jmp_buf jb, *saved = err_jmp;
    SEXP out = NULL;
    err_msg[0] = 0;
    err_jmp = &jb;
    if (setjmp(jb) == 0) {
        switch (nargs) {
        case 0: out = ((SEXP(*)(void))fn)(); break;
        case 1: out = ((SEXP(*)(SEXP))fn)(a[0]); break;
        case 2: out = ((SEXP(*)(SEXP, SEXP))fn)(a[0], a[1]); break;
        case 3: out = ((SEXP(*)(SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2]); break;
        case 4: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2], a[3]); break;
        case 5: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2], a[3], a[4]); break;
        case 9: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
        case 10: out = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]); break;
        default: snprintf(err_msg, sizeof err_msg, "rstub_call: %d arguments not supported", nargs); break;
        }
    } else {
        out = NULL;         /* Rf_error unwound to here */
        protect_depth = 0;  /* R resets the protect stack on error */
    }
    err_jmp = saved;
    while (ralloc_list) { /* R reclaims R_alloc memory at the end of .Call */
        void **blk = (void **)ralloc_list;
        ralloc_list = blk[0];
        free(blk);
    }
    return out;
}

/* "garbage collection": run the finalizers of external pointers, free every object */
int rstub_reset(void)
{
    int finalized = 0;
    for (SEXP s = all_objects; s; s = s->next)
        if (s->type == RSTUB_EXTPTR && s->fin && s->ext) { s->fin(s); finalized++; }
    while (all_objects) {
        SEXP s = all_objects;
        all_objects = s->next;
        free(s->data);
        free(s);
    }
    protect_depth = protect_max = protect_underflow = 0;
    return finalized;
}
