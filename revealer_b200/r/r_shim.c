/*
 * r_shim.c -- the thin .Call shim between R and librevealer_b200.so (include/revealer_b200.h).
 *
 * Build where R exists (not in the build container, which has no R):
 *     R CMD SHLIB r_shim.c -I../../include -L.. -lrevealer_b200 -o revealer_b200_r.so
 * and load from R with dyn.load("revealer_b200_r.so")  (see REVEALER_b200.R).
 *
 * No math here (north_star: "no Rcpp-side math and no CPU fallback"): argument checking, the
 * external-pointer lifetime of the rvl_ctx, allocation of the R result vectors, and the
 * conversion of a non-zero rvl_status into an R error AFTER temporaries are released
 * (Rf_error longjmps).  R API calls happen only on the calling (main R) thread.
 *
 * Replaces, inside REVEALER.v1 (REVEALER_library.v5C.R, live copy-B lines):
 *   the double loop                      LIB:1846-1856
 *   order()/top hit                      LIB:1858-1864
 *   seed update + assoc(target, seed)    LIB:2167-2171
 *   cmi.seed.iter0                       LIB:1833
 */
#include <R.h>
#include <Rinternals.h>
#include <stdint.h>
#include <string.h>

#include "revealer_b200.h"

static void rvl_r_finalizer(SEXP ptr)
{
    rvl_ctx *ctx = (rvl_ctx *)R_ExternalPtrAddr(ptr);
    if (ctx) {
        rvl_destroy(ctx); /* frees HBM when R garbage-collects the handle */
        R_ClearExternalPtr(ptr);
    }
}

static rvl_ctx *get_ctx(SEXP ptr)
{
    rvl_ctx *ctx = (rvl_ctx *)R_ExternalPtrAddr(ptr);
    if (!ctx) Rf_error("revealer_b200: context already released");
    return ctx;
}

/* .Call("rvlR_create", device) -> external pointer */
SEXP rvlR_create(SEXP device)
{
    rvl_ctx *ctx = NULL;
    int rc = rvl_create(&ctx, Rf_asInteger(device));
    if (rc != RVL_OK) Rf_error("revealer_b200: no usable CUDA device %d (status %d); there is no CPU fallback", Rf_asInteger(device), rc);
    SEXP ptr = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, rvl_r_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* .Call("rvlR_search", ctx, target, m2, seed, max.n.iter, negative, n.grid, bandwidth.mult,
 *       bandwidth.shrink, target.rand)
 *   target       double[n]          (already sorted as LIB:1626-1633)
 *   m2           double F x n matrix, R's column-major layout, entries 0/1   (m.2)
 *   seed         double[n] of 0/1                                            (seed)
 *   target.rand  NULL or double n.perm x n matrix                            (LIB:1843-1844)
 * Returns list(cmi = F x iter (unsorted, rows of m.2), top = int[iter] (1-based), top.score,
 *              seed.iter = iter x n, cmi.seed = double[iter], cmi.seed.iter0,
 *              cmi.rand = F x n.perm x iter or NULL). */
SEXP rvlR_search(SEXP sctx, SEXP starget, SEXP sm2, SEXP sseed, SEXP siters, SEXP snegative, SEXP sngrid,
                 SEXP smult, SEXP sshrink, SEXP srand)
{
    rvl_ctx *ctx = get_ctx(sctx);
    if (!Rf_isReal(starget) || !Rf_isReal(sm2) || !Rf_isMatrix(sm2) || !Rf_isReal(sseed))
        Rf_error("revealer_b200: target, m2 and seed must be double (m2 a matrix)");
    const R_xlen_t n = XLENGTH(starget);
    const int F = Rf_nrows(sm2);
    if (Rf_ncols(sm2) != n || XLENGTH(sseed) != n) Rf_error("revealer_b200: dimension mismatch (n = %ld)", (long)n);
    const int iters = Rf_asInteger(siters);
    int nperm = 0;
    if (!Rf_isNull(srand)) {
        if (!Rf_isReal(srand) || !Rf_isMatrix(srand) || Rf_ncols(srand) != n) Rf_error("revealer_b200: target.rand must be n.perm x n");
        nperm = Rf_nrows(srand);
    }
    const int T = 1 + nperm;

    rvl_opts o;
    rvl_default_options(&o);
    o.n_grid = Rf_asInteger(sngrid);
    o.negative = Rf_asLogical(snegative) ? 1 : 0;
    o.bandwidth_mult = Rf_asReal(smult);
    o.bandwidth_shrink = Rf_asReal(sshrink);

    /* targets row-major [T][n]; seed as bytes.  R_alloc memory is reclaimed by R even on error. */
    double *tg = (double *)R_alloc((size_t)T * n, sizeof(double));
    memcpy(tg, REAL(starget), sizeof(double) * n);
    for (int j = 0; j < nperm; j++)
        for (R_xlen_t s = 0; s < n; s++) tg[(size_t)(1 + j) * n + s] = REAL(srand)[j + (size_t)nperm * s];
    uint8_t *sd = (uint8_t *)R_alloc(n, 1);
    for (R_xlen_t s = 0; s < n; s++) {
        double v = REAL(sseed)[s];
        if (v != 0.0 && v != 1.0) Rf_error("revealer_b200: seed must be 0/1");
        sd[s] = (uint8_t)v;
    }

    int rc = rvl_set_options(ctx, &o);
    if (!rc) rc = rvl_load_matrix_f64_colmajor(ctx, REAL(sm2), F, n, F); /* validates 0/1, packs on device */
    if (!rc) rc = rvl_set_targets(ctx, tg, n, T);
    if (!rc && nperm) rc = rvl_set_null_targets(ctx, 1);
    if (!rc) rc = rvl_set_seed(ctx, sd, n, 0);
    if (rc) Rf_error("revealer_b200: %s", rvl_last_error(ctx));

    /* outputs: the objects REVEALER.v1 allocates at LIB:1837-1842 */
    double *cmi_all = (double *)R_alloc((size_t)T * iters * F, sizeof(double));
    int64_t *top = (int64_t *)R_alloc((size_t)T * iters, sizeof(int64_t));
    double *top_score = (double *)R_alloc((size_t)T * iters, sizeof(double));
    uint8_t *seed_iter = (uint8_t *)R_alloc((size_t)T * iters * n, 1);
    double *cmi_seed = (double *)R_alloc((size_t)T * iters, sizeof(double));
    double *ic0 = (double *)R_alloc((size_t)T, sizeof(double));
    rvl_result res = { cmi_all, top, top_score, seed_iter, cmi_seed, ic0 };
    rc = rvl_search(ctx, iters, &res);
    if (rc) Rf_error("revealer_b200: %s", rvl_last_error(ctx));

    SEXP out = PROTECT(Rf_allocVector(VECSXP, 7));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 7));
    const char *nm[7] = { "cmi", "top", "top.score", "seed.iter", "cmi.seed", "cmi.seed.iter0", "cmi.rand" };
    for (int i = 0; i < 7; i++) SET_STRING_ELT(names, i, Rf_mkChar(nm[i]));
    Rf_setAttrib(out, R_NamesSymbol, names);

    SEXP cmi = PROTECT(Rf_allocMatrix(REALSXP, F, iters));          /* column it = scores of iteration it */
    memcpy(REAL(cmi), cmi_all, sizeof(double) * (size_t)iters * F); /* target 0: [iters][F] == F x iters col-major */
    SET_VECTOR_ELT(out, 0, cmi);
    SEXP rtop = PROTECT(Rf_allocVector(INTSXP, iters));
    SEXP rts = PROTECT(Rf_allocVector(REALSXP, iters));
    SEXP rcs = PROTECT(Rf_allocVector(REALSXP, iters));
    SEXP rsi = PROTECT(Rf_allocMatrix(REALSXP, iters, (int)n));
    for (int it = 0; it < iters; it++) {
        INTEGER(rtop)[it] = (int)top[it] + 1; /* 1-based row of m.2 */
        REAL(rts)[it] = top_score[it];
        REAL(rcs)[it] = cmi_seed[it];
        for (R_xlen_t s = 0; s < n; s++) REAL(rsi)[it + (size_t)iters * s] = (double)seed_iter[(size_t)it * n + s];
    }
    SET_VECTOR_ELT(out, 1, rtop);
    SET_VECTOR_ELT(out, 2, rts);
    SET_VECTOR_ELT(out, 3, rsi);
    SET_VECTOR_ELT(out, 4, rcs);
    SET_VECTOR_ELT(out, 5, Rf_ScalarReal(ic0[0]));
    if (nperm) {
        SEXP dims = PROTECT(Rf_allocVector(INTSXP, 3));
        INTEGER(dims)[0] = F; INTEGER(dims)[1] = nperm; INTEGER(dims)[2] = iters;
        SEXP cr = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)F * nperm * iters));
        for (int it = 0; it < iters; it++)
            for (int j = 0; j < nperm; j++)
                memcpy(REAL(cr) + ((size_t)it * nperm + j) * F, cmi_all + ((size_t)(1 + j) * iters + it) * F,
                       sizeof(double) * F);
        Rf_setAttrib(cr, R_DimSymbol, dims);
        SET_VECTOR_ELT(out, 6, cr);
        UNPROTECT(2);
    } else {
        SET_VECTOR_ELT(out, 6, R_NilValue);
    }
    UNPROTECT(7);
    return out;
}

/* .Call("rvlR_search_multi", devices, target, m2, seed, max.n.iter, negative, n.grid, bandwidth.mult,
 *       bandwidth.shrink)
 * The same search with the rows of m.2 sharded over several GPUs of this R process (n.gpus > 1):
 * one context per device, rows [lo, lo+cnt) of the column-major R matrix passed in place (pointer
 * offset, ld = F: no copy), rvl_group_connect + rvl_group_search, scores scattered back into F x iter.
 * No permutation targets in this form. */
SEXP rvlR_search_multi(SEXP sdevices, SEXP starget, SEXP sm2, SEXP sseed, SEXP siters, SEXP snegative, SEXP sngrid,
                       SEXP smult, SEXP sshrink)
{
    if (!Rf_isInteger(sdevices) || !Rf_isReal(starget) || !Rf_isReal(sm2) || !Rf_isMatrix(sm2) || !Rf_isReal(sseed))
        Rf_error("revealer_b200: bad argument types");
    const int world = (int)XLENGTH(sdevices);
    const R_xlen_t n = XLENGTH(starget);
    const int F = Rf_nrows(sm2);
    const int iters = Rf_asInteger(siters);
    if (world < 1 || world > 64 || Rf_ncols(sm2) != n || XLENGTH(sseed) != n) Rf_error("revealer_b200: dimension mismatch");
    rvl_opts o;
    rvl_default_options(&o);
    o.n_grid = Rf_asInteger(sngrid);
    o.negative = Rf_asLogical(snegative) ? 1 : 0;
    o.bandwidth_mult = Rf_asReal(smult);
    o.bandwidth_shrink = Rf_asReal(sshrink);
    uint8_t *sd = (uint8_t *)R_alloc(n, 1);
    for (R_xlen_t s = 0; s < n; s++) {
        double v = REAL(sseed)[s];
        if (v != 0.0 && v != 1.0) Rf_error("revealer_b200: seed must be 0/1");
        sd[s] = (uint8_t)v;
    }
    rvl_ctx *ctxs[64];
    int made = 0, rc = 0;
    const char *msg = "";
    for (int r = 0; r < world && !rc; r++) {
        rc = rvl_create(&ctxs[r], INTEGER(sdevices)[r]);
        if (rc) { msg = "no usable CUDA device"; break; }
        made++;
        int64_t lo = ((int64_t)r * F) / world, hi = ((int64_t)(r + 1) * F) / world;
        rc = rvl_set_options(ctxs[r], &o);
        if (!rc) rc = rvl_set_shard(ctxs[r], r, world, lo, F);
        if (!rc) rc = rvl_load_matrix_f64_colmajor(ctxs[r], REAL(sm2) + lo, hi - lo, n, F);
        if (!rc) rc = rvl_set_targets(ctxs[r], REAL(starget), n, 1);
        if (!rc) rc = rvl_set_seed(ctxs[r], sd, n, 0);
        if (rc) msg = rvl_last_error(ctxs[r]);
    }
    if (!rc) { rc = rvl_group_connect(ctxs, world); if (rc) msg = rvl_last_error(ctxs[0]); }
    if (!rc) { rc = rvl_group_search(ctxs, world, iters); if (rc) msg = rvl_last_error(ctxs[0]); }
    SEXP out = R_NilValue;
    if (!rc) {
        out = PROTECT(Rf_allocVector(VECSXP, 5));
        SEXP names = PROTECT(Rf_allocVector(STRSXP, 5));
        const char *nm[5] = { "cmi", "top", "seed.iter", "cmi.seed", "cmi.seed.iter0" };
        for (int i = 0; i < 5; i++) SET_STRING_ELT(names, i, Rf_mkChar(nm[i]));
        Rf_setAttrib(out, R_NamesSymbol, names);
        SEXP cmi = PROTECT(Rf_allocMatrix(REALSXP, F, iters));
        SEXP rtop = PROTECT(Rf_allocVector(INTSXP, iters));
        SEXP rcs = PROTECT(Rf_allocVector(REALSXP, iters));
        SEXP rsi = PROTECT(Rf_allocMatrix(REALSXP, iters, (int)n));
        for (int r = 0; r < world && !rc; r++) {
            int64_t lo = ((int64_t)r * F) / world, cnt = ((int64_t)(r + 1) * F) / world - lo;
            double *loc = (double *)R_alloc((size_t)iters * (cnt ? cnt : 1), sizeof(double));
            int64_t *top = (int64_t *)R_alloc(iters, sizeof(int64_t));
            uint8_t *si = (uint8_t *)R_alloc((size_t)iters * n, 1);
            double *cs = (double *)R_alloc(iters, sizeof(double));
            double ic0 = 0.0;
            rvl_result res = { loc, top, NULL, si, cs, &ic0 };
            rc = rvl_search_fetch(ctxs[r], &res);
            if (rc) { msg = rvl_last_error(ctxs[r]); break; }
            for (int it = 0; it < iters; it++)
                memcpy(REAL(cmi) + (size_t)F * it + lo, loc + (size_t)it * cnt, sizeof(double) * (size_t)cnt);
            if (r == 0) {
                for (int it = 0; it < iters; it++) {
                    INTEGER(rtop)[it] = (int)top[it] + 1;
                    REAL(rcs)[it] = cs[it];
                    for (R_xlen_t s = 0; s < n; s++) REAL(rsi)[it + (size_t)iters * s] = (double)si[(size_t)it * n + s];
                }
                SET_VECTOR_ELT(out, 4, Rf_ScalarReal(ic0));
            }
        }
        SET_VECTOR_ELT(out, 0, cmi);
        SET_VECTOR_ELT(out, 1, rtop);
        SET_VECTOR_ELT(out, 2, rsi);
        SET_VECTOR_ELT(out, 3, rcs);
        UNPROTECT(6);
    }
    /* copy the message before the contexts that own it go away, then raise */
    char buf[512];
    strncpy(buf, msg, sizeof buf - 1);
    buf[sizeof buf - 1] = 0;
    for (int r = 0; r < made; r++) rvl_destroy(ctxs[r]);
    if (rc) Rf_error("revealer_b200: %s (status %d)", buf, rc);
    return out;
}

/* .Call("rvlR_assoc", ctx, target, y) -> assoc(target, y, "IC")   LIB:2491-2501 */
SEXP rvlR_assoc(SEXP sctx, SEXP starget, SEXP sy)
{
    rvl_ctx *ctx = get_ctx(sctx);
    const R_xlen_t n = XLENGTH(starget);
    if (!Rf_isReal(starget) || !Rf_isReal(sy) || XLENGTH(sy) != n) Rf_error("revealer_b200: bad assoc arguments");
    uint8_t *y = (uint8_t *)R_alloc(n, 1);
    for (R_xlen_t s = 0; s < n; s++) {
        double v = REAL(sy)[s];
        if (v != 0.0 && v != 1.0) Rf_error("revealer_b200: y must be 0/1");
        y[s] = (uint8_t)v;
    }
    double out = 0.0;
    int rc = rvl_set_targets(ctx, REAL(starget), n, 1);
    if (!rc) rc = rvl_assoc(ctx, 0, y, n, &out);
    if (rc) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    return Rf_ScalarReal(out);
}

/* .Call("rvlR_pvalues", ctx, cmi, cmi.rand) -> p.val   LIB:1869 */
SEXP rvlR_pvalues(SEXP sctx, SEXP scmi, SEXP srand)
{
    rvl_ctx *ctx = get_ctx(sctx);
    R_xlen_t F = XLENGTH(scmi), R = XLENGTH(srand);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, F));
    int rc = rvl_pvalues(ctx, REAL(scmi), F, REAL(srand), R, REAL(out));
    UNPROTECT(1);
    if (rc) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    return out;
}
