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
#include <unistd.h>

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
 *   m2           double F x n matrix, R's column-major layout, entries 0/1   (m.2); or NULL: use the matrix
 *                already resident in ctx (rvlR_load_gct + rvlR_select_rows)
 *   seed         double[n] of 0/1                                            (seed)
 *   target.rand  NULL or double n.perm x n matrix                            (LIB:1843-1844)
 * Returns list(cmi = F x iter (unsorted, rows of m.2), top = int[iter] (1-based), top.score,
 *              seed.iter = iter x n, cmi.seed = double[iter], cmi.seed.iter0,
 *              cmi.rand = F x n.perm x iter or NULL). */
SEXP rvlR_search(SEXP sctx, SEXP starget, SEXP sm2, SEXP sseed, SEXP siters, SEXP snegative, SEXP sngrid,
                 SEXP smult, SEXP sshrink, SEXP srand)
{
    rvl_ctx *ctx = get_ctx(sctx);
    const int resident = Rf_isNull(sm2);
    if (!Rf_isReal(starget) || !Rf_isReal(sseed) || (!resident && (!Rf_isReal(sm2) || !Rf_isMatrix(sm2))))
        Rf_error("revealer_b200: target, m2 and seed must be double (m2 a matrix or NULL)");
    const R_xlen_t n = XLENGTH(starget);
    int F;
    if (resident) {
        int64_t rows = 0, samples = 0;
        if (rvl_get_matrix_dims(ctx, &rows, &samples)) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
        if (samples != n) Rf_error("revealer_b200: resident matrix has %ld samples, target %ld", (long)samples, (long)n);
        F = (int)rows;
    } else {
        F = Rf_nrows(sm2);
        if (Rf_ncols(sm2) != n) Rf_error("revealer_b200: dimension mismatch (n = %ld)", (long)n);
    }
    if (XLENGTH(sseed) != n) Rf_error("revealer_b200: dimension mismatch (n = %ld)", (long)n);
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
    if (!rc && !resident) rc = rvl_load_matrix_f64_colmajor(ctx, REAL(sm2), F, n, F); /* validates 0/1, packs on device */
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
 *       bandwidth.shrink, target.rand)
 * The same search with the rows of m.2 sharded over several GPUs of this R process (REVEALER.v1's n.gpus > 1):
 * one context per device, rows [lo, lo+cnt) of the column-major R matrix passed in place (pointer offset,
 * ld = F: only the block's rows are read and shipped), rvl_group_connect + rvl_group_search (one persistent
 * launch per device; the devices meet through peer memory), scores scattered back into F x iter.
 * target.rand: NULL or n.perm x n permuted targets (LIB:1843-1844) -> cmi.rand = F x n.perm x iter. */
SEXP rvlR_search_multi(SEXP sdevices, SEXP starget, SEXP sm2, SEXP sseed, SEXP siters, SEXP snegative, SEXP sngrid,
                       SEXP smult, SEXP sshrink, SEXP srand)
{
    if (!Rf_isInteger(sdevices) || !Rf_isReal(starget) || !Rf_isReal(sm2) || !Rf_isMatrix(sm2) || !Rf_isReal(sseed))
        Rf_error("revealer_b200: bad argument types");
    const int world = (int)XLENGTH(sdevices);
    const R_xlen_t n = XLENGTH(starget);
    const int F = Rf_nrows(sm2);
    const int iters = Rf_asInteger(siters);
    if (world < 1 || world > 64 || Rf_ncols(sm2) != n || XLENGTH(sseed) != n || iters < 1) Rf_error("revealer_b200: dimension mismatch");
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
    double *tg = (double *)R_alloc((size_t)T * n, sizeof(double)); /* targets row-major [T][n] */
    memcpy(tg, REAL(starget), sizeof(double) * n);
    for (int j = 0; j < nperm; j++)
        for (R_xlen_t s = 0; s < n; s++) tg[(size_t)(1 + j) * n + s] = REAL(srand)[j + (size_t)nperm * s];
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
        if (!rc) { /* the shards are loaded one after the other by this one process: every load may use all the cores */
            long cores = sysconf(_SC_NPROCESSORS_ONLN);
            rc = rvl_set_host_threads(ctxs[r], cores < 1 ? 1 : (cores > 32 ? 32 : (int)cores));
        }
        if (!rc) rc = rvl_load_matrix_f64_colmajor(ctxs[r], REAL(sm2) + lo, hi - lo, n, F);
        if (!rc) rc = rvl_set_targets(ctxs[r], tg, n, T);
        if (!rc && nperm) rc = rvl_set_null_targets(ctxs[r], 1);
        if (!rc) rc = rvl_set_seed(ctxs[r], sd, n, 0);
        if (rc) msg = rvl_last_error(ctxs[r]);
    }
    if (!rc) { rc = rvl_group_connect(ctxs, world); if (rc) msg = rvl_last_error(ctxs[0]); }
    if (!rc) { rc = rvl_group_search(ctxs, world, iters); if (rc) msg = rvl_last_error(ctxs[0]); }
    SEXP out = R_NilValue;
    int nprot = 0;
    if (!rc) {
        out = PROTECT(Rf_allocVector(VECSXP, 7));
        SEXP names = PROTECT(Rf_allocVector(STRSXP, 7));
        const char *nm[7] = { "cmi", "top", "top.score", "seed.iter", "cmi.seed", "cmi.seed.iter0", "cmi.rand" };
        for (int i = 0; i < 7; i++) SET_STRING_ELT(names, i, Rf_mkChar(nm[i]));
        Rf_setAttrib(out, R_NamesSymbol, names);
        SEXP cmi = PROTECT(Rf_allocMatrix(REALSXP, F, iters));
        SEXP rtop = PROTECT(Rf_allocVector(INTSXP, iters));
        SEXP rts = PROTECT(Rf_allocVector(REALSXP, iters));
        SEXP rcs = PROTECT(Rf_allocVector(REALSXP, iters));
        SEXP rsi = PROTECT(Rf_allocMatrix(REALSXP, iters, (int)n));
        nprot = 7;
        SEXP cr = R_NilValue;
        if (nperm) {
            SEXP dims = PROTECT(Rf_allocVector(INTSXP, 3));
            INTEGER(dims)[0] = F; INTEGER(dims)[1] = nperm; INTEGER(dims)[2] = iters;
            cr = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)F * nperm * iters));
            Rf_setAttrib(cr, R_DimSymbol, dims);
            nprot += 2;
        }
        for (int r = 0; r < world && !rc; r++) {
            int64_t lo = ((int64_t)r * F) / world, cnt = ((int64_t)(r + 1) * F) / world - lo;
            double *loc = (double *)R_alloc((size_t)T * iters * (cnt ? cnt : 1), sizeof(double)); /* [T][iters][cnt] */
            int64_t *top = (int64_t *)R_alloc((size_t)T * iters, sizeof(int64_t));
            double *ts = (double *)R_alloc((size_t)T * iters, sizeof(double));
            uint8_t *si = (uint8_t *)R_alloc((size_t)T * iters * n, 1);
            double *cs = (double *)R_alloc((size_t)T * iters, sizeof(double));
            double *ic0 = (double *)R_alloc((size_t)T, sizeof(double));
            rvl_result res = { loc, top, ts, si, cs, ic0 };
            rc = rvl_search_fetch(ctxs[r], &res);
            if (rc) { msg = rvl_last_error(ctxs[r]); break; }
            for (int it = 0; it < iters; it++) {
                memcpy(REAL(cmi) + (size_t)F * it + lo, loc + (size_t)it * cnt, sizeof(double) * (size_t)cnt);
                for (int j = 0; j < nperm; j++)
                    memcpy(REAL(cr) + ((size_t)it * nperm + j) * F + lo, loc + ((size_t)(1 + j) * iters + it) * cnt,
                           sizeof(double) * (size_t)cnt);
            }
            if (r == 0) {
                for (int it = 0; it < iters; it++) {
                    INTEGER(rtop)[it] = (int)top[it] + 1;
                    REAL(rts)[it] = ts[it];
                    REAL(rcs)[it] = cs[it];
                    for (R_xlen_t s = 0; s < n; s++) REAL(rsi)[it + (size_t)iters * s] = (double)si[(size_t)it * n + s];
                }
                SET_VECTOR_ELT(out, 5, Rf_ScalarReal(ic0[0]));
            }
        }
        SET_VECTOR_ELT(out, 0, cmi);
        SET_VECTOR_ELT(out, 1, rtop);
        SET_VECTOR_ELT(out, 2, rts);
        SET_VECTOR_ELT(out, 3, rsi);
        SET_VECTOR_ELT(out, 4, rcs);
        SET_VECTOR_ELT(out, 6, cr);
        UNPROTECT(nprot);
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

/* .Call("rvlR_pvalues_iter", ctx, iter) -> p.val of iteration `iter` (1-based) of the last rvlR_search on ctx, from the
 * scores still resident on the device (LIB:1868-1869); NA for a NaN score. */
SEXP rvlR_pvalues_iter(SEXP sctx, SEXP siter)
{
    rvl_ctx *ctx = get_ctx(sctx);
    int64_t rows = 0, samples = 0;
    if (rvl_get_matrix_dims(ctx, &rows, &samples)) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)rows));
    int rc = rvl_pvalues_iter(ctx, Rf_asInteger(siter) - 1, REAL(out));
    UNPROTECT(1);
    if (rc) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    return out;
}

/* .Call("rvlR_topk", ctx, iter, k) -> integer[k]: the first k rows (1-based) of order(cmi[, iter], decreasing = !negative)
 * of the last rvlR_search on ctx, ranked on the device (the n.markers rows of LIB:1858-1864). */
SEXP rvlR_topk(SEXP sctx, SEXP siter, SEXP sk)
{
    rvl_ctx *ctx = get_ctx(sctx);
    int64_t rows = 0, samples = 0;
    if (rvl_get_matrix_dims(ctx, &rows, &samples)) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    int64_t k = Rf_asInteger(sk);
    if (k < 0) k = 0;
    if (k > rows) k = rows;
    int64_t *idx = (int64_t *)R_alloc((size_t)(k ? k : 1), sizeof(int64_t));
    int rc = rvl_topk(ctx, Rf_asInteger(siter) - 1, 0, k, idx, NULL);
    if (rc) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    SEXP out = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)k));
    for (int64_t i = 0; i < k; i++) INTEGER(out)[i] = (int)idx[i] + 1;
    UNPROTECT(1);
    return out;
}

/* .Call("rvlR_assess", ctx, target, m, target.rand, n.grid)
 * The numerical core of REVEALER_assess_features.v1 (LIB:2870-2890): for every row of the 0/1 matrix m (F x n,
 * column-major; or NULL: the matrix resident in ctx), IC = mutual.inf.v2(target, row)$IC (LIB:2875) and the same
 * against every row of target.rand (n.perm x n, the sample(target) draws of LIB:2885; or NULL).
 * Returns list(IC = double[F], null.IC = F x n.perm or NULL).  Counting the p-value (LIB:2886-2890) stays in R. */
SEXP rvlR_assess(SEXP sctx, SEXP starget, SEXP sm, SEXP srand, SEXP sngrid)
{
    rvl_ctx *ctx = get_ctx(sctx);
    const int resident = Rf_isNull(sm);
    if (!Rf_isReal(starget) || (!resident && (!Rf_isReal(sm) || !Rf_isMatrix(sm))))
        Rf_error("revealer_b200: target and m must be double (m a matrix or NULL)");
    const R_xlen_t n = XLENGTH(starget);
    int64_t F, samples = n;
    if (resident) {
        if (rvl_get_matrix_dims(ctx, &F, &samples)) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    } else {
        F = Rf_nrows(sm);
        samples = Rf_ncols(sm);
    }
    if (samples != n) Rf_error("revealer_b200: dimension mismatch (n = %ld)", (long)n);
    int nperm = 0;
    if (!Rf_isNull(srand)) {
        if (!Rf_isReal(srand) || !Rf_isMatrix(srand) || Rf_ncols(srand) != n) Rf_error("revealer_b200: target.rand must be n.perm x n");
        nperm = Rf_nrows(srand);
    }
    rvl_opts o;
    rvl_default_options(&o);
    o.n_grid = Rf_asInteger(sngrid);
    int rc = rvl_set_options(ctx, &o);
    if (!rc && !resident) rc = rvl_load_matrix_f64_colmajor(ctx, REAL(sm), F, n, F);
    if (rc) Rf_error("revealer_b200: %s", rvl_last_error(ctx));

    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_STRING_ELT(names, 0, Rf_mkChar("IC"));
    SET_STRING_ELT(names, 1, Rf_mkChar("null.IC"));
    Rf_setAttrib(out, R_NamesSymbol, names);
    SEXP ic = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)F));
    SEXP null_ic = nperm ? PROTECT(Rf_allocMatrix(REALSXP, (int)F, nperm)) : PROTECT(R_NilValue);
    SET_VECTOR_ELT(out, 0, ic);
    SET_VECTOR_ELT(out, 1, null_ic);

    /* the target and its permutations go up in chunks of <= 60000 rows (the ABI takes <= 65535 targets); row 0 of
     * every chunk is the target itself, so that the chunk is "one vector and its permutations" (shared bandwidths) */
    enum { CHUNK = 60000 };
    const int first = nperm < CHUNK ? nperm : CHUNK;
    double *tg = (double *)R_alloc((size_t)(1 + first) * n, sizeof(double));
    double *sc = (double *)R_alloc((size_t)(1 + first) * (size_t)F, sizeof(double));
    uint8_t *zero = (uint8_t *)R_alloc(n, 1);
    memset(zero, 0, n);
    memcpy(tg, REAL(starget), sizeof(double) * n);
    for (int lo = 0; lo == 0 || lo < nperm; lo += CHUNK) {
        const int cnt = nperm - lo < CHUNK ? nperm - lo : CHUNK;
        for (int j = 0; j < cnt; j++)
            for (R_xlen_t s = 0; s < n; s++) tg[(size_t)(1 + j) * n + s] = REAL(srand)[(lo + j) + (size_t)nperm * s];
        rc = rvl_set_targets_permuted(ctx, tg, n, 1 + cnt);
        if (!rc) rc = rvl_set_seed(ctx, zero, n, 0); /* constant z: cond.assoc reduces to mutual.inf.v2 (LIB:2509) */
        if (!rc) rc = rvl_score_once(ctx, sc);
        if (rc) break;
        if (lo == 0) memcpy(REAL(ic), sc, sizeof(double) * (size_t)F);
        if (cnt) memcpy(REAL(null_ic) + (size_t)lo * (size_t)F, sc + (size_t)F, sizeof(double) * (size_t)cnt * (size_t)F);
    }
    UNPROTECT(4);
    if (rc) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    return out;
}

/* ---- GCT ingest (replaces MSIG.Gct2Frame + data.matrix for ds2, LIB:2616-2626 / 1582) -------------------- */

static void rvl_r_gct_finalizer(SEXP ptr)
{
    rvl_gct *g = (rvl_gct *)R_ExternalPtrAddr(ptr);
    if (g) {
        rvl_gct_close(g);
        R_ClearExternalPtr(ptr);
    }
}

static rvl_gct *get_gct(SEXP ptr)
{
    rvl_gct *g = (rvl_gct *)R_ExternalPtrAddr(ptr);
    if (!g) Rf_error("revealer_b200: GCT handle already released");
    return g;
}

/* .Call("rvlR_gct_open", path) -> external pointer (the file stays memory-mapped until it is collected) */
SEXP rvlR_gct_open(SEXP spath)
{
    if (!Rf_isString(spath) || XLENGTH(spath) != 1) Rf_error("revealer_b200: path must be one string");
    rvl_gct *g = NULL;
    if (rvl_gct_open(CHAR(STRING_ELT(spath, 0)), &g)) Rf_error("revealer_b200: %s", rvl_gct_last_error(NULL));
    SEXP ptr = PROTECT(R_MakeExternalPtr(g, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, rvl_r_gct_finalizer, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* names joined by '\n' in buf -> character vector of length k */
static SEXP split_names(char *buf, int64_t k)
{
    SEXP out = PROTECT(Rf_allocVector(STRSXP, (R_xlen_t)k));
    char *p = buf;
    for (int64_t i = 0; i < k; i++) {
        char *e = strchr(p, '\n');
        if (e) *e = 0;
        SET_STRING_ELT(out, (R_xlen_t)i, Rf_mkChar(p));
        p = e ? e + 1 : p + strlen(p);
    }
    UNPROTECT(1);
    return out;
}

/* .Call("rvlR_gct_info", gct) -> list(row.names, descs, names): what MSIG.Gct2Frame returns besides ds */
SEXP rvlR_gct_info(SEXP sgct)
{
    rvl_gct *g = get_gct(sgct);
    int64_t F = 0, n = 0;
    rvl_gct_dims(g, &F, &n);
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 3));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 3));
    const char *nm[3] = { "row.names", "descs", "names" };
    for (int i = 0; i < 3; i++) SET_STRING_ELT(names, i, Rf_mkChar(nm[i]));
    Rf_setAttrib(out, R_NamesSymbol, names);
    for (int field = 0; field < 2; field++) {
        int64_t need = rvl_gct_row_names(g, field, NULL, 0);
        if (need < 0) Rf_error("revealer_b200: %s", rvl_gct_last_error(g));
        char *buf = (char *)R_alloc((size_t)need + 1, 1);
        buf[0] = 0;
        rvl_gct_row_names(g, field, buf, (size_t)need + 1);
        SET_VECTOR_ELT(out, field, split_names(buf, F));
    }
    SEXP cols = PROTECT(Rf_allocVector(STRSXP, (R_xlen_t)n));
    for (int64_t j = 0; j < n; j++) {
        int len = rvl_gct_sample_name(g, j, NULL, 0);
        char *buf = (char *)R_alloc((size_t)len + 1, 1);
        rvl_gct_sample_name(g, j, buf, (size_t)len + 1);
        SET_STRING_ELT(cols, (R_xlen_t)j, Rf_mkChar(buf));
    }
    SET_VECTOR_ELT(out, 2, cols);
    UNPROTECT(3);
    return out;
}

/* .Call("rvlR_gct_row", gct, i) -> numeric[n.cols], row i (1-based): the target row of ds1 */
SEXP rvlR_gct_row(SEXP sgct, SEXP si)
{
    rvl_gct *g = get_gct(sgct);
    int64_t F = 0, n = 0;
    rvl_gct_dims(g, &F, &n);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n));
    int rc = rvl_gct_row_values(g, (int64_t)Rf_asInteger(si) - 1, REAL(out));
    UNPROTECT(1);
    if (rc) Rf_error("revealer_b200: %s", rvl_gct_last_error(g));
    return out;
}

/* .Call("rvlR_load_gct", ctx, gct, cols) : every row of the file, samples cols (1-based file columns, in the
 * order REVEALER.v1 wants them) -> the resident bit matrix of ctx.  Returns rowSums(m.2) (LIB:1656). */
SEXP rvlR_load_gct(SEXP sctx, SEXP sgct, SEXP scols)
{
    rvl_ctx *ctx = get_ctx(sctx);
    rvl_gct *g = get_gct(sgct);
    if (!Rf_isInteger(scols)) Rf_error("revealer_b200: cols must be integer");
    const R_xlen_t n = XLENGTH(scols);
    int64_t F = 0, nc = 0;
    rvl_gct_dims(g, &F, &nc);
    int32_t *cols = (int32_t *)R_alloc((size_t)n, sizeof(int32_t));
    for (R_xlen_t s = 0; s < n; s++) cols[s] = INTEGER(scols)[s] - 1;
    if (rvl_load_matrix_gct(ctx, g, cols, (int64_t)n, 0, F)) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    SEXP out = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)F));
    int rc = rvl_get_row_counts(ctx, (int32_t *)INTEGER(out));
    UNPROTECT(1);
    if (rc) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    return out;
}

/* .Call("rvlR_get_rows", ctx, rows) -> length(rows) x n numeric matrix of the resident rows (1-based) */
SEXP rvlR_get_rows(SEXP sctx, SEXP srows)
{
    rvl_ctx *ctx = get_ctx(sctx);
    if (!Rf_isInteger(srows)) Rf_error("revealer_b200: rows must be integer");
    const R_xlen_t k = XLENGTH(srows);
    int64_t F = 0, n = 0;
    if (rvl_get_matrix_dims(ctx, &F, &n)) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    int64_t *rows = (int64_t *)R_alloc((size_t)k + 1, sizeof(int64_t));
    for (R_xlen_t i = 0; i < k; i++) rows[i] = (int64_t)INTEGER(srows)[i] - 1;
    uint8_t *buf = (uint8_t *)R_alloc((size_t)k * (size_t)n + 1, 1);
    if (rvl_get_rows_u8(ctx, rows, (int64_t)k, buf)) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)k, (int)n));
    for (R_xlen_t i = 0; i < k; i++)
        for (int64_t s = 0; s < n; s++) REAL(out)[i + (size_t)k * s] = (double)buf[(size_t)i * n + s];
    UNPROTECT(1);
    return out;
}

/* .Call("rvlR_select_rows", ctx, rows) : m.2 <- m.2[rows, ] on the device (rows 1-based) */
SEXP rvlR_select_rows(SEXP sctx, SEXP srows)
{
    rvl_ctx *ctx = get_ctx(sctx);
    if (!Rf_isInteger(srows)) Rf_error("revealer_b200: rows must be integer");
    const R_xlen_t k = XLENGTH(srows);
    int64_t *rows = (int64_t *)R_alloc((size_t)k + 1, sizeof(int64_t));
    for (R_xlen_t i = 0; i < k; i++) rows[i] = (int64_t)INTEGER(srows)[i] - 1;
    if (rvl_select_rows(ctx, rows, (int64_t)k)) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    return R_NilValue;
}

/* .Call("rvlR_consolidate_identical", ctx) -> list(rows = int[k] (1-based rows of the resident matrix before the
 * call), counts = int[k]): consolidate.identical.features = "identical" (LIB:1727-1748) on the device. */
SEXP rvlR_consolidate_identical(SEXP sctx)
{
    rvl_ctx *ctx = get_ctx(sctx);
    int64_t F = 0, n = 0, k = 0;
    if (rvl_get_matrix_dims(ctx, &F, &n)) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    int64_t *rows = (int64_t *)R_alloc((size_t)F + 1, sizeof(int64_t));
    int32_t *counts = (int32_t *)R_alloc((size_t)F + 1, sizeof(int32_t));
    if (rvl_consolidate_identical(ctx, rows, counts, &k)) Rf_error("revealer_b200: %s", rvl_last_error(ctx));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SEXP names = PROTECT(Rf_allocVector(STRSXP, 2));
    SET_STRING_ELT(names, 0, Rf_mkChar("rows"));
    SET_STRING_ELT(names, 1, Rf_mkChar("counts"));
    Rf_setAttrib(out, R_NamesSymbol, names);
    SEXP r = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)k));
    SEXP c = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)k));
    for (int64_t i = 0; i < k; i++) {
        INTEGER(r)[i] = (int)rows[i] + 1;
        INTEGER(c)[i] = (int)counts[i];
    }
    SET_VECTOR_ELT(out, 0, r);
    SET_VECTOR_ELT(out, 1, c);
    UNPROTECT(4);
    return out;
}
