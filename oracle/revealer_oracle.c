/*
 * revealer_oracle.c -- CPU restatement of REVEALER's conditional-mutual-information search.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (revealer_b200/, librevealer_b200.so) never links, imports or falls back to anything here.
 *
 * PARITY UNPINNED: R, MASS 7.3-29 and misc3d 0.8-4 (r.package.info:82,88) are not installed in
 * the build container or on the GPU box, and the reference ships no numeric golden vectors (only
 * a PDF with 2-significant-figure scores whose inputs are unreachable URLs,
 * gpunit/example_data_test.yml:6,9).  This file therefore restates, operation by operation and
 * in R's evaluation order, the published algorithms the reference's call sites reach:
 *
 *   LIB = /root/reference/REVEALER_library.v5C.R  (live copy B; copy-A line = copy-B line - 1458)
 *
 *   orc_assoc            <- assoc             LIB:2491-2501
 *   orc_cond_assoc       <- cond.assoc        LIB:2503-2522
 *   orc_mutual_inf_v2    <- mutual.inf.v2     LIB:2524-2561   (calls MASS::bcv, MASS::kde2d, stats::cor)
 *   orc_cond_mutual_inf  <- cond.mutual.inf   LIB:2564-2614   (calls MASS::bcv, misc3d::kde3d, stats::cor)
 *   orc_search           <- iteration loop    LIB:1846-1856, rank LIB:1858-1864,
 *                                             seed update LIB:2167-2171, seed IC LIB:1833
 *   orc_bcv              <- MASS::bcv 7.3-29  (R wrapper + C VR_den_bin / VR_bcv_bin, restated)
 *   orc_brent_fmin       <- stats::optimize   (R's Brent_fmin, the Forsythe-Malcolm-Moler fmin)
 *   orc_kde2d            <- MASS::kde2d       (restated)
 *   orc_kde3d            <- misc3d::kde3d     (restated)
 *   orc_cor / orc_var    <- stats::cor / var  (two-pass means, long double accumulators)
 *
 * tools/make_golden.R regenerates full-precision golden vectors on any machine that has R;
 * tests/test_golden_r.py consumes them when present and reports "skipped" otherwise.
 *
 * Deliberately literal: no bandwidth hoisting, no rank-4 shortcut for binary features, full
 * 25x25x25 density cube, all seven entropies.  That keeps it an independent check of the CUDA
 * kernels' algebra.  ORC_HOIST (orc_search flag) only caches the three bcv() calls, whose
 * inputs are identical across calls; the arithmetic of each call is unchanged.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_NGRID_MAX 64
typedef long double LDOUBLE;

/* The two KDE contractions are compiled twice (AVX2 and baseline x86-64, picked at load time): vector
 * lanes hold different grid points i, each with its own accumulator, so the result is bit-identical
 * to the scalar loop (-ffp-contract=off: no FMA in either clone). */
#if defined(__GNUC__) && defined(__x86_64__) && !defined(__clang__)
#define ORC_HOT __attribute__((target_clones("avx2", "default")))
#else
#define ORC_HOT
#endif

#define ORC_M_1_SQRT_2PI 0.398942280401432677939946059934
#define ORC_M_PI 3.141592653589793238462643383280

/* ------------------------------------------------------------------ stats::var / stats::cor */

/* R's MEAN macro in cov.c: long-double sum, then one refinement pass. */
static double orc_mean(const double *x, int n)
{
    LDOUBLE s = 0.0;
    for (int i = 0; i < n; i++) s += x[i];
    LDOUBLE tmp = s / n;
    if (isfinite((double)tmp)) {
        s = 0.0;
        for (int i = 0; i < n; i++) s += (x[i] - tmp);
        tmp = tmp + s / n;
    }
    return (double)tmp;
}

double orc_var(const double *x, int n)
{
    double xm = orc_mean(x, n);
    LDOUBLE s = 0.0;
    for (int i = 0; i < n; i++) s += (LDOUBLE)(x[i] - xm) * (x[i] - xm);
    return (double)(s / (n - 1));
}

/* Pearson correlation, complete cases, as cov_complete2 with cor=TRUE: clamp to [-1, 1]. */
double orc_cor(const double *x, const double *y, int n)
{
    double xm = orc_mean(x, n), ym = orc_mean(y, n);
    int n1 = n - 1;
    LDOUBLE sxx = 0.0, syy = 0.0, sxy = 0.0;
    for (int i = 0; i < n; i++) {
        sxx += (LDOUBLE)(x[i] - xm) * (x[i] - xm);
        syy += (LDOUBLE)(y[i] - ym) * (y[i] - ym);
        sxy += (LDOUBLE)(x[i] - xm) * (y[i] - ym);
    }
    double xsd = sqrt((double)(sxx / n1)), ysd = sqrt((double)(syy / n1));
    double r = (double)(sxy / n1);
    if (xsd == 0.0 || ysd == 0.0) return NAN;
    r = r / (xsd * ysd);
    if (r > 1.0) r = 1.0;
    if (r < -1.0) r = -1.0;
    return r;
}

/* ------------------------------------------------------------------ stats::optimize (Brent_fmin) */

typedef double (*orc_fn1)(double, void *);

double orc_brent_fmin(double ax, double bx, orc_fn1 f, void *info, double tol)
{
    const double c = (3. - sqrt(5.)) * .5;
    double a, b, d, e, p, q, r, u, v, w, x;
    double t2, fu, fv, fw, fx, xm, eps, tol1, tol3;

    eps = sqrt(DBL_EPSILON);
    a = ax; b = bx;
    v = a + c * (b - a);
    w = v; x = v;
    d = 0.; e = 0.;
    fx = f(x, info);
    fv = fx; fw = fx;
    tol3 = tol / 3.;

    for (;;) {
        xm = (a + b) * .5;
        tol1 = eps * fabs(x) + tol3;
        t2 = tol1 * 2.;
        if (fabs(x - xm) <= t2 - (b - a) * .5) break;
        p = 0.; q = 0.; r = 0.;
        if (fabs(e) > tol1) { /* fit parabola */
            r = (x - w) * (fx - fv);
            q = (x - v) * (fx - fw);
            p = (x - v) * q - (x - w) * r;
            q = (q - r) * 2.;
            if (q > 0.) p = -p; else q = -q;
            r = e;
            e = d;
        }
        if (fabs(p) >= fabs(q * .5 * r) || p <= q * (a - x) || p >= q * (b - x)) {
            /* golden-section step */
            if (x < xm) e = b - x; else e = a - x;
            d = c * e;
        } else { /* parabolic-interpolation step */
            d = p / q;
            u = x + d;
            if (u - a < t2 || b - u < t2) {
                d = tol1;
                if (x >= xm) d = -d;
            }
        }
        if (fabs(d) >= tol1) u = x + d;
        else if (d > 0.) u = x + tol1;
        else u = x - tol1;

        fu = f(u, info);

        if (fu <= fx) {
            if (u < x) b = x; else a = x;
            v = w; w = x; x = u;
            fv = fw; fw = fx; fx = fu;
        } else {
            if (u < x) a = u; else b = u;
            if (fu <= fw || w == x) {
                v = w; fv = fw;
                w = u; fw = fu;
            } else if (fu <= fv || v == x || v == w) {
                v = u; fv = fu;
            }
        }
    }
    return x;
}

/* ------------------------------------------------------------------ MASS::bcv */

#define ORC_NB 1000
#define ORC_DELMAX 1000

/* VR_den_bin: histogram of |bin(x_i) - bin(x_j)| over all pairs, bin = (int)(x/dd). */
static void orc_den_bin(int n, int nb, double *d, const double *x, int *cnt)
{
    for (int i = 0; i < nb; i++) cnt[i] = 0;
    double xmin = x[0], xmax = x[0];
    for (int i = 1; i < n; i++) {
        if (x[i] < xmin) xmin = x[i];
        if (x[i] > xmax) xmax = x[i];
    }
    double rang = (xmax - xmin) * 1.01;
    double dd = rang / nb;
    *d = dd;
    /* the original recomputes jj = (int)(x[j]/dd) inside the pair loop; the values are the same */
    int *bin = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; i++) bin[i] = (int)(x[i] / dd);
    for (int i = 1; i < n; i++) {
        int ii = bin[i];
        for (int j = 0; j < i; j++) cnt[abs(ii - bin[j])]++;
    }
    free(bin);
}

typedef struct { int n, nb; double d; const int *cnt; } orc_bcv_info;

/* VR_bcv_bin.  64*n*n is evaluated in double: the C int product of the original overflows
 * for n >= 5793 (undefined behaviour there); identical for every n below that. */
static double orc_bcv_obj(double h, void *vp)
{
    const orc_bcv_info *p = (const orc_bcv_info *)vp;
    double hh = h / 4, sum = 0.0;
    for (int i = 0; i < p->nb; i++) {
        double delta = i * p->d / hh;
        delta *= delta;
        if (delta >= ORC_DELMAX) break;
        double term = exp(-delta / 4) * (delta * delta - 12 * delta + 12);
        sum += term * p->cnt[i];
    }
    double nn = (double)p->n;
    return 1 / (2 * nn * hh * sqrt(ORC_M_PI)) + sum / (64 * nn * nn * hh * sqrt(ORC_M_PI));
}

double orc_bcv(const double *x, int n)
{
    int cnt[ORC_NB];
    orc_bcv_info info;
    double hmax = 1.144 * sqrt(orc_var(x, n)) * pow((double)n, -1.0 / 5) * 4;
    double lower = 0.1 * hmax, upper = hmax;
    orc_den_bin(n, ORC_NB, &info.d, x, cnt);
    info.n = n; info.nb = ORC_NB; info.cnt = cnt;
    return orc_brent_fmin(lower, upper, orc_bcv_obj, &info, 0.1 * lower);
}

/* ------------------------------------------------------------------ helpers */

static void orc_range(const double *x, int n, double *lo, double *hi)
{
    double a = x[0], b = x[0];
    for (int i = 1; i < n; i++) { if (x[i] < a) a = x[i]; if (x[i] > b) b = x[i]; }
    *lo = a; *hi = b;
}

/* seq(from, to, length.out = m): from + i*by, last element pinned to `to` (seq.default). */
static void orc_seq(double from, double to, int m, double *g)
{
    double by = (to - from) / (m - 1);
    for (int i = 0; i < m; i++) g[i] = from + i * by;
    g[m - 1] = to;
}

static int orc_is_constant(const double *x, int n)
{
    for (int i = 1; i < n; i++) if (x[i] != x[0]) return 0;
    return 1;
}

static double orc_sign(double r) { return (r > 0) - (r < 0); }

/* dnorm(x, mean, sd) for finite args: M_1_SQRT_2PI * exp(-0.5*z*z) / sd */
static double orc_dnorm(double x, double mu, double sd)
{
    double z = (x - mu) / sd;
    return ORC_M_1_SQRT_2PI * exp(-0.5 * z * z) / sd;
}

/* ------------------------------------------------------------------ MASS::kde2d */

/* z is ng x ng, column-major like R: z[i + ng*j]. h is kde2d's `h` argument (divided by 4 inside). */
ORC_HOT static void orc_kde2d(const double *x, const double *y, int n, const double h[2], int ng,
                      double *gx, double *gy, double *z)
{
    double xlo, xhi, ylo, yhi;
    orc_range(x, n, &xlo, &xhi);
    orc_range(y, n, &ylo, &yhi);
    orc_seq(xlo, xhi, ng, gx);
    orc_seq(ylo, yhi, ng, gy);
    double hx = h[0] / 4, hy = h[1] / 4;
    double *mx = (double *)malloc(sizeof(double) * (size_t)ng * n * 2);
    double *my = mx + (size_t)ng * n;
    for (int s = 0; s < n; s++)
        for (int i = 0; i < ng; i++) {
            double ax = (gx[i] - x[s]) / hx, ay = (gy[i] - y[s]) / hy;
            mx[i + (size_t)ng * s] = ORC_M_1_SQRT_2PI * exp(-0.5 * ax * ax);
            my[i + (size_t)ng * s] = ORC_M_1_SQRT_2PI * exp(-0.5 * ay * ay);
        }
    /* z = tcrossprod(mx, my)/(n hx hy): every cell is the sum over s = 0..n-1, in that order, of
     * mx[i,s]*my[j,s].  The loops run s-outer / i-inner over ng independent accumulators (one per
     * i) so that the compiler can keep them in vector registers; no cell's summation order changes. */
    for (int j = 0; j < ng; j++) {
        double acc[ORC_NGRID_MAX];
        for (int i = 0; i < ng; i++) acc[i] = 0.0;
        for (int s = 0; s < n; s++) {
            const double w = my[j + (size_t)ng * s];
            const double *col = mx + (size_t)ng * s;
            for (int i = 0; i < ng; i++) acc[i] += col[i] * w;
        }
        for (int i = 0; i < ng; i++) z[i + ng * j] = acc[i] / (n * hx * hy);
    }
    free(mx);
}

/* ------------------------------------------------------------------ misc3d::kde3d */

/* d is ng^3, column-major: d[i + ng*(j + ng*k)].  h used directly as the sd of each axis. */
ORC_HOT static void orc_kde3d(const double *x, const double *y, const double *zz, int n, const double h[3],
                      int ng, double *gx, double *gy, double *gz, double *d)
{
    double lo, hi;
    orc_range(x, n, &lo, &hi);  orc_seq(lo, hi, ng, gx);
    orc_range(y, n, &lo, &hi);  orc_seq(lo, hi, ng, gy);
    orc_range(zz, n, &lo, &hi); orc_seq(lo, hi, ng, gz);
    double *mx = (double *)malloc(sizeof(double) * (size_t)ng * n * 4);
    double *my = mx + (size_t)ng * n, *mz = my + (size_t)ng * n, *t = mz + (size_t)ng * n;
    for (int s = 0; s < n; s++)
        for (int i = 0; i < ng; i++) {
            mx[i + (size_t)ng * s] = orc_dnorm(gx[i], x[s], h[0]);
            my[i + (size_t)ng * s] = orc_dnorm(gy[i], y[s], h[1]);
            mz[i + (size_t)ng * s] = orc_dnorm(gz[i], zz[s], h[2]);
        }
    for (int k = 0; k < ng; k++) {
        /* tmy.nz.zk = t(my)/nx * mz[k,]  (n x ng), then v[,,k] = mx %*% tmy.nz.zk: every cell is the
         * sum over s = 0..n-1, in that order, of mx[i,s]*t[s,j].  As in orc_kde2d the loops run
         * s-outer / i-inner over ng independent accumulators; no cell's summation order changes. */
        for (int j = 0; j < ng; j++)
            for (int s = 0; s < n; s++)
                t[s + (size_t)n * j] = my[j + (size_t)ng * s] / n * mz[k + (size_t)ng * s];
        for (int j = 0; j < ng; j++) {
            double acc[ORC_NGRID_MAX];
            const double *tj = t + (size_t)n * j;
            for (int i = 0; i < ng; i++) acc[i] = 0.0;
            for (int s = 0; s < n; s++) {
                const double w = tj[s];
                const double *col = mx + (size_t)ng * s;
                for (int i = 0; i < ng; i++) acc[i] += col[i] * w;
            }
            for (int i = 0; i < ng; i++) d[i + ng * (j + ng * k)] = acc[i];
        }
    }
    free(mx);
}

/* ------------------------------------------------------------------ mutual.inf.v2  LIB:2524-2561 */

typedef struct { double MI, SMI, HXY, HX, HY, NMI, IC; } orc_mi_t;

/* delta: the two bcv() bandwidths (before the rho shrink); pass NULL to compute them here,
 * as the default argument at LIB:2524 does on every call. */
int orc_mutual_inf_v2(const double *x, const double *y, int n, int ng, const double *delta_in,
                      orc_mi_t *out)
{
    if (ng > ORC_NGRID_MAX) return -1;
    double delta[2];
    if (delta_in) { delta[0] = delta_in[0]; delta[1] = delta_in[1]; }
    else { delta[0] = orc_bcv(x, n); delta[1] = orc_bcv(y, n); }

    double rho = orc_cor(x, y, n);
    double rho2 = fabs(rho);
    delta[0] = delta[0] * (1 + (-0.75) * rho2);
    delta[1] = delta[1] * (1 + (-0.75) * rho2);

    double gx[ORC_NGRID_MAX], gy[ORC_NGRID_MAX];
    int G2 = ng * ng;
    double *F = (double *)malloc(sizeof(double) * (size_t)G2 * 2);
    double *P = F + G2;
    orc_kde2d(x, y, n, delta, ng, gx, gy, F);
    LDOUBLE tot = 0.0;
    for (int c = 0; c < G2; c++) { F[c] = F[c] + DBL_EPSILON; tot += F[c]; }
    double dx = gx[1] - gx[0], dy = gy[1] - gy[0];
    double norm = (double)tot * dx * dy;
    for (int c = 0; c < G2; c++) P[c] = F[c] / norm;

    double PX[ORC_NGRID_MAX], PY[ORC_NGRID_MAX];
    for (int i = 0; i < ng; i++) {           /* rowSums(PXY)*dy */
        LDOUBLE s = 0.0;
        for (int j = 0; j < ng; j++) s += P[i + ng * j];
        PX[i] = (double)s * dy;
    }
    for (int j = 0; j < ng; j++) {           /* colSums(PXY)*dx */
        LDOUBLE s = 0.0;
        for (int i = 0; i < ng; i++) s += P[i + ng * j];
        PY[j] = (double)s * dx;
    }
    LDOUBLE s = 0.0;
    for (int c = 0; c < G2; c++) s += P[c] * log(P[c]);
    out->HXY = -(double)s * dx * dy;
    s = 0.0; for (int i = 0; i < ng; i++) s += PX[i] * log(PX[i]);
    out->HX = -(double)s * dx;
    s = 0.0; for (int j = 0; j < ng; j++) s += PY[j] * log(PY[j]);
    out->HY = -(double)s * dy;

    s = 0.0;                                  /* MI <- sum(PXY * log(PXY/(PX*PY)))*dx*dy */
    for (int j = 0; j < ng; j++)
        for (int i = 0; i < ng; i++) {
            double pxy = P[i + ng * j];
            s += pxy * log(pxy / (PX[i] * PY[j]));
        }
    out->MI = (double)s * dx * dy;
    out->SMI = orc_sign(rho) * out->MI;
    out->IC = orc_sign(rho) * sqrt(1 - exp(-2 * out->MI));
    out->NMI = orc_sign(rho) * ((out->HX + out->HY) / out->HXY - 1);
    free(F);
    return 0;
}

/* ------------------------------------------------------------------ cond.mutual.inf  LIB:2564-2614 */

typedef struct { double CMI, MI, SCMI, SMI, HXY, HXYZ, IC, CIC; } orc_cmi_t;

/* delta_in: 0.25*c(bcv(x), bcv(y), bcv(z)) or NULL to compute it here (LIB:2564 default arg). */
int orc_cond_mutual_inf(const double *x, const double *y, const double *z, int n, int ng,
                        const double *delta_in, orc_cmi_t *out)
{
    if (ng > ORC_NGRID_MAX) return -1;
    double delta[3];
    if (delta_in) { for (int a = 0; a < 3; a++) delta[a] = delta_in[a]; }
    else {
        delta[0] = 0.25 * orc_bcv(x, n);
        delta[1] = 0.25 * orc_bcv(y, n);
        delta[2] = 0.25 * orc_bcv(z, n);
    }
    double rho = orc_cor(x, y, n);
    double rho2 = (rho < 0) ? 0 : rho;
    for (int a = 0; a < 3; a++) delta[a] = delta[a] * (1 + (-0.75) * rho2);

    double X[ORC_NGRID_MAX], Y[ORC_NGRID_MAX], Z[ORC_NGRID_MAX];
    size_t G3 = (size_t)ng * ng * ng;
    int G2 = ng * ng;
    double *P = (double *)malloc(sizeof(double) * (G3 + 3 * (size_t)G2));
    double *PXZ = P + G3, *PYZ = PXZ + G2, *PXY = PYZ + G2;
    orc_kde3d(x, y, z, n, delta, ng, X, Y, Z, P);
    LDOUBLE tot = 0.0;
    for (size_t c = 0; c < G3; c++) { P[c] = P[c] + DBL_EPSILON; tot += P[c]; }
    double dx = X[1] - X[0], dy = Y[1] - Y[0], dz = Z[1] - Z[0];
    double norm = (double)tot * dx * dy * dz;
    for (size_t c = 0; c < G3; c++) P[c] = P[c] / norm;

#define PC(i, j, k) P[(i) + ng * ((j) + (size_t)ng * (k))]
    double PZ[ORC_NGRID_MAX], PX[ORC_NGRID_MAX], PY[ORC_NGRID_MAX];
    for (int k = 0; k < ng; k++)
        for (int i = 0; i < ng; i++) {       /* PXZ[i,k] = sum_j * dy */
            LDOUBLE s = 0.0;
            for (int j = 0; j < ng; j++) s += PC(i, j, k);
            PXZ[i + ng * k] = (double)s * dy;
        }
    for (int k = 0; k < ng; k++)
        for (int j = 0; j < ng; j++) {       /* PYZ[j,k] = sum_i * dx */
            LDOUBLE s = 0.0;
            for (int i = 0; i < ng; i++) s += PC(i, j, k);
            PYZ[j + ng * k] = (double)s * dx;
        }
    for (int k = 0; k < ng; k++) {           /* PZ[k] = sum_ij * dx*dy */
        LDOUBLE s = 0.0;
        for (int j = 0; j < ng; j++) for (int i = 0; i < ng; i++) s += PC(i, j, k);
        PZ[k] = (double)s * dx * dy;
    }
    for (int j = 0; j < ng; j++)
        for (int i = 0; i < ng; i++) {       /* PXY[i,j] = sum_k * dz */
            LDOUBLE s = 0.0;
            for (int k = 0; k < ng; k++) s += PC(i, j, k);
            PXY[i + ng * j] = (double)s * dz;
        }
    for (int i = 0; i < ng; i++) {           /* PX[i] = sum_jk * dy*dz */
        LDOUBLE s = 0.0;
        for (int k = 0; k < ng; k++) for (int j = 0; j < ng; j++) s += PC(i, j, k);
        PX[i] = (double)s * dy * dz;
    }
    for (int j = 0; j < ng; j++) {           /* PY[j] = sum_ik * dx*dz */
        LDOUBLE s = 0.0;
        for (int k = 0; k < ng; k++) for (int i = 0; i < ng; i++) s += PC(i, j, k);
        PY[j] = (double)s * dx * dz;
    }
#undef PC
    LDOUBLE s = 0.0;
    for (size_t c = 0; c < G3; c++) s += P[c] * log(P[c]);
    double HXYZ = -(double)s * dx * dy * dz;
    s = 0.0; for (int c = 0; c < G2; c++) s += PXZ[c] * log(PXZ[c]);
    double HXZ = -(double)s * dx * dz;
    s = 0.0; for (int c = 0; c < G2; c++) s += PYZ[c] * log(PYZ[c]);
    double HYZ = -(double)s * dy * dz;
    s = 0.0; for (int k = 0; k < ng; k++) s += PZ[k] * log(PZ[k]);
    double HZ = -(double)s * dz;
    s = 0.0; for (int c = 0; c < G2; c++) s += PXY[c] * log(PXY[c]);
    double HXY = -(double)s * dx * dy;
    s = 0.0; for (int i = 0; i < ng; i++) s += PX[i] * log(PX[i]);
    double HX = -(double)s * dx;
    s = 0.0; for (int j = 0; j < ng; j++) s += PY[j] * log(PY[j]);
    double HY = -(double)s * dy;

    out->MI = HX + HY - HXY;
    out->CMI = HXZ + HYZ - HXYZ - HZ;
    out->SMI = orc_sign(rho) * out->MI;
    out->SCMI = orc_sign(rho) * out->CMI;
    out->HXY = HXY;
    out->HXYZ = HXYZ;
    out->IC = orc_sign(rho) * sqrt(1 - exp(-2 * out->MI));
    out->CIC = orc_sign(rho) * sqrt(1 - exp(-2 * out->CMI));
    free(P);
    return 0;
}

/* ------------------------------------------------------------------ assoc / cond.assoc */

/* assoc(x, y, "IC")  LIB:2491-2501 */
double orc_assoc(const double *x, const double *y, int n, int ng)
{
    if (orc_is_constant(x, n) || orc_is_constant(y, n)) return 0.0;
    orc_mi_t r;
    orc_mutual_inf_v2(x, y, n, ng, NULL, &r);
    return r.IC;
}

/* cond.assoc(x, y, z, "IC")  LIB:2503-2522 */
double orc_cond_assoc(const double *x, const double *y, const double *z, int n, int ng)
{
    if (orc_is_constant(x, n) || orc_is_constant(y, n)) return 0.0;
    if (orc_is_constant(z, n)) {
        orc_mi_t r;
        orc_mutual_inf_v2(x, y, n, ng, NULL, &r);
        return r.IC;
    }
    orc_cmi_t r;
    orc_cond_mutual_inf(x, y, z, n, ng, NULL, &r);
    return r.CIC;
}

/* Same dispatcher with caller-supplied raw bcv values bw = {bcv(x), bcv(y), bcv(z)}:
 * the arithmetic of the call is unchanged, only the three searches are not repeated. */
static double orc_cond_assoc_bw(const double *x, const double *y, const double *z, int n, int ng,
                                const double bw[3])
{
    if (orc_is_constant(x, n) || orc_is_constant(y, n)) return 0.0;
    if (orc_is_constant(z, n)) {
        orc_mi_t r;
        double d2[2] = { bw[0], bw[1] };
        orc_mutual_inf_v2(x, y, n, ng, d2, &r);
        return r.IC;
    }
    orc_cmi_t r;
    double d3[3] = { 0.25 * bw[0], 0.25 * bw[1], 0.25 * bw[2] };
    orc_cond_mutual_inf(x, y, z, n, ng, d3, &r);
    return r.CIC;
}

/* Score a batch of feature rows against one target and one seed.
 * m2: F x n, ROW-major doubles here (row k contiguous = m.2[k,]).  LIB:1849-1852.
 * hoist = 0: literal reference cost (3 bcv per call); 1: bcv(x), bcv(z) once, bcv(y) per row. */
void orc_score_rows(const double *target, const double *m2, int64_t F, int n, const double *seed,
                    int ng, int hoist, int nthreads, double *out)
{
    double bwx = 0, bwz = 0;
    if (hoist) { bwx = orc_bcv(target, n); bwz = orc_is_constant(seed, n) ? 0.0 : orc_bcv(seed, n); }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t k = 0; k < F; k++) {
        const double *y = m2 + (size_t)k * n;
        if (hoist) {
            double bw[3] = { bwx, 0.0, bwz };
            if (!orc_is_constant(y, n)) bw[1] = orc_bcv(y, n);
            out[k] = orc_cond_assoc_bw(target, y, seed, n, ng, bw);
        } else {
            out[k] = orc_cond_assoc(target, y, seed, n, ng);
        }
    }
}

/* As orc_score_rows(hoist = 1), additionally returning the MI / CMI each score was formed from
 * (out_info): the quantity whose absolute rounding noise tests/test_noise_band.py measures. */
void orc_score_rows_info(const double *target, const double *m2, int64_t F, int n, const double *seed,
                         int ng, int nthreads, double *out, double *out_info)
{
    const double bwx = orc_bcv(target, n);
    const int zconst = orc_is_constant(seed, n), xconst = orc_is_constant(target, n);
    const double bwz = zconst ? 0.0 : orc_bcv(seed, n);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t k = 0; k < F; k++) {
        const double *y = m2 + (size_t)k * n;
        out[k] = 0.0;
        out_info[k] = 0.0;
        if (xconst || orc_is_constant(y, n)) continue;
        const double bwy = orc_bcv(y, n);
        if (zconst) {
            orc_mi_t r;
            double d2[2] = { bwx, bwy };
            orc_mutual_inf_v2(target, y, n, ng, d2, &r);
            out[k] = r.IC;
            out_info[k] = r.MI;
        } else {
            orc_cmi_t r;
            double d3[3] = { 0.25 * bwx, 0.25 * bwy, 0.25 * bwz };
            orc_cond_mutual_inf(target, y, seed, n, ng, d3, &r);
            out[k] = r.CIC;
            out_info[k] = r.CMI;
        }
    }
}

/* order(v, decreasing = !negative)[1]: stable, NaN last -> lowest index among exact extrema.
 * LIB:1858-1862.  Returns -1 when every score is NaN (R would still return an index; the
 * reference never reaches that state on valid input). */
int64_t orc_top1(const double *v, int64_t F, int negative)
{
    int64_t best = -1;
    for (int64_t k = 0; k < F; k++) {
        if (isnan(v[k])) continue;
        if (best < 0) { best = k; continue; }
        if (negative ? (v[k] < v[best]) : (v[k] > v[best])) best = k;
    }
    if (best < 0 && F > 0) best = 0;
    return best;
}

/* Full stable order (for top-k checks): idx_out gets the permutation R's order() returns (0-based). */
typedef struct { double v; int64_t i; } orc_kv;
static int orc_negative_flag;
static int orc_cmp(const void *a, const void *b)
{
    const orc_kv *p = (const orc_kv *)a, *q = (const orc_kv *)b;
    int pn = isnan(p->v), qn = isnan(q->v);
    if (pn || qn) { if (pn && qn) return (p->i > q->i) - (p->i < q->i); return pn - qn; }
    if (p->v != q->v) {
        int lt = p->v < q->v;
        return orc_negative_flag ? (lt ? -1 : 1) : (lt ? 1 : -1);
    }
    return (p->i > q->i) - (p->i < q->i);
}
void orc_order(const double *v, int64_t F, int negative, int64_t *idx_out)
{
    orc_kv *t = (orc_kv *)malloc(sizeof(orc_kv) * (size_t)F);
    for (int64_t k = 0; k < F; k++) { t[k].v = v[k]; t[k].i = k; }
    orc_negative_flag = negative;
    qsort(t, (size_t)F, sizeof(orc_kv), orc_cmp);
    for (int64_t k = 0; k < F; k++) idx_out[k] = t[k].i;
    free(t);
}

/* The iteration loop, LIB:1846-1864 + 2167-2171 (no permutations, no plots).
 *   cmi_out      F x iters, column-major like R (cmi_out[k + F*it]), UNSORTED row order of m2
 *   top_out      iters winners (0-based row index)
 *   seed_iter    iters x n row-major: the seed AFTER each iteration's merge (seed.iter[iter,])
 *   cmi_seed     iters: assoc(target, seed) after each merge  (LIB:2170)
 *   ic_seed0     assoc(target, initial seed)                  (LIB:1833)
 * seed is updated in place. */
void orc_search(const double *target, const double *m2, int64_t F, int n, double *seed, int iters,
                int negative, int ng, int hoist, int nthreads, double *cmi_out, int64_t *top_out,
                double *seed_iter, double *cmi_seed, double *ic_seed0)
{
    *ic_seed0 = orc_assoc(target, seed, n, ng);
    for (int it = 0; it < iters; it++) {
        double *col = cmi_out + (size_t)F * it;
        orc_score_rows(target, m2, F, n, seed, ng, hoist, nthreads, col);
        int64_t top = orc_top1(col, F, negative);
        top_out[it] = top;
        const double *row = m2 + (size_t)top * n;
        for (int s = 0; s < n; s++) seed[s] = (row[s] > seed[s]) ? row[s] : seed[s];
        memcpy(seed_iter + (size_t)it * n, seed, sizeof(double) * n);
        cmi_seed[it] = orc_assoc(target, seed, n, ng);
    }
}

/* p.val[i] = sum(cmi.rand > cmi[i]) / (F*n.perm)   LIB:1868-1869 (O(F^2 n.perm) in R).  A NaN score makes every
 * comparison NA in R and sum() of NAs is NA: NaN here.  NaN null scores are NA comparisons as well -- R's sum would be
 * NA for EVERY row then; the reference never removes them (no na.rm), so neither case is reachable on valid input;
 * the null-side NaNs are simply not counted. */
void orc_pvals(const double *cmi, int64_t F, const double *cmi_rand, int64_t nrand, double *pval)
{
    for (int64_t i = 0; i < F; i++) {
        if (isnan(cmi[i])) { pval[i] = NAN; continue; }
        int64_t c = 0;
        for (int64_t r = 0; r < nrand; r++) c += (cmi_rand[r] > cmi[i]);
        pval[i] = (double)c / (double)nrand;
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
