/*
 * oracle_hp.c -- the same two estimators as revealer_oracle.c, evaluated in EXTENDED precision.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as revealer_oracle.c: only tests/ may load it).
 *
 * Purpose: measure the rounding-noise band of the reference's own double-precision arithmetic.
 * IC = sqrt(1 - exp(-2 CMI)) and CMI = HXZ + HYZ - HXYZ - HZ is a difference of entropies of
 * magnitude ~5, each a sum of ~15 000 rounded P log P terms: a double-precision evaluation (R's, the
 * oracle's, the GPU's) knows CMI only to an ABSOLUTE accuracy kappa, so a score s carries a relative
 * error of about kappa / s^2.  This file evaluates the published formulas (LIB:2524-2561, 2564-2614 +
 * MASS::kde2d / misc3d::kde3d, restated as in revealer_oracle.c) with every intermediate in `HP`
 *   default          long double (x87, 64-bit mantissa: 2^-64 ~ 5e-20)
 *   -DORC_HP_QUAD    __float128 (libquadmath, 113-bit mantissa)
 * from the SAME inputs and the SAME bandwidths (bcv is a discrete Brent search whose result is an
 * input here, not something a higher precision would refine).  |double - HP| is then the noise of
 * the double evaluation; tests/test_noise_band.py prints it and derives the parity tolerance from it.
 */
#include "revealer_oracle.c"

#ifdef ORC_HP_QUAD
#include <quadmath.h>
typedef __float128 HP;
#define HP_EXP expq
#define HP_LOG logq
#define HP_SQRT sqrtq
#define HP_FABS fabsq
#define HP_C(x) x##Q
#else
typedef long double HP;
#define HP_EXP expl
#define HP_LOG logl
#define HP_SQRT sqrtl
#define HP_FABS fabsl
#define HP_C(x) x##L
#endif

#define HP_1_SQRT_2PI HP_C(0.398942280401432677939946059934381868475858631164934657665925)
#define HP_EPS ((HP)DBL_EPSILON) /* .Machine$double.eps is a double constant of the reference */

int orchp_bits(void)
{
#ifdef ORC_HP_QUAD
    return 113;
#else
    return LDBL_MANT_DIG;
#endif
}

static HP hp_cor(const double *x, const double *y, int n)
{
    HP sx = 0, sy = 0;
    for (int i = 0; i < n; i++) { sx += x[i]; sy += y[i]; }
    HP xm = sx / n, ym = sy / n, sxx = 0, syy = 0, sxy = 0;
    for (int i = 0; i < n; i++) {
        HP a = x[i] - xm, b = y[i] - ym;
        sxx += a * a; syy += b * b; sxy += a * b;
    }
    HP r = sxy / HP_SQRT(sxx * syy);
    if (r > 1) r = 1;
    if (r < -1) r = -1;
    return r;
}

static void hp_seq(double from, double to, int m, HP *g)
{
    HP by = ((HP)to - (HP)from) / (m - 1);
    for (int i = 0; i < m; i++) g[i] = (HP)from + i * by;
    g[m - 1] = to;
}

static HP hp_sign(HP r) { return (HP)((r > 0) - (r < 0)); }

/* mutual.inf.v2 (LIB:2524-2561) in HP.  delta_in = {bcv(x), bcv(y)} (doubles, from the double oracle).
 * out[0] = MI, out[1] = IC. */
int orchp_mutual_inf_v2(const double *x, const double *y, int n, int ng, const double *delta_in, double *out)
{
    if (ng > ORC_NGRID_MAX) return -1;
    HP rho = hp_cor(x, y, n);
    HP shrink = 1 + HP_C(-0.75) * HP_FABS(rho);
    HP hx = (HP)delta_in[0] * shrink / 4, hy = (HP)delta_in[1] * shrink / 4;
    double lo, hi;
    HP gx[ORC_NGRID_MAX], gy[ORC_NGRID_MAX];
    orc_range(x, n, &lo, &hi); hp_seq(lo, hi, ng, gx);
    orc_range(y, n, &lo, &hi); hp_seq(lo, hi, ng, gy);
    int G2 = ng * ng;
    HP *mx = (HP *)malloc(sizeof(HP) * ((size_t)ng * n * 2 + (size_t)G2));
    HP *my = mx + (size_t)ng * n, *P = my + (size_t)ng * n;
    for (int s = 0; s < n; s++)
        for (int i = 0; i < ng; i++) {
            HP ax = (gx[i] - x[s]) / hx, ay = (gy[i] - y[s]) / hy;
            mx[i + (size_t)ng * s] = HP_1_SQRT_2PI * HP_EXP(-ax * ax / 2);
            my[i + (size_t)ng * s] = HP_1_SQRT_2PI * HP_EXP(-ay * ay / 2);
        }
    HP tot = 0;
    for (int j = 0; j < ng; j++) {
        HP acc[ORC_NGRID_MAX];
        for (int i = 0; i < ng; i++) acc[i] = 0;
        for (int s = 0; s < n; s++) {
            HP w = my[j + (size_t)ng * s];
            const HP *col = mx + (size_t)ng * s;
            for (int i = 0; i < ng; i++) acc[i] += col[i] * w;
        }
        for (int i = 0; i < ng; i++) { P[i + ng * j] = acc[i] / (n * hx * hy) + HP_EPS; tot += P[i + ng * j]; }
    }
    HP dx = gx[1] - gx[0], dy = gy[1] - gy[0];
    HP norm = tot * dx * dy;
    for (int c = 0; c < G2; c++) P[c] = P[c] / norm;
    HP PX[ORC_NGRID_MAX], PY[ORC_NGRID_MAX];
    for (int i = 0; i < ng; i++) { HP s = 0; for (int j = 0; j < ng; j++) s += P[i + ng * j]; PX[i] = s * dy; }
    for (int j = 0; j < ng; j++) { HP s = 0; for (int i = 0; i < ng; i++) s += P[i + ng * j]; PY[j] = s * dx; }
    HP s = 0;
    for (int j = 0; j < ng; j++)
        for (int i = 0; i < ng; i++) { HP p = P[i + ng * j]; s += p * HP_LOG(p / (PX[i] * PY[j])); }
    HP mi = s * dx * dy;
    out[0] = (double)mi;
    out[1] = (double)(hp_sign(rho) * HP_SQRT(1 - HP_EXP(-2 * mi)));
    free(mx);
    return 0;
}

/* cond.mutual.inf (LIB:2564-2614) in HP.  delta_in = 0.25 * {bcv(x), bcv(y), bcv(z)} as LIB:2564 forms it.
 * out[0] = CMI, out[1] = CIC. */
int orchp_cond_mutual_inf(const double *x, const double *y, const double *z, int n, int ng, const double *delta_in,
                          double *out)
{
    if (ng > ORC_NGRID_MAX) return -1;
    HP rho = hp_cor(x, y, n);
    HP rho2 = rho < 0 ? 0 : rho;
    HP h[3];
    for (int a = 0; a < 3; a++) h[a] = (HP)delta_in[a] * (1 + HP_C(-0.75) * rho2);
    double lo, hi;
    HP X[ORC_NGRID_MAX], Y[ORC_NGRID_MAX], Z[ORC_NGRID_MAX];
    orc_range(x, n, &lo, &hi); hp_seq(lo, hi, ng, X);
    orc_range(y, n, &lo, &hi); hp_seq(lo, hi, ng, Y);
    orc_range(z, n, &lo, &hi); hp_seq(lo, hi, ng, Z);
    size_t G3 = (size_t)ng * ng * ng;
    int G2 = ng * ng;
    HP *P = (HP *)malloc(sizeof(HP) * (G3 + 2 * (size_t)G2 + (size_t)ng * n * 4));
    HP *PXZ = P + G3, *PYZ = PXZ + G2, *mx = PYZ + G2;
    HP *my = mx + (size_t)ng * n, *mz = my + (size_t)ng * n, *t = mz + (size_t)ng * n;
    for (int s = 0; s < n; s++)
        for (int i = 0; i < ng; i++) {
            HP a = (X[i] - x[s]) / h[0], b = (Y[i] - y[s]) / h[1], c = (Z[i] - z[s]) / h[2];
            mx[i + (size_t)ng * s] = HP_1_SQRT_2PI * HP_EXP(-a * a / 2) / h[0];
            my[i + (size_t)ng * s] = HP_1_SQRT_2PI * HP_EXP(-b * b / 2) / h[1];
            mz[i + (size_t)ng * s] = HP_1_SQRT_2PI * HP_EXP(-c * c / 2) / h[2];
        }
    HP tot = 0;
    for (int k = 0; k < ng; k++) {
        for (int j = 0; j < ng; j++)
            for (int s = 0; s < n; s++) t[s + (size_t)n * j] = my[j + (size_t)ng * s] / n * mz[k + (size_t)ng * s];
        for (int j = 0; j < ng; j++) {
            HP acc[ORC_NGRID_MAX];
            const HP *tj = t + (size_t)n * j;
            for (int i = 0; i < ng; i++) acc[i] = 0;
            for (int s = 0; s < n; s++) {
                HP w = tj[s];
                const HP *col = mx + (size_t)ng * s;
                for (int i = 0; i < ng; i++) acc[i] += col[i] * w;
            }
            for (int i = 0; i < ng; i++) {
                HP v = acc[i] + HP_EPS;
                P[i + ng * (j + (size_t)ng * k)] = v;
                tot += v;
            }
        }
    }
    HP dx = X[1] - X[0], dy = Y[1] - Y[0], dz = Z[1] - Z[0];
    HP norm = tot * dx * dy * dz;
    for (size_t c = 0; c < G3; c++) P[c] = P[c] / norm;
#define PC(i, j, k) P[(i) + ng * ((j) + (size_t)ng * (k))]
    HP PZ[ORC_NGRID_MAX];
    for (int k = 0; k < ng; k++)
        for (int i = 0; i < ng; i++) { HP s = 0; for (int j = 0; j < ng; j++) s += PC(i, j, k); PXZ[i + ng * k] = s * dy; }
    for (int k = 0; k < ng; k++)
        for (int j = 0; j < ng; j++) { HP s = 0; for (int i = 0; i < ng; i++) s += PC(i, j, k); PYZ[j + ng * k] = s * dx; }
    for (int k = 0; k < ng; k++) {
        HP s = 0;
        for (int j = 0; j < ng; j++) for (int i = 0; i < ng; i++) s += PC(i, j, k);
        PZ[k] = s * dx * dy;
    }
#undef PC
    HP s = 0;
    for (size_t c = 0; c < G3; c++) s += P[c] * HP_LOG(P[c]);
    HP HXYZ = -s * dx * dy * dz;
    s = 0; for (int c = 0; c < G2; c++) s += PXZ[c] * HP_LOG(PXZ[c]);
    HP HXZ = -s * dx * dz;
    s = 0; for (int c = 0; c < G2; c++) s += PYZ[c] * HP_LOG(PYZ[c]);
    HP HYZ = -s * dy * dz;
    s = 0; for (int k = 0; k < ng; k++) s += PZ[k] * HP_LOG(PZ[k]);
    HP HZ = -s * dz;
    HP cmi = HXZ + HYZ - HXYZ - HZ;
    out[0] = (double)cmi;
    out[1] = (double)(hp_sign(rho) * HP_SQRT(1 - HP_EXP(-2 * cmi)));
    free(P);
    return 0;
}

/* cond.assoc(target, m2[k,], seed) for a batch of rows, HP arithmetic, bandwidths from the double oracle.
 * out_score[k] = IC or CIC; out_info[k] = MI or CMI (the quantity whose absolute error is the noise band). */
void orchp_score_rows(const double *target, const double *m2, int64_t F, int n, const double *seed, int ng,
                      int nthreads, double *out_score, double *out_info)
{
    const double bwx = orc_bcv(target, n);
    const int zconst = orc_is_constant(seed, n);
    const double bwz = zconst ? 0.0 : orc_bcv(seed, n);
    const int xconst = orc_is_constant(target, n);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t k = 0; k < F; k++) {
        const double *y = m2 + (size_t)k * n;
        double r[2] = { 0.0, 0.0 };
        if (!xconst && !orc_is_constant(y, n)) {
            const double bwy = orc_bcv(y, n);
            if (zconst) {
                double d2[2] = { bwx, bwy };
                orchp_mutual_inf_v2(target, y, n, ng, d2, r);
            } else {
                double d3[3] = { 0.25 * bwx, 0.25 * bwy, 0.25 * bwz };
                orchp_cond_mutual_inf(target, y, seed, n, ng, d3, r);
            }
        }
        out_info[k] = r[0];
        out_score[k] = r[1];
    }
}
