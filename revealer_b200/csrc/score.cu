// score.cu -- the REVEALER scoring pass and the per-iteration tail, hand-written for sm_100a.
//
//   k_score      cmi[k] = cond.assoc(target, m.2[k,], seed, "IC") for every local row k and
//                every target at once                                  (LIB:1849-1856)
//   k_argmax     order(cmi, decreasing = !negative)[1] on the shard    (LIB:1858-1862)
//   k_advance    global winner from `world` candidate records, seed <- max(seed, m.2[top,])
//                (LIB:2167-2168), seed state (dispatch of cond.assoc LIB:2507-2517), x-axis table
//   k_assoc      cmi.seed[iter] <- assoc(target, seed)                 (LIB:2170), side stream
//   k_search     all of the above for every iteration in ONE persistent cooperative launch, the candidate exchange
//                between ranks done with peer-memory stores over NVLink inside the kernel (optional: rvl_set_fused)
//
// Algebra.  With y, z in {0,1}^n the 25^3 KDE cube of misc3d::kde3d is
//   d[i,j,k] = kappa * sum_{a,b in {0,1}} A_ab[i] * gy_a[j] * gz_b[k],
//   A_ab[i]  = sum_{s: y_s=a, z_s=b} exp(-beta (i - u_s)^2),  u_s = (x_s - xmin)/dx,
//   beta     = dx^2 / (2 hx^2),  gy_a[j] = exp(-(j/(G-1) - a)^2 / (2 hy^2)),  gz_b likewise,
//   kappa    = 1 / (n (2 pi)^{3/2} hx hy hz).
// The reference then adds .Machine$double.eps, normalises and takes entropies (LIB:2581-2605);
// in terms of Q = D + eps/kappa (D = d/kappa) and S = sum Q the cell-volume factors cancel:
//   CMI = HXZ + HYZ - HXYZ - HZ = ( L(Q_ijk) - L(Q_ik) - L(Q_jk) + L(Q_k) ) / S,  L(.) = sum q log q,
// with the marginals Q_ik, Q_jk, Q_k and S available in closed form from the rank-4 structure.
// Columns (j,k) whose density cannot change fl(D + eps') (D < eps' * 2^-55 for every i) contribute
// exactly G * eps' log eps' and are not visited.
//
// One warp owns one (target, feature) evaluation: no block barrier inside an evaluation, all
// reductions are fixed-order warp shuffles, so identical rows give bit-identical scores whatever
// their index, block or GPU -- which is what makes "lowest row index wins ties" reproducible.
#include "rvl_common.cuh"
#include <cfloat>
#include <cmath>
#include <cstring>

#define RVL_SCORE_WARPS 8 // warps per block of k_advance / k_assoc (fixes their summation order)
#ifndef RVL_KSCORE_WARPS
#define RVL_KSCORE_WARPS 16 // most independent worker warps per block of k_score (one block per SM; the host launches
                            // fewer when the per-warp scratch of a long bit row would not fit, rvl_score_plan)
#endif
#define RVL_TWO_PI 6.283185307179586476925286766559

// (rc_i, -log rc_i) of rvl_log_pos, filled per device by rvl_upload_log_table
__device__ double2 rvl_logtab_g[RVL_LOGTAB_ENTRIES];

// ---- per-warp scratch in shared memory ----------------------------------------------------
#ifndef RVL_LIST_MAX
#define RVL_LIST_MAX 512    // minority-class samples listed per row (longer rows walk their bit words instead)
#endif
#ifndef RVL_ITEMS_CAP
#define RVL_ITEMS_CAP 1024  // (j,k) columns of pass C: four-term ones from the front, one-class (soft) ones from the back
#endif
struct __align__(16) RvlWarpScratch {
    double A[RVL_MAXG][4];     // A[i][a*2+b]: x-axis kernel sums of the four (y,z) classes
    double gy[2][RVL_MAXG];    // gy[a][j]
    double gz[2][RVL_MAXG];    // gz[b][k]
    // y-axis tables of the entropy epilogue, indexed by the distance d from a y-class's own end of the axis:
    // prefix sums PG = sum g, PGL = sum g log g, PL = sum log g (d' <= d); suffix sums SG1..3 = sum g^1..3 (d' >= d)
    double tab[6][RVL_MAXG + 1];
    float lgf[RVL_MAXG];       // log2(g[d] / eps') as float: the column constant of the soft-floor cells (pass C)
    union {
        uint32_t list[RVL_LIST_MAX];   // minority-class samples of the row being scored (before the epilogue)
        uint16_t items[RVL_ITEMS_CAP]; // density columns evaluated cell by cell (epilogue): j << 5 | k, or class << 15 | d << 5 | k
    };
};

// FAST: rvl_log_pos (normal positive arguments only); otherwise the library log.
template <bool FAST> __device__ __forceinline__ double rvl_logq(double q) { return FAST ? rvl_log_pos(q) : log(q); }
template <bool FAST> __device__ __forceinline__ double rvl_xlogx(double q) { return q * rvl_logq<FAST>(q); }

// "a is ahead of b" in R's stable order(decreasing = !negative): NaN last, ties by lower row.
// (value, row) pairs under this relation form a total order, so reductions may be nested freely:
// warp (k_score) -> shard (k_argmax) -> world (k_advance) all give order()[1].
__device__ __forceinline__ bool rvl_ahead(double va, int64_t ia, double vb, int64_t ib, int negative)
{
    if (ib < 0) return ia >= 0;
    if (ia < 0) return false;
    bool na = isnan(va), nb = isnan(vb);
    if (na || nb) {
        if (na && nb) return ia < ib;
        return nb;
    }
    if (va != vb) return negative ? (va < vb) : (va > vb);
    return ia < ib;
}

// Programmatic dependent launch (sm_90+): a kernel of the per-iteration chain k_score -> k_argmax -> k_advance -> k_score
// may be made resident before its predecessor has finished (launch attribute programmaticStreamSerialization, see
// rvl_launch_chain); it then does what does not depend on the predecessor (launch latency, block scheduling, the log-table
// fill) and blocks in rvl_pdl_wait() until the predecessor grid has completed and its writes are visible.  Both are no-ops
// in a launch without the attribute.
__device__ __forceinline__ void rvl_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void rvl_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

// Loads of data that ANOTHER block of the same launch may have written (k_search: the x-axis table, the seeds, the
// per-iteration seed state, the candidate records): COH = true goes to L2 (ld.global.cg) instead of a possibly stale
// L1 line; COH = false is the plain load of the split-phase kernels, whose inputs are constant during a launch.
template <bool COH> __device__ __forceinline__ double rvl_ld(const double *p) { return COH ? __ldcg(p) : *p; }
template <bool COH> __device__ __forceinline__ uint64_t rvl_ld(const uint64_t *p)
{
    return COH ? (uint64_t)__ldcg(reinterpret_cast<const unsigned long long *>(p)) : *p;
}
template <bool COH> __device__ __forceinline__ RvlTarget rvl_ld_target(const RvlTarget *p)
{
    if (!COH) return *p;
    static_assert(sizeof(RvlTarget) % 8 == 0, "RvlTarget is copied in 8-byte words");
    RvlTarget t;
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(p);
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(&t);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(RvlTarget) / 8); i++) dst[i] = __ldcg(src + i);
    return t;
}

// Accumulate the class-masked Gaussian sums for grid point `lane` over samples [s_begin, s_end)
// with stride s_step.  yrow / zrow are bit rows (zrow may be NULL: single z class).
template <bool COH>
__device__ __forceinline__ void rvl_accumulate_dense(const double *__restrict__ u, const uint64_t *__restrict__ yrow,
                                                     const uint64_t *zrow, int s_begin, int s_end,
                                                     int s_step, double gi, double beta, double acc[4])
{
#define RVL_DENSE_ADD(ss, ee)                                                          \
    {                                                                                  \
        const int yb = (int)((yrow[(ss) >> 6] >> ((ss) & 63)) & 1ull);                 \
        const int zb = zrow ? (int)((rvl_ld<COH>(zrow + ((ss) >> 6)) >> ((ss) & 63)) & 1ull) : 0; \
        if (yb) { if (zb) acc[3] += (ee); else acc[2] += (ee); }                       \
        else    { if (zb) acc[1] += (ee); else acc[0] += (ee); }                       \
    }
    int s = s_begin;
    for (; s + 3 * s_step < s_end; s += 4 * s_step) { // four exp chains in flight; accumulation order unchanged
        const int s1 = s + s_step, s2 = s + 2 * s_step, s3 = s + 3 * s_step;
        const double t0 = gi - u[s], t1 = gi - u[s1], t2 = gi - u[s2], t3 = gi - u[s3];
        const double e0 = rvl_exp_neg(-beta * t0 * t0), e1 = rvl_exp_neg(-beta * t1 * t1),
                     e2 = rvl_exp_neg(-beta * t2 * t2), e3 = rvl_exp_neg(-beta * t3 * t3);
        RVL_DENSE_ADD(s, e0) RVL_DENSE_ADD(s1, e1) RVL_DENSE_ADD(s2, e2) RVL_DENSE_ADD(s3, e3)
    }
    for (; s < s_end; s += s_step) {
        const double t = gi - u[s];
        const double e = rvl_exp_neg(-beta * t * t);
        RVL_DENSE_ADD(s, e)
    }
#undef RVL_DENSE_ADD
}

// 3-D epilogue: CMI from A (in scratch), axis kernels gy/gz, eps' = eps/kappa.  Whole warp.
template <bool FAST>
__device__ double rvl_cmi_epilogue(RvlWarpScratch &S, int G, double epsq, double theta, int lane, int *nlive_out)
{
    const unsigned FULL = 0xffffffffu;
    // per-class totals and maxima over the x grid
    double sa[4], amax[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        double v = (lane < G) ? S.A[lane][c] : 0.0;
        sa[c] = rvl_warp_sum(v);
        amax[c] = rvl_warp_max(v);
    }
    double gyv = (lane < G) ? S.gy[0][lane] : 0.0;
    double gzv = (lane < G) ? S.gz[0][lane] : 0.0;
    double Gy = rvl_warp_sum(gyv); // = sum_j gy_1[j] as well (mirror image)
    double Gz = rvl_warp_sum(gzv);

    // live (j,k) columns.  A column is skipped when its largest possible density is below
    //   max(eps' * 2^-55, theta * S / G^3):
    // the first bound is exact (fl(D + eps') == eps'); the second replaces cells of mass < theta/G^3
    // by the eps' floor, which moves CMI by at most theta * (|log eps'| + 1) in total.
    const double Stot = (sa[0] + sa[1] + sa[2] + sa[3]) * Gy * Gz + (double)(G * G) * (double)G * epsq;
    const double dead_thr = fmax(epsq * 0x1p-55, theta * Stot / ((double)(G * G) * (double)G));
    int GG = G * G, nlive = 0;
    for (int p0 = 0; p0 < GG; p0 += 32) {
        int p = p0 + lane;
        bool live = false;
        if (p < GG) {
            int j = p / G, k = p - j * G;
            double dmax = amax[0] * S.gy[0][j] * S.gz[0][k] + amax[1] * S.gy[0][j] * S.gz[1][k] +
                          amax[2] * S.gy[1][j] * S.gz[0][k] + amax[3] * S.gy[1][j] * S.gz[1][k];
            live = dmax >= dead_thr;
        }
        unsigned m = __ballot_sync(FULL, live);
        if (live) S.items[nlive + __popc(m & ((1u << lane) - 1u))] = (uint16_t)p;
        nlive += __popc(m);
    }
    __syncwarp();
    *nlive_out = nlive;

    // L3 = sum over all cells of Q log Q;  Lyz = sum over (j,k) of Q_jk log Q_jk with
    // Q_jk = sum_i Q_ijk in closed form (skipped columns: Q_jk = G eps' to rounding)
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0, accyz = 0.0;
    const double gE = (double)G * epsq;
    for (int idx = lane; idx < nlive; idx += 32) {
        int p = S.items[idx];
        int j = p / G, k = p - j * G;
        double w00 = S.gy[0][j] * S.gz[0][k], w01 = S.gy[0][j] * S.gz[1][k];
        double w10 = S.gy[1][j] * S.gz[0][k], w11 = S.gy[1][j] * S.gz[1][k];
        double qyz = fma(sa[0], w00, fma(sa[1], w01, fma(sa[2], w10, fma(sa[3], w11, gE))));
        accyz = fma(qyz, rvl_logq<FAST>(qyz), accyz);
        // four independent logs in flight per lane (the FP64 chains are ~16 dependent ops long)
#define RVL_CELL(ii) fma(S.A[ii][0], w00, fma(S.A[ii][1], w01, fma(S.A[ii][2], w10, fma(S.A[ii][3], w11, epsq))))
        int i = 0;
#pragma unroll 1
        for (; i + 3 < G; i += 4) {
            double q0 = RVL_CELL(i), q1 = RVL_CELL(i + 1), q2 = RVL_CELL(i + 2), q3 = RVL_CELL(i + 3);
            double l0 = rvl_logq<FAST>(q0), l1 = rvl_logq<FAST>(q1), l2 = rvl_logq<FAST>(q2), l3 = rvl_logq<FAST>(q3);
            acc0 = fma(q0, l0, acc0);
            acc1 = fma(q1, l1, acc1);
            acc2 = fma(q2, l2, acc2);
            acc3 = fma(q3, l3, acc3);
        }
#pragma unroll 1
        for (; i < G; i++) {
            double q0 = RVL_CELL(i);
            acc0 = fma(q0, rvl_logq<FAST>(q0), acc0);
        }
#undef RVL_CELL
    }
    const double ndead = (double)(GG - nlive);
    double L3 = rvl_warp_sum((acc0 + acc1) + (acc2 + acc3)) + ndead * (double)G * rvl_xlogx<FAST>(epsq);
    double Lyz = rvl_warp_sum(accyz) + ndead * rvl_xlogx<FAST>(gE);

    // marginals: Q_ik (XZ), Q_k (Z)
    double accxz = 0.0;
    for (int p = lane; p < GG; p += 32) {
        int i = p / G, k = p - i * G;
        double qxz = fma((S.A[i][0] + S.A[i][2]) * Gy, S.gz[0][k], fma((S.A[i][1] + S.A[i][3]) * Gy, S.gz[1][k], gE));
        accxz += rvl_xlogx<FAST>(qxz);
    }
    double accz = 0.0;
    if (lane < G) {
        double qz = fma((sa[0] + sa[2]) * Gy, S.gz[0][lane], fma((sa[1] + sa[3]) * Gy, S.gz[1][lane], (double)GG * epsq));
        accz = rvl_xlogx<FAST>(qz);
    }
    double Lxz = rvl_warp_sum(accxz), Lz = rvl_warp_sum(accz);
    return (L3 - Lxz - Lyz + Lz) / Stot;
}

// ---- closed forms along the y axis ---------------------------------------------------------------
// The densities above one (x,z) cell -- and the entries of the y-z marginal above one z grid point -- have the form
//   Q_j = g[d_M] XM + g[d_m] Xm + E,   g[d] = exp(-kappa d^2) the y-axis kernel at distance d from a y-class's own end
// of the axis, d_M + d_m = G-1.  g falls off so fast that along j a class term t = g[d] X / E is, from its own end outwards,
//   big    t >= 16:      Q log Q = D log D + E (log D + 1) + O(E^2/D),  log D = log g[d] + log X    -- no log needed
//   mid    16 > t > 1/2: evaluated directly
//   small  t <= 1/2:     Q log Q = E log E + D (log E + 1) + D^2/(2E) - D^3/(6E^2) + O(E t^4/12)
// wherever the other class's term is invisible to fl(Q).  Summed over the j of a regime these are O(1) from six tables
// over d: prefix sums of g, g log g, log g (`big`), suffix sums of g, g^2, g^3 (`small`).  The neglected terms sum to
// < 1e-17 on CMI (tools/proto/epilogue_c.py checks the closed forms against the literal cube in extended precision;
// tests: production == literal cube).  Regime boundaries come from an approximate float sqrt: a term one step off its
// regime is still well inside the expansions' range, and direct evaluation is exact everywhere.
struct RvlRegime {          // per evaluation and per floor E
    double E, lE1;          // floor, log E + 1
    double h2, h3;          // 1/(2E), 1/(6E^2)
    double Lhi;             // log(16 E)
    float inv_kappa;        // 1/kappa = 2 ((G-1) hy)^2
    float x_lo, x_dead;     // log(16 / 0.5) / kappa, log(16 * 2^54) / kappa
};

__device__ __forceinline__ RvlRegime rvl_regime_make(double E, double lE, double inv_kappa)
{
    RvlRegime R;
    R.E = E; R.lE1 = lE + 1.0;
    R.h2 = 0.5 / E;
    R.h3 = (1.0 / 6.0) / (E * E);
    R.Lhi = lE + 2.772588722239781;                       // log 16
    R.inv_kappa = (float)inv_kappa;
    R.x_lo = (float)(3.4657359027997265 * inv_kappa);     // log 32
    R.x_dead = (float)(40.20253647247683 * inv_kappa);    // log(16 * 2^54)
    return R;
}

// floor(sqrt(x)) clamped to [-1, hi] (x < 0: -1) from the approximate square root: one MUFU instead of the IEEE sequence.
__device__ __forceinline__ int rvl_sqrt_bound(float x, int hi)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(x, 0.f)));
    return x < 0.f ? -1 : min(hi, (int)r);
}

// One cell: class values XM / Xm (>= 0) with logs lM / lm.  Adds the closed-form parts to accX (terms carrying X) and
// accE (coefficient of E), the number of j whose E log E is still owed to cntE, and returns the mask (bit = j' = d_M)
// of the entries that must be evaluated directly.  `valid` = 0 turns the lane into a no-op.
__device__ __forceinline__ unsigned rvl_cell_regimes(const RvlWarpScratch &S, const RvlRegime &R, int G, bool valid, double XM,
                                                     double Xm, double lM, double lm, double &accX, double &accE, int &cntE)
{
    const int G1 = G - 1;
    // last d with t >= 16 (a1), t > 1/2 (a2), t >= 2^-54 (a3); -1: none
    const float xM = (float)(lM - R.Lhi) * R.inv_kappa, xm = (float)(lm - R.Lhi) * R.inv_kappa;
    const int a1M = rvl_sqrt_bound(xM, G1), a2M = rvl_sqrt_bound(xM + R.x_lo, G1), a3M = rvl_sqrt_bound(xM + R.x_dead, G1);
    const int a1m = rvl_sqrt_bound(xm, G1), a2m = rvl_sqrt_bound(xm + R.x_lo, G1), a3m = rvl_sqrt_bound(xm + R.x_dead, G1);
    // a class's closed forms stop where the other class becomes visible
    const int limM = G - 2 - a3m, limm = G - 2 - a3M;
    const int bM = min(a1M, limM), bm = min(a1m, limm);
    const int s1M = a2M + 1, s1m = a2m + 1;
    unsigned closedM = 0u, closedm = 0u; // in each class's own d coordinates
    if (valid) {
        if (bM >= 0) { // big, d_M in [0, bM]
            accX = fma(XM, fma(lM, S.tab[0][bM], S.tab[1][bM]), accX);
            accE += fma((double)(bM + 1), lM + 1.0, S.tab[2][bM]);
            closedM = (2u << bM) - 1u;
        }
        if (bm >= 0) {
            accX = fma(Xm, fma(lm, S.tab[0][bm], S.tab[1][bm]), accX);
            accE += fma((double)(bm + 1), lm + 1.0, S.tab[2][bm]);
            closedm = (2u << bm) - 1u;
        }
        if (s1M <= limM) { // small, d_M in [s1M, limM] (through the gap where both classes are dead)
            const double g1 = S.tab[3][s1M] - S.tab[3][limM + 1], g2 = S.tab[4][s1M] - S.tab[4][limM + 1],
                         g3 = S.tab[5][s1M] - S.tab[5][limM + 1];
            accX = fma(XM, fma(XM, fma(-XM * R.h3, g3, g2 * R.h2), g1 * R.lE1), accX);
            closedM |= ((2u << limM) - 1u) & ~((1u << s1M) - 1u);
        }
        if (s1m <= limm) {
            const double g1 = S.tab[3][s1m] - S.tab[3][limm + 1], g2 = S.tab[4][s1m] - S.tab[4][limm + 1],
                         g3 = S.tab[5][s1m] - S.tab[5][limm + 1];
            accX = fma(Xm, fma(Xm, fma(-Xm * R.h3, g3, g2 * R.h2), g1 * R.lE1), accX);
            closedm |= ((2u << limm) - 1u) & ~((1u << s1m) - 1u);
        }
    }
    const unsigned all = (G >= 32) ? 0xffffffffu : ((1u << G) - 1u);
    const unsigned direct = valid ? (all & ~closedM & ~(__brev(closedm) >> (32 - G))) : 0u;
    cntE += valid ? (G - (bM >= 0 ? bM + 1 : 0) - (bm >= 0 ? bm + 1 : 0) - __popc(direct)) : 0;
    return direct;
}

// bits lo..hi (inclusive) of a 32-bit mask; empty when lo > hi or hi < 0
__device__ __forceinline__ unsigned rvl_bits(int lo, int hi)
{
    lo = max(lo, 0);
    return (lo <= hi) ? (((2u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
}

// ---- soft-floor cells: the part of pass C that needs no double-precision log ------------------------------------
// A density column in which only ONE y-class is visible has Q_i = g X_i + E (X = V or W of pass A, E = eps') and
//   sum_i Q_i log Q_i = g (log g SX + LX) + E (G log g + LLX)            closed form, pass A's per-k sums of X, X log X, log X
//                     + E ln2 sum_i (1 + t_i) log2(1 + 1/t_i),           t_i = g X_i / E
// exactly.  The second line is O(E) per cell whatever t is -- it IS the whole effect of the eps floor on the column --
// so single precision carries it: summed over the cube it is ~2e-14 of CMI (at most 5e-12: G^3 cells at the floor), and a
// relative error of 4e-4 there moves CMI by ~1e-17, < 2e-15 in the worst case (tools/proto/epilogue_softfloor.py: the
// column sums come out closer to a long-double evaluation than the cell-by-cell double-precision logs were).
// With d = log2 t = lf[cls][k][i] + lgf[dist] (float copies of log2 X from pass A and of log2(g / E)), a = |d|, e = 2^-a
// (one MUFU.EX2) and P(e) = log2(1 + e) / e (degree-3 minimax on [0, 1], relative error 3.9e-4):
//   d > 0  (cell above the floor):   (1 + t) log2(1 + 1/t) = (1 + e) P(e)
//   d <= 0 (cell at the floor):      (1 + t) log2(1 + 1/t) = (1 + e) (a + e P(e))
// -- 11 FP32-pipe instructions and one MUFU per cell against 11 FP64 + table gather + I2F for Q = fma(..), log Q, Q log Q.
// A clamped X (pass A takes the log of max(X, 1e-290)) stays consistent: closed form and correction use the same log X.
__device__ __forceinline__ float rvl_soft_cell(float d)
{
    const float a = fabsf(d);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-a));
    float p = fmaf(e, -0.10725461691617966f, 0.3664810359477997f);
    p = fmaf(p, e, -0.701747477054596f);
    p = fmaf(p, e, 1.4421281814575195f);
    float r = p;
    if (!(d > 0.f)) r = fmaf(e, p, a); // one predicated FFMA
    return fmaf(e, r, r);
}

// stride (in floats) of one k row of the float table lf[class][k][i]: a multiple of 4 (16-byte rows, LDS.128) whose
// quarter is odd, so that the rows of 8 consecutive k start in 8 distinct 16-byte bank groups
__host__ __device__ __forceinline__ int rvl_lf_stride(int G) { return (((G + 3) >> 2) | 1) * 4; }
__host__ __device__ __forceinline__ size_t rvl_lf_bytes(int G) { return sizeof(float) * 2 * (size_t)G * rvl_lf_stride(G); }

// 3-D epilogue, production form.  Same quantity as rvl_cmi_epilogue, organised around the majority
// y-class `am` (the class holding most samples; a = 0 for the usual sparse alteration row):
//   Q_ijk = g_M[j] * V_k[i] + g_m[j] * W_k[i] + eps',   V_k[i] = A_{am,0}[i] gz_0[k] + A_{am,1}[i] gz_1[k]
// (W likewise for the minority class).  Pass A (lanes = k) builds V, W once per (i,k) with their logs, the per-k sums
// of X, X log X, log X for both classes, a float copy of log2 X (lf), and takes the x-z marginal log.  Pass B (lanes = k)
// sorts the G density columns above every k, in j' = distance from the majority class's end of the y axis, into
//   dead        neither class term can reach the pruning threshold        -> constant G eps' log eps'
//   majority    minority term cannot change fl(Q) (g_m maxW < 2^-54 eps'): the closed form above with X = V; where also
//               eps' <= g_M minV / 16 ("factorised") the correction is G eps' to first order (neglected < eps'/32 per
//               cell, ~1e-16 on CMI), elsewhere it is summed cell by cell in single precision (pass C, soft cells)
//   minority    majority term cannot change fl(Q): the same with X = W, always with the cell-by-cell correction
//   four-term   both classes visible: G cells, one double-precision log each (pass C)
// The y-axis kernel g[d] = exp(-kappa d^2) is monotone in the distance d from a class's end of the axis, so all sets
// are INTERVALS whose ends follow in closed form from log maxV, log minV, log maxW (five sqrt bounds per lane): the
// closed forms are summed with prefix tables of the y-axis kernel, the columns of pass C are written to the lists at
// offsets from a warp scan, and the y-z marginal comes from the cell closed forms above (cells = k).
// lf: this warp's float table [2][G][rvl_lf_stride(G)].
template <bool FAST>
__device__ double rvl_cmi_epilogue_f(RvlWarpScratch &S, float *__restrict__ lf, int G, double epsq, double theta, double hy, int am,
                                     int lane, int *nfour_out, int *nfact_out, int *nsoft_out, int *nyz_out)
{
    const unsigned FULL = 0xffffffffu;
    const int M0 = am * 2, M1 = am * 2 + 1, m0 = (1 - am) * 2, m1 = (1 - am) * 2 + 1;
    double sa[4];
#pragma unroll
    for (int c = 0; c < 4; c++) sa[c] = rvl_warp_sum((lane < G) ? S.A[lane][c] : 0.0);
    const bool active = lane < G;
    const double g = active ? S.gy[0][lane] : 0.0;
    const double Gy = rvl_warp_sum(g);
    const double Gz = rvl_warp_sum(active ? S.gz[0][lane] : 0.0);
    const int GG = G * G, G1 = G - 1;
    const double gE = (double)G * epsq;
    const double Stot = (sa[0] + sa[1] + sa[2] + sa[3]) * Gy * Gz + (double)GG * (double)G * epsq;
    const double dead_thr = fmax(epsq * 0x1p-55, theta * Stot / ((double)GG * (double)G));
    const int kk = active ? lane : 0;
    const double gz0 = S.gz[0][kk], gz1 = S.gz[1][kk];
    const double lEps = rvl_logq<FAST>(epsq);
    const int stride = rvl_lf_stride(G);

    // ---- y-axis tables over d = lane (the argument of exp in rvl_fill_axis is log g)
    const double inv = 1.0 / ((double)G1 * hy);
    {
        const double t = (double)lane * inv;
        const double lg = active ? -0.5 * t * t : 0.0;
        double p0 = g, p1 = g * lg, p2 = lg, s1 = g, s2 = g * g, s3 = g * g * g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double u0 = __shfl_up_sync(FULL, p0, o), u1 = __shfl_up_sync(FULL, p1, o), u2 = __shfl_up_sync(FULL, p2, o);
            const double d1 = __shfl_down_sync(FULL, s1, o), d2 = __shfl_down_sync(FULL, s2, o), d3 = __shfl_down_sync(FULL, s3, o);
            if (lane >= o) { p0 += u0; p1 += u1; p2 += u2; }
            if (lane + o < 32) { s1 += d1; s2 += d2; s3 += d3; }
        }
        S.tab[0][lane] = p0; S.tab[1][lane] = p1; S.tab[2][lane] = p2;
        S.tab[3][lane] = s1; S.tab[4][lane] = s2; S.tab[5][lane] = s3;
        if (lane == 0) { S.tab[3][32] = 0.0; S.tab[4][32] = 0.0; S.tab[5][32] = 0.0; }
        S.lgf[lane] = (float)((lg - lEps) * 1.4426950408889634);
    }

    // ---- pass A: per-k tables over i, and the x-z marginal
    // min / max of V and max of W are only used in the (conservative) column tests of pass B, so they are
    // tracked on the high words (non-negative doubles order like their bit patterns: one integer min/max
    // instead of a NaN-aware FP64 compare-select) and widened outwards afterwards.
    double LV = 0.0, LLV = 0.0, LW = 0.0, LLW = 0.0, accxz = 0.0; // (sum_i V and sum_i W are sM and sm below, in closed form)
    int minVh = 0x7ff00000, maxVh = 0, maxWh = 0;
    float *lfV = lf + kk * stride, *lfW = lf + (G + kk) * stride;
    const double2 *pM = reinterpret_cast<const double2 *>(&S.A[0][M0]), *pm = reinterpret_cast<const double2 *>(&S.A[0][m0]);
#pragma unroll 1
    for (int i = 0; i < G; i++) {
        const double2 aM = pM[2 * i], an = pm[2 * i]; // the (z = 0, z = 1) sums of the majority / minority y-class at grid point i
        const double V = fma(aM.x, gz0, aM.y * gz1);
        const double W = fma(an.x, gz0, an.y * gz1);
        const int Vh = __double2hiint(V), Wh = __double2hiint(W);
        minVh = min(minVh, Vh);
        maxVh = max(maxVh, Vh);
        maxWh = max(maxWh, Wh);
        const double qxz = fma(Gy, V + W, gE);
        // X >= 0 below 2^-963 ~ 1.3e-290 is lifted to at least that by an integer max on its high word: keeps the fast log in
        // range, and closed form and correction use the same log X (see rvl_soft_cell)
        const double Vs = __hiloint2double(max(Vh, 0x03c00000), __double2loint(V));
        const double Ws = __hiloint2double(max(Wh, 0x03c00000), __double2loint(W));
        const double lv = rvl_logq<FAST>(Vs), lw = rvl_logq<FAST>(Ws), lq = rvl_logq<FAST>(qxz);
        LV = fma(V, lv, LV);
        LLV += lv;
        LW = fma(W, lw, LW);
        LLW += lw;
        // (lanes >= G repeat lane 0's k and store the same two values to the same two addresses: no branch needed)
        lfV[i] = (float)lv * 1.4426950408889634f;
        lfW[i] = (float)lw * 1.4426950408889634f;
        accxz = fma(qxz, lq, accxz);
    }
    const double minV = __hiloint2double(minVh, 0);                 // <= the true minimum
    const double maxV = __hiloint2double(maxVh, (int)0xffffffff);   // >= the true maxima
    const double maxW = __hiloint2double(maxWh, (int)0xffffffff);
    if (!active) accxz = 0.0;
    const double sM = am ? fma(sa[2], gz0, sa[3] * gz1) : fma(sa[0], gz0, sa[1] * gz1);
    const double sm = am ? fma(sa[0], gz0, sa[1] * gz1) : fma(sa[2], gz0, sa[3] * gz1);
    __syncwarp();

    // ---- pass B: the column intervals of this k
    const double inv_kappa = 2.0 * ((double)G1 * hy) * ((double)G1 * hy);
    const float ik = (float)inv_kappa;
    const double lmaxV = rvl_logq<FAST>(fmax(maxV, 1e-290)), lminV = rvl_logq<FAST>(fmax(minV, 1e-290)),
                 lmaxW = rvl_logq<FAST>(fmax(maxW, 1e-290));
    const double Ld = rvl_logq<FAST>(fmax(0.5 * dead_thr, 1e-300)); // each class term below half the threshold
    const double L54 = lEps - 37.429947750237048;                  // log(2^-54 eps'): below it a term cannot change fl(Q)
    const int iM = rvl_sqrt_bound((float)(lmaxV - L54) * ik, G1);  // last j' where the majority term is visible at all
    const int im = rvl_sqrt_bound((float)(lmaxW - L54) * ik, G1);  // last d_m where the minority term is visible at all
    // last j' / d_m where a class term can reach the threshold -- never beyond where it is visible: an invisible term
    // leaves fl(Q) = eps' exactly, which is what a dead column contributes
    const int vM = min(rvl_sqrt_bound((float)(lmaxV - Ld) * ik, G1), iM);
    const int vm = min(rvl_sqrt_bound((float)(lmaxW - Ld) * ik, G1), im);
    const int fM = rvl_sqrt_bound((float)(lminV - (lEps + 2.772588722239781)) * ik, G1); // last j' with g_M minV >= 16 eps'
    const int jM = active ? min(vM, G - 2 - im) : -1;              // majority columns: j' in [0, jM] (minority invisible)
    const int hm = active ? min(vm, G - 2 - iM) : -1;              // minority columns: d_m in [0, hm] (majority invisible)
    const int bF = active ? min(fM, G - 2 - im) : -1;              // of the majority columns, factorised: j' in [0, bF]
    const int bC = max(bF, jM);
    const unsigned mSoftM = rvl_bits(bF + 1, jM);                   // majority columns with the cell-by-cell correction
    const unsigned mMin = (hm >= 0) ? rvl_bits(G1 - hm, G1) : 0u;   // minority columns, in j' coordinates
    const unsigned m4 = active ? (rvl_bits(G1 - im, iM) & (rvl_bits(0, vM) | rvl_bits(G1 - vm, G1))) : 0u; // both visible
    const int c4 = __popc(m4), c2M = __popc(mSoftM), c2m = hm + 1;
    const int ndead = active ? G - __popc(rvl_bits(0, bC) | mMin | m4) : 0;
    double accF = 0.0;
    if (bC >= 0) // closed form over the majority columns: g (log g SV + LV) + eps' (G log g + LLV), + G eps' when factorised
        accF = fma(sM, S.tab[1][bC], LV * S.tab[0][bC]) +
               epsq * (fma((double)G, S.tab[2][bC], (double)(bC + 1) * LLV) + (double)((bF + 1) * G));
    if (hm >= 0)
        accF += fma(sm, S.tab[1][hm], LW * S.tab[0][hm]) + epsq * fma((double)G, S.tab[2][hm], (double)(hm + 1) * LLW);
    // list offsets: four-term columns (j << 5 | k) from the front of S.items; soft columns (class << 15 | d << 5 | k, d the
    // distance from the class's own end of the axis) from the back, the majority class's first
    int o4 = c4, oM = c2M, om = c2m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u4 = __shfl_up_sync(FULL, o4, o), uM = __shfl_up_sync(FULL, oM, o), um = __shfl_up_sync(FULL, om, o);
        if (lane >= o) { o4 += u4; oM += uM; om += um; }
    }
    const int n4 = __shfl_sync(FULL, o4, 31), n2M = __shfl_sync(FULL, oM, 31), n2 = n2M + __shfl_sync(FULL, om, 31);
    o4 -= c4;
    oM -= c2M;
    om += n2M - c2m;
    for (unsigned m = m4; m; m &= m - 1u) {
        const int jp = __ffs((int)m) - 1, j = am ? G1 - jp : jp;
        S.items[o4++] = (uint16_t)((j << 5) | lane);
    }
    for (int d = bF + 1; d <= jM; d++) S.items[RVL_ITEMS_CAP - 1 - (oM++)] = (uint16_t)((d << 5) | lane);
    for (int d = 0; d <= hm; d++) S.items[RVL_ITEMS_CAP - 1 - (om++)] = (uint16_t)((d << 5) | lane | 0x8000);
    *nfour_out = n4;
    *nfact_out = rvl_warp_sum_i(max(bF + 1, 0));
    *nsoft_out = n2;

    // ---- y-z marginal: Q_jk = g[d_M] sM_k + g[d_m] sm_k + G eps', cells = k
    double accyz = 0.0;
    int nyz;
    {
        const double lgE = rvl_logq<FAST>(gE);
        const RvlRegime Ryz = rvl_regime_make(gE, lgE, inv_kappa);
        const double lM = rvl_logq<FAST>(fmax(sM, 1e-290)), lm = rvl_logq<FAST>(fmax(sm, 1e-290));
        double aX = 0.0, aE = 0.0;
        int cE = 0;
        unsigned mask = rvl_cell_regimes(S, Ryz, G, active, sM, sm, lM, lm, aX, aE, cE);
        nyz = __popc(mask);
        while (mask) { // a handful per lane
            const int jp = __ffs((int)mask) - 1;
            mask &= mask - 1u;
            const double q = fma(S.gy[0][jp], sM, fma(S.gy[0][G1 - jp], sm, gE));
            accyz = fma(q, rvl_logq<FAST>(q), accyz);
        }
        accyz += aX + gE * aE + (double)cE * (gE * lgE);
    }
    *nyz_out = rvl_warp_sum_i(nyz);
    __syncwarp();

    // ---- pass C, four-term columns: lanes = columns; four grid points per trip, level by level: the four cells'
    // FMA chains and the four logs advance side by side
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    for (int idx = lane; idx < n4; idx += 32) {
        const int p = S.items[idx];
        const int j = p >> 5, k = p & 31;
        const double w00 = S.gy[0][j] * S.gz[0][k], w01 = S.gy[0][j] * S.gz[1][k];
        const double w10 = S.gy[1][j] * S.gz[0][k], w11 = S.gy[1][j] * S.gz[1][k];
        int i = 0;
#pragma unroll 1
        for (; i + 3 < G; i += 4) {
            double2 lo = *reinterpret_cast<const double2 *>(&S.A[i][0]), hi = *reinterpret_cast<const double2 *>(&S.A[i][2]);
            const double q0 = fma(lo.x, w00, fma(lo.y, w01, fma(hi.x, w10, fma(hi.y, w11, epsq))));
            lo = *reinterpret_cast<const double2 *>(&S.A[i + 1][0]); hi = *reinterpret_cast<const double2 *>(&S.A[i + 1][2]);
            const double q1 = fma(lo.x, w00, fma(lo.y, w01, fma(hi.x, w10, fma(hi.y, w11, epsq))));
            lo = *reinterpret_cast<const double2 *>(&S.A[i + 2][0]); hi = *reinterpret_cast<const double2 *>(&S.A[i + 2][2]);
            const double q2 = fma(lo.x, w00, fma(lo.y, w01, fma(hi.x, w10, fma(hi.y, w11, epsq))));
            lo = *reinterpret_cast<const double2 *>(&S.A[i + 3][0]); hi = *reinterpret_cast<const double2 *>(&S.A[i + 3][2]);
            const double q3 = fma(lo.x, w00, fma(lo.y, w01, fma(hi.x, w10, fma(hi.y, w11, epsq))));
            double l0, l1, l2, l3;
            if (FAST) rvl_log_pos4(q0, q1, q2, q3, l0, l1, l2, l3);
            else { l0 = log(q0); l1 = log(q1); l2 = log(q2); l3 = log(q3); }
            acc0 = fma(q0, l0, acc0);
            acc1 = fma(q1, l1, acc1);
            acc2 = fma(q2, l2, acc2);
            acc3 = fma(q3, l3, acc3);
        }
#pragma unroll 1
        for (; i < G; i++) {
            const double q0 = fma(S.A[i][0], w00, fma(S.A[i][1], w01, fma(S.A[i][2], w10, fma(S.A[i][3], w11, epsq))));
            acc0 = fma(q0, rvl_logq<FAST>(q0), acc0);
        }
    }
    // ---- pass C, soft cells: the single-precision correction of the one-class columns (see rvl_soft_cell); four grid
    // points per 16-byte load, one float sum per column widened to double
    double accS = 0.0;
    for (int idx = lane; idx < n2; idx += 32) {
        const int e = S.items[RVL_ITEMS_CAP - 1 - idx];
        const int k = e & 31;
        const float lg = S.lgf[(e >> 5) & 31];
        const float *row = lf + (((e >> 15) ? G : 0) + k) * stride;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int i = 0;
#pragma unroll 1
        for (; i + 3 < G; i += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(row + i);
            a0 += rvl_soft_cell(v.x + lg);
            a1 += rvl_soft_cell(v.y + lg);
            a2 += rvl_soft_cell(v.z + lg);
            a3 += rvl_soft_cell(v.w + lg);
        }
        for (; i < G; i++) a0 += rvl_soft_cell(row[i] + lg);
        accS += (double)((a0 + a1) + (a2 + a3));
    }
    const double L3 = rvl_warp_sum((acc0 + acc1) + (acc2 + acc3) + fma(epsq * 0.6931471805599453094, accS, accF)) +
                      (double)rvl_warp_sum_i(ndead) * ((double)G * (epsq * lEps));
    const double Lyz = rvl_warp_sum(accyz);
    const double Lxz = rvl_warp_sum(accxz);
    double accz = 0.0;
    if (active) {
        double qz = fma((sa[0] + sa[2]) * Gy, gz0, fma((sa[1] + sa[3]) * Gy, gz1, (double)GG * epsq));
        accz = rvl_xlogx<FAST>(qz);
    }
    const double Lz = rvl_warp_sum(accz);
    return (L3 - Lxz - Lyz + Lz) / Stot;
}

// 2-D epilogue (mutual.inf.v2, LIB:2536-2550): MI from B_a[i] = A[i][a*2] and gy.  Whole warp.
template <bool FAST>
__device__ double rvl_mi_epilogue(RvlWarpScratch &S, int G, double epsq, int lane)
{
    double b0 = (lane < G) ? S.A[lane][0] : 0.0, b1 = (lane < G) ? S.A[lane][2] : 0.0;
    double sb0 = rvl_warp_sum(b0), sb1 = rvl_warp_sum(b1);
    double gyv = (lane < G) ? S.gy[0][lane] : 0.0;
    double Gy = rvl_warp_sum(gyv);
    int GG = G * G;
    double gE = (double)G * epsq;
    double acc = 0.0;
    for (int p = lane; p < GG; p += 32) {
        int i = p / G, j = p - i * G;
        double q = fma(S.A[i][0], S.gy[0][j], fma(S.A[i][2], S.gy[1][j], epsq));
        acc += rvl_xlogx<FAST>(q);
    }
    double accx = 0.0, accy = 0.0;
    if (lane < G) {
        double qx = fma(b0 + b1, Gy, gE);
        double qy = fma(sb0, S.gy[0][lane], fma(sb1, S.gy[1][lane], gE));
        accx = rvl_xlogx<FAST>(qx);
        accy = rvl_xlogx<FAST>(qy);
    }
    double L2 = rvl_warp_sum(acc), Lx = rvl_warp_sum(accx), Ly = rvl_warp_sum(accy);
    double Stot = (sb0 + sb1) * Gy + (double)GG * epsq;
    return (L2 - Lx - Ly) / Stot + log(Stot);
}

__device__ __forceinline__ double rvl_sign(double r) { return (double)((r > 0.0) - (r < 0.0)); }

// Fill the axis kernel of a non-constant binary variable: g[0][j] = exp(-(j/(G-1))^2/(2h^2)),
// g[1][j] = g[0][G-1-j].
__device__ __forceinline__ void rvl_fill_axis(double g[2][RVL_MAXG], int G, double h, int lane)
{
    if (lane < G) {
        double t = (double)lane / (double)(G - 1) / h;
        double v = rvl_exp_neg(-0.5 * t * t);
        g[0][lane] = v;
        g[1][G - 1 - lane] = v;
    }
}

// Pearson correlation of x with a 0/1 vector having n1 ones, s1c = sum_{y=1}(x - mean).
__device__ __forceinline__ double rvl_rho(double s1c, int n1, int n, double sxx)
{
    double r = s1c / sqrt(sxx * ((double)n1 * (double)(n - n1) / (double)n));
    return fmin(1.0, fmax(-1.0, r));
}

// One warp: popcount of a bit row and sum of the centred target over its set bits (fixed order: lane
// owns words lane, lane+32, ...; then the shuffle tree).  Optionally copies the row to shared memory.
__device__ __forceinline__ void rvl_row_stats(const uint64_t *__restrict__ row, int words, const double *__restrict__ xc,
                                              int lane, uint64_t *yrow, int &n1, double &s1c)
{
    n1 = 0;
    s1c = 0.0;
    for (int w = lane; w < words; w += 32) {
        uint64_t v = row[w];
        if (yrow) yrow[w] = v;
        n1 += __popcll(v);
        while (v) {
            int b = __ffsll((long long)v) - 1;
            v &= v - 1;
            s1c += xc[w * 64 + b];
        }
    }
    n1 = rvl_warp_sum_i(n1);
    s1c = rvl_warp_sum(s1c);
}

// Largest rho^+ and |rho| over the loaded rows, per target (non-negative doubles order like their bit
// patterns, so atomicMax on the bits is exact).  grid = (row blocks, targets), one warp per row.
__global__ void k_rho_max(RvlParams P, const uint64_t *__restrict__ bits, const double *__restrict__ xc_all,
                          RvlTarget *tg_all, const int64_t *__restrict__ uniq, int64_t U)
{
    const int t = blockIdx.y, lane = threadIdx.x & 31;
    const int64_t nrows = uniq ? U : P.F;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    const double sxx = tg_all[t].sxx;
    if (tg_all[t].is_const) return;
    double pos = 0.0, ab = 0.0;
    for (int64_t q = wid; q < nrows; q += nw) {
        const int64_t f = uniq ? uniq[q] : q;
        int n1;
        double s1c;
        rvl_row_stats(bits + (size_t)f * P.words, P.words, xc_all + (size_t)t * P.n, lane, nullptr, n1, s1c);
        if (n1 == 0 || n1 == P.n) continue;
        double r = rvl_rho(s1c, n1, P.n, sxx);
        pos = fmax(pos, r);
        ab = fmax(ab, fabs(r));
    }
    if (lane == 0) {
        atomicMax(&tg_all[t].rho_pos_max_bits, (unsigned long long)__double_as_longlong(pos));
        atomicMax(&tg_all[t].rho_abs_max_bits, (unsigned long long)__double_as_longlong(ab));
    }
}

// Finish one evaluation once A (dense sums) is in scratch.  Returns IC or CIC.
// mode: RVL_MODE_IC (2-D) or RVL_MODE_CIC (3-D).  c is the bandwidth shrink factor already applied
// to hx (beta); here it is applied to hy / hz.
// am: majority y-class (0 unless the row has more ones than zeros).  nitems_out: cube cells evaluated with a log of
// their own; nyz_out: the same for the y-z marginal; nfact_out: columns in factorised form; ntwo_out: one-class columns
// whose eps-floor correction is summed cell by cell in single precision (soft cells).  lf: the warp's float table.
__device__ double rvl_finish(RvlWarpScratch &S, float *lf, const RvlParams &P, int mode, double rho, double c, double hx,
                             double bcv_y, double bcv_z, int am, int lane, int *nitems_out, int *nyz_out, int *nfact_out,
                             int *ntwo_out)
{
    *nitems_out = 0;
    *nyz_out = 0;
    *nfact_out = 0;
    *ntwo_out = 0;
    int G = P.G;
    double dn = (double)P.n;
    if (mode == RVL_MODE_IC) {
        double hy = bcv_y * c / 4; // kde2d: h <- h/4
        rvl_fill_axis(S.gy, G, hy, lane);
        __syncwarp();
        double epsq = RVL_DBL_EPS * (RVL_TWO_PI * dn * hx * hy);
        double mi = (epsq >= 1e-290) ? rvl_mi_epilogue<true>(S, G, epsq, lane) : rvl_mi_epilogue<false>(S, G, epsq, lane);
        return rvl_sign(rho) * sqrt(1.0 - exp(-2.0 * mi));
    }
    double hy = P.bw_mult * bcv_y * c, hz = P.bw_mult * bcv_z * c;
    rvl_fill_axis(S.gy, G, hy, lane);
    rvl_fill_axis(S.gz, G, hz, lane);
    __syncwarp();
    double epsq = RVL_DBL_EPS * (dn * RVL_TWO_PI * sqrt(RVL_TWO_PI) * hx * hy * hz);
    // every Q lies in [eps', 4 n G^3]: normal doubles unless eps' itself is tiny (degenerate bandwidths)
    double cmi;
    if (epsq < 1e-140) { // also keeps 1/(6 E^2) of the small-regime closed forms finite
        int nl;
        cmi = rvl_cmi_epilogue<false>(S, G, epsq, P.prune_theta, lane, &nl);
        *nitems_out = nl * G;
    } else if (P.literal) { // validation mode: the cube cell by cell
        int nl;
        cmi = rvl_cmi_epilogue<true>(S, G, epsq, P.prune_theta, lane, &nl);
        *nitems_out = nl * G;
    } else {
        int nd, nf, n2;
        cmi = rvl_cmi_epilogue_f<true>(S, lf, G, epsq, P.prune_theta, hy, am, lane, &nd, &nf, &n2, nyz_out);
        *nitems_out = nd * G;
        *nfact_out = nf;
        *ntwo_out = n2;
    }
    return rvl_sign(rho) * sqrt(1.0 - exp(-2.0 * cmi));
}

// ---- x-axis sums: the bandwidth-interpolation table ---------------------------------------------
// The x-axis bandwidth of a feature is hx = hx0 * c with c = 1 + shrink * rho^+ (LIB:2571-2573), so
//   T_b(c)[i] = sum_{s: z_s = b} exp(-beta0 * t * (i - u_s)^2),   t = 1/c^2,
// the per-z-class total over ALL samples, depends on the feature only through the scalar t.  As a
// function of tau = ln t it is a sum of exp(-a_s e^tau): entire, and 8-point Lagrange interpolation
// on nodes spaced 1/128 in tau reproduces it to ~1e-16 of the class size (checked in
// tests/test_ttable_interp.py against direct summation).  The table is built once per (target, search) by
// k_advance -- M * n * G exps, the cost of ~M dense passes instead of one per feature -- then updated in place
// after every merge, and k_score only visits the minority y-class of each row bit by bit.
// Node tt_pad is tau = 0 exactly (c = 1: every feature with rho <= 0), so those use it unweighted.
// Layout: tt[target][node][b][RVL_MAXG].

// Reduce the per-warp bests k_score left in part[nparts] (one target) and publish the candidate
// record [score][global row][bit row].  With a peer exchange (X.peers) the record goes straight into
// every rank's buffer over NVLink and is followed by a stamp; otherwise into `send` for the caller's
// all-gather.  Called by a whole block (any size that is a multiple of 32).
template <bool COH>
__device__ void rvl_publish_best(const RvlParams &P, const RvlCandHead *v, int nparts,
                                 const uint64_t *__restrict__ bits, unsigned char *__restrict__ send, size_t rec_bytes,
                                 const RvlPeerXchg &X, int t)
{
    __shared__ double shv[32];
    __shared__ long long shi[32];
    __syncthreads(); // k_search calls this once per target: shv / shi are reused
    double bv = 0.0;
    int64_t bi = -1;
    for (int k0 = threadIdx.x; k0 < nparts; k0 += 8 * blockDim.x) { // eight loads in flight per thread, then the compares
        double hs[8];
        int64_t hr[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int k = k0 + q * blockDim.x;
            hr[q] = -1;
            hs[q] = 0.0;
            if (k < nparts) {
                hs[q] = rvl_ld<COH>(&v[k].score);
                hr[q] = (int64_t)rvl_ld<COH>(reinterpret_cast<const uint64_t *>(&v[k].row));
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (rvl_ahead(hs[q], hr[q], bv, bi, P.negative)) { bv = hs[q]; bi = hr[q]; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        long long oi = __shfl_xor_sync(0xffffffffu, (long long)bi, o);
        if (rvl_ahead(ov, oi, bv, bi, P.negative)) { bv = ov; bi = oi; }
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { shv[w] = bv; shi[w] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        int nw = blockDim.x >> 5;
        bv = (lane < nw) ? shv[lane] : 0.0;
        bi = (lane < nw) ? shi[lane] : -1;
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            long long oi = __shfl_xor_sync(0xffffffffu, (long long)bi, o);
            if (rvl_ahead(ov, oi, bv, bi, P.negative)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { shv[0] = bv; shi[0] = bi; }
    }
    __syncthreads();
    bv = shv[0]; bi = shi[0];
    const int ndst = X.peers ? P.world : 1;
    const size_t slot = ((size_t)X.parity * P.world + P.rank) * P.T + t;
    for (int r = 0; r < ndst; r++) {
        unsigned char *rec = X.peers ? X.peers[r] + slot * rec_bytes : send + (size_t)t * rec_bytes;
        RvlCandHead *h = reinterpret_cast<RvlCandHead *>(rec);
        uint64_t *rb = reinterpret_cast<uint64_t *>(rec + sizeof(RvlCandHead));
        if (threadIdx.x == 0) { h->score = bv; h->row = (bi < 0) ? -1 : (P.row_offset + bi); }
        for (int i = threadIdx.x; i < P.words; i += blockDim.x) rb[i] = (bi < 0) ? 0ull : bits[(size_t)bi * P.words + i];
    }
    if (X.peers) {
        if (P.world > 1) __threadfence_system(); // records visible to every GPU before any stamp
        else __threadfence();
        __syncthreads();
        if (threadIdx.x < P.world) {
            volatile unsigned int *st =
                reinterpret_cast<volatile unsigned int *>(X.peers[threadIdx.x] + X.stamps_off) + slot;
            *st = X.stamp;
        }
    }
}


// k_advance = everything between two scoring passes, in one launch (grid = table nodes x targets):
//   * merge (when `recv` is given): every block resolves, for its target's seed owner, the same global
//     winner among the `world` candidate records (max score, ties -> lowest GLOBAL row, NaN last;
//     LIB:1858-1862) and forms the new seed = old seed | winner row (LIB:2168) on the fly; the block of
//     node 0 publishes it (seeds_out, seed.iter[iter,], top, top.score).  Seeds are double-buffered so
//     that no block reads what another writes.
//   * seed state: popcount -> cond.assoc's dispatch (LIB:2507-2517); node-0 block writes it to RvlTarget,
//     resets the target's per-warp best slots and (target 0) k_score's work counter.
//   * x-axis table: rebuilt from all samples (first iteration, or when the dispatch mode changed), or
//     updated in place from the samples the winner just moved from z-class 0 to z-class 1
//     (T_1 += d, T_0 -= d with d summed over those few samples): O(n1) instead of O(n) per node.
// In-place update of one table node by ONE warp: the samples the winner moved from z-class 0 to z-class 1 (the new
// bits of the seed) leave T_0 and enter T_1.  Lanes own grid points; the new samples are taken 32 at a time: lane l
// locates the l-th new bit and loads its target coordinate (all loads in flight together), then the warp walks the
// batch in ascending sample order with shuffles -- a fixed order, so the last bit of every sum is the same on every
// rank, in every run and in both search drivers.  t0_old / t1_old: the node's current values for this lane's grid point.
template <bool COH>
__device__ __forceinline__ void rvl_table_incr_warp(const double *__restrict__ u, const uint64_t *zold, const uint64_t *wrow,
                                                    int words, double beta, double *out, double t0_old, double t1_old, int lane)
{
    const unsigned FULL = 0xffffffffu;
    const double gi = (double)lane;
    double d = 0.0;
    for (int base = 0; base < words; base += 32) {
        const int wl = base + lane;
        const uint64_t mine = (wl < words) ? (rvl_ld<COH>(wrow + wl) & ~rvl_ld<COH>(zold + wl)) : 0ull; // entering z-class 1
        const int cnt = __popcll(mine);
        int pre = cnt; // inclusive prefix of the bit counts over lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(FULL, pre, o);
            if (lane >= o) pre += up;
        }
        const int total = __shfl_sync(FULL, pre, 31);
        pre -= cnt;
        const unsigned owners = __ballot_sync(FULL, cnt > 0);
        for (int b0 = 0; b0 < total; b0 += 32) {
            // lane l takes new bit number b0 + l (ascending sample order): find the word that holds it
            const int j = b0 + lane;
            uint64_t wv = 0ull;
            int rank = 0, wi = 0;
            for (unsigned rest = owners; rest; rest &= rest - 1u) {
                const int src = __ffs((int)rest) - 1;
                const int ps = __shfl_sync(FULL, pre, src), cs = __shfl_sync(FULL, cnt, src);
                const uint64_t ms = __shfl_sync(FULL, mine, src);
                if (j >= ps && j < ps + cs) { wv = ms; rank = j - ps; wi = base + src; }
            }
            double uval = 0.0;
            if (j < total) {
                const unsigned lo = (unsigned)wv, hi = (unsigned)(wv >> 32);
                const int clo = __popc(lo);
                const int bit = rank < clo ? (int)__fns(lo, 0, rank + 1) : 32 + (int)__fns(hi, 0, rank - clo + 1);
                uval = u[wi * 64 + bit];
            }
            const int nb = min(32, total - b0);
#pragma unroll 4
            for (int k = 0; k < nb; k++) {
                const double dd = gi - __shfl_sync(FULL, uval, k);
                d += rvl_exp_neg(-beta * dd * dd);
            }
        }
    }
    out[lane] = t0_old - d;
    out[RVL_MAXG + lane] = t1_old + d;
}

// k_search: every block waits ONCE per iteration for the stamps of all `world` ranks and all seed-owning targets (its
// first warp polls, with a short back-off so that a few hundred pollers do not crowd the L2 line the peers' stores and
// stamps land in); afterwards the candidate records of the iteration are readable by every thread of the block.
__device__ __forceinline__ void rvl_wait_stamps(const RvlParams &P, const RvlPeerXchg &X, int nreal)
{
    if (threadIdx.x < 32) {
        for (int idx = threadIdx.x; idx < nreal * P.world; idx += 32) {
            const int r = idx / nreal, t = idx - r * nreal;
            volatile unsigned int *st = reinterpret_cast<volatile unsigned int *>(X.local + X.stamps_off) +
                                        ((size_t)X.parity * P.world + r) * P.T + t;
            const long long t0 = clock64();
            while (*st != X.stamp) {
                __nanosleep(100);
                if (clock64() - t0 > 4000000000ll) { // ~2 s: a peer is gone; fail loudly instead of hanging
                    *X.err = 1;
                    break;
                }
            }
        }
        if (P.world > 1) __threadfence_system();
        else __threadfence();
    }
    __syncthreads();
}

// One (table node m, target t) task of the merge step done by ONE warp (k_search): wait for the ranks' stamps, resolve
// the global winner, node-0 duties (new seed, seed.iter, top, seed state), in-place table update.  Returns false when
// the table of this target must be REBUILT (the dispatch mode changed): that takes a block (rvl_advance_task).
template <bool COH>
__device__ bool rvl_advance_warp(const RvlParams &P, RvlTarget *tg_all, const double *__restrict__ u_all, const uint64_t *seeds_in,
                                 uint64_t *seeds_out, const double *__restrict__ bcvtab, double *tt, size_t rec_bytes, int iter,
                                 int iters, int first_null, int64_t *__restrict__ top, double *__restrict__ top_score,
                                 uint64_t *__restrict__ seed_iter, const RvlPeerXchg &X, int t, int m, bool last_iter, int lane)
{
    // m >= 0: table node m of target t.  m == -1: the target's bookkeeping instead (new seed, seed.iter, top, seed state)
    // -- a task of its own so that no table node waits for it.
    const unsigned FULL = 0xffffffffu;
    const RvlTarget tg = tg_all[t]; // only the per-target constants are used (never nz / mode / bcv_z)
    const int owner = tg.seed_of;
    const bool is_null = (first_null > 0 && t >= first_null);
    const uint64_t *zold = seeds_in + (size_t)owner * P.words;
    const unsigned char *recv = X.local + (size_t)X.parity * P.world * P.T * rec_bytes;
    // what does not depend on the winner is fetched before the wait: this node's current values
    double *out = tt + (((size_t)t * P.tt_nodes + (m < 0 ? 0 : m)) * 2) * RVL_MAXG;
    const bool table = !(P.dense_x || last_iter) && m >= 0;
    const double t0_old = table ? rvl_ld<COH>(out + lane) : 0.0, t1_old = table ? rvl_ld<COH>(out + RVL_MAXG + lane) : 0.0;
    // the records of this iteration sit in the local buffer once every rank's stamp for (parity, rank, owner) is there
    double bv = 0.0;
    int64_t bi = -1;
    int br = 0;
    for (int r0 = 0; r0 < P.world; r0 += 32) {
        const int r = r0 + lane;
        double hs = 0.0;
        int64_t hr = -1;
        if (r < P.world) { // (the block has waited for every stamp of this iteration: rvl_wait_stamps)
            const unsigned char *rec = recv + ((size_t)r * P.T + owner) * rec_bytes;
            hs = rvl_ld<true>(reinterpret_cast<const double *>(rec));
            hr = (int64_t)rvl_ld<true>(reinterpret_cast<const uint64_t *>(rec + sizeof(double)));
        }
        if (rvl_ahead(hs, hr, bv, bi, P.negative)) { bv = hs; bi = hr; br = r; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { // the same total order on every lane, warp and rank: order()[1]
        const double ov = __shfl_xor_sync(FULL, bv, o);
        const long long oi = __shfl_xor_sync(FULL, (long long)bi, o);
        const int orr = __shfl_xor_sync(FULL, br, o);
        if (rvl_ahead(ov, oi, bv, bi, P.negative)) { bv = ov; bi = oi; br = orr; }
    }
    const uint64_t *wrow = reinterpret_cast<const uint64_t *>(recv + ((size_t)br * P.T + owner) * rec_bytes + sizeof(RvlCandHead));
    int c0 = 0, c1 = 0;
    for (int w = lane; w < P.words; w += 32) {
        const uint64_t zo = rvl_ld<COH>(zold + w), zn = zo | rvl_ld<true>(wrow + w);
        c0 += __popcll(zo);
        c1 += __popcll(zn);
    }
    const int nz_old = rvl_warp_sum_i(c0), nz = rvl_warp_sum_i(c1);
    const int mode_old = tg.is_const ? RVL_MODE_ZERO : ((nz_old == 0 || nz_old == P.n) ? RVL_MODE_IC : RVL_MODE_CIC);
    const int mode = tg.is_const ? RVL_MODE_ZERO : ((nz == 0 || nz == P.n) ? RVL_MODE_IC : RVL_MODE_CIC);
    if (m < 0) {
        if (lane == 0) {
            tg_all[t].nz = nz;
            tg_all[t].bcv_z = bcvtab[nz];
            tg_all[t].mode = mode;
            if (!is_null) {
                top[(size_t)t * iters + iter] = bi;
                top_score[(size_t)t * iters + iter] = bv;
            }
        }
        if (!is_null) { // apply(rbind(seed, m.2[top,]), 2, "max")  LIB:2168
            uint64_t *zo = seeds_out + (size_t)t * P.words;
            uint64_t *zi = seed_iter + ((size_t)t * iters + iter) * P.words;
            for (int w = lane; w < P.words; w += 32) {
                const uint64_t v = rvl_ld<COH>(zold + w) | rvl_ld<true>(wrow + w);
                zo[w] = v;
                zi[w] = v;
            }
        }
        return true;
    }
    if (!table) return true; // no table kept / nobody will read it
    if (mode == RVL_MODE_ZERO) {
        out[lane] = 0.0;
        out[RVL_MAXG + lane] = 0.0;
        return true;
    }
    const double rmax = __longlong_as_double((long long)(mode == RVL_MODE_IC ? tg.rho_abs_max_bits : tg.rho_pos_max_bits));
    const double cmin = 1.0 + P.bw_shrink * fmin(rmax * (1.0 + 1e-9) + 1e-12, 1.0);
    const int needed = (int)ceil(-2.0 * log(cmin) / P.tt_dt) + 2 * P.tt_pad + 1;
    if (m >= needed) return true;
    const bool incr = mode == RVL_MODE_CIC &&
                      (mode_old == RVL_MODE_CIC || (mode_old == RVL_MODE_IC && nz_old == 0 && P.bw_mult == 0.25));
    if (!incr) return false;
    const double hx0 = P.bw_mult * tg.bcv;
    const double dx = (tg.xmax - tg.xmin) / (double)(P.G - 1);
    const double beta0 = 0.5 * (dx / hx0) * (dx / hx0);
    const double tau = (double)(m - P.tt_pad) * P.tt_dt;
    const double beta = (m == P.tt_pad) ? beta0 : beta0 * exp(tau);
    rvl_table_incr_warp<COH>(u_all + (size_t)t * P.n, zold, wrow, P.words, beta, out, t0_old, t1_old, lane);
    return true;
}

// Block-shared workspace of rvl_advance_task: static in k_advance, carved out of the (idle) per-warp scratch in k_search.
struct __align__(16) RvlAdvanceShared {
    double part[RVL_SCORE_WARPS][2][RVL_MAXG];
    double su[1024];      // staged samples of the full build
    uint64_t sz[16];
    int s_nz_old, s_nz_new, s_win;
};

// One (table node m, target t) task of the step between two scoring passes; see k_advance.  Whole block.
template <bool COH>
__device__ void rvl_advance_task(const RvlParams &P, RvlTarget *tg_all, const double *__restrict__ u_all, const uint64_t *seeds_in,
                                 uint64_t *seeds_out, const double *__restrict__ bcvtab, double *tt, RvlCandHead *slots, int nslots,
                                 unsigned long long *work, const unsigned char *recv, size_t rec_bytes, int iter, int iters,
                                 int first_null, int64_t *__restrict__ top, double *__restrict__ top_score,
                                 uint64_t *__restrict__ seed_iter, int allow_incr, const RvlPeerXchg &X, int t, int m, RvlAdvanceShared &sh,
                                 bool full_only = false)
{
    double (&part)[RVL_SCORE_WARPS][2][RVL_MAXG] = sh.part;
    int &s_nz_old = sh.s_nz_old, &s_nz_new = sh.s_nz_new, &s_win = sh.s_win;
    __syncthreads(); // k_search runs several tasks per block: the shared arrays below are reused
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const RvlTarget tg = tg_all[t]; // only the per-target constants are used below (never nz / mode / bcv_z)
    const int owner = tg.seed_of;
    const bool is_null = (first_null > 0 && t >= first_null);
    const uint64_t *zold = seeds_in + (size_t)owner * P.words;
    const uint64_t *wrow = nullptr; // the winner's bit row (merge only)
    if (recv && X.peers) {
        // peer exchange: the records of this iteration sit in the local buffer once every rank's stamp
        // for (parity, rank, owner) has arrived
        recv = X.local + (size_t)X.parity * P.world * P.T * rec_bytes;
        if (threadIdx.x < P.world) {
            volatile unsigned int *st = reinterpret_cast<volatile unsigned int *>(X.local + X.stamps_off) +
                                        ((size_t)X.parity * P.world + threadIdx.x) * P.T + owner;
            const long long t0 = clock64();
            while (*st != X.stamp) {
                if (clock64() - t0 > 4000000000ll) { // ~2 s: a peer is gone; fail loudly instead of hanging
                    *X.err = 1;
                    break;
                }
            }
        }
        __threadfence_system();
        __syncthreads();
    }
    if (recv) {
        if (threadIdx.x == 0) {
            double bv = 0.0;
            int64_t bi = -1;
            int br = 0;
            for (int r = 0; r < P.world; r++) {
                const unsigned char *rec = recv + ((size_t)r * P.T + owner) * rec_bytes;
                const double hs = rvl_ld<COH>(reinterpret_cast<const double *>(rec));
                const int64_t hr = (int64_t)rvl_ld<COH>(reinterpret_cast<const uint64_t *>(rec + sizeof(double)));
                if (rvl_ahead(hs, hr, bv, bi, P.negative)) { bv = hs; bi = hr; br = r; }
            }
            s_win = br;
            if (m == 0 && !is_null) {
                top[(size_t)t * iters + iter] = bi;
                top_score[(size_t)t * iters + iter] = bv;
            }
        }
        __syncthreads();
        wrow = reinterpret_cast<const uint64_t *>(recv + ((size_t)s_win * P.T + owner) * rec_bytes + sizeof(RvlCandHead));
    }
    if (warp == 0) {
        int c0 = 0, c1 = 0;
        for (int w = lane; w < P.words; w += 32) {
            const uint64_t zo = rvl_ld<COH>(zold + w), zn = wrow ? (zo | rvl_ld<COH>(wrow + w)) : zo;
            c0 += __popcll(zo);
            c1 += __popcll(zn);
        }
        c0 = rvl_warp_sum_i(c0);
        c1 = rvl_warp_sum_i(c1);
        if (lane == 0) { s_nz_old = c0; s_nz_new = c1; }
    }
    __syncthreads();
    const int nz_old = s_nz_old, nz = s_nz_new;
    const int mode_old = tg.is_const ? RVL_MODE_ZERO : ((nz_old == 0 || nz_old == P.n) ? RVL_MODE_IC : RVL_MODE_CIC);
    const int mode = tg.is_const ? RVL_MODE_ZERO : ((nz == 0 || nz == P.n) ? RVL_MODE_IC : RVL_MODE_CIC);
    if (m == 0) {
        if (threadIdx.x == 0) {
            tg_all[t].nz = nz;
            tg_all[t].bcv_z = bcvtab[nz];
            tg_all[t].mode = mode;
            if (work && t == 0) *work = 0ull;
        }
        for (int i = threadIdx.x; slots && i < nslots; i += blockDim.x) {
            RvlCandHead h;
            h.score = 0.0;
            h.row = -1;
            slots[(size_t)t * nslots + i] = h;
        }
        if (recv && !is_null) { // apply(rbind(seed, m.2[top,]), 2, "max")  LIB:2168
            uint64_t *zo = seeds_out + (size_t)t * P.words;
            uint64_t *zi = seed_iter + ((size_t)t * iters + iter) * P.words;
            for (int w = threadIdx.x; w < P.words; w += blockDim.x) {
                const uint64_t v = rvl_ld<COH>(zold + w) | rvl_ld<COH>(wrow + w);
                zo[w] = v;
                zi[w] = v;
            }
        }
    }
    if (P.dense_x) return; // literal mode keeps no table
    double *out = tt + (((size_t)t * P.tt_nodes + m) * 2) * RVL_MAXG;
    if (mode == RVL_MODE_ZERO) {
        if (threadIdx.x < 2 * RVL_MAXG) out[threadIdx.x] = 0.0;
        return;
    }
    // nodes beyond the largest shrink any loaded row can produce are never read (k_rho_max)
    const double rmax = __longlong_as_double((long long)(mode == RVL_MODE_IC ? tg.rho_abs_max_bits : tg.rho_pos_max_bits));
    const double cmin = 1.0 + P.bw_shrink * fmin(rmax * (1.0 + 1e-9) + 1e-12, 1.0);
    const int needed = (int)ceil(-2.0 * log(cmin) / P.tt_dt) + 2 * P.tt_pad + 1;
    if (m >= needed) return;
    const double hx0 = (mode == RVL_MODE_IC) ? tg.bcv / 4 : P.bw_mult * tg.bcv;
    const double dx = (tg.xmax - tg.xmin) / (double)(P.G - 1);
    const double beta0 = 0.5 * (dx / hx0) * (dx / hx0);
    const double tau = (double)(m - P.tt_pad) * P.tt_dt;
    const double beta = (m == P.tt_pad) ? beta0 : beta0 * exp(tau);
    const double *u = u_all + (size_t)t * P.n;
    const double gi = (double)lane;

    // in-place update: valid when the table on record was built for the same x bandwidth and the same
    // assignment of the untouched samples -- CIC before and after, or the all-zero seed (every sample in
    // class 0, IC dispatch) whose x bandwidth bcv/4 coincides with bw_mult*bcv when bw_mult = 0.25
    const bool incr = allow_incr && recv && mode == RVL_MODE_CIC &&
                      (mode_old == RVL_MODE_CIC || (mode_old == RVL_MODE_IC && nz_old == 0 && P.bw_mult == 0.25));
    if (full_only && incr) return; // k_search: the per-warp pass has done this task
    double *su = sh.su;
    uint64_t *sz = sh.sz;
    if (incr) { // O(new bits) per node: one warp is plenty (the same routine serves k_search's per-warp tasks)
        if (warp == 0)
            rvl_table_incr_warp<COH>(u, zold, wrow, P.words, beta, out, rvl_ld<COH>(out + lane), rvl_ld<COH>(out + RVL_MAXG + lane), lane);
        return;
    }

    // full build: samples staged through shared memory in chunks of 1024 (coalesced loads, broadcast reads)
    const bool use_z = (mode == RVL_MODE_CIC);
    double a0 = 0.0, a1 = 0.0;
    for (int base = 0; base < P.n; base += 1024) {
        const int cnt = min(1024, P.n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) su[i] = u[base + i];
        if (threadIdx.x < 16) {
            const int w = (base >> 6) + threadIdx.x;
            uint64_t zz = 0ull;
            if (use_z && w < P.words) zz = wrow ? (rvl_ld<COH>(zold + w) | rvl_ld<COH>(wrow + w)) : rvl_ld<COH>(zold + w);
            sz[threadIdx.x] = zz;
        }
        __syncthreads();
        // four independent exp chains in flight per lane; fixed summation order (s ascending per warp)
        int s = warp;
        for (; s + 3 * nwarps < cnt; s += 4 * nwarps) {
            const int s1 = s + nwarps, s2 = s + 2 * nwarps, s3 = s + 3 * nwarps;
            const double d0 = gi - su[s], d1 = gi - su[s1], d2 = gi - su[s2], d3 = gi - su[s3];
            const double e0 = rvl_exp_neg(-beta * d0 * d0), e1 = rvl_exp_neg(-beta * d1 * d1),
                         e2 = rvl_exp_neg(-beta * d2 * d2), e3 = rvl_exp_neg(-beta * d3 * d3);
            if ((sz[s >> 6] >> (s & 63)) & 1ull) a1 += e0; else a0 += e0;
            if ((sz[s1 >> 6] >> (s1 & 63)) & 1ull) a1 += e1; else a0 += e1;
            if ((sz[s2 >> 6] >> (s2 & 63)) & 1ull) a1 += e2; else a0 += e2;
            if ((sz[s3 >> 6] >> (s3 & 63)) & 1ull) a1 += e3; else a0 += e3;
        }
        for (; s < cnt; s += nwarps) {
            const double d = gi - su[s];
            const double e = rvl_exp_neg(-beta * d * d);
            if ((sz[s >> 6] >> (s & 63)) & 1ull) a1 += e; else a0 += e;
        }
    }
    part[warp][0][lane] = a0;
    part[warp][1][lane] = a1;
    __syncthreads();
    if (threadIdx.x < 2 * RVL_MAXG) {
        int b = threadIdx.x >> 5;
        double sum = 0.0;
        for (int w = 0; w < nwarps; w++) sum += part[w][b][lane];
        out[b * RVL_MAXG + lane] = sum;
    }
}

__global__ void __launch_bounds__(RVL_SCORE_WARPS * 32)
k_advance(RvlParams P, RvlTarget *tg_all, const double *__restrict__ u_all, const uint64_t *__restrict__ seeds_in,
          uint64_t *__restrict__ seeds_out, const double *__restrict__ bcvtab, double *__restrict__ tt,
          RvlCandHead *__restrict__ slots, int nslots, unsigned long long *__restrict__ work,
          const unsigned char *__restrict__ recv, size_t rec_bytes, int iter, int iters, int first_null,
          int64_t *__restrict__ top, double *__restrict__ top_score, uint64_t *__restrict__ seed_iter, int allow_incr,
          RvlPeerXchg X)
{
    __shared__ RvlAdvanceShared sh;
    rvl_pdl_launch_dependents();
    rvl_pdl_wait();
    rvl_advance_task<false>(P, tg_all, u_all, seeds_in, seeds_out, bcvtab, tt, slots, nslots, work, recv, rec_bytes, iter, iters,
                            first_null, top, top_score, seed_iter, allow_incr, X, blockIdx.y, blockIdx.x, sh);
}

// Per-class totals at shrink factor c for grid point `lane`.
template <bool COH>
__device__ __forceinline__ void rvl_tt_lookup(const RvlParams &P, const double *tt_t, double c, int lane,
                                              double &tot0, double &tot1)
{
    if (c == 1.0) {
        const double *nd = tt_t + (size_t)P.tt_pad * 2 * RVL_MAXG;
        tot0 = rvl_ld<COH>(nd + lane);
        tot1 = rvl_ld<COH>(nd + RVL_MAXG + lane);
        return;
    }
    double pos = -2.0 * log(c) / P.tt_dt;
    double fl = floor(pos);
    double x = pos - fl;
    int m0 = (int)fl + P.tt_pad - (RVL_TT_STENCIL / 2 - 1);
    m0 = max(0, min(m0, P.tt_nodes - RVL_TT_STENCIL));
    x = pos - (double)(m0 - P.tt_pad + (RVL_TT_STENCIL / 2 - 1)); // offset of pos from stencil node 3
    // Lagrange weight of stencil node q = lane (q = 0..7, offsets q-3), then broadcast
    double wl = 0.0;
    if (lane < RVL_TT_STENCIL) {
        double num = 1.0, den = 1.0;
#pragma unroll
        for (int r = 0; r < RVL_TT_STENCIL; r++) {
            if (r != lane) {
                num *= x - (double)(r - (RVL_TT_STENCIL / 2 - 1));
                den *= (double)(lane - r);
            }
        }
        wl = num / den;
    }
    tot0 = 0.0;
    tot1 = 0.0;
    const double *nd = tt_t + (size_t)m0 * 2 * RVL_MAXG;
#pragma unroll
    for (int q = 0; q < RVL_TT_STENCIL; q++) {
        double wq = __shfl_sync(0xffffffffu, wl, q);
        tot0 = fma(wq, rvl_ld<COH>(nd + (size_t)q * 2 * RVL_MAXG + lane), tot0);
        tot1 = fma(wq, rvl_ld<COH>(nd + (size_t)q * 2 * RVL_MAXG + RVL_MAXG + lane), tot1);
    }
}

// ---- k_score --------------------------------------------------------------------------------
// Persistent kernel: grid = SMs x resident CTAs, block = RVL_KSCORE_WARPS warps.  Every warp is an
// independent worker: it draws the next (target, row) item from a global counter (target-major, so
// neighbouring items share the target's small arrays in L1), scores it, and keeps its running best per
// target.  No block barrier anywhere: evaluation cost varies with the number of live density columns,
// and a barrier or a static row split would leave FP64 pipes idle at every CTA tail.
// Dynamic smem: per-warp scratch + the row's bits.  Besides the score of every row, each warp leaves
// its best (score, local row) per target in part[target][global warp id] (pre-set to "none" by
// k_advance) so that the rank step never re-reads the F scores.
// The scoring pass of one warp (see k_score): draws (target, row) items from *work until they run out.  COH: the seeds, the
// per-iteration seed state and the x-axis table may have been written by other blocks of the SAME launch (k_search).
template <bool COH>
__device__ __forceinline__ void rvl_score_phase(const RvlParams &P, const uint64_t *__restrict__ bits, const double *__restrict__ u_all,
                                                const double *__restrict__ xc_all, const RvlTarget *tg_all, const uint64_t *seeds,
                                                const double *__restrict__ bcvtab, const double *tt, double *__restrict__ scores,
                                                RvlCandHead *part, unsigned long long *work, unsigned long long *ctr,
                                                const int64_t *__restrict__ uniq, int64_t U, RvlWarpScratch &S, uint64_t *yrow, float *lf,
                                                int slot, int nslots, int lane)
{
    const int n = P.n, words = P.words, G = P.G;
    const unsigned long long nrows = (unsigned long long)(uniq ? U : P.F);
    const unsigned long long total = nrows * (unsigned long long)P.T;
    int t = -1, mode = RVL_MODE_ZERO;
    RvlTarget tg;
    const double *xc = nullptr, *u = nullptr;
    const uint64_t *zrow = nullptr;
    double *out = nullptr;
    double best_v = 0.0;
    int64_t best_i = -1;

    // (Claiming the next item one evaluation ahead -- the atomic in flight during the evaluation -- and fetching the next
    // row's index and bit words before the epilogue were both measured SLOWER on B200: 0.973 / 0.993 vs 0.964 ms per launch.)
    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(work, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        const int it = (P.T == 1) ? 0 : (int)(item / nrows);
        const int64_t q = (int64_t)(item - (unsigned long long)it * nrows);
        const int64_t f = uniq ? uniq[q] : q;
        if (it != t) { // new target: flush the running best, switch the per-target pointers
            if (t >= 0 && lane == 0) { RvlCandHead h; h.score = best_v; h.row = best_i; part[(size_t)t * nslots + slot] = h; }
            t = it;
            tg = rvl_ld_target<COH>(tg_all + t);
            mode = tg.mode;
            xc = xc_all + (size_t)t * n;
            u = u_all + (size_t)t * n;
            zrow = seeds + (size_t)tg.seed_of * words;
            out = scores + (size_t)t * P.F;
            best_v = 0.0;
            best_i = -1;
        }
        // bit row -> smem, popcount, masked target sum
        int n1;
        double s1c;
        rvl_row_stats(bits + (size_t)f * words, words, xc, lane, yrow, n1, s1c);
        __syncwarp();
        if (mode == RVL_MODE_ZERO || n1 == 0 || n1 == n) { // LIB:2507: constant x or y -> 0
            if (lane == 0) out[f] = 0.0;
            if (rvl_ahead(0.0, f, best_v, best_i, P.negative)) { best_v = 0.0; best_i = f; }
            continue;
        }
        double rho = rvl_rho(s1c, n1, n, tg.sxx);
        double c, hx;
        if (mode == RVL_MODE_IC) { // LIB:2530-2532: abs(rho); kde2d divides h by 4
            c = 1.0 + P.bw_shrink * fabs(rho);
            hx = tg.bcv * c / 4;
        } else {                   // LIB:2571-2573: positive part of rho
            c = 1.0 + P.bw_shrink * (rho < 0.0 ? 0.0 : rho);
            hx = P.bw_mult * tg.bcv * c;
        }
        double dx = (tg.xmax - tg.xmin) / (double)(G - 1);
        double beta = 0.5 * (dx / hx) * (dx / hx);

        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        int nexp;
        if (P.dense_x) {
            // validation mode: every sample visited for every feature
            if (lane < G)
                rvl_accumulate_dense<COH>(u, yrow, mode == RVL_MODE_CIC ? zrow : nullptr, 0, n, 1, (double)lane, beta, acc);
            nexp = n;
        } else {
            // x-axis sums = (per-class totals over ALL samples at this feature's bandwidth, interpolated
            // from the per-iteration table) - (the minority y-class, visited bit by bit)
            double tot0, tot1;
            rvl_tt_lookup<COH>(P, tt + (size_t)t * P.tt_nodes * 2 * RVL_MAXG, c, lane, tot0, tot1);
            const bool flip = n1 > n - n1; // visit the zeros of y instead when they are fewer
            double m0 = 0.0, m1 = 0.0;     // minority-class sums split by z
            const double gi = (double)lane;
            const int last = (n - 1) >> 6;
            const int mc = flip ? n - n1 : n1; // minority-class size
            const bool use_z = (mode == RVL_MODE_CIC);
            if (mc <= RVL_LIST_MAX) {
                // The minority samples as a compact list (sample index, z bit in bit 31): every lane turns the
                // words it owns into entries at its prefix offset, then all lanes walk the list together --
                // no per-word loop over the mostly empty words of a sparse row.
                int mine = 0;
                for (int w = lane; w <= last; w += 32) {
                    uint64_t v = yrow[w];
                    if (flip) { v = ~v; if (w == last && (n & 63)) v &= (1ull << (n & 63)) - 1ull; }
                    mine += __popcll(v);
                }
                int off = mine; // inclusive scan over lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int up = __shfl_up_sync(0xffffffffu, off, o);
                    if (lane >= o) off += up;
                }
                off -= mine;
                for (int w = lane; w <= last; w += 32) {
                    uint64_t v = yrow[w];
                    if (flip) { v = ~v; if (w == last && (n & 63)) v &= (1ull << (n & 63)) - 1ull; }
                    const uint64_t zw = use_z ? rvl_ld<COH>(zrow + w) : 0ull;
                    while (v) {
                        const int b = __ffsll((long long)v) - 1;
                        v &= v - 1;
                        S.list[off++] = (uint32_t)(w * 64 + b) | ((uint32_t)((zw >> b) & 1ull) << 31);
                    }
                }
                __syncwarp();
                // 32 samples at a time: every lane fetches one entry and its target coordinate (32 independent loads in flight
                // instead of one dependent global load per sample), then the walk takes them from the lanes by shuffle --
                // same samples, same order, two exp chains in flight
                for (int q0 = 0; q0 < mc; q0 += 32) {
                    const int cnt = min(32, mc - q0);
                    const uint32_t el = (lane < cnt) ? S.list[q0 + lane] : 0u;
                    const double ul = (lane < cnt) ? u[el & 0x7fffffffu] : 0.0;
                    int j = 0;
                    for (; j + 1 < cnt; j += 2) {
                        const uint32_t e0 = __shfl_sync(0xffffffffu, el, j), e1 = __shfl_sync(0xffffffffu, el, j + 1);
                        const double d0 = gi - __shfl_sync(0xffffffffu, ul, j), d1 = gi - __shfl_sync(0xffffffffu, ul, j + 1);
                        const double x0 = rvl_exp_neg(-beta * d0 * d0), x1 = rvl_exp_neg(-beta * d1 * d1);
                        if (e0 >> 31) m1 += x0; else m0 += x0;
                        if (e1 >> 31) m1 += x1; else m0 += x1;
                    }
                    if (j < cnt) {
                        const uint32_t e0 = __shfl_sync(0xffffffffu, el, j);
                        const double d0 = gi - __shfl_sync(0xffffffffu, ul, j);
                        const double x0 = rvl_exp_neg(-beta * d0 * d0);
                        if (e0 >> 31) m1 += x0; else m0 += x0;
                    }
                }
                __syncwarp();
            } else {
                for (int w = 0; w <= last; w++) {
                    uint64_t v = yrow[w];
                    if (flip) {
                        v = ~v;
                        if (w == last && (n & 63)) v &= (1ull << (n & 63)) - 1ull;
                    }
                    const uint64_t zw = use_z ? rvl_ld<COH>(zrow + w) : 0ull;
                    while (v) { // warp-uniform loop: every lane sees the same word
                        int b = __ffsll((long long)v) - 1;
                        v &= v - 1;
                        double d = gi - u[w * 64 + b];
                        double e = rvl_exp_neg(-beta * d * d);
                        if ((zw >> b) & 1ull) m1 += e; else m0 += e;
                    }
                }
            }
            double r0 = fmax(tot0 - m0, 0.0), r1 = fmax(tot1 - m1, 0.0); // the majority class
            if (flip) { acc[0] = m0; acc[1] = m1; acc[2] = r0; acc[3] = r1; }
            else      { acc[0] = r0; acc[1] = r1; acc[2] = m0; acc[3] = m1; }
            nexp = flip ? n - n1 : n1;
        }
        if (lane < G) {
            S.A[lane][0] = acc[0]; S.A[lane][1] = acc[1]; S.A[lane][2] = acc[2]; S.A[lane][3] = acc[3];
        }
        __syncwarp();
        int nitems, nyz, nfact, ntwo;
        double score = rvl_finish(S, lf, P, mode, rho, c, hx, bcvtab[n1], tg.bcv_z, n1 > n - n1 ? 1 : 0, lane, &nitems, &nyz, &nfact,
                                  &ntwo);
        if (lane == 0) {
            out[f] = score;
            if (ctr) { // instrumentation (rvl_set_profiling): work actually executed, bench.py's roofline numerator
                atomicAdd(&ctr[0], 1ull);                          // evaluations that ran the KDE path
                atomicAdd(&ctr[1], (unsigned long long)(nitems / G)); // (j,k) density columns evaluated cell by cell with a double-precision log
                atomicAdd(&ctr[2], (unsigned long long)nexp * G);  // x-axis exp() evaluations
                atomicAdd(&ctr[3], (unsigned long long)nfact);     // (j,k) columns in factorised form
                atomicAdd(&ctr[4], (unsigned long long)ntwo);      // one-class columns corrected cell by cell in single precision
                // log() calls and plain FP64 flops (outside exp / log) of this evaluation.  CIC: pass A 3 logs per (i,k),
                // one per cell of a four-term column, 8 per lane in pass B and the y-z marginal + its direct entries, G for
                // the z marginal, 1 in the table lookup; FMAs of pass A (~34 flop per cell), 10 flop per four-term cell.
                // The soft cells (ntwo * G) run on the FP32 pipe: ctr[7].  IC / literal: one log per cell.
                const unsigned long long GG = (unsigned long long)G * G;
                const bool cols = (mode == RVL_MODE_CIC) && !P.literal;
                atomicAdd(&ctr[5], cols ? 3ull * GG + nitems + 9ull * G + nyz + 1ull
                                        : (mode == RVL_MODE_CIC ? (unsigned long long)nitems + 2ull * GG + G : GG + 2ull * G) + 1ull);
                atomicAdd(&ctr[6], (cols ? 34ull * GG + 10ull * nitems + 2ull * (unsigned long long)ntwo + 70ull * G
                                         : 10ull * (unsigned long long)nitems + 6ull * GG) + (unsigned long long)nexp * G * 6ull);
                atomicAdd(&ctr[7], (unsigned long long)ntwo * G);  // soft cells (11 FP32 instructions + 1 MUFU each)
            }
        }
        if (rvl_ahead(score, f, best_v, best_i, P.negative)) { best_v = score; best_i = f; }
        __syncwarp();
    }
    // every lane of a warp holds the same best_v / best_i
    if (t >= 0 && lane == 0) { RvlCandHead h; h.score = best_v; h.row = best_i; part[(size_t)t * nslots + slot] = h; }
}

// This warp's slice of the dynamic shared memory.  Its 32-bit shared-window address is made opaque to the compiler: left
// alone it recomputes tid / 32 * bytes + base inside the hot loops of the epilogue (6-7 instructions per trip) instead of
// keeping it; going through the shared window keeps the accesses LDS / STS rather than generic loads.
__device__ __forceinline__ unsigned char *rvl_warp_smem(unsigned char *smem_raw, int warp, const RvlParams &P)
{
    unsigned int off = (unsigned int)__cvta_generic_to_shared(smem_raw) +
                       (unsigned int)warp * (unsigned int)(sizeof(RvlWarpScratch) + rvl_lf_bytes(P.G) + sizeof(uint64_t) * (size_t)P.words);
    asm volatile("" : "+r"(off));
    return reinterpret_cast<unsigned char *>(__cvta_shared_to_generic((size_t)off));
}

#ifndef RVL_SCORE_MIN_CTAS
#define RVL_SCORE_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(RVL_KSCORE_WARPS * 32, RVL_SCORE_MIN_CTAS)
k_score(RvlParams P, const uint64_t *__restrict__ bits, const double *__restrict__ u_all,
        const double *__restrict__ xc_all, const RvlTarget *__restrict__ tg_all,
        const uint64_t *__restrict__ seeds, const double *__restrict__ bcvtab, const double *__restrict__ tt,
        double *__restrict__ scores, RvlCandHead *__restrict__ part, unsigned long long *__restrict__ work,
        unsigned long long *__restrict__ ctr, const int64_t *__restrict__ uniq, int64_t U)
{
    // uniq[0..U): the rows to score (first row of every group of identical rows); NULL = all F rows
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;

    rvl_pdl_launch_dependents();
    rvl_logtab_fill(rvl_logtab_g); // the only block barrier of the kernel
    rvl_pdl_wait();                // seeds, seed state, table, best slots and the work counter come from k_advance

    // dynamic shared memory, per warp: scratch | float table of the soft cells [2][G][stride] | bit row (rvl_score_plan)
    unsigned char *base = rvl_warp_smem(smem_raw, warp, P);
    RvlWarpScratch &S = *reinterpret_cast<RvlWarpScratch *>(base);
    float *lf = reinterpret_cast<float *>(base + sizeof(RvlWarpScratch));
    uint64_t *yrow = reinterpret_cast<uint64_t *>(base + sizeof(RvlWarpScratch) + rvl_lf_bytes(P.G));
    rvl_score_phase<false>(P, bits, u_all, xc_all, tg_all, seeds, bcvtab, tt, scores, part, work, ctr, uniq, U, S, yrow, lf, slot, nslots,
                           lane);
}

// ---- k_search: the whole search in ONE persistent cooperative launch ---------------------------------------
// grid = SMs x resident CTAs (the k_score grid), launched cooperatively so that every block is resident.  Per iteration:
//   score     every warp draws (target, row) items from this iteration's counter (rvl_score_phase) and leaves its
//             best per target in part[];
//   rank      the block that finishes last (ticket on done[it]) reduces part[] per seed-owning target and stores the
//             candidate record [score][global row][bit row] + a stamp into EVERY rank's exchange buffer -- peer memory
//             over NVLink (CUDA IPC / peer access), the local buffer included; no collective call, no host;
//   merge     every block waits for the `world` stamps of the targets whose tasks it owns, resolves the same global
//             winner (ties -> lowest global row, NaN last), and runs its share of the (table node, target) tasks of
//             rvl_advance_task: seed OR-merge, seed state, in-place x-axis table update;
//   barrier   grid-wide (atomic counter), after which the next scoring pass sees the new seed, state and table.
// Compared with the split-phase launches (k_score, k_argmax, k_advance per iteration) this removes 3 launches and
// their gaps per iteration and the log-table fill of every k_score launch; the arithmetic and its order are the
// same functions, so the results are bit-identical (tests: fused == split-phase).
// Data written and read by different blocks inside the launch (table, seeds, seed state, part[], records) is read with
// ld.global.cg (rvl_ld<true>); writers fence before the ticket / stamp / barrier.
__device__ __forceinline__ void rvl_grid_barrier(unsigned int *bar, unsigned int target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (*reinterpret_cast<volatile unsigned int *>(bar) < target) { }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long rvl_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#ifndef RVL_KSEARCH_COH
#define RVL_KSEARCH_COH true
#endif
__global__ void __launch_bounds__(RVL_KSCORE_WARPS * 32, RVL_SCORE_MIN_CTAS)
k_search(const RvlSearchArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_last;
    const RvlParams &P = a.P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int slot = blockIdx.x * nwarps + warp, nslots = gridDim.x * nwarps;
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[6] = rvl_globaltimer(); // kernel entry (block 0)
    rvl_logtab_fill(rvl_logtab_g);
    const int nodes = P.dense_x ? 1 : P.tt_nodes;
    const int ntasks = nodes * P.T;
    unsigned int nbar = 0;
    for (int it = 0; it < a.iters; it++) {
        const uint64_t *seeds_cur = (it & 1) ? a.seed_b : a.seed_a;
        uint64_t *seeds_nxt = (it & 1) ? a.seed_a : a.seed_b;
        // ---- score
        if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[it * 8 + 0] = rvl_globaltimer();
        for (int t = lane; t < P.T; t += 32) { // this warp's best slots: "none" for the targets it will not touch
            RvlCandHead h;
            h.score = 0.0;
            h.row = -1;
            a.part[(size_t)t * nslots + slot] = h;
        }
        __syncwarp();
        {
            unsigned char *base = rvl_warp_smem(smem_raw, warp, P);
            RvlWarpScratch &S = *reinterpret_cast<RvlWarpScratch *>(base);
            float *lf = reinterpret_cast<float *>(base + sizeof(RvlWarpScratch));
            uint64_t *yrow = reinterpret_cast<uint64_t *>(base + sizeof(RvlWarpScratch) + rvl_lf_bytes(P.G));
            rvl_score_phase<RVL_KSEARCH_COH>(P, a.bits, a.u_all, a.xc_all, a.tg_all, seeds_cur, a.bcvtab, a.tt,
                                             a.scores + (size_t)it * P.T * P.F, a.part, a.work + it, a.ctr, a.uniq, a.U, S, yrow, lf,
                                             slot, nslots, lane);
        }
        __threadfence(); // this warp's part[] entries before the block's ticket
        __syncthreads();
        if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[it * 8 + 1] = rvl_globaltimer(); // block 0 done scoring
        // ---- rank: the last block to finish publishes this rank's candidate records
        RvlPeerXchg X = a.X;
        X.stamp = (a.epoch << 16) | (unsigned int)(it + 1);
        X.parity = (a.seq0 + it) & 1;
        if (threadIdx.x == 0) s_last = (atomicAdd(a.done + it, 1u) == gridDim.x - 1u);
        __syncthreads();
#ifndef RVL_KSEARCH_SCORE_ONLY
        if (s_last) {
            if (a.dbg && threadIdx.x == 0) a.dbg[it * 8 + 2] = rvl_globaltimer(); // every block done scoring
            __threadfence();
            for (int t = 0; t < a.nreal; t++)
                rvl_publish_best<true>(P, a.part + (size_t)t * nslots, nslots, a.bits, nullptr, a.rec_bytes, X, t);
            if (a.dbg && threadIdx.x == 0) a.dbg[it * 8 + 3] = rvl_globaltimer(); // records + stamps out
        }
        // ---- merge: one (node, target) task per WARP -- stamps, winner, seed, in-place table update -- all tasks of the
        // grid in flight at once; a target whose dispatch mode changed (table rebuild, O(n G) per node) is left to a
        // block-level pass of rvl_advance_task.  Nothing of the table is touched after the last iteration.
        const bool last_iter = (it + 1 == a.iters);
        bool rebuild = false;
        rvl_wait_stamps(P, X, a.nreal);
        for (int q = slot; q < ntasks + P.T; q += nslots) { // the last T tasks are the targets' bookkeeping (m = -1)
            const int t = q < ntasks ? q / nodes : q - ntasks, m = q < ntasks ? q - t * nodes : -1;
            if (!rvl_advance_warp<true>(P, a.tg_all, a.u_all, seeds_cur, seeds_nxt, a.bcvtab, a.tt, a.rec_bytes, it, a.iters,
                                        a.first_null, a.top, a.top_score, a.seed_iter, X, t, m, last_iter, lane))
                rebuild = true;
        }
        if (rebuild && lane == 0) atomicOr(a.need + it, 1u);
#else
        const bool last_iter = (it + 1 == a.iters);
#endif
        if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[it * 8 + 4] = rvl_globaltimer(); // block 0's tasks done
        if (!last_iter) {
            nbar++;
            rvl_grid_barrier(a.bar, nbar * gridDim.x);
            if (__ldcg(a.need + it)) { // rare: some target's table must be rebuilt from all samples -- block-level tasks
                for (int q = blockIdx.x; q < ntasks; q += gridDim.x) {
                    const int t = q / nodes, m = q - t * nodes;
                    rvl_advance_task<true>(P, a.tg_all, a.u_all, seeds_cur, seeds_nxt, a.bcvtab, a.tt, nullptr, 0, nullptr, X.local,
                                           a.rec_bytes, it, a.iters, a.first_null, a.top, a.top_score, a.seed_iter, 1, X, t, m,
                                           *reinterpret_cast<RvlAdvanceShared *>(smem_raw), true); // the per-warp scratch is idle
                }
                nbar++;
                rvl_grid_barrier(a.bar, nbar * gridDim.x);
            }
        }
        if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[it * 8 + 5] = rvl_globaltimer(); // barrier passed
    }
}

// assoc(x_t, y, "IC") by one block (8 warps): warps split the samples, lanes own grid points.
// LIB:2491-2501 -> mutual.inf.v2.  Result written by thread 0.
struct RvlAssocShared {
    RvlWarpScratch S;
    double part[RVL_SCORE_WARPS][RVL_MAXG][2];
    double red[32];
    int redi[32];
};

__device__ double rvl_assoc_block(const RvlParams &P, const RvlTarget &tg, const double *__restrict__ u,
                                  const double *__restrict__ xc, const uint64_t *__restrict__ y,
                                  const double *__restrict__ bcvtab, RvlAssocShared &sh)
{
    const int n = P.n, G = P.G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // popcount + masked sum, fixed order: words strided over threads, then block tree
    int n1 = 0;
    double s1c = 0.0;
    for (int w = threadIdx.x; w < P.words; w += blockDim.x) {
        uint64_t v = y[w];
        n1 += __popcll(v);
        while (v) {
            int b = __ffsll((long long)v) - 1;
            v &= v - 1;
            s1c += xc[w * 64 + b];
        }
    }
    n1 = rvl_warp_sum_i(n1);
    s1c = rvl_warp_sum(s1c);
    __syncthreads();
    if (lane == 0) { sh.red[warp] = s1c; sh.redi[warp] = n1; }
    __syncthreads();
    n1 = 0; s1c = 0.0;
    for (int i = 0; i < nwarps; i++) { n1 += sh.redi[i]; s1c += sh.red[i]; }
    if (tg.is_const || n1 == 0 || n1 == n) return 0.0;

    double rho = rvl_rho(s1c, n1, n, tg.sxx);
    double c = 1.0 + P.bw_shrink * fabs(rho);
    double hx = tg.bcv * c / 4;
    double dx = (tg.xmax - tg.xmin) / (double)(G - 1);
    double beta = 0.5 * (dx / hx) * (dx / hx);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    if (lane < G) rvl_accumulate_dense<false>(u, y, nullptr, warp, n, nwarps, (double)lane, beta, acc);
    if (lane < G) { sh.part[warp][lane][0] = acc[0]; sh.part[warp][lane][1] = acc[2]; }
    __syncthreads();
    double score = 0.0;
    if (warp == 0) {
        if (lane < G) {
            double b0 = 0.0, b1 = 0.0;
            for (int w = 0; w < nwarps; w++) { b0 += sh.part[w][lane][0]; b1 += sh.part[w][lane][1]; }
            sh.S.A[lane][0] = b0; sh.S.A[lane][1] = 0.0; sh.S.A[lane][2] = b1; sh.S.A[lane][3] = 0.0;
        }
        __syncwarp();
        int nitems, nyz, nfact, ntwo;
        score = rvl_finish(sh.S, nullptr, P, RVL_MODE_IC, rho, c, hx, bcvtab[n1], 0.0, 0, lane, &nitems, &nyz, &nfact, &ntwo);
    }
    return score; // valid in warp 0
}

// grid = number of (target, vector) pairs; y_all holds one bit row per pair.
__global__ void __launch_bounds__(RVL_SCORE_WARPS * 32)
k_assoc(RvlParams P, const RvlTarget *__restrict__ tg_all, const int *__restrict__ target_of,
        const double *__restrict__ u_all, const double *__restrict__ xc_all, const uint64_t *__restrict__ y_all,
        const double *__restrict__ bcvtab, double *__restrict__ out, size_t y_stride, size_t out_stride, int t_div)
{
    __shared__ RvlAssocShared sh;
    rvl_logtab_fill(rvl_logtab_g);
    int q = blockIdx.x;
    int t = target_of ? target_of[q] : (t_div > 0 ? q / t_div : q); // t_div: t_div consecutive vectors per target
    double v = rvl_assoc_block(P, tg_all[t], u_all + (size_t)t * P.n, xc_all + (size_t)t * P.n,
                               y_all + (size_t)q * y_stride, bcvtab, sh);
    if (threadIdx.x == 0) out[(size_t)q * out_stride] = v;
}

// ---- rank: local best per target ---------------------------------------------------------------

// One block per target: reduce the per-CTA bests k_score left in part[target][nparts] and write the
// candidate record [score][global row][bit row] into the send buffer.

// One block per target (the NCCL / caller-driven exchange; with the stamped peer exchange the same
// reduction runs inside k_advance and this launch disappears).
__global__ void k_argmax(RvlParams P, const RvlCandHead *__restrict__ part, int nparts,
                         const uint64_t *__restrict__ bits, unsigned char *__restrict__ send, size_t rec_bytes,
                         RvlPeerXchg X)
{
    rvl_pdl_launch_dependents();
    rvl_pdl_wait();
    rvl_publish_best<false>(P, part + (size_t)blockIdx.x * nparts, nparts, bits, send, rec_bytes, X, blockIdx.x);
}

// ---- bit packing ------------------------------------------------------------------------------

// Column-major doubles (R matrix) -> bit rows, for columns [s0, s0 + ncols) where s0 % 64 == 0.
// chunk[k + ld * (s - s0)].  One thread per (row, word).  Sets *flag on any entry not 0 or 1.
__global__ void k_pack_f64_colmajor(const double *__restrict__ chunk, int64_t F, int64_t ld, int s0, int ncols,
                                    int words, uint64_t *__restrict__ bits, int *__restrict__ flag)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int w = blockIdx.y; // word within the chunk
    if (k >= F) return;
    uint64_t v = 0;
    int bad = 0;
    int base = w * 64;
    for (int b = 0; b < 64 && base + b < ncols; b++) {
        double e = chunk[k + ld * (int64_t)(base + b)];
        if (e == 1.0) v |= (1ull << b);
        else if (!(e == 0.0)) bad = 1;
    }
    bits[(size_t)k * words + (s0 >> 6) + w] = v;
    if (bad) atomicOr(flag, 1);
}

// Words packed on the host (hostpack.cpp): packed[(w - w0) * F + k] -> bits[k][w] for w in [w0, w1).
__global__ void k_scatter_words(const uint64_t *__restrict__ packed, int64_t F, int w0, int w1, int words,
                                uint64_t *__restrict__ bits)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = w0 + blockIdx.y;
    if (k >= F || w >= w1) return;
    bits[(size_t)k * words + w] = packed[(size_t)(w - w0) * F + k];
}

// Row-major bytes -> bit rows.  One warp per (row, 32 words) tile; lane = word.
__global__ void k_pack_u8_rowmajor(const uint8_t *__restrict__ m, int64_t F, int n, int words,
                                   uint64_t *__restrict__ bits, int *__restrict__ flag)
{
    int64_t k = blockIdx.x;
    int bad = 0;
    for (int w = threadIdx.x; w < words; w += blockDim.x) {
        uint64_t v = 0;
        int base = w * 64;
        for (int b = 0; b < 64 && base + b < n; b++) {
            uint8_t e = m[(size_t)k * n + base + b];
            if (e == 1) v |= (1ull << b);
            else if (e != 0) bad = 1;
        }
        bits[(size_t)k * words + w] = v;
    }
    if (bad) atomicOr(flag, 1);
}

// Padding bits beyond n must be zero in caller-packed rows.
__global__ void k_check_padding(const uint64_t *__restrict__ bits, int64_t F, int n, int words, int *__restrict__ flag)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= F) return;
    int full = n >> 6, rem = n & 63;
    int bad = 0;
    if (rem && (bits[(size_t)k * words + full] >> rem)) bad = 1;
    for (int w = full + (rem ? 1 : 0); w < words; w++)
        if (bits[(size_t)k * words + w]) bad = 1;
    if (bad) atomicOr(flag, 1);
}

__global__ void k_row_counts(const uint64_t *__restrict__ bits, int64_t F, int words, int32_t *__restrict__ out)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= F) return;
    int c = 0;
    for (int w = 0; w < words; w++) c += __popcll(bits[(size_t)k * words + w]);
    out[k] = c;
}

// dst[i] = src[rows[i]]: row selection (count filter, seed / excluded rows removed) without leaving the device.
__global__ void k_gather_rows(const uint64_t *__restrict__ src, const int64_t *__restrict__ rows, int64_t k, int words,
                              uint64_t *__restrict__ dst)
{
    const int64_t tot = k * words;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / words;
        const int w = (int)(i - r * words);
        dst[i] = src[(size_t)rows[r] * words + w];
    }
}

// dst[i][s] = bit s of row rows[i], one byte per sample.
__global__ void k_unpack_rows(const uint64_t *__restrict__ src, const int64_t *__restrict__ rows, int64_t k, int n, int words,
                              uint8_t *__restrict__ dst)
{
    const int64_t tot = k * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n;
        const int s = (int)(i - r * n);
        dst[i] = (uint8_t)((src[(size_t)rows[r] * words + (s >> 6)] >> (s & 63)) & 1ull);
    }
}

// ---- exact-duplicate rows (amplicon blocks of identical rows are normal) --
// Identical rows have identical scores, and R's stable order() ranks the lowest index first, so only the
// first row of every group of identical rows is scored; the others receive a copy of its score.

__global__ void k_row_hash(const uint64_t *__restrict__ bits, int64_t F, int words, uint64_t *__restrict__ hash,
                           int64_t *__restrict__ idx)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= F) return;
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int w = 0; w < words; w++) {
        h ^= bits[(size_t)k * words + w];
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 29;
    }
    hash[k] = h;
    idx[k] = k;
}

// After a stable sort by hash: position p is a group head when its hash differs from p-1's.
__global__ void k_dedup_heads(const uint64_t *__restrict__ hash, int64_t F, int64_t *__restrict__ headpos)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= F) return;
    headpos[p] = (p == 0 || hash[p] != hash[p - 1]) ? p : 0; // max-scanned into "position of my group's head"
}

// rep[row] = first row of the hash group if the bit rows really are equal, else the row itself.
__global__ void k_dedup_verify(const uint64_t *__restrict__ bits, int64_t F, int words, const int64_t *__restrict__ idx,
                               const int64_t *__restrict__ headpos, int64_t *__restrict__ rep)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= F) return;
    int64_t row = idx[p], head = idx[headpos[p]];
    bool same = true;
    if (head != row)
        for (int w = 0; w < words; w++) same = same && (bits[(size_t)row * words + w] == bits[(size_t)head * words + w]);
    rep[row] = same ? head : row;
}

// scores[r][f] = scores[r][rep[f]] for every duplicate row f, over `rows` score vectors of length F.
__global__ void k_fill_duplicates(double *__restrict__ scores, int64_t F, int64_t rows, const int64_t *__restrict__ rep)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    int64_t r = rep[f];
    if (r == f) return;
    for (int64_t v = blockIdx.y; v < rows; v += gridDim.y) scores[(size_t)v * F + f] = scores[(size_t)v * F + r];
}

// Permutation p-values (LIB:1869): pval[i] = #(rand > cmi[i]) / n_rand with `rand` sorted ascending
// and NaNs removed: count = n_valid - upper_bound(rand, cmi[i]).
__global__ void k_pvalues(const double *__restrict__ cmi, int64_t F, const double *__restrict__ rand_sorted,
                          int64_t n_valid, int64_t n_rand, double *__restrict__ pval)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F) return;
    double v = cmi[i];
    if (isnan(v)) { pval[i] = v; return; } // (cmi.rand > NaN) is NA in R and so is its sum: NaN in, NaN out
    int64_t lo = 0, hi = n_valid;
    while (lo < hi) { // first index with rand > v
        int64_t mid = (lo + hi) >> 1;
        if (rand_sorted[mid] > v) hi = mid; else lo = mid + 1;
    }
    pval[i] = (double)(n_valid - lo) / (double)n_rand;
}

// ---- launchers ---------------------------------------------------------------------------------

// Host side of rvl_log_pos's table for the current device: c_i = midpoint of sub-interval i of the folded
// mantissa range, rc_i = fl(1/c_i), lc_i = fl(-log rc_i) from the 80-bit logl.
extern "C" cudaError_t rvl_upload_log_table(void)
{
    double2 h[RVL_LOGTAB_ENTRIES];
    for (int i = 0; i < RVL_LOGTAB_ENTRIES; i++) {
        const uint64_t a_hi = 0x3fe6a09eull + (uint64_t)i * 4096ull, b_hi = a_hi + 4096ull;
        const uint64_t ab = a_hi << 32, bb = b_hi << 32;
        double a, b;
        memcpy(&a, &ab, 8);
        memcpy(&b, &bb, 8);
        const double c = 0.5 * (a + b);
        h[i].x = 1.0 / c;
        h[i].y = (double)(-logl((long double)h[i].x));
    }
    return cudaMemcpyToSymbol(rvl_logtab_g, h, sizeof h);
}

// Block shape of k_score / k_search for bit rows of `words` words and an n_grid of G: the most worker warps (<= 16, one
// block per SM) whose scratch -- RvlWarpScratch + the bit row + the float table of the soft cells, per warp -- fits in
// `budget` bytes of dynamic shared memory, at most max_warps when that is > 0.  Returns the bytes to request and sets
// *warps_out; 0 when not even one warp fits.
extern "C" size_t rvl_score_plan(int words, int G, size_t budget, int max_warps, int *warps_out)
{
    const size_t per_warp = sizeof(RvlWarpScratch) + sizeof(uint64_t) * (size_t)words + rvl_lf_bytes(G);
    int w = (int)(budget / per_warp);
    if (w > RVL_KSCORE_WARPS) w = RVL_KSCORE_WARPS;
    if (max_warps > 0 && w > max_warps) w = max_warps;
    *warps_out = w;
    if (w < 1) return 0;
    size_t smem = per_warp * (size_t)w;
    if (smem < sizeof(RvlAdvanceShared)) smem = sizeof(RvlAdvanceShared); // k_search overlays the merge workspace on the scratch
    return smem;
}

// Sets the dynamic shared memory limit and returns the persistent grid: SMs x resident CTAs per SM.
extern "C" cudaError_t rvl_score_configure(size_t smem, int warps, int device, int *grid_out)
{
    cudaError_t e = cudaFuncSetAttribute(k_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_score, warps * 32, smem);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    *grid_out = sms * per_sm;
    return cudaSuccess;
}

// Persistent grid of k_search for `smem` bytes of dynamic shared memory: SMs x resident blocks (every block must be
// resident: the launch is cooperative); 0 when the device cannot launch cooperatively.
extern "C" cudaError_t rvl_search_configure(size_t smem, int warps, int device, int *grid_out)
{
    *grid_out = 0;
    int coop = 0;
    cudaError_t e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    if (e != cudaSuccess) return e;
    if (!coop) return cudaSuccess;
    e = cudaFuncSetAttribute(k_search, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_search, warps * 32, smem);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    *grid_out = sms * per_sm;
    return cudaSuccess;
}

// Launch of one kernel of the per-iteration chain, optionally as a programmatic dependent of the kernel before it in the stream.
static int g_rvl_pdl = 1; // on by default; RVL_PDL=0 in the environment turns it off (A/B: tools/ab_pdl.sh)
extern "C" void rvl_set_pdl(int on) { g_rvl_pdl = on; }
template <typename... KArgs, typename... Args>
static void rvl_launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_rvl_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

extern "C" cudaError_t rvl_launch_search(const RvlSearchArgs *a, int grid_x, int warps, size_t smem, cudaStream_t st)
{
    void *args[] = {(void *)a};
    return cudaLaunchCooperativeKernel((const void *)k_search, dim3(grid_x), dim3(warps * 32), args, smem, st);
}

extern "C" void rvl_launch_score(const RvlParams *P, int grid_x, int warps, size_t smem, const uint64_t *bits,
                                 const double *u_all, const double *xc_all, const RvlTarget *tg,
                                 const uint64_t *seeds, const double *bcvtab, const double *tt, double *scores,
                                 RvlCandHead *part, unsigned long long *work, unsigned long long *ctr,
                                 const int64_t *uniq, int64_t U, cudaStream_t st)
{
    rvl_launch_chain(k_score, dim3(grid_x), dim3(warps * 32), smem, st, *P, bits, u_all, xc_all, tg, seeds, bcvtab, tt, scores, part,
                     work, ctr, uniq, U);
}

extern "C" void rvl_launch_advance(const RvlParams *P, RvlTarget *tg, const double *u_all, const uint64_t *seeds_in,
                                   uint64_t *seeds_out, const double *bcvtab, double *tt, RvlCandHead *slots,
                                   int nslots, unsigned long long *work, const unsigned char *recv, size_t rec_bytes,
                                   int iter, int iters, int first_null, int64_t *top, double *top_score,
                                   uint64_t *seed_iter, int allow_incr, const RvlPeerXchg *X, cudaStream_t st)
{
    dim3 grid(P->dense_x ? 1 : P->tt_nodes, P->T);
    rvl_launch_chain(k_advance, grid, dim3(RVL_SCORE_WARPS * 32), 0, st, *P, tg, u_all, seeds_in, seeds_out, bcvtab, tt, slots, nslots,
                     work, recv, rec_bytes, iter, iters, first_null, top, top_score, seed_iter, allow_incr, *X);
}

// FP64 FMA peak micro-benchmark (MEASURED_PEAKS.json has no FP64 figure): 8 independent DFMA chains
// per thread.  out[tid] keeps the compiler honest.
__global__ void k_fp64_peak(int iters, double seedv, double *__restrict__ out)
{
    double a0 = seedv, a1 = seedv + 1, a2 = seedv + 2, a3 = seedv + 3, a4 = seedv + 4, a5 = seedv + 5, a6 = seedv + 6,
           a7 = seedv + 7;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" void rvl_launch_fp64_peak(int blocks, int threads, int iters, double *out, cudaStream_t st)
{
    k_fp64_peak<<<blocks, threads, 0, st>>>(iters, 1.0, out);
}

extern "C" void rvl_launch_assoc(const RvlParams *P, int npairs, const RvlTarget *tg, const int *target_of,
                                 const double *u_all, const double *xc_all, const uint64_t *y_all,
                                 const double *bcvtab, double *out, size_t y_stride, size_t out_stride, int t_div,
                                 cudaStream_t st)
{
    k_assoc<<<npairs, RVL_SCORE_WARPS * 32, 0, st>>>(*P, tg, target_of, u_all, xc_all, y_all, bcvtab, out, y_stride,
                                                     out_stride, t_div);
}

extern "C" void rvl_launch_argmax(const RvlParams *P, const RvlCandHead *part, int nparts, const uint64_t *bits,
                                  unsigned char *send, size_t rec_bytes, const RvlPeerXchg *X, cudaStream_t st)
{
    rvl_launch_chain(k_argmax, dim3(P->T), dim3(1024), 0, st, *P, part, nparts, bits, send, rec_bytes, *X);
}

extern "C" void rvl_launch_pack_f64_colmajor(const double *chunk, int64_t F, int64_t ld, int s0, int ncols,
                                             int words, uint64_t *bits, int *flag, cudaStream_t st)
{
    dim3 grid((unsigned)((F + 255) / 256), (unsigned)((ncols + 63) / 64));
    k_pack_f64_colmajor<<<grid, 256, 0, st>>>(chunk, F, ld, s0, ncols, words, bits, flag);
}

extern "C" void rvl_launch_pack_u8_rowmajor(const uint8_t *m, int64_t F0, int64_t Fc, int n, int words,
                                            uint64_t *bits, int *flag, cudaStream_t st)
{
    k_pack_u8_rowmajor<<<(unsigned)Fc, 64, 0, st>>>(m, Fc, n, words, bits + (size_t)F0 * words, flag);
}

extern "C" void rvl_launch_gather_rows(const uint64_t *src, const int64_t *rows, int64_t k, int words, uint64_t *dst,
                                       cudaStream_t st)
{
    int64_t tot = k * words;
    int blocks = (int)((tot + 255) / 256 < 148 * 16 ? (tot + 255) / 256 : 148 * 16);
    if (blocks < 1) blocks = 1;
    k_gather_rows<<<blocks, 256, 0, st>>>(src, rows, k, words, dst);
}

extern "C" void rvl_launch_unpack_rows(const uint64_t *src, const int64_t *rows, int64_t k, int n, int words, uint8_t *dst,
                                       cudaStream_t st)
{
    int64_t tot = k * n;
    int blocks = (int)((tot + 255) / 256 < 148 * 16 ? (tot + 255) / 256 : 148 * 16);
    if (blocks < 1) blocks = 1;
    k_unpack_rows<<<blocks, 256, 0, st>>>(src, rows, k, n, words, dst);
}

extern "C" void rvl_launch_scatter_words(const uint64_t *packed, int64_t F, int w0, int w1, int words, uint64_t *bits,
                                         cudaStream_t st)
{
    if (w1 <= w0 || F <= 0) return;
    dim3 grid((unsigned)((F + 255) / 256), (unsigned)(w1 - w0));
    k_scatter_words<<<grid, 256, 0, st>>>(packed, F, w0, w1, words, bits);
}

extern "C" void rvl_launch_check_padding(const uint64_t *bits, int64_t F, int n, int words, int *flag,
                                         cudaStream_t st)
{
    k_check_padding<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(bits, F, n, words, flag);
}

extern "C" void rvl_launch_row_counts(const uint64_t *bits, int64_t F, int words, int32_t *out, cudaStream_t st)
{
    k_row_counts<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(bits, F, words, out);
}

extern "C" void rvl_launch_rho_max(const RvlParams *P, const uint64_t *bits, const double *xc_all, RvlTarget *tg,
                                   const int64_t *uniq, int64_t U, cudaStream_t st)
{
    int64_t nrows = uniq ? U : P->F;
    if (nrows <= 0) return;
    int64_t blocks = (nrows + 63) / 64; // 8 warps per block, ~8 rows per warp
    if (blocks > 2048) blocks = 2048;
    dim3 grid((unsigned)blocks, P->T);
    k_rho_max<<<grid, 256, 0, st>>>(*P, bits, xc_all, tg, uniq, U);
}

extern "C" void rvl_launch_row_hash(const uint64_t *bits, int64_t F, int words, uint64_t *hash, int64_t *idx,
                                    cudaStream_t st)
{
    k_row_hash<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(bits, F, words, hash, idx);
}

extern "C" void rvl_launch_dedup_heads(const uint64_t *hash, int64_t F, int64_t *headpos, cudaStream_t st)
{
    k_dedup_heads<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(hash, F, headpos);
}

extern "C" void rvl_launch_dedup_verify(const uint64_t *bits, int64_t F, int words, const int64_t *idx,
                                        const int64_t *headpos, int64_t *rep, cudaStream_t st)
{
    k_dedup_verify<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(bits, F, words, idx, headpos, rep);
}

extern "C" void rvl_launch_fill_duplicates(double *scores, int64_t F, int64_t rows, const int64_t *rep,
                                           cudaStream_t st)
{
    dim3 grid((unsigned)((F + 255) / 256), (unsigned)(rows < 1024 ? rows : 1024));
    k_fill_duplicates<<<grid, 256, 0, st>>>(scores, F, rows, rep);
}

extern "C" void rvl_launch_pvalues(const double *cmi, int64_t F, const double *rand_sorted, int64_t n_valid,
                                   int64_t n_rand, double *pval, cudaStream_t st)
{
    k_pvalues<<<(unsigned)((F + 255) / 256), 256, 0, st>>>(cmi, F, rand_sorted, n_valid, n_rand, pval);
}
