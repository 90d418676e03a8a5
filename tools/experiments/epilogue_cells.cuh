// epilogue_cells.cuh -- EXPERIMENT, NOT BUILT.  The "cell regime" entropy epilogue tried in round 2 and dropped.
//
// Idea (validated numerically by tools/proto/epilogue_c.py and on the GPU: all parity tests green, production ==
// literal cube to 2e-14 on CMI): classify every density of the 25^3 cube per (i,k) cell along j into big / mid /
// small relative to the eps' floor; big and small ranges have O(1) closed forms from six prefix / suffix tables of
// the y-axis kernel, only the mid cells (t = D/eps' in (1/2, 16)) need a log.  Logs per evaluation at C3 drop from
// ~7 200 to ~2 970 (FP64-pipe instructions 316 M -> 218 M per launch).
//
// Why it lost (ncu, gpurun_out/prof_r2a, C3 1 GPU): k_score went from 1.30 ms to 1.72 ms per launch.  The kernel is
// issue-bound, not FP64-bound: warp instructions per evaluation rose from 10.8 k to 15.3 k -- the per-cell
// classification (six float sqrt bounds: 2.3 k instructions per evaluation; closed forms and masks: 2.8 k), the
// ballot-compaction of the direct cells (1.9 k) and their decode + V/W rebuild (1.1 k) cost more than the 4 300
// logs they save (about 18 instructions per 32 logs).  A lean version (approximate sqrt, scan-based compaction) was
// estimated at ~9.5 k instructions per evaluation, against ~8.5 k for keeping the column epilogue and replacing its
// serial classification pass by closed-form ranges, which is what score.cu does now.
//
// Kept because the closed forms (rvl_cell_regimes) are reused by score.cu for the y-z marginal, and as a record.

// ---- 3-D epilogue, production form: cell regimes ------------------------------------------------
// Same quantity as rvl_cmi_epilogue, organised around the (i,k) cells of the x-z plane and the majority
// y-class `am` (the class holding most samples; a = 0 for the usual sparse alteration row).  With
//   V = A_{am,0}[i] gz_0[k] + A_{am,1}[i] gz_1[k]   (W likewise for the minority class)
// the G densities above cell (i,k) are  Q_j = g[d_M] V + g[d_m] W + eps',  g[d] = exp(-kappa d^2) the y-axis kernel
// at distance d from a class's own end of the axis, d_M + d_m = G-1.  g falls off so fast that along j a class
// term t = g[d] X / eps' (X = V or W) is, from its own end outwards,
//   big    t >= 16:    Q log Q = D log D + eps' (log D + 1) + O(eps'^2 / D),  log D = log g[d] + log X  -- no log
//   mid    16 > t > 1/2: evaluated directly (one log)
//   small  t <= 1/2:   Q log Q = eps' log eps' + D (log eps' + 1) + D^2/(2 eps') - D^3/(6 eps'^2) + O(eps' t^4/12)
// (the other class's term being invisible to fl(Q) wherever a closed form is used; where both classes are
// visible the cell goes to the direct list).  Summed over the j of a regime these are O(1) per cell from six
// per-evaluation tables over d (prefix sums of g, g log g, log g for `big`, suffix sums of g, g^2, g^3 for `small`),
// so one evaluation needs log X for every cell (pass A, with the x-z marginal's log) and a log only for the few
// `mid` cells: ~300 instead of ~5 600 at n = 1000.  The neglected terms sum to < 1e-17 on CMI (tools/proto/
// epilogue_c.py checks the closed forms against the literal cube in extended precision); regime boundaries come
// from a float sqrt -- a cell that lands one step off its regime is still inside the expansions' range, and direct
// evaluation is exact everywhere, so their rounding cannot matter.
// The y-z marginal  Q_jk = g[d_M] sM_k + g[d_m] sm_k + G eps'  has the same form (cells = k) and reuses the routine.

struct RvlRegime {          // per evaluation (and per floor E: eps' for the cube, G eps' for the y-z marginal)
    double E, lE, lE1;      // floor, log E, log E + 1
    double h2, h3;          // 1/(2E), 1/(6E^2)
    double Lhi;             // log(16 E)
    float inv_kappa;        // 1/kappa = 2 ((G-1) hy)^2
    float x_lo, x_dead;     // log(16 / 0.5) / kappa, log(16 * 2^54) / kappa
};

__device__ __forceinline__ RvlRegime rvl_regime_make(double E, double lE, double inv_kappa)
{
    RvlRegime R;
    R.E = E; R.lE = lE; R.lE1 = lE + 1.0;
    R.h2 = 0.5 / E;
    R.h3 = (1.0 / 6.0) / (E * E);
    R.Lhi = lE + 2.772588722239781;                       // log 16
    R.inv_kappa = (float)inv_kappa;
    R.x_lo = (float)(3.4657359027997265 * inv_kappa);     // log 32
    R.x_dead = (float)(40.20253647247683 * inv_kappa);    // log(16 * 2^54)
    return R;
}

// One cell: class values XM / Xm (>= 0) with logs lM / lm.  Adds the closed-form parts to accX (terms carrying X) and
// accE (coefficient of E), the number of j whose E log E is still owed to cntE, and returns the mask (bit = j' = d_M)
// of the densities that must be evaluated directly.  `valid` = 0 turns the lane into a no-op.
__device__ __forceinline__ unsigned rvl_cell_regimes(const RvlWarpScratch &S, const RvlRegime &R, int G, bool valid, double XM,
                                                     double Xm, double lM, double lm, double &accX, double &accE, int &cntE)
{
    const int G1 = G - 1;
    // last d with t >= 16 (a1), t > 1/2 (a2), t >= 2^-54 (a3); -1: none
    const float xM = (float)(lM - R.Lhi) * R.inv_kappa, xm = (float)(lm - R.Lhi) * R.inv_kappa;
    const float xM2 = xM + R.x_lo, xM3 = xM + R.x_dead, xm2 = xm + R.x_lo, xm3 = xm + R.x_dead;
    const int a1M = xM < 0.f ? -1 : min(G1, (int)__fsqrt_rn(xM)), a2M = xM2 < 0.f ? -1 : min(G1, (int)__fsqrt_rn(xM2));
    const int a3M = xM3 < 0.f ? -1 : min(G1, (int)__fsqrt_rn(xM3));
    const int a1m = xm < 0.f ? -1 : min(G1, (int)__fsqrt_rn(xm)), a2m = xm2 < 0.f ? -1 : min(G1, (int)__fsqrt_rn(xm2));
    const int a3m = xm3 < 0.f ? -1 : min(G1, (int)__fsqrt_rn(xm3));
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

// Four listed cells per lane: rebuild V, W (the expressions of pass A), the density, its log.
template <bool FAST>
__device__ __forceinline__ void rvl_items4(const RvlWarpScratch &S, int G, int M0, int m0, double epsq, int base, int count, int lane,
                                           double &acc0, double &acc1, double &acc2, double &acc3)
{
    double q[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int idx = base + r * 32 + lane;
        const int it = S.items[idx < count ? idx : base]; // idx >= count: a duplicate, weight 0 below
        const int i = it >> 10, k = (it >> 5) & 31, j = it & 31;
        const double2 aM = *reinterpret_cast<const double2 *>(&S.A[i][M0]), am_ = *reinterpret_cast<const double2 *>(&S.A[i][m0]);
        const double z0 = S.gz[0][k], z1 = S.gz[1][k];
        const double V = fma(aM.x, z0, aM.y * z1), W = fma(am_.x, z0, am_.y * z1);
        q[r] = fma(S.gy[0][j], V, fma(S.gy[0][G - 1 - j], W, epsq));
    }
    double l0, l1, l2, l3;
    if (FAST) rvl_log_pos4(q[0], q[1], q[2], q[3], l0, l1, l2, l3);
    else { l0 = log(q[0]); l1 = log(q[1]); l2 = log(q[2]); l3 = log(q[3]); }
    acc0 = fma(base + lane < count ? q[0] : 0.0, l0, acc0);
    acc1 = fma(base + 32 + lane < count ? q[1] : 0.0, l1, acc1);
    acc2 = fma(base + 64 + lane < count ? q[2] : 0.0, l2, acc2);
    acc3 = fma(base + 96 + lane < count ? q[3] : 0.0, l3, acc3);
}

template <bool FAST>
__device__ double rvl_cmi_epilogue_c(RvlWarpScratch &S, int G, double epsq, double hy, int am, int lane, int *nitems_out,
                                     int *nyz_out)
{
    const unsigned FULL = 0xffffffffu;
    // y-class am occupies the j = 0 end of the y axis when am == 0 (gy[0][j] = g[j]), the j = G-1 end otherwise; in
    // terms of d_M (distance from the majority class's end) the density above a cell is g[d_M] V + g[G-1-d_M] W + eps'
    const int M0 = am * 2, m0 = (1 - am) * 2;
    double sa[4];
#pragma unroll
    for (int c = 0; c < 4; c++) sa[c] = rvl_warp_sum((lane < G) ? S.A[lane][c] : 0.0);
    const bool active = lane < G;
    const double g = active ? S.gy[0][lane] : 0.0;
    const double Gy = rvl_warp_sum(g);
    const double Gz = rvl_warp_sum(active ? S.gz[0][lane] : 0.0);
    const int GG = G * G;
    const double gE = (double)G * epsq;
    const double Stot = (sa[0] + sa[1] + sa[2] + sa[3]) * Gy * Gz + (double)GG * (double)G * epsq;

    // ---- y-axis tables over d = lane (the argument of exp in rvl_fill_axis is log g)
    const double inv = 1.0 / ((double)(G - 1) * hy);
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
    }
    __syncwarp();
    const double inv_kappa = 2.0 * ((double)(G - 1) * hy) * ((double)(G - 1) * hy);
    const double lEps = rvl_logq<FAST>(epsq);
    const RvlRegime R = rvl_regime_make(epsq, lEps, inv_kappa);

    // ---- y-z marginal: cells = k (one lane each)
    double accX = 0.0, accE = 0.0, accyz = 0.0;
    int cntE = 0, nyz = 0;
    {
        const int kk = active ? lane : 0;
        const double gz0 = S.gz[0][kk], gz1 = S.gz[1][kk];
        const double sM = am ? fma(sa[2], gz0, sa[3] * gz1) : fma(sa[0], gz0, sa[1] * gz1);
        const double sm = am ? fma(sa[0], gz0, sa[1] * gz1) : fma(sa[2], gz0, sa[3] * gz1);
        const double lgE = rvl_logq<FAST>(gE);
        const RvlRegime Ryz = rvl_regime_make(gE, lgE, inv_kappa);
        const double lM = rvl_logq<FAST>(fmax(sM, 1e-290)), lm = rvl_logq<FAST>(fmax(sm, 1e-290));
        double aX = 0.0, aE = 0.0;
        int cE = 0;
        unsigned mask = rvl_cell_regimes(S, Ryz, G, active, sM, sm, lM, lm, aX, aE, cE);
        nyz = __popc(mask);
        while (mask) { // a handful per lane
            const int j = __ffs((int)mask) - 1;
            mask &= mask - 1u;
            const double q = fma(S.gy[0][j], sM, fma(S.gy[0][G - 1 - j], sm, gE));
            accyz = fma(q, rvl_logq<FAST>(q), accyz);
        }
        accyz += aX + gE * aE + (double)cE * (gE * lgE);
    }
    const double Lyz = rvl_warp_sum(accyz);
    nyz = rvl_warp_sum_i(nyz);

    // ---- pass A: every (i,k) cell once -- V, W, their logs, the x-z marginal, the closed forms, the direct list
    double accxz = 0.0, acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int nlist = 0, nitems = 0;
    const unsigned below = (1u << lane) - 1u;
    const unsigned mdiv = 65536u / (unsigned)G + 1u; // p / G for p < 1024
    for (int p0 = 0; p0 < GG; p0 += 32) {
        const int p = p0 + lane;
        const bool valid = p < GG;
        const int pc = valid ? p : 0;
        const int i = (int)(((unsigned)pc * mdiv) >> 16), k = pc - i * G;
        const double2 aM = *reinterpret_cast<const double2 *>(&S.A[i][M0]), am_ = *reinterpret_cast<const double2 *>(&S.A[i][m0]);
        const double z0 = S.gz[0][k], z1 = S.gz[1][k];
        const double V = fma(aM.x, z0, aM.y * z1), W = fma(am_.x, z0, am_.y * z1);
        const double qxz = fma(Gy, V + W, gE);
        const double lq = rvl_logq<FAST>(qxz), lV = rvl_logq<FAST>(fmax(V, 1e-290)), lW = rvl_logq<FAST>(fmax(W, 1e-290));
        accxz = fma(valid ? qxz : 0.0, lq, accxz);
        unsigned mask = rvl_cell_regimes(S, R, G, valid, V, W, lV, lW, accX, accE, cntE);
        // append this trip's direct cells to the list, one ballot round per item of the busiest lane
        const int code = (i << 10) | (k << 5);
        while (true) {
            const unsigned has = __ballot_sync(FULL, mask != 0u);
            if (!has) break;
            if (mask) {
                const int j = __ffs((int)mask) - 1;
                mask &= mask - 1u;
                S.items[nlist + __popc(has & below)] = (uint16_t)(code | j);
            }
            nlist += __popc(has);
        }
        __syncwarp();
        while (nlist >= 128) { // drain from the end of the list in full batches
            nlist -= 128;
            rvl_items4<FAST>(S, G, M0, m0, epsq, nlist, nlist + 128, lane, acc0, acc1, acc2, acc3);
            nitems += 128;
        }
        __syncwarp();
    }
    nitems += nlist;
    for (int base = 0; base < nlist; base += 128) rvl_items4<FAST>(S, G, M0, m0, epsq, base, nlist, lane, acc0, acc1, acc2, acc3);
    *nitems_out = nitems;
    *nyz_out = nyz;

    const double L3 = rvl_warp_sum((acc0 + acc1) + (acc2 + acc3) + accX + epsq * accE) +
                      (double)rvl_warp_sum_i(cntE) * (epsq * lEps);
    const double Lxz = rvl_warp_sum(accxz);
    double accz = 0.0;
    if (active) {
        const double qz = fma((sa[0] + sa[2]) * Gy, S.gz[0][lane], fma((sa[1] + sa[3]) * Gy, S.gz[1][lane], (double)GG * epsq));
        accz = rvl_xlogx<FAST>(qz);
    }
    const double Lz = rvl_warp_sum(accz);
    return (L3 - Lxz - Lyz + Lz) / Stot;
}

