// rvl_common.cuh -- shared device-side structures of the REVEALER CMI-search core.
//
// Vocabulary follows the reference (REVEALER_library.v5C.R, "LIB", copy-B line numbers):
//   target  x  : double[n]            (LIB:1626 `target`)
//   feature y  : one bit-row of m.2   (LIB:1851 `m.2[k,]`)
//   seed    z  : bit vector           (LIB:1691-1700 `seed`)
//   n_grid  G  : KDE grid points per axis (25)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define RVL_MAXG 32

// Per-target constants (computed once by rvl_set_targets) and per-iteration seed state.
struct RvlTarget {
    // --- constant per target
    double mean;      // mean(x)
    double sxx;       // sum((x-mean)^2)
    double xmin, xmax;
    double bcv;       // MASS::bcv(x)                        LIB:2524 / 2564 default args
    int32_t is_const; // length(unique(x)) == 1              LIB:2495 / 2507
    int32_t seed_of;  // index of the target whose seed this target is scored against
                      // (itself, or 0 for permutation-null targets LIB:1853-1855)
    // largest rho^+ and |rho| = cor(x, row) over the loaded rows (k_rho_max): bounds the range of the
    // bandwidth shrink factor c, hence how many nodes of the x-axis table are ever read
    unsigned long long rho_pos_max_bits, rho_abs_max_bits;
    unsigned long long fingerprint; // order-independent hash of the multiset of target values
    // --- per iteration (written by k_advance / k_search's merge tasks)
    int32_t nz;       // popcount(seed)
    int32_t mode;     // RVL_MODE_*
    double bcv_z;     // bcv(seed) from the binary table
};

enum { RVL_MODE_ZERO = 0,  // constant target: every score is 0          LIB:2507
       RVL_MODE_IC = 1,    // constant seed: mutual.inf.v2(x,y)$IC       LIB:2509-2511
       RVL_MODE_CIC = 2 }; // cond.mutual.inf(x,y,z)$CIC                 LIB:2517

// One candidate record per (rank, target) exchanged each iteration.
// Layout in bytes: [score f64][global row i64][bits: words * u64]
struct RvlCandHead {
    double score;
    int64_t row;
};

// Peer-memory exchange of the candidate records (instead of an NCCL all-gather): every rank owns one
// buffer  [2 parities][world][T] records  +  [2][world][T] uint32 stamps, mapped into every other
// rank with CUDA IPC.  k_argmax stores its record into slot [parity][rank][t] of EVERY rank's buffer over
// NVLink, fences, then stores the stamp; k_advance polls the stamps in its own (local) buffer.
// peers == NULL disables the path.
struct RvlPeerXchg {
    unsigned char *const *peers; // device array [world]: base of every rank's buffer (own entry = local)
    unsigned char *local;        // this rank's buffer
    size_t stamps_off;           // byte offset of the stamp array inside a buffer
    unsigned int stamp;          // (search epoch << 16) | (iteration + 1)
    int parity;                  // iteration & 1
    int *err;                    // set to 1 when a stamp did not arrive in time
};

struct RvlParams {
    int32_t n;        // samples
    int32_t words;    // uint64 per bit row (row stride)
    int32_t G;        // n_grid
    int32_t negative; // target.match == "negative"
    double bw_mult;   // 0.25
    double bw_shrink; // -0.75
    int64_t F;        // local rows
    int64_t row_offset;
    int32_t T;        // targets
    int32_t world, rank;
    // bandwidth-interpolation table (see score.cu, "x-axis sums"): nodes tau_m = (m - tt_pad) * tt_dt
    int32_t tt_nodes, tt_pad;
    double tt_dt;
    int32_t dense_x;  // 1: x-axis sums recomputed per feature over all samples (validation mode, or
                      //    chosen automatically when so few rows are scored that a table does not pay)
    int32_t literal;  // 1: literal entropy epilogue, no factorised columns (validation mode)
    double prune_theta; // density columns holding < theta/G^3 of the mass are replaced by the eps floor
};

// Arguments of k_search (score.cu): the whole search in one persistent cooperative launch.
struct RvlSearchArgs {
    RvlParams P;
    const uint64_t *bits;
    const double *u_all, *xc_all;
    RvlTarget *tg_all;
    uint64_t *seed_a, *seed_b;     // iteration `it` reads (it & 1 ? b : a) and writes the other
    const double *bcvtab;
    double *tt;
    double *scores;                // [iters][T][F]
    RvlCandHead *part;
    unsigned long long *work;      // [iters] item counters, zeroed before the launch
    unsigned int *done;            // [iters] block tickets, zeroed before the launch
    unsigned int *bar;             // grid barrier counter, zeroed before the launch
    unsigned int *need;            // [iters] set when an iteration's merge needs a table rebuild, zeroed before the launch
    unsigned long long *ctr;
    const int64_t *uniq;
    int64_t U;
    size_t rec_bytes;
    int iters, first_null, nreal;  // nreal: targets that own a seed (the others are permutation nulls)
    int64_t *top;
    double *top_score;
    uint64_t *seed_iter;
    RvlPeerXchg X;                 // peers / local / stamps_off / err; stamp and parity are set per iteration here
    unsigned int epoch;
    int seq0;                      // exchanges issued before this search (parity continues across searches)
    unsigned long long *dbg;       // NULL, or [iters][8] globaltimer stamps of the phases (tools/time_phases.py)
};

#define RVL_TT_STENCIL 8   // Lagrange points
#define RVL_TT_PAD 4       // nodes beyond each end of [0, tau_max]
#define RVL_TT_INV_DT 128  // nodes per unit of tau = -2 ln c

#define RVL_DBL_EPS 2.220446049250313e-16

// ---- log for the entropy epilogue ----------------------------------------------------------------
// log(q) for a positive, finite, NORMAL double (callers guarantee the range once per evaluation).
// Table-driven and branch-free: the mantissa is folded to m in [sqrt(1/2), sqrt(2)); the top 8 bits of
// the folded mantissa field select c_i (midpoint of the 256 sub-intervals) with rc_i = fl(1/c_i) and
// lc_i = -log(rc_i) (rounded from an 80-bit evaluation on the host, rvl_upload_log_table); then
//   r = fma(m, rc_i, -1) = m rc_i - 1, |r| <= 2^-9,   log q = k ln2 + lc_i + (r - r^2/2 + r^3/3 - r^4/4 + r^5/5)
// (truncation r^6/6 <= 9e-18).  Error: <= 2 ulp for |log q| > 0.5, <= 8e-17 absolute below (emulated in
// tests/test_fastlog.py).  9 FP64-pipe instructions (was 18 with the division-based form), dependency
// depth 5.  The table sits in shared memory, every entry replicated 8 times ([entry][lane & 7], 16 B
// each = 32 KB per block): the 8 lanes of each quarter-warp of an LDS.128 then hit 8 distinct 16-byte
// bank groups whatever their entries, so the gather is conflict-free.
#define RVL_LOGTAB_ENTRIES 256
#ifndef RVL_LOGTAB_COPIES
#define RVL_LOGTAB_COPIES 8
#endif
// byte offset of table entry (hs >> 12) & 255: entries are RVL_LOGTAB_COPIES * 16 bytes apart
#define RVL_LOGTAB_OFF(hs) (((hs) >> (RVL_LOGTAB_COPIES == 8 ? 5 : RVL_LOGTAB_COPIES == 4 ? 6 : 7)) & (255 * RVL_LOGTAB_COPIES * 16))
static __constant__ double RVL_LG[5] = {1.0 / 3, -0.5, 1.0 / 5, -0.25, 0.6931471805599453094};

// The block's copy of the table (one instance per kernel that calls it).
__device__ __forceinline__ double2 *rvl_logtab_smem()
{
    __shared__ __align__(16) double2 tab[RVL_LOGTAB_ENTRIES * RVL_LOGTAB_COPIES];
    return tab;
}

// Whole block: replicate the 256-entry global table into shared memory.  Ends with a barrier.
__device__ __forceinline__ void rvl_logtab_fill(const double2 *__restrict__ g)
{
    double2 *tab = rvl_logtab_smem();
    for (int e = threadIdx.x; e < RVL_LOGTAB_ENTRIES; e += blockDim.x) { // one global load per entry, then its copies
        const double2 v = g[e];
#pragma unroll
        for (int c = 0; c < RVL_LOGTAB_COPIES; c++) tab[e * RVL_LOGTAB_COPIES + c] = v;
    }
    __syncthreads();
}

// The mantissa m = q 2^-k is never formed: the scaling is moved onto the table value, f = rc_i 2^-k (an integer
// subtraction on rc_i's high word), and r = fma(q, f, -1) -- the same product m rc_i bit for bit (powers of two
// scale exactly; q in [2^-1000, 2^1000] keeps f normal), one instruction and one FP64 <- integer operand word
// fewer per log than assembling m from q's two words.
__device__ __forceinline__ double rvl_log_pos(double q)
{
    const int hs = __double2hiint(q) + (0x3ff00000 - 0x3fe6a09e);
    const int k = (hs >> 20) - 0x3ff;
    const char *base = reinterpret_cast<const char *>(rvl_logtab_smem()) + (threadIdx.x & (RVL_LOGTAB_COPIES - 1)) * 16;
    const double2 e = *reinterpret_cast<const double2 *>(base + RVL_LOGTAB_OFF(hs)); // entry (hs >> 12) & 255, 128 B apart
    const double f = __hiloint2double(__double2hiint(e.x) - (k << 20), __double2loint(e.x));
    const double r = fma(q, f, -1.0);
    const double r2 = r * r;
    const double pa = fma(r, RVL_LG[0], RVL_LG[1]);
    const double pb = fma(r, RVL_LG[2], RVL_LG[3]);
    const double p = fma(pb, r2, pa);
    const double lm = fma(r2, p, r);
    return fma((double)k, RVL_LG[4], e.y) + lm;
}

// Four logs at once, written stage by stage (all index computations, all table loads, all r, ...): the
// four dependency chains are independent, and laying them out side by side is what lets the scheduler
// issue one chain's FP64 instruction while the others wait out their ~9-cycle latency.  Same arithmetic
// as rvl_log_pos, bit for bit.
__device__ __forceinline__ void rvl_log_pos4(const double q0, const double q1, const double q2, const double q3,
                                             double &l0, double &l1, double &l2, double &l3)
{
    const char *base = reinterpret_cast<const char *>(rvl_logtab_smem()) + (threadIdx.x & (RVL_LOGTAB_COPIES - 1)) * 16;
    const int s0 = __double2hiint(q0) + (0x3ff00000 - 0x3fe6a09e), s1 = __double2hiint(q1) + (0x3ff00000 - 0x3fe6a09e),
              s2 = __double2hiint(q2) + (0x3ff00000 - 0x3fe6a09e), s3 = __double2hiint(q3) + (0x3ff00000 - 0x3fe6a09e);
    const double2 e0 = *reinterpret_cast<const double2 *>(base + RVL_LOGTAB_OFF(s0));
    const double2 e1 = *reinterpret_cast<const double2 *>(base + RVL_LOGTAB_OFF(s1));
    const double2 e2 = *reinterpret_cast<const double2 *>(base + RVL_LOGTAB_OFF(s2));
    const double2 e3 = *reinterpret_cast<const double2 *>(base + RVL_LOGTAB_OFF(s3));
    const int k0 = (s0 >> 20) - 0x3ff, k1 = (s1 >> 20) - 0x3ff, k2 = (s2 >> 20) - 0x3ff, k3 = (s3 >> 20) - 0x3ff;
    const double f0 = __hiloint2double(__double2hiint(e0.x) - (k0 << 20), __double2loint(e0.x));
    const double f1 = __hiloint2double(__double2hiint(e1.x) - (k1 << 20), __double2loint(e1.x));
    const double f2 = __hiloint2double(__double2hiint(e2.x) - (k2 << 20), __double2loint(e2.x));
    const double f3 = __hiloint2double(__double2hiint(e3.x) - (k3 << 20), __double2loint(e3.x));
    const double r0 = fma(q0, f0, -1.0), r1 = fma(q1, f1, -1.0), r2 = fma(q2, f2, -1.0), r3 = fma(q3, f3, -1.0);
    const double b0 = fma((double)k0, RVL_LG[4], e0.y), b1 = fma((double)k1, RVL_LG[4], e1.y),
                 b2 = fma((double)k2, RVL_LG[4], e2.y), b3 = fma((double)k3, RVL_LG[4], e3.y);
    const double z0 = r0 * r0, z1 = r1 * r1, z2 = r2 * r2, z3 = r3 * r3;
    const double a0 = fma(r0, RVL_LG[0], RVL_LG[1]), a1 = fma(r1, RVL_LG[0], RVL_LG[1]),
                 a2 = fma(r2, RVL_LG[0], RVL_LG[1]), a3 = fma(r3, RVL_LG[0], RVL_LG[1]);
    const double c0 = fma(r0, RVL_LG[2], RVL_LG[3]), c1 = fma(r1, RVL_LG[2], RVL_LG[3]),
                 c2 = fma(r2, RVL_LG[2], RVL_LG[3]), c3 = fma(r3, RVL_LG[2], RVL_LG[3]);
    const double p0 = fma(c0, z0, a0), p1 = fma(c1, z1, a1), p2 = fma(c2, z2, a2), p3 = fma(c3, z3, a3);
    const double t0 = fma(z0, p0, r0), t1 = fma(z1, p1, r1), t2 = fma(z2, p2, r2), t3 = fma(z3, p3, r3);
    l0 = b0 + t0;
    l1 = b1 + t1;
    l2 = b2 + t2;
    l3 = b3 + t3;
}

// exp(a) for a <= 0 (Gaussian kernel weights).  Branch-free: k = round(a log2 e) by the 1.5*2^52 trick,
// r = a - k ln2 (hi/lo split, FMA), degree-12 Taylor polynomial on |r| <= 0.347, 2^k by an integer add
// to the exponent field.  Results below ~2e-307 are flushed to 0 (they are > 280 orders of magnitude
// under anything that can matter next to the eps floor).  <= 3 ulp (emulated in tests/test_fastlog.py);
// 17 FP64-pipe instructions with the coefficients taken from constant memory.
static __constant__ double RVL_EXPC[11] = {1.0 / 2,         1.0 / 6,          1.0 / 24,          1.0 / 120,
                                           1.0 / 720,       1.0 / 5040,       1.0 / 40320,       1.0 / 362880,
                                           1.0 / 3628800,   1.0 / 39916800,   1.0 / 479001600};
__device__ __forceinline__ double rvl_exp_neg(double a)
{
    const double t = fma(a, 1.4426950408889634074, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, a);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double q = RVL_EXPC[10];
    q = fma(q, r, RVL_EXPC[9]);
    q = fma(q, r, RVL_EXPC[8]);
    q = fma(q, r, RVL_EXPC[7]);
    q = fma(q, r, RVL_EXPC[6]);
    q = fma(q, r, RVL_EXPC[5]);
    q = fma(q, r, RVL_EXPC[4]);
    q = fma(q, r, RVL_EXPC[3]);
    q = fma(q, r, RVL_EXPC[2]);
    q = fma(q, r, RVL_EXPC[1]);
    q = fma(q, r, RVL_EXPC[0]);
    const double p = fma(r * r, q, r) + 1.0;
    const double s = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return a < -707.0 ? 0.0 : s;
}

__device__ __forceinline__ double rvl_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double rvl_warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int rvl_warp_sum_i(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
