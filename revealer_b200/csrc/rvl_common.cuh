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
    // --- per iteration (written by k_seed_prep)
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

#define RVL_TT_STENCIL 8   // Lagrange points
#define RVL_TT_PAD 4       // nodes beyond each end of [0, tau_max]
#define RVL_TT_INV_DT 128  // nodes per unit of tau = -2 ln c

#define RVL_DBL_EPS 2.220446049250313e-16

// Coefficients of rvl_log_pos in constant memory so that DFMA takes them as c[bank][offset] operands
// (no per-use UMOV pair): R(z) = Lg1 z + ... + Lg7 z^7 ~ (log(1+s) - log(1-s) - 2s)/s on
// z = s^2 <= 0.0295 (the minimax fit published with fdlibm's e_log.c, |error| < 2^-58.45), then ln 2.
static __constant__ double RVL_LG[8] = {6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
                                 2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
                                 1.479819860511658591e-01, 0.6931471805599453094};

// log(q) for a positive, finite, NORMAL double (callers guarantee the range once per evaluation).
// Branch-free: mantissa folded to [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1) by a MUFU reciprocal seed
// + two Newton steps, log m = 2s + s R(s^2), k*ln2 added last.
// <= 4 ulp against numpy's log over [1e-87, 1e17] (emulated in tests/test_fastlog.py);
// 18 FP64-pipe instructions, no special-case branches.
__device__ __forceinline__ double rvl_log_pos(double q)
{
    int hi = __double2hiint(q);
    const int lo = __double2loint(q);
    hi += 0x3ff00000 - 0x3fe6a09e;
    const int k = (hi >> 20) - 0x3ff;
    hi = (hi & 0x000fffff) + 0x3fe6a09e;
    const double m = __hiloint2double(hi, lo);
    const double f = m - 1.0, t = m + 1.0;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(t));
    double e = fma(-t, r, 1.0);
    r = fma(r, e, r);
    e = fma(-t, r, 1.0);
    r = fma(r, e, r);
    const double s = f * r;
    const double z = s * s;
    double p = RVL_LG[6];
    p = fma(p, z, RVL_LG[5]);
    p = fma(p, z, RVL_LG[4]);
    p = fma(p, z, RVL_LG[3]);
    p = fma(p, z, RVL_LG[2]);
    p = fma(p, z, RVL_LG[1]);
    p = fma(p, z, RVL_LG[0]);
    const double lm = fma(s * z, p, s + s);
    return fma((double)k, RVL_LG[7], lm);
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
