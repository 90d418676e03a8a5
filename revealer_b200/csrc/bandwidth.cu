// bandwidth.cu -- device-side MASS::bcv (biased cross-validation bandwidth) for the targets and
// for every possible binary vector of length n.  Compiled with -fmad=false so that the Brent
// iteration sees the same IEEE operation sequence as the published C routines it restates
// (VR_den_bin / VR_bcv_bin of MASS 7.3-29 and R's Brent_fmin behind stats::optimize); the
// bandwidth returned by optimize(tol = 0.01*hmax) depends on the exact iteration path.
//
// Reference call sites: delta = c(bcv(x), bcv(y)) LIB:2524; 0.25*c(bcv(x), bcv(y), bcv(z)) LIB:2564
// -- evaluated on EVERY cond.assoc call in R; here once per target, once per (n, n1).
#include "rvl_common.cuh"
#include <cfloat>
#include <cmath>

#define RVL_NB 1000
#define RVL_DELMAX 1000.0
#define RVL_PI 3.141592653589793238462643383280

// R's Brent_fmin (Forsythe-Malcolm-Moler fmin), decision rules as in SURVEY Appendix B.
template <class Fn>
__device__ double rvl_brent_fmin(double ax, double bx, const Fn &f, double tol)
{
    const double c = (3. - sqrt(5.)) * .5;
    double a, b, d, e, p, q, r, u, v, w, x;
    double t2, fu, fv, fw, fx, xm, eps, tol1, tol3;
    eps = sqrt(DBL_EPSILON);
    a = ax; b = bx;
    v = a + c * (b - a);
    w = v; x = v;
    d = 0.; e = 0.;
    fx = f(x);
    fv = fx; fw = fx;
    tol3 = tol / 3.;
    for (int guard = 0; guard < 1000; guard++) {
        xm = (a + b) * .5;
        tol1 = eps * fabs(x) + tol3;
        t2 = tol1 * 2.;
        if (fabs(x - xm) <= t2 - (b - a) * .5) break;
        p = 0.; q = 0.; r = 0.;
        if (fabs(e) > tol1) {
            r = (x - w) * (fx - fv);
            q = (x - v) * (fx - fw);
            p = (x - v) * q - (x - w) * r;
            q = (q - r) * 2.;
            if (q > 0.) p = -p; else q = -q;
            r = e;
            e = d;
        }
        if (fabs(p) >= fabs(q * .5 * r) || p <= q * (a - x) || p >= q * (b - x)) {
            if (x < xm) e = b - x; else e = a - x;
            d = c * e;
        } else {
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
        fu = f(u);
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

// VR_bcv_bin for a histogram with only bins 0 and `hi` populated (binary vectors: hi = 990).
// Skipped bins add term*0 = +0.0 in the original loop, which leaves the running sum unchanged.
struct RvlBinaryBcvObjective {
    double n, d, cnt0, cnthi;
    int hi;
    __device__ double operator()(double h) const
    {
        double hh = h / 4, sum = 0.0;
        sum += 12.0 * cnt0; // i = 0: delta = 0, term = exp(-0)*(0 - 0 + 12)
        double delta = hi * d / hh;
        delta *= delta;
        if (delta < RVL_DELMAX) {
            double term = exp(-delta / 4) * (delta * delta - 12 * delta + 12);
            sum += term * cnthi;
        }
        return 1 / (2 * n * hh * sqrt(RVL_PI)) + sum / (64 * n * n * hh * sqrt(RVL_PI));
    }
};

// table[n1] = bcv(y) for a 0/1 vector y of length n with n1 ones; table[0] = table[n] = 0.
__global__ void k_bcv_binary_table(int n, double *__restrict__ table)
{
    int n1 = blockIdx.x * blockDim.x + threadIdx.x;
    if (n1 > n) return;
    if (n1 == 0 || n1 == n) { table[n1] = 0.0; return; }
    double dn = (double)n, d1 = (double)n1, d0 = (double)(n - n1);
    double var = d1 * d0 / (dn * (dn - 1.0));
    double hmax = 1.144 * sqrt(var) * pow(dn, -1.0 / 5) * 4;
    double lower = 0.1 * hmax, upper = hmax;
    RvlBinaryBcvObjective f;
    f.n = dn;
    double rang = (1.0 - 0.0) * 1.01;
    f.d = rang / RVL_NB;
    f.hi = (int)(1.0 / f.d);
    f.cnt0 = d1 * (d1 - 1.0) * 0.5 + d0 * (d0 - 1.0) * 0.5;
    f.cnthi = d1 * d0;
    table[n1] = rvl_brent_fmin(lower, upper, f, 0.1 * lower);
}

// ---- per-target statistics ------------------------------------------------------------

__device__ double rvl_block_sum(double v, double *sh)
{
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = rvl_warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += sh[i];
    return t;
}

// One block per target: range, mean (R's two-pass MEAN), sum of squares; writes the grid-unit
// coordinates u = (x - xmin)/dx (dx = range/(G-1)), the centred values xc = x - mean and the
// VR_den_bin bin index (int)(x/dd).
__global__ void k_target_stats(const double *__restrict__ x_all, int n, int G, RvlTarget *__restrict__ tg,
                               double *__restrict__ u_all, double *__restrict__ xc_all,
                               int *__restrict__ bin_all, int *__restrict__ nonfinite)
{
    __shared__ double sh[32];
    __shared__ double shmin[32], shmax[32];
    int t = blockIdx.x;
    const double *x = x_all + (size_t)t * n;
    double lo = INFINITY, hi = -INFINITY, s = 0.0;
    int bad = 0;
    unsigned long long fp = 0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v = x[i];
        if (!isfinite(v)) bad = 1;
        lo = fmin(lo, v); hi = fmax(hi, v); s += v;
        unsigned long long h = (unsigned long long)__double_as_longlong(v == 0.0 ? 0.0 : v); // -0 == +0
        h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
        fp += h; // a sum of hashes does not depend on the order of the samples
    }
    atomicAdd(&tg[t].fingerprint, fp);
    if (bad) atomicOr(nonfinite, 1);
    double tot = rvl_block_sum(s, sh);
    double mean = tot / n;
    s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += x[i] - mean;
    tot = rvl_block_sum(s, sh);
    mean = mean + tot / n;
    s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { double d = x[i] - mean; s += d * d; }
    double sxx = rvl_block_sum(s, sh);
    // min / max
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __syncthreads();
    if (lane == 0) { shmin[w] = lo; shmax[w] = hi; }
    __syncthreads();
    lo = shmin[0]; hi = shmax[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); i++) { lo = fmin(lo, shmin[i]); hi = fmax(hi, shmax[i]); }

    double dx = (hi - lo) / (G - 1);
    double dd = (hi - lo) * 1.01 / RVL_NB;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v = x[i];
        u_all[(size_t)t * n + i] = (hi > lo) ? (v - lo) / dx : 0.0;
        xc_all[(size_t)t * n + i] = v - mean;
        bin_all[(size_t)t * n + i] = (hi > lo) ? (int)(v / dd) : 0;
    }
    if (threadIdx.x == 0) {
        tg[t].mean = mean;
        tg[t].sxx = sxx;
        tg[t].xmin = lo;
        tg[t].xmax = hi;
        tg[t].is_const = (hi == lo) ? 1 : 0;
        tg[t].seed_of = t; // every target is scored against its own seed unless rvl_set_null_targets says otherwise
    }
}

// VR_den_bin pair histogram: cnt[|bin_i - bin_j|]++ over all pairs j < i.  Integer counts, so
// the result is exact and order-independent.  grid = (row chunks, targets).
__global__ void k_pair_histogram(const int *__restrict__ bin_all, int n, int *__restrict__ cnt_all)
{
    __shared__ int hist[RVL_NB];
    int t = blockIdx.y;
    const int *bin = bin_all + (size_t)t * n;
    for (int i = threadIdx.x; i < RVL_NB; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = blockIdx.x + 1; i < n; i += gridDim.x) {
        int bi = bin[i];
        for (int j = threadIdx.x; j < i; j += blockDim.x) {
            int dlt = bi - bin[j];
            dlt = dlt < 0 ? -dlt : dlt;
            if (dlt < RVL_NB) atomicAdd(&hist[dlt], 1);
        }
    }
    __syncthreads();
    int *cnt = cnt_all + (size_t)t * RVL_NB;
    for (int i = threadIdx.x; i < RVL_NB; i += blockDim.x)
        if (hist[i]) atomicAdd(&cnt[i], hist[i]);
}

// VR_bcv_bin over the 1000-bin histogram, evaluated by one block: the threads compute the products
// term_i * cnt_i side by side, then one thread adds them in bin order -- the same operands in the
// same order as the original's sequential loop, so the value (and with it the Brent path) is the
// one a single thread would get.  delta grows with i, so "break at the first delta >= DELMAX" is
// "sum the bins below the first such i".
struct RvlHistBcvObjective {
    double n, d;
    const int *cnt;
    double *prod; // shared [RVL_NB]
    int *stop;    // shared: first bin with delta >= DELMAX
    double *res;  // shared: the sum
    __device__ double operator()(double h) const
    {
        double hh = h / 4;
        if (threadIdx.x == 0) *stop = RVL_NB;
        __syncthreads();
        for (int i = threadIdx.x; i < RVL_NB; i += blockDim.x) {
            double delta = i * d / hh;
            delta *= delta;
            if (delta >= RVL_DELMAX) { atomicMin(stop, i); continue; }
            int c = cnt[i];
            // an empty bin adds term*0 = +0.0, which leaves the running sum as it is: skip its exp
            prod[i] = c ? exp(-delta / 4) * (delta * delta - 12 * delta + 12) * c : 0.0;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double sum = 0.0;
            int m = *stop;
            for (int i = 0; i < m; i++) sum += prod[i];
            *res = sum;
        }
        __syncthreads();
        double sum = *res;
        return 1 / (2 * n * hh * sqrt(RVL_PI)) + sum / (64 * n * n * hh * sqrt(RVL_PI));
    }
};

// One block per target runs the Brent search on its 1000-bin histogram; every thread follows the
// same (uniform) control flow, the objective is the block-wide evaluation above.
__global__ void __launch_bounds__(256) k_bcv_target(int n, int T, const int *__restrict__ cnt_all, RvlTarget *__restrict__ tg)
{
    __shared__ double prod[RVL_NB];
    __shared__ int cnt[RVL_NB];
    __shared__ int stop;
    __shared__ double res;
    int t = blockIdx.x;
    if (t >= T) return;
    if (tg[t].is_const) { if (threadIdx.x == 0) tg[t].bcv = 0.0; return; }
    for (int i = threadIdx.x; i < RVL_NB; i += blockDim.x) cnt[i] = cnt_all[(size_t)t * RVL_NB + i];
    __syncthreads();
    double dn = (double)n;
    double var = tg[t].sxx / (dn - 1.0);
    double hmax = 1.144 * sqrt(var) * pow(dn, -1.0 / 5) * 4;
    double lower = 0.1 * hmax, upper = hmax;
    RvlHistBcvObjective f;
    f.n = dn;
    f.d = (tg[t].xmax - tg[t].xmin) * 1.01 / RVL_NB;
    f.cnt = cnt;
    f.prod = prod;
    f.stop = &stop;
    f.res = &res;
    double h = rvl_brent_fmin(lower, upper, f, 0.1 * lower);
    if (threadIdx.x == 0) tg[t].bcv = h;
}

// ---- host-callable launchers (called from capi.cu) --------------------------------------

extern "C" void rvl_launch_bcv_binary_table(int n, double *table, cudaStream_t st)
{
    int threads = 128, blocks = (n + 1 + threads - 1) / threads;
    k_bcv_binary_table<<<blocks, threads, 0, st>>>(n, table);
}

// Targets 1..T-1 declared to be permutations of target 0 (LIB:1843-1844, 2885): bcv() is invariant
// under permutation, so they inherit target 0's bandwidth.  The claim is checked with the
// order-independent fingerprint; *flag |= 2 when it does not hold.
__global__ void k_bcv_copy_permuted(int T, RvlTarget *__restrict__ tg, int *__restrict__ flag)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 1 || t >= T) return;
    if (tg[t].fingerprint != tg[0].fingerprint) atomicOr(flag, 2);
    tg[t].bcv = tg[0].bcv;
}

extern "C" void rvl_launch_target_setup(const double *x_all, int n, int T, int G, RvlTarget *tg, double *u_all,
                                        double *xc_all, int *bin_all, int *cnt_all, int *nonfinite,
                                        int permuted, cudaStream_t st)
{
    k_target_stats<<<T, 256, 0, st>>>(x_all, n, G, tg, u_all, xc_all, bin_all, nonfinite);
    const int Th = permuted ? 1 : T; // targets that need their own pair histogram + Brent search
    cudaMemsetAsync(cnt_all, 0, sizeof(int) * (size_t)Th * RVL_NB, st);
    int chunks = n < 512 ? n : 512;
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, Th);
    k_pair_histogram<<<grid, 256, 0, st>>>(bin_all, n, cnt_all);
    k_bcv_target<<<Th, 256, 0, st>>>(n, Th, cnt_all, tg);
    if (permuted && T > 1) k_bcv_copy_permuted<<<(T + 127) / 128, 128, 0, st>>>(T, tg, nonfinite);
}
