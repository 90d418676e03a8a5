// capi.cu -- the C ABI declared in include/revealer_b200.h: context, uploads, the search driver.
// Host code here is plumbing only (allocation, copies, launch order); all arithmetic of the path
// runs in the kernels of score.cu / bandwidth.cu.  There is no CPU fallback.
#include "../../include/revealer_b200.h"
#include "rvl_common.cuh"

#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <sched.h>

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/partition.h>
#include <thrust/sort.h>
#include <thrust/scan.h>
#include <thrust/transform.h>
#include <thrust/copy.h>
#include <thrust/gather.h>
#include <thrust/functional.h>
#include <thrust/iterator/counting_iterator.h>

// dynamic shared memory k_score may ask for: 227 KB per block minus its static part (the 32 KB log table
// of rvl_log_pos + 1 KB of slack)
#ifndef RVL_SCORE_DYN_SMEM_MAX
#define RVL_SCORE_DYN_SMEM_MAX ((size_t)(227 - 33) * 1024)
#endif

// launchers (score.cu, bandwidth.cu)
extern "C" {
cudaError_t rvl_upload_log_table(void);
// hostpack.cpp
void *rvl_hostpack_begin(const double *m2, int64_t F, int64_t n, int64_t ld, uint64_t *packed, int nthreads);
int64_t rvl_hostpack_claim_front(void *h);
int rvl_hostpack_end(void *h, int64_t *first_host_word);
void rvl_launch_scatter_words(const uint64_t *packed, int64_t F, int w0, int w1, int words, uint64_t *bits, cudaStream_t st);
// gct.cu
int64_t rvl_gct_row_names_impl(rvl_gct *g, int field, char *buf, size_t cap);
int rvl_gct_open_impl(const char *path, rvl_gct **out, char *errbuf, size_t errcap);
void rvl_gct_close_impl(rvl_gct *g);
int rvl_gct_dims_impl(const rvl_gct *g, int64_t *nrow, int64_t *ncol);
int rvl_gct_sample_name_impl(const rvl_gct *g, int64_t j, char *buf, size_t cap);
int rvl_gct_row_field_impl(rvl_gct *g, int64_t i, int field, char *buf, size_t cap);
int rvl_gct_row_values_impl(rvl_gct *g, int64_t i, double *out);
const char *rvl_gct_error_impl(const rvl_gct *g);
int rvl_gct_parse_to_bits(rvl_gct *g, const int32_t *cols, int64_t n, int64_t first_row, int64_t n_rows, int words,
                          uint64_t *d_bits, int *d_flag, cudaStream_t st, int64_t *launches, void **scratch_io,
                          char *errbuf, size_t errcap);
void rvl_gct_scratch_free(void *p);
void rvl_launch_gather_rows(const uint64_t *src, const int64_t *rows, int64_t k, int words, uint64_t *dst, cudaStream_t st);
void rvl_launch_unpack_rows(const uint64_t *src, const int64_t *rows, int64_t k, int n, int words, uint8_t *dst, cudaStream_t st);
void rvl_launch_bcv_binary_table(int n, double *table, cudaStream_t st);
void rvl_launch_target_setup(const double *x_all, int n, int T, int G, RvlTarget *tg, double *u_all, double *xc_all,
                             int *bin_all, int *cnt_all, int *nonfinite, int permuted, cudaStream_t st);
size_t rvl_score_plan(int words, int G, size_t budget, int max_warps, int *warps_out);
void rvl_set_pdl(int on);
cudaError_t rvl_score_configure(size_t smem, int warps, int device, int *grid_out);
void rvl_launch_score(const RvlParams *P, int grid_x, int warps, size_t smem, const uint64_t *bits, const double *u_all,
                      const double *xc_all, const RvlTarget *tg, const uint64_t *seeds, const double *bcvtab,
                      const double *tt, double *scores, RvlCandHead *part, unsigned long long *work,
                      unsigned long long *ctr, const int64_t *uniq, int64_t U, cudaStream_t st);
void rvl_launch_advance(const RvlParams *P, RvlTarget *tg, const double *u_all, const uint64_t *seeds_in,
                        uint64_t *seeds_out, const double *bcvtab, double *tt, RvlCandHead *slots, int nslots,
                        unsigned long long *work, const unsigned char *recv, size_t rec_bytes, int iter, int iters,
                        int first_null, int64_t *top, double *top_score, uint64_t *seed_iter, int allow_incr,
                        const RvlPeerXchg *X, cudaStream_t st);
void rvl_launch_fp64_peak(int blocks, int threads, int iters, double *out, cudaStream_t st);
void rvl_launch_assoc(const RvlParams *P, int npairs, const RvlTarget *tg, const int *target_of, const double *u_all,
                      const double *xc_all, const uint64_t *y_all, const double *bcvtab, double *out,
                      size_t y_stride, size_t out_stride, int t_div, cudaStream_t st);
cudaError_t rvl_search_configure(size_t smem, int warps, int device, int *grid_out);
cudaError_t rvl_launch_search(const RvlSearchArgs *a, int grid_x, int warps, size_t smem, cudaStream_t st);
void rvl_launch_argmax(const RvlParams *P, const RvlCandHead *part, int nparts, const uint64_t *bits,
                       unsigned char *send, size_t rec_bytes, const RvlPeerXchg *X, cudaStream_t st);
void rvl_launch_pack_f64_colmajor(const double *chunk, int64_t F, int64_t ld, int s0, int ncols, int words,
                                  uint64_t *bits, int *flag, cudaStream_t st);
void rvl_launch_pack_u8_rowmajor(const uint8_t *m, int64_t F0, int64_t Fc, int n, int words, uint64_t *bits,
                                 int *flag, cudaStream_t st);
void rvl_launch_check_padding(const uint64_t *bits, int64_t F, int n, int words, int *flag, cudaStream_t st);
void rvl_launch_row_counts(const uint64_t *bits, int64_t F, int words, int32_t *out, cudaStream_t st);
void rvl_launch_rho_max(const RvlParams *P, const uint64_t *bits, const double *xc_all, RvlTarget *tg,
                        const int64_t *uniq, int64_t U, cudaStream_t st);
void rvl_launch_row_hash(const uint64_t *bits, int64_t F, int words, uint64_t *hash, int64_t *idx, cudaStream_t st);
void rvl_launch_dedup_heads(const uint64_t *hash, int64_t F, int64_t *headpos, cudaStream_t st);
void rvl_launch_dedup_verify(const uint64_t *bits, int64_t F, int words, const int64_t *idx, const int64_t *headpos,
                             int64_t *rep, cudaStream_t st);
void rvl_launch_fill_duplicates(double *scores, int64_t F, int64_t rows, const int64_t *rep, cudaStream_t st);
void rvl_launch_pvalues(const double *cmi, int64_t F, const double *rand_sorted, int64_t n_valid, int64_t n_rand,
                        double *pval, cudaStream_t st);
}

struct rvl_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    rvl_opts opts;

    // shard
    int rank = 0, world = 1;
    int64_t row_offset = 0, F_global = -1;

    // matrix
    int64_t F = -1;
    int n = 0, words = 0;
    uint64_t *d_bits = nullptr;
    size_t bits_cap = 0;   // capacity of d_bits in uint64 (buffers are kept across loads: a context that is
    size_t rows_cap = 0;   // re-loaded -- an R session, bench.py's e2e loop -- does not re-allocate)
    // upload staging (two halves, filled by the copy stream while the main stream packs the other)
    unsigned char *d_stage[2] = {nullptr, nullptr};
    size_t stage_cap = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_packed[2] = {nullptr, nullptr};
    // scratch of the duplicate-row grouping + pooled temporary storage for thrust
    uint64_t *d_hash = nullptr;
    int64_t *d_idx = nullptr, *d_head = nullptr;
    struct pool_block { char *p; size_t bytes; bool used; };
    std::vector<pool_block> pool;
    void *gct_scratch = nullptr; // staging of the GCT ingest (gct.cu), kept across loads
    // cooperative upload of R doubles: host packer threads (hostpack.cpp) + their pinned / device word buffers
    int host_threads = -1;       // -1: automatic, 0: off (DMA + device packing only)
    uint64_t *h_packed = nullptr, *d_packed = nullptr;
    size_t packed_cap = 0;
    uint64_t last_h2d_bytes = 0, last_host_packed_words = 0; // of the last rvl_load_matrix_f64_colmajor
    // exact-duplicate rows: rep[f] = first identical row; uniq = rows with rep[f] == f
    int64_t *d_rep = nullptr, *d_uniq = nullptr;
    int64_t U = 0;
    int dedup = 1;
    bool fill_pending = false; // duplicate rows of the score buffer not filled in yet
    bool rho_dirty = true;     // RvlTarget.rho_*_max_bits must be recomputed (matrix or targets changed)

    // targets
    int T = 0;
    int first_null = 0;
    int G_targets = 0; // n_grid the target tables were built for
    double *d_x = nullptr, *d_u = nullptr, *d_xc = nullptr, *d_bcvtab = nullptr;
    int tg_alloc_T = 0, tg_alloc_n = 0, tg_alloc_words = 0; // shape the target buffers were allocated for (kept when it repeats)
    int *d_bin = nullptr, *d_cnt = nullptr, *d_flag = nullptr;
    RvlTarget *d_tg = nullptr;

    // per-iteration x-axis table [T][nodes][2][RVL_MAXG] (score.cu, k_advance)
    double *d_tt = nullptr;
    size_t tt_cap = 0;
    int dense_x = 0;
    double prune_theta = 1e-14;

    // seeds
    uint64_t *d_seed0 = nullptr; // initial seeds [T][words]
    uint64_t *d_seed = nullptr;  // working seeds  [T][words] (current)
    uint64_t *d_seed_alt = nullptr; // the other half of the double buffer (k_advance reads one, writes the other)
    bool have_seed = false;

    // search state
    int iters = 0;
    double *d_scores = nullptr; // [iters][T][F]  (rvl_score_once uses slot 0)
    size_t scores_cap = 0;
    int64_t *d_top = nullptr;
    double *d_top_score = nullptr, *d_cmi_seed = nullptr, *d_ic0 = nullptr;
    uint64_t *d_seed_iter = nullptr;
    int res_cap_T = 0, res_cap_iters = 0, res_cap_words = 0;
    unsigned char *d_send = nullptr, *d_recv = nullptr; // exchange buffers
    unsigned char *d_own_xchg = nullptr;                // owned fallback (world == 1)
    size_t own_xchg_bytes = 0;

    // peer-memory exchange of the candidate records (rvl_p2p_export / rvl_p2p_import)
    unsigned char *d_xchg = nullptr;   // this rank's buffer: [2][world][T] records + [2][world][T] stamps
    size_t xchg_bytes = 0, xchg_stamps_off = 0;
    int xchg_world = 0, xchg_T = 0, xchg_words = 0;
    unsigned char **d_peers = nullptr; // device array [world] of every rank's buffer
    std::vector<void *> opened;        // peer mappings to close
    bool p2p = false;
    int *d_xerr = nullptr;
    unsigned int epoch = 0;

    // fused search (k_search): one persistent cooperative launch per search
    int fused = 0;                  // rvl_set_fused (default off: measured 2-3 % slower than the split-phase launches)
    int search_grid_x = -1;         // -1: not configured; 0: cooperative launch unavailable
    size_t search_smem = 0;
    int search_warps = 0;
    unsigned long long *d_swork = nullptr; // [sctr_cap] per-iteration item counters
    unsigned int *d_sdone = nullptr;       // [sctr_cap] per-iteration block tickets, then the barrier counter
    int sctr_cap = 0;
    unsigned char *d_self_xchg = nullptr;  // world == 1: the "exchange" buffer the last block publishes into
    unsigned char **d_self_peers = nullptr;
    size_t self_xchg_bytes = 0, self_stamps_off = 0;
    int self_T = 0, self_words = 0;
    int xseq = 0, seq0 = 0;         // exchanges issued so far / before the current search: the parity slot runs on across searches
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> search_events; // profiling: one pair per k_search launch
    size_t search_events_used = 0;
    int last_search_iters = 0;
    unsigned long long *d_phase_dbg = nullptr; // rvl_debug_phase_times: [64][8] globaltimer stamps of the last k_search

    // per-CTA bests of the last k_score launch [T][score_grid_x]; side stream for the seed IC
    RvlCandHead *d_part = nullptr;
    size_t part_cap = 0;
    int score_grid_x = 0; // persistent grid of k_score (SMs x resident CTAs)
    int score_warps = 0;  // worker warps per block (rvl_score_plan)
    int score_warps_fixed = 0; // > 0: block width forced (environment RVL_SCORE_WARPS, tuning only)
    size_t score_smem = 0;
    unsigned long long *d_work = nullptr; // k_score's item counter
    bool work_dirty = true;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_pending = false;

    // instrumentation (rvl_set_profiling): work counters of k_score and a CUDA-event pair around each of its launches
    int profiling = 0;
    unsigned long long *d_ctr = nullptr; // [0] KDE evaluations, [1] four-term (j,k) columns, [2] x-axis exps, [3] factorised, [4] one-class soft columns, [5] logs, [6] flops, [7] soft cells
    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> score_events;
    size_t score_events_used = 0;
    double score_ms = 0.0;
    int64_t score_launches = 0;
};

static void p2p_release(rvl_ctx *c);
static void quiesce_side(rvl_ctx *c);

static int fail(rvl_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, RVL_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define GUARD()                                                                  \
    if (!ctx) return RVL_ERR_ARG;                                                \
    {                                                                            \
        cudaError_t e_ = cudaSetDevice(ctx->device);                             \
        if (e_ != cudaSuccess) return fail(ctx, RVL_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e_)); \
    }

template <class T> static cudaError_t dalloc(T **p, size_t count)
{
    *p = nullptr;
    return cudaMalloc((void **)p, sizeof(T) * (count ? count : 1));
}
template <class T> static void dfree(T *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}

// The x-axis table costs about as much as 70-360 literal passes per target and iteration; with only
// a handful of rows to score (REVEALER_assess_features: a few rows x 10 000 permuted targets) the
// literal per-row pass is cheaper.
static bool use_dense_x(const rvl_ctx *c)
{
    int64_t rows = c->dedup ? c->U : c->F;
    return c->dense_x || (rows >= 0 && rows <= 64);
}

static RvlParams make_params(const rvl_ctx *c)
{
    RvlParams P;
    P.n = c->n; P.words = c->words; P.G = c->opts.n_grid; P.negative = c->opts.negative;
    P.bw_mult = c->opts.bandwidth_mult; P.bw_shrink = c->opts.bandwidth_shrink;
    P.F = c->F; P.row_offset = c->row_offset; P.T = c->T; P.world = c->world; P.rank = c->rank;
    // tau = -2 ln c ranges over [0, -2 ln(1 + shrink)]; RVL_TT_PAD nodes beyond each end
    double tau_max = -2.0 * log(1.0 + c->opts.bandwidth_shrink);
    P.tt_dt = 1.0 / RVL_TT_INV_DT;
    P.tt_pad = RVL_TT_PAD;
    P.tt_nodes = (int)ceil(tau_max / P.tt_dt) + 2 * RVL_TT_PAD + 1;
    if (P.tt_nodes < RVL_TT_STENCIL) P.tt_nodes = RVL_TT_STENCIL;
    P.dense_x = use_dense_x(c) ? 1 : 0;
    P.literal = c->dense_x;
    P.prune_theta = c->prune_theta;
    return P;
}

extern "C" int rvl_abi_version(void) { return RVL_ABI_VERSION; }

extern "C" int rvl_default_options(rvl_opts *o)
{
    if (!o) return RVL_ERR_ARG;
    o->n_grid = 25; o->negative = 0; o->bandwidth_mult = 0.25; o->bandwidth_shrink = -0.75;
    return RVL_OK;
}

extern "C" int rvl_create(rvl_ctx **out, int device)
{
    if (!out) return RVL_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return RVL_ERR_CUDA;
    rvl_ctx *ctx = new (std::nothrow) rvl_ctx();
    if (!ctx) return RVL_ERR_ARG;
    ctx->device = device;
    rvl_default_options(&ctx->opts);
    if (cudaSetDevice(device) != cudaSuccess || rvl_upload_log_table() != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void **)&ctx->d_flag, sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&ctx->d_xerr, sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->d_xerr, 0, sizeof(int)) != cudaSuccess ||
        cudaMalloc((void **)&ctx->d_ctr, 8 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(ctx->d_ctr, 0, 8 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx;
        return RVL_ERR_CUDA;
    }
    ctx->own_stream = true;
    if (const char *w = getenv("RVL_SCORE_WARPS")) ctx->score_warps_fixed = atoi(w); // tuning: force the block width of k_score
    if (const char *w = getenv("RVL_PDL")) rvl_set_pdl(atoi(w));
    *out = ctx;
    return RVL_OK;
}

// Forget the loaded matrix; the device buffers stay for the next load unless `release`.
static void free_matrix(rvl_ctx *c, bool release = false)
{
    c->U = 0;
    c->rho_dirty = true;
    c->F = -1;
    if (!release) return;
    dfree(c->d_bits); c->bits_cap = 0;
    dfree(c->d_rep); dfree(c->d_uniq); dfree(c->d_hash); dfree(c->d_idx); dfree(c->d_head);
    c->rows_cap = 0;
    dfree(c->d_stage[0]); dfree(c->d_stage[1]); c->stage_cap = 0;
    for (auto &b : c->pool) cudaFree(b.p);
    c->pool.clear();
    if (c->gct_scratch) rvl_gct_scratch_free(c->gct_scratch);
    c->gct_scratch = nullptr;
    if (c->h_packed) cudaFreeHost(c->h_packed);
    c->h_packed = nullptr;
    dfree(c->d_packed);
    c->packed_cap = 0;
}
static void free_targets(rvl_ctx *c)
{
    dfree(c->d_x); dfree(c->d_u); dfree(c->d_xc); dfree(c->d_bcvtab); dfree(c->d_bin); dfree(c->d_cnt);
    dfree(c->d_tg); dfree(c->d_seed0); dfree(c->d_seed); dfree(c->d_seed_alt); dfree(c->d_tt);
    c->tt_cap = 0;
    c->tg_alloc_T = c->tg_alloc_n = c->tg_alloc_words = 0;
    c->T = 0; c->have_seed = false;
    c->rho_dirty = true;
}
static void free_results(rvl_ctx *c)
{
    dfree(c->d_scores); c->scores_cap = 0;
    dfree(c->d_top); dfree(c->d_top_score); dfree(c->d_cmi_seed); dfree(c->d_ic0); dfree(c->d_seed_iter);
    c->res_cap_T = c->res_cap_iters = c->res_cap_words = 0;
    dfree(c->d_own_xchg); c->own_xchg_bytes = 0;
}

extern "C" void rvl_destroy(rvl_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_matrix(ctx, true); free_targets(ctx); free_results(ctx);
    p2p_release(ctx);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
        if (ctx->ev_packed[i]) cudaEventDestroy(ctx->ev_packed[i]);
    }
    dfree(ctx->d_xerr);
    dfree(ctx->d_flag);
    dfree(ctx->d_ctr);
    for (auto &p : ctx->score_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    dfree(ctx->d_part);
    dfree(ctx->d_work);
    dfree(ctx->d_swork); dfree(ctx->d_sdone); dfree(ctx->d_self_xchg); dfree(ctx->d_self_peers); dfree(ctx->d_phase_dbg);
    for (auto &p : ctx->search_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *rvl_last_error(const rvl_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int rvl_set_stream(rvl_ctx *ctx, void *cuda_stream)
{
    GUARD();
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return RVL_OK;
}

extern "C" int rvl_set_options(rvl_ctx *ctx, const rvl_opts *o)
{
    GUARD();
    if (!o) return fail(ctx, RVL_ERR_ARG, "opts is NULL");
    if (o->n_grid < 2 || o->n_grid > RVL_MAX_GRID) return fail(ctx, RVL_ERR_ARG, "n_grid must be in [2, %d]", RVL_MAX_GRID);
    if (!(o->bandwidth_mult > 0.0)) return fail(ctx, RVL_ERR_ARG, "bandwidth_mult must be > 0");
    if (!(o->bandwidth_shrink > -1.0) || !(o->bandwidth_shrink <= 0.0))
        return fail(ctx, RVL_ERR_ARG, "bandwidth_shrink must be in (-1, 0]");
    if (ctx->T > 0 && o->n_grid != ctx->G_targets)
        return fail(ctx, RVL_ERR_STATE, "n_grid must be set before rvl_set_targets");
    ctx->opts = *o;
    return RVL_OK;
}

extern "C" int rvl_get_upload_stats(rvl_ctx *ctx, uint64_t *out_h2d_bytes, uint64_t *out_host_packed_words)
{
    GUARD();
    if (out_h2d_bytes) *out_h2d_bytes = ctx->last_h2d_bytes;
    if (out_host_packed_words) *out_host_packed_words = ctx->last_host_packed_words;
    return RVL_OK;
}

extern "C" int rvl_set_host_threads(rvl_ctx *ctx, int n)
{
    GUARD();
    if (n > 256) return fail(ctx, RVL_ERR_ARG, "host threads must be <= 256 (negative: automatic, 0: off)");
    ctx->host_threads = n < 0 ? -1 : n;
    return RVL_OK;
}

extern "C" int rvl_set_shard(rvl_ctx *ctx, int rank, int world, int64_t row_offset, int64_t F_global)
{
    GUARD();
    if (world < 1 || rank < 0 || rank >= world || row_offset < 0) return fail(ctx, RVL_ERR_ARG, "bad shard spec");
    ctx->rank = rank; ctx->world = world; ctx->row_offset = row_offset; ctx->F_global = F_global;
    return RVL_OK;
}

// ---- matrix ---------------------------------------------------------------------------------

// sort key of a row for the k_score work list: minus its minority-class count (densest first)
struct rvl_cost_key {
    const int32_t *cnt;
    int n;
    __host__ __device__ int32_t operator()(int64_t f) const
    {
        int32_t c = cnt[f];
        return -(c < n - c ? c : n - c);
    }
};

struct rvl_is_rep {
    const int64_t *rep;
    __host__ __device__ bool operator()(int64_t f) const { return rep[f] == f; }
};

// Temporary storage of the thrust calls below comes from a small pool kept in the context
// (cudaMalloc / cudaFree per call would cost more than the sorts themselves).
struct rvl_pool_alloc {
    typedef char value_type;
    rvl_ctx *c;
    char *allocate(std::ptrdiff_t n)
    {
        size_t need = n > 0 ? (size_t)n : 1;
        rvl_ctx::pool_block *best = nullptr;
        for (auto &b : c->pool)
            if (!b.used && b.bytes >= need && (!best || b.bytes < best->bytes)) best = &b;
        if (best) { best->used = true; return best->p; }
        char *p = nullptr;
        size_t bytes = (need + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        if (cudaMalloc((void **)&p, bytes) != cudaSuccess) throw std::bad_alloc();
        c->pool.push_back({p, bytes, true});
        return p;
    }
    void deallocate(char *p, size_t)
    {
        for (auto &b : c->pool)
            if (b.p == p) { b.used = false; return; }
    }
};

// Group identical rows (hash -> stable sort -> exact compare); see score.cu "exact-duplicate rows".
static int build_dedup(rvl_ctx *ctx)
{
    int64_t F = ctx->F;
    ctx->U = F;
    if (F <= 0) return RVL_OK;
    uint64_t *d_hash = ctx->d_hash;
    int64_t *d_idx = ctx->d_idx, *d_head = ctx->d_head;
    int rc = RVL_OK;
    try {
        rvl_pool_alloc alloc{ctx};
        auto pol = thrust::cuda::par(alloc).on(ctx->stream);
        rvl_launch_row_hash(ctx->d_bits, F, ctx->words, d_hash, d_idx, ctx->stream);
        thrust::stable_sort_by_key(pol, thrust::device_ptr<uint64_t>(d_hash), thrust::device_ptr<uint64_t>(d_hash) + F,
                                   thrust::device_ptr<int64_t>(d_idx));
        rvl_launch_dedup_heads(d_hash, F, d_head, ctx->stream);
        thrust::inclusive_scan(pol, thrust::device_ptr<int64_t>(d_head), thrust::device_ptr<int64_t>(d_head) + F,
                               thrust::device_ptr<int64_t>(d_head), thrust::maximum<int64_t>());
        rvl_launch_dedup_verify(ctx->d_bits, F, ctx->words, d_idx, d_head, ctx->d_rep, ctx->stream);
        rvl_is_rep pred{ctx->d_rep};
        auto end = thrust::copy_if(pol, thrust::counting_iterator<int64_t>(0), thrust::counting_iterator<int64_t>(F),
                                   thrust::device_ptr<int64_t>(ctx->d_uniq), pred);
        ctx->U = end - thrust::device_ptr<int64_t>(ctx->d_uniq);
        // schedule the densest rows first (their evaluations are the longest: more live density columns),
        // so that the persistent k_score grid drains on short items -- order affects timing only
        int32_t *d_cnt = reinterpret_cast<int32_t *>(d_head); // reuse: F int64 >= F int32
        int32_t *d_key = reinterpret_cast<int32_t *>(d_idx);
        rvl_launch_row_counts(ctx->d_bits, F, ctx->words, d_cnt, ctx->stream);
        rvl_cost_key key{d_cnt, ctx->n};
        thrust::transform(pol, thrust::device_ptr<int64_t>(ctx->d_uniq), thrust::device_ptr<int64_t>(ctx->d_uniq) + ctx->U,
                          thrust::device_ptr<int32_t>(d_key), key);
        thrust::stable_sort_by_key(pol, thrust::device_ptr<int32_t>(d_key), thrust::device_ptr<int32_t>(d_key) + ctx->U,
                                   thrust::device_ptr<int64_t>(ctx->d_uniq));
        ctx->launches += 9;
    } catch (...) {
        rc = fail(ctx, RVL_ERR_CUDA, "duplicate-row grouping failed on the device");
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (!rc && e != cudaSuccess) rc = fail(ctx, RVL_ERR_CUDA, "duplicate-row grouping: %s", cudaGetErrorString(e));
    if (rc) free_matrix(ctx);
    return rc;
}

// Device buffers for an F x words matrix and its per-row side arrays (grown, never shrunk).
static int reserve_matrix(rvl_ctx *ctx, int64_t F)
{
    size_t need = (size_t)F * ctx->words;
    if (need > ctx->bits_cap || !ctx->d_bits) {
        dfree(ctx->d_bits);
        ctx->bits_cap = 0;
        CK(dalloc(&ctx->d_bits, need));
        ctx->bits_cap = need ? need : 1;
    }
    if ((size_t)F > ctx->rows_cap || !ctx->d_rep) {
        dfree(ctx->d_rep); dfree(ctx->d_uniq); dfree(ctx->d_hash); dfree(ctx->d_idx); dfree(ctx->d_head);
        ctx->rows_cap = 0;
        CK(dalloc(&ctx->d_rep, (size_t)F));
        CK(dalloc(&ctx->d_uniq, (size_t)F));
        CK(dalloc(&ctx->d_hash, (size_t)F));
        CK(dalloc(&ctx->d_idx, (size_t)F));
        CK(dalloc(&ctx->d_head, (size_t)F));
        ctx->rows_cap = F > 0 ? (size_t)F : 1;
    }
    return RVL_OK;
}

// Two staging halves of `bytes` each + the copy stream and events that pipeline H2D against packing.
static int reserve_stage(rvl_ctx *ctx, size_t bytes)
{
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            CK(cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_packed[i], cudaEventDisableTiming));
        }
    }
    if (bytes > ctx->stage_cap) {
        dfree(ctx->d_stage[0]); dfree(ctx->d_stage[1]);
        ctx->stage_cap = 0;
        CK(dalloc(&ctx->d_stage[0], bytes));
        CK(dalloc(&ctx->d_stage[1], bytes));
        ctx->stage_cap = bytes;
    }
    return RVL_OK;
}

static int begin_matrix(rvl_ctx *ctx, int64_t F, int64_t n)
{
    quiesce_side(ctx);
    if (F < 0 || n < 2 || n > (1 << 24)) return fail(ctx, RVL_ERR_ARG, "need F >= 0 and 2 <= n <= 2^24 (F=%lld n=%lld)", (long long)F, (long long)n);
    if (ctx->T > 0 && ctx->n != (int)n) return fail(ctx, RVL_ERR_ARG, "matrix has n=%lld but targets have n=%d", (long long)n, ctx->n);
    {
        int w = 0;
        if (!rvl_score_plan((int)(((n + 63) / 64 + 1) & ~1ll), RVL_MAX_GRID, RVL_SCORE_DYN_SMEM_MAX, 0, &w))
            return fail(ctx, RVL_ERR_ARG, "n=%lld samples exceed the score kernel's shared-memory budget (one bit row of n/8 bytes per worker warp; n <= ~1 400 000)", (long long)n);
    }
    free_matrix(ctx);
    ctx->n = (int)n;
    ctx->words = (int)(((n + 63) / 64 + 1) & ~1ll); // rows padded to 16 B
    return RVL_OK;
}

static int check_flag(rvl_ctx *ctx, const char *what)
{
    int h = 0;
    CK(cudaMemcpyAsync(&h, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (h) return fail(ctx, RVL_ERR_NOT_BINARY, "%s", what);
    return RVL_OK;
}

extern "C" int rvl_load_matrix_f64_colmajor(rvl_ctx *ctx, const double *m2, int64_t F, int64_t n, int64_t ld)
{
    GUARD();
    if (!m2 && F > 0) return fail(ctx, RVL_ERR_ARG, "m2 is NULL");
    if (ld < F) return fail(ctx, RVL_ERR_ARG, "ld < F");
    int rc = begin_matrix(ctx, F, n);
    if (rc) return rc;
    rc = reserve_matrix(ctx, F);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->d_bits, 0, sizeof(uint64_t) * (size_t)F * ctx->words, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    ctx->F = F;
    if (F == 0) return RVL_OK;
    // Packer threads: PCIe carries 8 bytes for every bit the device keeps, so while the DMA engine ships the
    // leading 64-sample words as doubles, host threads turn the trailing ones into bits (hostpack.cpp: the
    // same packing and 0/1 validation as k_pack_f64_colmajor, 1/64 of the bytes to send).  The two sides
    // claim words from opposite ends until they meet, so the split follows their actual speeds -- with
    // pageable (unpinned) R memory the packers end up doing most of it.
    const int64_t W = (n + 63) / 64;
    int nthreads = ctx->host_threads;
    if (nthreads < 0) {
        const int total = (int)std::thread::hardware_concurrency();
        int avail = total;
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof set, &set) == 0) avail = CPU_COUNT(&set);
        // one process per GPU: a process already confined to its share of the cores (bench.py binds every rank to the
        // cores local to its GPU) uses all of them; an unconfined one takes 1/world of the machine
        nthreads = (total > 0 && avail < total) ? avail : avail / (ctx->world > 0 ? ctx->world : 1);
        if (nthreads > 32) nthreads = 32;
        if ((double)F * (double)n < 4e6 || W < 2) nthreads = 0; // small matrices: not worth starting threads
    }
    void *plan = nullptr;
    if (nthreads > 0) {
        const size_t need = (size_t)W * (size_t)F;
        if (need > ctx->packed_cap) {
            if (ctx->h_packed) cudaFreeHost(ctx->h_packed);
            ctx->h_packed = nullptr;
            dfree(ctx->d_packed);
            ctx->packed_cap = 0;
            CK(cudaMallocHost((void **)&ctx->h_packed, sizeof(uint64_t) * need));
            CK(dalloc(&ctx->d_packed, need));
            ctx->packed_cap = need;
        }
        plan = rvl_hostpack_begin(m2, F, n, ld, ctx->h_packed, nthreads);
    }
    // the DMA side: column chunks of whole bit words (~64 MB each; exactly one word when packers run, so that the
    // two sides meet on a word boundary) through two staging halves -- the copy stream fills one while the main
    // stream packs the other, so the DMA engine never waits for a pack kernel
    int64_t cols = (64ll << 20) / (8 * F);
    cols = (cols / 64) * 64;
    if (cols < 64 || plan) cols = 64;
    int64_t stage_cols = cols < n ? cols : n;
    // the staging halves hold the F rows of each column densely (pitch F), whatever the caller's ld: a row block of a
    // larger R matrix (ld > F) is gathered by a pitched copy, so neither PCIe nor the pack kernel ever sees the
    // other shards' rows and nothing beyond m2 + ld*(n-1) + F is read
    rc = reserve_stage(ctx, sizeof(double) * (size_t)F * stage_cols);
    cudaError_t ce = cudaSuccess;
    if (!rc) ce = cudaEventRecord(ctx->ev_packed[0], ctx->stream); // staging halves are free once earlier work is done
    if (!rc && ce == cudaSuccess) ce = cudaEventRecord(ctx->ev_packed[1], ctx->stream);
    int half = 0;
    int64_t s0 = 0;
    ctx->last_h2d_bytes = 0;
    ctx->last_host_packed_words = 0;
    while (!rc && ce == cudaSuccess) {
        if (plan) {
            const int64_t w = rvl_hostpack_claim_front(plan);
            if (w < 0) break;
            s0 = w * 64;
        } else if (s0 >= n) break;
        const int64_t nc = (n - s0 < cols) ? (n - s0) : cols;
        double *stage = reinterpret_cast<double *>(ctx->d_stage[half]);
        ce = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_packed[half], 0);
        if (ce == cudaSuccess) {
            if (ld == F)
                ce = cudaMemcpyAsync(stage, m2 + ld * s0, sizeof(double) * (size_t)F * nc, cudaMemcpyHostToDevice, ctx->copy_stream);
            else
                ce = cudaMemcpy2DAsync(stage, sizeof(double) * (size_t)F, m2 + ld * s0, sizeof(double) * (size_t)ld,
                                       sizeof(double) * (size_t)F, (size_t)nc, cudaMemcpyHostToDevice, ctx->copy_stream);
        }
        if (ce == cudaSuccess) ce = cudaEventRecord(ctx->ev_copied[half], ctx->copy_stream);
        if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[half], 0);
        if (ce != cudaSuccess) break;
        ctx->last_h2d_bytes += sizeof(double) * (uint64_t)F * (uint64_t)nc;
        rvl_launch_pack_f64_colmajor(stage, F, F, (int)s0, (int)nc, ctx->words, ctx->d_bits, ctx->d_flag, ctx->stream);
        ce = cudaEventRecord(ctx->ev_packed[half], ctx->stream);
        ctx->launches++;
        half ^= 1;
        if (!plan) s0 += cols;
        else if (ce == cudaSuccess) ce = cudaEventSynchronize(ctx->ev_copied[half ^ 1]); // claim the next word only when this copy is done
    }
    int host_bad = 0;
    if (plan) { // always join the packers, also on an error above
        int64_t first = W;
        host_bad = rvl_hostpack_end(plan, &first);
        if (!rc && ce == cudaSuccess && first < W) {
            ce = cudaMemcpyAsync(ctx->d_packed, ctx->h_packed + (size_t)first * F, sizeof(uint64_t) * (size_t)(W - first) * F,
                                 cudaMemcpyHostToDevice, ctx->stream);
            if (ce == cudaSuccess) {
                rvl_launch_scatter_words(ctx->d_packed, F, (int)first, (int)W, ctx->words, ctx->d_bits, ctx->stream);
                ctx->launches++;
                ctx->last_h2d_bytes += sizeof(uint64_t) * (uint64_t)(W - first) * (uint64_t)F;
                ctx->last_host_packed_words = (uint64_t)(W - first);
            }
        }
    }
    if (rc) { free_matrix(ctx); return rc; }
    if (ce != cudaSuccess) { free_matrix(ctx); return fail(ctx, RVL_ERR_CUDA, "upload: %s", cudaGetErrorString(ce)); }
    if (host_bad) {
        cudaStreamSynchronize(ctx->stream);
        free_matrix(ctx);
        return fail(ctx, RVL_ERR_NOT_BINARY, "feature matrix has an entry that is not 0 or 1 (binary features only; no fallback)");
    }
    rc = check_flag(ctx, "feature matrix has an entry that is not 0 or 1 (binary features only; no fallback)");
    if (rc) { free_matrix(ctx); return rc; }
    return build_dedup(ctx);
}

extern "C" int rvl_load_matrix_u8_rowmajor(rvl_ctx *ctx, const uint8_t *m2, int64_t F, int64_t n)
{
    GUARD();
    if (!m2 && F > 0) return fail(ctx, RVL_ERR_ARG, "m2 is NULL");
    int rc = begin_matrix(ctx, F, n);
    if (rc) return rc;
    rc = reserve_matrix(ctx, F);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    ctx->F = F;
    if (F == 0) return RVL_OK;
    int64_t rows = (64ll << 20) / n; // row blocks of ~64 MB through the two staging halves (as above)
    if (rows < 1) rows = 1;
    if (rows > F) rows = F;
    rc = reserve_stage(ctx, (size_t)rows * n);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev_packed[0], ctx->stream));
    CK(cudaEventRecord(ctx->ev_packed[1], ctx->stream));
    int half = 0;
    for (int64_t f0 = 0; f0 < F; f0 += rows, half ^= 1) {
        int64_t fc = (F - f0 < rows) ? (F - f0) : rows;
        uint8_t *stage = ctx->d_stage[half];
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_packed[half], 0));
        CK(cudaMemcpyAsync(stage, m2 + (size_t)f0 * n, (size_t)fc * n, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->ev_copied[half], ctx->copy_stream));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[half], 0));
        rvl_launch_pack_u8_rowmajor(stage, f0, fc, (int)n, ctx->words, ctx->d_bits, ctx->d_flag, ctx->stream);
        CK(cudaEventRecord(ctx->ev_packed[half], ctx->stream));
        ctx->launches++;
    }
    rc = check_flag(ctx, "feature matrix has an entry that is not 0 or 1 (binary features only; no fallback)");
    if (rc) { free_matrix(ctx); return rc; }
    return build_dedup(ctx);
}

static int load_bits_common(rvl_ctx *ctx, const void *src, bool src_device, int64_t F, int64_t n, int64_t words)
{
    if (!src && F > 0) return fail(ctx, RVL_ERR_ARG, "bits is NULL");
    int rc = begin_matrix(ctx, F, n);
    if (rc) return rc;
    if (words < (n + 63) / 64) return fail(ctx, RVL_ERR_ARG, "words < ceil(n/64)");
    rc = reserve_matrix(ctx, F);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    ctx->F = F;
    if (F == 0) return RVL_OK;
    cudaMemcpyKind kind = src_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    int64_t cw = words < ctx->words ? words : ctx->words;
    if (cw < ctx->words) CK(cudaMemsetAsync(ctx->d_bits, 0, sizeof(uint64_t) * (size_t)F * ctx->words, ctx->stream));
    CK(cudaMemcpy2DAsync(ctx->d_bits, sizeof(uint64_t) * ctx->words, src, sizeof(uint64_t) * words,
                         sizeof(uint64_t) * cw, (size_t)F, kind, ctx->stream));
    rvl_launch_check_padding(ctx->d_bits, F, (int)n, ctx->words, ctx->d_flag, ctx->stream);
    ctx->launches++;
    rc = check_flag(ctx, "bit-packed matrix has non-zero padding bits beyond sample n");
    if (rc) { free_matrix(ctx); return rc; }
    return build_dedup(ctx);
}

extern "C" int rvl_load_matrix_bits(rvl_ctx *ctx, const uint64_t *bits, int64_t F, int64_t n, int64_t words)
{
    GUARD();
    return load_bits_common(ctx, bits, false, F, n, words);
}

extern "C" int rvl_load_matrix_bits_device(rvl_ctx *ctx, const void *dev_bits, int64_t F, int64_t n, int64_t words)
{
    GUARD();
    return load_bits_common(ctx, dev_bits, true, F, n, words);
}

// ---- GCT ingest, row selection ---------------------------------------------

static thread_local std::string g_gct_err;

// (the host side of the GCT handle allocates strings and vectors: nothing may be thrown across the ABI)
extern "C" int rvl_gct_open(const char *path, rvl_gct **out)
{
    char buf[512] = {0};
    int rc = -2;
    try {
        rc = rvl_gct_open_impl(path, out, buf, sizeof buf);
        g_gct_err = buf;
    } catch (...) {
        if (out) *out = nullptr;
    }
    return rc ? RVL_ERR_ARG : RVL_OK;
}
extern "C" void rvl_gct_close(rvl_gct *g) { rvl_gct_close_impl(g); }
extern "C" const char *rvl_gct_last_error(const rvl_gct *g)
{
    if (!g) return g_gct_err.c_str();
    return rvl_gct_error_impl(g);
}
extern "C" int rvl_gct_dims(const rvl_gct *g, int64_t *nrow, int64_t *ncol) { return rvl_gct_dims_impl(g, nrow, ncol) ? RVL_ERR_ARG : RVL_OK; }
extern "C" int rvl_gct_sample_name(const rvl_gct *g, int64_t j, char *buf, size_t cap) { return rvl_gct_sample_name_impl(g, j, buf, cap); }
extern "C" int rvl_gct_row_name(rvl_gct *g, int64_t i, char *buf, size_t cap)
{
    try { return rvl_gct_row_field_impl(g, i, 0, buf, cap); } catch (...) { return RVL_ERR_ARG; }
}
extern "C" int rvl_gct_row_description(rvl_gct *g, int64_t i, char *buf, size_t cap)
{
    try { return rvl_gct_row_field_impl(g, i, 1, buf, cap); } catch (...) { return RVL_ERR_ARG; }
}
extern "C" int64_t rvl_gct_row_names(rvl_gct *g, int field, char *buf, size_t cap)
{
    try { return rvl_gct_row_names_impl(g, field, buf, cap); } catch (...) { return RVL_ERR_ARG; }
}
extern "C" int rvl_gct_row_values(rvl_gct *g, int64_t i, double *out)
{
    try { return rvl_gct_row_values_impl(g, i, out) ? RVL_ERR_ARG : RVL_OK; } catch (...) { return RVL_ERR_ARG; }
}

extern "C" int rvl_load_matrix_gct(rvl_ctx *ctx, rvl_gct *g, const int32_t *cols, int64_t n, int64_t first_row, int64_t n_rows)
{
    GUARD();
    if (!g || !cols) return fail(ctx, RVL_ERR_ARG, "gct / cols is NULL");
    int rc = begin_matrix(ctx, n_rows, n);
    if (rc) return rc;
    rc = reserve_matrix(ctx, n_rows);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    ctx->F = n_rows;
    char buf[512] = {0};
    int prc = rvl_gct_parse_to_bits(g, cols, n, first_row, n_rows, ctx->words, ctx->d_bits, ctx->d_flag, ctx->stream,
                                    &ctx->launches, &ctx->gct_scratch, buf, sizeof buf);
    if (prc) {
        free_matrix(ctx);
        return fail(ctx, prc == -1 ? RVL_ERR_CUDA : (prc == -3 ? RVL_ERR_NOT_BINARY : RVL_ERR_ARG), "GCT ingest: %s", buf);
    }
    if (n_rows == 0) return RVL_OK;
    return build_dedup(ctx);
}

// Keep rows `rows[0..k)` (local indices, any order, no repeats required) of the loaded matrix, in that
// order: the device-side form of m.2 <- m.2[retain, ] (count filter LIB:1655-1671, seed rows LIB:1704-1706,
// excluded features LIB:1716-1720).
extern "C" int rvl_select_rows(rvl_ctx *ctx, const int64_t *rows, int64_t k)
{
    GUARD();
    if (ctx->F < 0) return fail(ctx, RVL_ERR_STATE, "feature matrix not loaded");
    if (k < 0 || (!rows && k > 0)) return fail(ctx, RVL_ERR_ARG, "bad row selection");
    for (int64_t i = 0; i < k; i++)
        if (rows[i] < 0 || rows[i] >= ctx->F) return fail(ctx, RVL_ERR_ARG, "row %lld out of range", (long long)rows[i]);
    quiesce_side(ctx);
    uint64_t *d_new = nullptr;
    int64_t *d_rows = nullptr;
    CK(dalloc(&d_new, (size_t)k * ctx->words));
    CK(dalloc(&d_rows, (size_t)k));
    if (k > 0) {
        CK(cudaMemcpyAsync(d_rows, rows, sizeof(int64_t) * (size_t)k, cudaMemcpyHostToDevice, ctx->stream));
        rvl_launch_gather_rows(ctx->d_bits, d_rows, k, ctx->words, d_new, ctx->stream);
        ctx->launches++;
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_rows);
    if (e != cudaSuccess) { cudaFree(d_new); return fail(ctx, RVL_ERR_CUDA, "row selection: %s", cudaGetErrorString(e)); }
    dfree(ctx->d_bits);
    ctx->d_bits = d_new;
    ctx->bits_cap = (size_t)k * ctx->words;
    if (ctx->bits_cap == 0) ctx->bits_cap = 1;
    ctx->F = k;
    ctx->rho_dirty = true;
    int rc = reserve_matrix(ctx, k);
    if (rc) return rc;
    return build_dedup(ctx);
}

// Lexicographic order of bit rows as R orders the strings paste(m.2[k,], collapse = "") (LIB:1728-1729):
// sample 0 decides first.  Sample s sits at bit s%64 of word s/64, so within a word the lowest differing
// bit decides: compare the bit-reversed words as unsigned numbers.
struct rvl_row_lex_less {
    const uint64_t *bits;
    int words;
    __device__ bool operator()(int64_t a, int64_t b) const
    {
        const uint64_t *ra = bits + (size_t)a * words, *rb = bits + (size_t)b * words;
        for (int w = 0; w < words; w++) {
            const uint64_t x = ra[w], y = rb[w];
            if (x != y) return __brevll(x) < __brevll(y);
        }
        return false;
    }
};
struct rvl_row_differs { // position p of the sorted order starts a new group
    const uint64_t *bits;
    const int64_t *order;
    int words;
    __device__ bool operator()(int64_t p) const
    {
        if (p == 0) return true;
        const uint64_t *ra = bits + (size_t)order[p] * words, *rb = bits + (size_t)order[p - 1] * words;
        for (int w = 0; w < words; w++)
            if (ra[w] != rb[w]) return true;
        return false;
    }
};

// consolidate.identical.features = "identical" (LIB:1727-1748) on the device: rows are put in the order of their
// 0/1 strings (stable, so the first row of a group of identical rows is the one with the lowest index, as with
// R's order()), every group keeps its first row, and the matrix becomes those rows in that order.
// out_rows[i] = original (local) index of kept row i, out_counts[i] = size of its group -- the " (count)" suffix
// the reference appends to the row name.  out arrays hold F entries; *out_k rows are kept.
extern "C" int rvl_consolidate_identical(rvl_ctx *ctx, int64_t *out_rows, int32_t *out_counts, int64_t *out_k)
{
    GUARD();
    if (ctx->F < 0) return fail(ctx, RVL_ERR_STATE, "feature matrix not loaded");
    if (!out_rows || !out_counts || !out_k) return fail(ctx, RVL_ERR_ARG, "output is NULL");
    const int64_t F = ctx->F;
    *out_k = 0;
    if (F == 0) return RVL_OK;
    quiesce_side(ctx);
    int64_t *d_order = nullptr, *d_heads = nullptr;
    CK(dalloc(&d_order, (size_t)F));
    CK(dalloc(&d_heads, (size_t)F + 1));
    int64_t k = 0;
    int rc = RVL_OK;
    std::vector<int64_t> heads, order;
    try {
        rvl_pool_alloc alloc{ctx};
        auto pol = thrust::cuda::par(alloc).on(ctx->stream);
        thrust::device_ptr<int64_t> ord(d_order), hd(d_heads);
        thrust::copy(pol, thrust::counting_iterator<int64_t>(0), thrust::counting_iterator<int64_t>(F), ord);
        thrust::stable_sort(pol, ord, ord + F, rvl_row_lex_less{ctx->d_bits, ctx->words});
        auto end = thrust::copy_if(pol, thrust::counting_iterator<int64_t>(0), thrust::counting_iterator<int64_t>(F), hd,
                                   rvl_row_differs{ctx->d_bits, d_order, ctx->words});
        k = end - hd;
        ctx->launches += 4;
        heads.resize((size_t)k + 1);
        order.resize((size_t)F);
        CK(cudaMemcpyAsync(heads.data(), d_heads, sizeof(int64_t) * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(order.data(), d_order, sizeof(int64_t) * (size_t)F, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        rc = fail(ctx, RVL_ERR_CUDA, "identical-row consolidation failed on the device");
    }
    cudaFree(d_order); cudaFree(d_heads);
    if (rc) return rc;
    heads[(size_t)k] = F;
    for (int64_t i = 0; i < k; i++) {
        out_rows[i] = order[(size_t)heads[(size_t)i]];
        out_counts[i] = (int32_t)(heads[(size_t)i + 1] - heads[(size_t)i]);
    }
    *out_k = k;
    return rvl_select_rows(ctx, out_rows, k);
}

// Rows `rows[0..k)` of the loaded matrix as one byte per sample, out[k][n] (the seed rows m.2[seed.names, ],
// LIB:1693-1700).
extern "C" int rvl_get_rows_u8(rvl_ctx *ctx, const int64_t *rows, int64_t k, uint8_t *out)
{
    GUARD();
    if (ctx->F < 0) return fail(ctx, RVL_ERR_STATE, "feature matrix not loaded");
    if (k < 0 || (k > 0 && (!rows || !out))) return fail(ctx, RVL_ERR_ARG, "bad row request");
    if (k == 0) return RVL_OK;
    for (int64_t i = 0; i < k; i++)
        if (rows[i] < 0 || rows[i] >= ctx->F) return fail(ctx, RVL_ERR_ARG, "row %lld out of range", (long long)rows[i]);
    uint8_t *d_out = nullptr;
    int64_t *d_rows = nullptr;
    CK(dalloc(&d_out, (size_t)k * ctx->n));
    CK(dalloc(&d_rows, (size_t)k));
    cudaMemcpyAsync(d_rows, rows, sizeof(int64_t) * (size_t)k, cudaMemcpyHostToDevice, ctx->stream);
    rvl_launch_unpack_rows(ctx->d_bits, d_rows, k, ctx->n, ctx->words, d_out, ctx->stream);
    ctx->launches++;
    cudaMemcpyAsync(out, d_out, (size_t)k * ctx->n, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out); cudaFree(d_rows);
    if (e != cudaSuccess) return fail(ctx, RVL_ERR_CUDA, "rvl_get_rows_u8: %s", cudaGetErrorString(e));
    return RVL_OK;
}

// ---- targets and seeds -----------------------------------------------------------------------

static int set_targets_impl(rvl_ctx *ctx, const double *targets, int64_t n, int64_t n_targets, int permuted);

extern "C" int rvl_set_targets(rvl_ctx *ctx, const double *targets, int64_t n, int64_t n_targets)
{
    GUARD();
    return set_targets_impl(ctx, targets, n, n_targets, 0);
}

extern "C" int rvl_set_targets_permuted(rvl_ctx *ctx, const double *targets, int64_t n, int64_t n_targets)
{
    GUARD();
    return set_targets_impl(ctx, targets, n, n_targets, 1);
}

static int set_targets_impl(rvl_ctx *ctx, const double *targets, int64_t n, int64_t n_targets, int permuted)
{
    quiesce_side(ctx);
    if (!targets || n < 2 || n_targets < 1 || n_targets > 65535) return fail(ctx, RVL_ERR_ARG, "bad targets (n=%lld, n_targets=%lld; need n >= 2, 1 <= n_targets <= 65535)", (long long)n, (long long)n_targets);
    if (ctx->F >= 0 && ctx->n != (int)n) return fail(ctx, RVL_ERR_ARG, "targets have n=%lld but matrix has n=%d", (long long)n, ctx->n);
    {
        int w = 0;
        if (!rvl_score_plan((int)(((n + 63) / 64 + 1) & ~1ll), RVL_MAX_GRID, RVL_SCORE_DYN_SMEM_MAX, 0, &w))
            return fail(ctx, RVL_ERR_ARG, "n=%lld samples exceed the score kernel's shared-memory budget (one bit row of n/8 bytes per worker warp; n <= ~1 400 000)", (long long)n);
    }
    int T = (int)n_targets;
    const int words_now = (ctx->F < 0) ? (int)(((n + 63) / 64 + 1) & ~1ll) : ctx->words;
    if (ctx->d_x && ctx->tg_alloc_T == T && ctx->tg_alloc_n == (int)n && ctx->tg_alloc_words == words_now) {
        // same shape as the last call (a caller scoring one target after the other, the end-to-end step of bench.py): the ten
        // device buffers and the x-axis table stay -- no cudaFree / cudaMalloc round trips
        ctx->T = 0; ctx->have_seed = false; ctx->rho_dirty = true;
    } else {
        free_targets(ctx);
    }
    if (ctx->F < 0) { ctx->n = (int)n; ctx->words = words_now; }
    if (!ctx->d_x) {
        CK(dalloc(&ctx->d_x, (size_t)T * n));
        CK(dalloc(&ctx->d_u, (size_t)T * n));
        CK(dalloc(&ctx->d_xc, (size_t)T * n));
        CK(dalloc(&ctx->d_bin, (size_t)T * n));
        CK(dalloc(&ctx->d_cnt, (size_t)T * 1000));
        CK(dalloc(&ctx->d_bcvtab, (size_t)n + 1));
        CK(dalloc(&ctx->d_tg, (size_t)T));
        CK(dalloc(&ctx->d_seed0, (size_t)T * ctx->words));
        CK(dalloc(&ctx->d_seed, (size_t)T * ctx->words));
        CK(dalloc(&ctx->d_seed_alt, (size_t)T * ctx->words));
        ctx->tg_alloc_T = T; ctx->tg_alloc_n = (int)n; ctx->tg_alloc_words = ctx->words;
    }
    CK(cudaMemsetAsync(ctx->d_tg, 0, sizeof(RvlTarget) * T, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_x, targets, sizeof(double) * (size_t)T * n, cudaMemcpyHostToDevice, ctx->stream));
    rvl_launch_bcv_binary_table((int)n, ctx->d_bcvtab, ctx->stream);
    rvl_launch_target_setup(ctx->d_x, (int)n, T, ctx->opts.n_grid, ctx->d_tg, ctx->d_u, ctx->d_xc, ctx->d_bin,
                            ctx->d_cnt, ctx->d_flag, permuted, ctx->stream);
    ctx->launches += 4 + (permuted ? 1 : 0);
    CK(cudaGetLastError());
    ctx->T = T;
    ctx->G_targets = ctx->opts.n_grid;
    ctx->first_null = 0;
    ctx->have_seed = false;
    int h = 0;
    CK(cudaMemcpyAsync(&h, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (h & 1) { free_targets(ctx); return fail(ctx, RVL_ERR_NONFINITE, "target has NA/NaN/Inf values"); }
    if (h & 2) { free_targets(ctx); return fail(ctx, RVL_ERR_ARG, "rvl_set_targets_permuted: a target is not a permutation of target 0"); }
    // (seed_of defaults to self: written by k_target_stats)
    return RVL_OK;
}

extern "C" int rvl_set_null_targets(rvl_ctx *ctx, int64_t first_null)
{
    GUARD();
    if (ctx->T <= 0) return fail(ctx, RVL_ERR_STATE, "set targets first");
    if (first_null >= ctx->T) return fail(ctx, RVL_ERR_ARG, "first_null >= n_targets");
    int fn = first_null > 0 ? (int)first_null : 0;
    if (fn > 1) return fail(ctx, RVL_ERR_ARG, "null targets are scored against target 0's seed: first_null must be 1 (or <= 0 to disable)");
    std::vector<RvlTarget> tg(ctx->T);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(tg.data(), ctx->d_tg, sizeof(RvlTarget) * ctx->T, cudaMemcpyDeviceToHost));
    for (int t = 0; t < ctx->T; t++) tg[t].seed_of = (fn > 0 && t >= fn) ? 0 : t;
    CK(cudaMemcpy(ctx->d_tg, tg.data(), sizeof(RvlTarget) * ctx->T, cudaMemcpyHostToDevice));
    ctx->first_null = fn;
    return RVL_OK;
}

extern "C" int rvl_set_seed(rvl_ctx *ctx, const uint8_t *seed, int64_t n, int per_target)
{
    GUARD();
    quiesce_side(ctx);
    if (ctx->T <= 0) return fail(ctx, RVL_ERR_STATE, "set targets before the seed");
    if (!seed || n != ctx->n) return fail(ctx, RVL_ERR_ARG, "seed length %lld != n %d", (long long)n, ctx->n);
    int T = ctx->T, words = ctx->words;
    std::vector<uint64_t> packed((size_t)T * words, 0ull);
    for (int t = 0; t < T; t++) {
        const uint8_t *s = per_target ? seed + (size_t)t * n : seed;
        for (int64_t i = 0; i < n; i++) {
            if (s[i] == 1) packed[(size_t)t * words + (i >> 6)] |= (1ull << (i & 63));
            else if (s[i] != 0) return fail(ctx, RVL_ERR_NOT_BINARY, "seed entry %lld is not 0 or 1", (long long)i);
        }
    }
    CK(cudaMemcpyAsync(ctx->d_seed0, packed.data(), sizeof(uint64_t) * packed.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_seed = true;
    return RVL_OK;
}

// ---- search ------------------------------------------------------------------------------------

// Persistent-grid configuration of k_score for the current (n, T): shared memory, grid = SMs x
// resident CTAs, the per-warp best slots and the work counter.
// Worker warps per block of k_score / k_search and the dynamic shared memory they need (0: a bit row does not fit).
static size_t plan_score(rvl_ctx *ctx, int *warps)
{
    size_t smem = rvl_score_plan(ctx->words, ctx->opts.n_grid, RVL_SCORE_DYN_SMEM_MAX, 0, warps);
    if (!smem) return 0;
    // (A narrower block for small shards -- fewer, faster warps so that the last round of whole evaluations is fuller --
    // was measured and loses: 12 500 rows per GPU take 144 us with 16 warps per SM, 150-157 with 15-13, 153 with 12.)
    const int want = ctx->score_warps_fixed;
    if (want > 0 && want < *warps) smem = rvl_score_plan(ctx->words, ctx->opts.n_grid, RVL_SCORE_DYN_SMEM_MAX, want, warps);
    return smem;
}

static int configure_score(rvl_ctx *ctx)
{
    int warps = 0;
    size_t smem = plan_score(ctx, &warps);
    if (!smem) return fail(ctx, RVL_ERR_ARG, "n=%d needs more than 194 KB of dynamic shared memory per worker warp", ctx->n);
    if (smem != ctx->score_smem || warps != ctx->score_warps || ctx->score_grid_x <= 0) {
        int grid = 0;
        CK(rvl_score_configure(smem, warps, ctx->device, &grid));
        ctx->score_smem = smem;
        ctx->score_warps = warps;
        ctx->score_grid_x = grid;
    }
    size_t pneed = (size_t)ctx->T * (size_t)ctx->score_grid_x * ctx->score_warps;
    if (pneed > ctx->part_cap) {
        dfree(ctx->d_part);
        CK(dalloc(&ctx->d_part, pneed));
        ctx->part_cap = pneed;
    }
    if (!ctx->d_work) CK(dalloc(&ctx->d_work, 1));
    return RVL_OK;
}

// Per-iteration seed state + the x-axis table that depends on it.
static int prep_iteration(rvl_ctx *ctx)
{
    RvlParams P = make_params(ctx);
    size_t need = (size_t)ctx->T * P.tt_nodes * 2 * RVL_MAXG;
    if (need > ctx->tt_cap) {
        dfree(ctx->d_tt);
        CK(dalloc(&ctx->d_tt, need));
        ctx->tt_cap = need;
    }
    int rc = configure_score(ctx);
    if (rc) return rc;
    if (ctx->rho_dirty && !use_dense_x(ctx)) { // range of the bandwidth shrink over the loaded rows (once per matrix/targets)
        CK(cudaMemset2DAsync((char *)ctx->d_tg + offsetof(RvlTarget, rho_pos_max_bits), sizeof(RvlTarget), 0,
                             2 * sizeof(unsigned long long), (size_t)ctx->T, ctx->stream));
        if (ctx->F > 0) {
            rvl_launch_rho_max(&P, ctx->d_bits, ctx->d_xc, ctx->d_tg, ctx->U < ctx->F ? ctx->d_uniq : nullptr, ctx->U,
                               ctx->stream);
            ctx->launches++;
        }
        ctx->rho_dirty = false;
    }
    const int nslots = ctx->score_grid_x * ctx->score_warps;
    // no merge here (recv = NULL): seed state + full table build for the seeds as they are
    RvlPeerXchg Xnone;
    memset(&Xnone, 0, sizeof Xnone);
    rvl_launch_advance(&P, ctx->d_tg, ctx->d_u, ctx->d_seed, ctx->d_seed_alt, ctx->d_bcvtab, ctx->d_tt, ctx->d_part, nslots,
                       ctx->d_work, nullptr, 0, 0, 1, ctx->first_null, nullptr, nullptr, nullptr, 0, &Xnone, ctx->stream);
    ctx->work_dirty = false;
    ctx->launches++;
    CK(cudaGetLastError());
    return RVL_OK;
}

static int ready(rvl_ctx *ctx)
{
    if (ctx->F < 0) return fail(ctx, RVL_ERR_STATE, "feature matrix not loaded");
    if (ctx->T <= 0) return fail(ctx, RVL_ERR_STATE, "targets not set");
    if (!ctx->have_seed) return fail(ctx, RVL_ERR_STATE, "seed not set");
    if (ctx->opts.n_grid != ctx->G_targets) return fail(ctx, RVL_ERR_STATE, "n_grid changed after rvl_set_targets");
    return RVL_OK;
}

// Inputs are about to be replaced: let the side-stream seed ICs that still read them finish.
static void quiesce_side(rvl_ctx *c)
{
    if (c->side_pending) {
        cudaStreamSynchronize(c->side);
        c->side_pending = false;
    }
}

static void p2p_release(rvl_ctx *c)
{
    for (void *p : c->opened) cudaIpcCloseMemHandle(p);
    c->opened.clear();
    dfree(c->d_xchg);
    dfree(c->d_peers);
    c->p2p = false;
    c->xchg_bytes = 0;
}

static RvlPeerXchg make_xchg(const rvl_ctx *c, int iter)
{
    RvlPeerXchg X;
    X.peers = c->p2p ? c->d_peers : nullptr;
    X.local = c->d_xchg;
    X.stamps_off = c->xchg_stamps_off;
    X.stamp = (c->epoch << 16) | (unsigned int)(iter + 1);
    X.parity = (c->seq0 + iter) & 1; // the slot alternates across searches too: reuse is ordered by the stamp chain
    X.err = c->d_xerr;
    return X;
}

extern "C" size_t rvl_candidate_bytes(const rvl_ctx *ctx)
{
    if (!ctx || ctx->T <= 0) return 0;
    return (size_t)ctx->T * (sizeof(RvlCandHead) + sizeof(uint64_t) * (size_t)ctx->words);
}

extern "C" int rvl_set_exchange_buffers(rvl_ctx *ctx, void *dev_send, void *dev_recv)
{
    GUARD();
    ctx->d_send = (unsigned char *)dev_send;
    ctx->d_recv = (unsigned char *)dev_recv;
    return RVL_OK;
}

static int ensure_results(rvl_ctx *ctx, int iters)
{
    int T = ctx->T;
    size_t need = (size_t)T * iters * (size_t)(ctx->F > 0 ? ctx->F : 1);
    if (need > ctx->scores_cap) {
        dfree(ctx->d_scores);
        CK(dalloc(&ctx->d_scores, need));
        ctx->scores_cap = need;
    }
    if (T > ctx->res_cap_T || iters > ctx->res_cap_iters || ctx->words > ctx->res_cap_words) {
        dfree(ctx->d_top); dfree(ctx->d_top_score); dfree(ctx->d_cmi_seed); dfree(ctx->d_ic0); dfree(ctx->d_seed_iter);
        CK(dalloc(&ctx->d_top, (size_t)T * iters));
        CK(dalloc(&ctx->d_top_score, (size_t)T * iters));
        CK(dalloc(&ctx->d_cmi_seed, (size_t)T * iters));
        CK(dalloc(&ctx->d_ic0, (size_t)T));
        CK(dalloc(&ctx->d_seed_iter, (size_t)T * iters * ctx->words));
        ctx->res_cap_T = T; ctx->res_cap_iters = iters; ctx->res_cap_words = ctx->words;
    }
    return RVL_OK;
}

// Device score layout is [iters][T][F]; `base` is one iteration's [T][F] slot.
static int launch_score(rvl_ctx *ctx, double *base)
{
    RvlParams P = make_params(ctx);
    if (ctx->F == 0) return RVL_OK; // k_advance already left every best slot at "none"
    const bool prof = ctx->profiling != 0;
    if (prof && ctx->score_events_used == ctx->score_events.size()) {
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        ctx->score_events.push_back({a, b});
    }
    if (ctx->work_dirty) CK(cudaMemsetAsync(ctx->d_work, 0, sizeof(unsigned long long), ctx->stream));
    ctx->work_dirty = true; // k_advance re-arms the counter; a repeated launch without it must do so here
    if (prof) CK(cudaEventRecord(ctx->score_events[ctx->score_events_used].first, ctx->stream));
    rvl_launch_score(&P, ctx->score_grid_x, ctx->score_warps, ctx->score_smem, ctx->d_bits, ctx->d_u, ctx->d_xc, ctx->d_tg, ctx->d_seed,
                     ctx->d_bcvtab, ctx->d_tt, base, ctx->d_part, ctx->d_work, prof ? ctx->d_ctr : nullptr,
                     ctx->dedup ? ctx->d_uniq : nullptr, ctx->U, ctx->stream);
    ctx->fill_pending = ctx->dedup && ctx->U < ctx->F;
    if (prof) CK(cudaEventRecord(ctx->score_events[ctx->score_events_used++].second, ctx->stream));
    ctx->launches++;
    CK(cudaGetLastError());
    return RVL_OK;
}

// Make the main stream wait for the side-stream seed ICs (before results are read or reused).
static int join_side(rvl_ctx *ctx)
{
    if (!ctx->side_pending) return RVL_OK;
    CK(cudaEventRecord(ctx->ev_join, ctx->side));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
    ctx->side_pending = false;
    return RVL_OK;
}

// A peer-exchange stamp that never arrived (k_advance's bounded wait) is an error, not a silent result.
static int check_xerr(rvl_ctx *ctx)
{
    int h = 0;
    CK(cudaMemcpy(&h, ctx->d_xerr, sizeof(int), cudaMemcpyDeviceToHost));
    if (h) {
        cudaMemset(ctx->d_xerr, 0, sizeof(int));
        return fail(ctx, RVL_ERR_CUDA, "peer exchange timed out: a rank did not deliver its candidate record");
    }
    return RVL_OK;
}

static int drain_score_events(rvl_ctx *ctx)
{
    for (size_t i = 0; i < ctx->search_events_used; i++) { // a k_search launch counts as one launch of the dominant kernel
        float ms = 0.f;
        cudaError_t e = cudaEventElapsedTime(&ms, ctx->search_events[i].first, ctx->search_events[i].second);
        if (e == cudaSuccess) { ctx->score_ms += ms; ctx->score_launches++; }
    }
    ctx->search_events_used = 0;
    for (size_t i = 0; i < ctx->score_events_used; i++) {
        float ms = 0.f;
        cudaError_t e = cudaEventElapsedTime(&ms, ctx->score_events[i].first, ctx->score_events[i].second);
        if (e == cudaSuccess) { ctx->score_ms += ms; ctx->score_launches++; }
    }
    ctx->score_events_used = 0;
    return RVL_OK;
}

extern "C" int rvl_search_begin(rvl_ctx *ctx, int iters)
{
    GUARD();
    int rc = ready(ctx);
    if (rc) return rc;
    if (iters < 1) return fail(ctx, RVL_ERR_ARG, "iters < 1");
    rc = join_side(ctx); // a previous search's seed ICs must land before buffers are reused
    if (rc) return rc;
    rc = ensure_results(ctx, iters);
    if (rc) return rc;
    ctx->iters = iters;
    size_t cb = rvl_candidate_bytes(ctx);
    if (ctx->world == 1 && !ctx->p2p && (!ctx->d_send || !ctx->d_recv)) {
        if (ctx->own_xchg_bytes < cb) {
            dfree(ctx->d_own_xchg);
            CK(dalloc(&ctx->d_own_xchg, cb));
            ctx->own_xchg_bytes = cb;
        }
        ctx->d_send = ctx->d_recv = ctx->d_own_xchg; // world of one: the record is its own gather
    }
    if (ctx->p2p && (ctx->xchg_world != ctx->world || ctx->xchg_T != ctx->T || ctx->xchg_words != ctx->words))
        return fail(ctx, RVL_ERR_STATE, "peer exchange was set up for another shape: call rvl_p2p_export/import again");
    if (!ctx->p2p && (!ctx->d_send || !ctx->d_recv))
        return fail(ctx, RVL_ERR_STATE, "world > 1 needs rvl_set_exchange_buffers or rvl_p2p_import");
    ctx->epoch = (ctx->epoch + 1) & 0x7fff;
    ctx->seq0 = ctx->xseq;
    ctx->xseq = (ctx->xseq + iters) & 0x3fffffff;
    RvlParams P = make_params(ctx);
    CK(cudaMemcpyAsync(ctx->d_seed, ctx->d_seed0, sizeof(uint64_t) * (size_t)ctx->T * ctx->words, cudaMemcpyDeviceToDevice, ctx->stream));
    rc = prep_iteration(ctx);
    if (rc) return rc;
    // cmi.seed.iter0 <- assoc(target, seed)  LIB:1833 -- off the critical path like the per-iteration seed ICs:
    // side stream, reading the immutable initial seeds
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    rvl_launch_assoc(&P, ctx->T, ctx->d_tg, nullptr, ctx->d_u, ctx->d_xc, ctx->d_seed0, ctx->d_bcvtab, ctx->d_ic0,
                     (size_t)ctx->words, 1, 0, ctx->side);
    ctx->side_pending = true;
    ctx->launches += 1;
    CK(cudaGetLastError());
    return RVL_OK;
}

extern "C" int rvl_iter_score(rvl_ctx *ctx, int iter)
{
    GUARD();
    if (iter < 0 || iter >= ctx->iters) return fail(ctx, RVL_ERR_ARG, "iter out of range");
    int rc = launch_score(ctx, ctx->d_scores + (size_t)iter * ctx->T * ctx->F);
    if (rc) return rc;
    RvlParams P2 = make_params(ctx);
    // (fusing this reduction into k_advance -- its `fuse_rank` path -- was measured 9 us per iteration
    //  slower on one GPU: 363 blocks then spin on the stamp that one 256-thread block has yet to write)
    RvlPeerXchg X = make_xchg(ctx, iter);
    rvl_launch_argmax(&P2, ctx->d_part, ctx->score_grid_x * ctx->score_warps, ctx->d_bits, ctx->d_send,
                      sizeof(RvlCandHead) + sizeof(uint64_t) * (size_t)ctx->words, &X, ctx->stream);
    ctx->launches++;
    CK(cudaGetLastError());
    return RVL_OK;
}

extern "C" int rvl_iter_merge(rvl_ctx *ctx, int iter)
{
    GUARD();
    if (iter < 0 || iter >= ctx->iters) return fail(ctx, RVL_ERR_ARG, "iter out of range");
    RvlParams P = make_params(ctx);
    // merge + seed state + x-axis table in one launch; seeds ping-pong between the two buffers
    RvlPeerXchg X = make_xchg(ctx, iter);
    rvl_launch_advance(&P, ctx->d_tg, ctx->d_u, ctx->d_seed, ctx->d_seed_alt, ctx->d_bcvtab, ctx->d_tt, ctx->d_part,
                       ctx->score_grid_x * ctx->score_warps, ctx->d_work, ctx->p2p ? ctx->d_xchg : ctx->d_recv,
                       sizeof(RvlCandHead) + sizeof(uint64_t) * (size_t)ctx->words, iter, ctx->iters, ctx->first_null,
                       ctx->d_top, ctx->d_top_score, ctx->d_seed_iter, 1, &X, ctx->stream);
    { uint64_t *tmp = ctx->d_seed; ctx->d_seed = ctx->d_seed_alt; ctx->d_seed_alt = tmp; }
    ctx->work_dirty = false;
    ctx->launches += 1;
    CK(cudaGetLastError());
    // cmi.seed[iter] <- assoc(target, seed) (LIB:2170) is off the critical path: side stream, reading
    // the seed_iter row k_advance just wrote (never modified again)
    int nreal = ctx->first_null > 0 ? ctx->first_null : ctx->T;
    CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
    rvl_launch_assoc(&P, nreal, ctx->d_tg, nullptr, ctx->d_u, ctx->d_xc, ctx->d_seed_iter + (size_t)iter * ctx->words,
                     ctx->d_bcvtab, ctx->d_cmi_seed + iter, (size_t)ctx->iters * ctx->words, (size_t)ctx->iters, 0,
                     ctx->side);
    ctx->launches += 1;
    ctx->side_pending = true;
    CK(cudaGetLastError());
    return RVL_OK;
}

extern "C" int rvl_synchronize(rvl_ctx *ctx)
{
    GUARD();
    int jrc = join_side(ctx);
    if (jrc) return jrc;
    CK(cudaStreamSynchronize(ctx->stream));
    drain_score_events(ctx);
    return check_xerr(ctx);
}

extern "C" int rvl_search_fetch(rvl_ctx *ctx, rvl_result *out)
{
    GUARD();
    if (!out) return fail(ctx, RVL_ERR_ARG, "result is NULL");
    int T = ctx->T, iters = ctx->iters, words = ctx->words, n = ctx->n;
    if (iters < 1) return fail(ctx, RVL_ERR_STATE, "no search has run");
    int jrc = join_side(ctx);
    if (jrc) return jrc;
    CK(cudaStreamSynchronize(ctx->stream));
    drain_score_events(ctx);
    jrc = check_xerr(ctx);
    if (jrc) return jrc;
    if (ctx->fill_pending && ctx->F > 0) { // duplicate rows get their group's score, once, for all iterations
        rvl_launch_fill_duplicates(ctx->d_scores, ctx->F, (int64_t)iters * T, ctx->d_rep, ctx->stream);
        ctx->launches++;
        ctx->fill_pending = false;
    }
    if (out->cmi && ctx->F > 0) {
        // device layout [iters][T][F] -> caller layout [T][iters][F]
        for (int t = 0; t < T; t++)
            CK(cudaMemcpy2DAsync(out->cmi + (size_t)t * iters * ctx->F, sizeof(double) * (size_t)ctx->F,
                                 ctx->d_scores + (size_t)t * ctx->F, sizeof(double) * (size_t)T * ctx->F,
                                 sizeof(double) * (size_t)ctx->F, (size_t)iters, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (out->top) CK(cudaMemcpyAsync(out->top, ctx->d_top, sizeof(int64_t) * (size_t)T * iters, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->top_score) CK(cudaMemcpyAsync(out->top_score, ctx->d_top_score, sizeof(double) * (size_t)T * iters, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->cmi_seed) CK(cudaMemcpyAsync(out->cmi_seed, ctx->d_cmi_seed, sizeof(double) * (size_t)T * iters, cudaMemcpyDeviceToHost, ctx->stream));
    if (out->ic_seed0) CK(cudaMemcpyAsync(out->ic_seed0, ctx->d_ic0, sizeof(double) * (size_t)T, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<uint64_t> sb;
    if (out->seed_iter) {
        sb.resize((size_t)T * iters * words);
        CK(cudaMemcpyAsync(sb.data(), ctx->d_seed_iter, sizeof(uint64_t) * sb.size(), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (out->seed_iter) {
        for (size_t r = 0; r < (size_t)T * iters; r++)
            for (int s = 0; s < n; s++) out->seed_iter[r * n + s] = (uint8_t)((sb[r * words + (s >> 6)] >> (s & 63)) & 1ull);
    }
    return RVL_OK;
}

// world == 1: the last block of k_search publishes into a local buffer laid out like a peer exchange buffer of one rank.
static int ensure_self_xchg(rvl_ctx *ctx)
{
    if (ctx->d_self_xchg && ctx->self_T == ctx->T && ctx->self_words == ctx->words) return RVL_OK;
    dfree(ctx->d_self_xchg);
    dfree(ctx->d_self_peers);
    size_t rec = sizeof(RvlCandHead) + sizeof(uint64_t) * (size_t)ctx->words;
    size_t recs = 2 * (size_t)ctx->T * rec;
    ctx->self_stamps_off = (recs + 255) & ~(size_t)255;
    ctx->self_xchg_bytes = ctx->self_stamps_off + 2 * (size_t)ctx->T * sizeof(unsigned int);
    CK(cudaMalloc((void **)&ctx->d_self_xchg, ctx->self_xchg_bytes));
    CK(cudaMemset(ctx->d_self_xchg, 0, ctx->self_xchg_bytes));
    CK(dalloc(&ctx->d_self_peers, 1));
    CK(cudaMemcpy(ctx->d_self_peers, &ctx->d_self_xchg, sizeof(unsigned char *), cudaMemcpyHostToDevice));
    ctx->self_T = ctx->T; ctx->self_words = ctx->words;
    return RVL_OK;
}

// Can this context run the search as one k_search launch?  (cooperative launch available; with world > 1 the candidate
// exchange must be the peer-memory one -- an NCCL all-gather cannot be called from inside a kernel.)
static bool fused_possible(rvl_ctx *ctx)
{
    if (!ctx->fused || ctx->F < 0 || ctx->T <= 0) return false;
    if (ctx->world > 1 && !ctx->p2p) return false;
    int warps = 0;
    size_t smem = plan_score(ctx, &warps);
    if (!smem) return false;
    if (ctx->search_grid_x < 0 || smem != ctx->search_smem || warps != ctx->search_warps) {
        int grid = 0;
        if (rvl_search_configure(smem, warps, ctx->device, &grid) != cudaSuccess) { cudaGetLastError(); grid = 0; }
        ctx->search_grid_x = grid;
        ctx->search_smem = smem;
        ctx->search_warps = warps;
    }
    return ctx->search_grid_x > 0;
}

// The whole search enqueued as ONE launch (k_search) + one batched launch of the seed ICs.  Asynchronous.
static int search_enqueue_fused(rvl_ctx *ctx, int iters)
{
    int rc = rvl_search_begin(ctx, iters); // seeds, initial seed state + table, assoc(target, initial seed), epoch / slot parity
    if (rc) return rc;
    if (ctx->world == 1 && !ctx->p2p) {
        rc = ensure_self_xchg(ctx);
        if (rc) return rc;
    }
    if (ctx->d_phase_dbg && iters > 64) return fail(ctx, RVL_ERR_ARG, "phase timing is armed: iters <= 64");
    if (iters > ctx->sctr_cap) {
        dfree(ctx->d_swork); dfree(ctx->d_sdone);
        ctx->sctr_cap = 0;
        CK(dalloc(&ctx->d_swork, (size_t)iters));
        CK(dalloc(&ctx->d_sdone, 2 * (size_t)iters + 1));
        ctx->sctr_cap = iters;
    }
    CK(cudaMemsetAsync(ctx->d_swork, 0, sizeof(unsigned long long) * (size_t)iters, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_sdone, 0, sizeof(unsigned int) * (2 * (size_t)iters + 1), ctx->stream));
    // the k_search grid equals the k_score grid the best slots were sized for (same block shape and shared memory)
    if ((size_t)ctx->T * (size_t)ctx->search_grid_x * ctx->search_warps > ctx->part_cap)
        return fail(ctx, RVL_ERR_STATE, "internal: k_search grid exceeds the best-slot buffer");
    RvlSearchArgs a;
    memset(&a, 0, sizeof a);
    a.P = make_params(ctx);
    a.bits = ctx->d_bits; a.u_all = ctx->d_u; a.xc_all = ctx->d_xc; a.tg_all = ctx->d_tg;
    a.seed_a = ctx->d_seed; a.seed_b = ctx->d_seed_alt;
    a.bcvtab = ctx->d_bcvtab; a.tt = ctx->d_tt; a.scores = ctx->d_scores; a.part = ctx->d_part;
    a.work = ctx->d_swork; a.done = ctx->d_sdone; a.bar = ctx->d_sdone + iters; a.need = ctx->d_sdone + iters + 1;
    a.ctr = ctx->profiling ? ctx->d_ctr : nullptr;
    a.uniq = ctx->dedup ? ctx->d_uniq : nullptr; a.U = ctx->U;
    a.rec_bytes = sizeof(RvlCandHead) + sizeof(uint64_t) * (size_t)ctx->words;
    a.iters = iters; a.first_null = ctx->first_null; a.nreal = ctx->first_null > 0 ? ctx->first_null : ctx->T;
    a.top = ctx->d_top; a.top_score = ctx->d_top_score; a.seed_iter = ctx->d_seed_iter;
    if (ctx->p2p) a.X = make_xchg(ctx, 0);
    else {
        a.X.peers = ctx->d_self_peers; a.X.local = ctx->d_self_xchg; a.X.stamps_off = ctx->self_stamps_off;
        a.X.err = ctx->d_xerr; a.X.stamp = 0; a.X.parity = 0;
    }
    a.epoch = ctx->epoch; a.seq0 = ctx->seq0;
    a.dbg = ctx->d_phase_dbg;
    const bool prof = ctx->profiling != 0;
    if (prof && ctx->search_events_used == ctx->search_events.size()) {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        ctx->search_events.push_back({e0, e1});
    }
    if (prof) CK(cudaEventRecord(ctx->search_events[ctx->search_events_used].first, ctx->stream));
    if (ctx->F > 0 || ctx->world > 1) CK(rvl_launch_search(&a, ctx->search_grid_x, ctx->search_warps, ctx->search_smem, ctx->stream));
    if (prof) CK(cudaEventRecord(ctx->search_events[ctx->search_events_used++].second, ctx->stream));
    ctx->launches++;
    if (iters & 1) { uint64_t *tmp = ctx->d_seed; ctx->d_seed = ctx->d_seed_alt; ctx->d_seed_alt = tmp; }
    ctx->work_dirty = true;
    ctx->fill_pending = ctx->dedup && ctx->U < ctx->F;
    // cmi.seed[iter] <- assoc(target, seed) (LIB:2170) for every iteration at once, from the seed.iter rows k_search wrote
    RvlParams P = a.P;
    rvl_launch_assoc(&P, a.nreal * iters, ctx->d_tg, nullptr, ctx->d_u, ctx->d_xc, ctx->d_seed_iter, ctx->d_bcvtab, ctx->d_cmi_seed,
                     (size_t)ctx->words, 1, iters, ctx->stream);
    ctx->launches++;
    CK(cudaGetLastError());
    return RVL_OK;
}

// The whole search of this shard, enqueued without waiting: one persistent launch (k_search) when possible, the
// split-phase launches otherwise.  With world > 1 this needs the peer exchange (rvl_p2p_import / rvl_group_connect)
// and every rank must call it; rvl_search_fetch / rvl_synchronize wait for the result.
extern "C" int rvl_search_run(rvl_ctx *ctx, int iters)
{
    GUARD();
    int rc = ready(ctx);
    if (rc) return rc;
    if (iters < 1) return fail(ctx, RVL_ERR_ARG, "iters < 1");
    if (ctx->world > 1 && !ctx->p2p)
        return fail(ctx, RVL_ERR_STATE, "rvl_search_run with world > 1 needs the peer exchange (rvl_p2p_import / rvl_group_connect); "
                                        "use rvl_iter_score / rvl_iter_merge around your own all-gather otherwise");
    if (fused_possible(ctx) && (ctx->F > 0 || ctx->world > 1)) return search_enqueue_fused(ctx, iters);
    rc = rvl_search_begin(ctx, iters);
    if (rc) return rc;
    for (int it = 0; it < iters; it++) { // enqueued back to back: no host round-trip inside the loop
        rc = rvl_iter_score(ctx, it);
        if (rc) return rc;
        rc = rvl_iter_merge(ctx, it);
        if (rc) return rc;
    }
    return RVL_OK;
}

// Development aid (tools/time_phases.py): out[iters][8] nanosecond stamps of the last fused search's phases -- [0] block 0
// starts scoring, [1] block 0 done scoring, [2] every block done, [3] records + stamps published, [4] block 0's merge
// tasks done, [5] grid barrier passed.  First call (out == NULL) arms the recording for the following searches.
extern "C" int rvl_debug_phase_times(rvl_ctx *ctx, unsigned long long *out, int iters)
{
    GUARD();
    if (!ctx->d_phase_dbg) {
        CK(dalloc(&ctx->d_phase_dbg, (size_t)64 * 8));
        CK(cudaMemset(ctx->d_phase_dbg, 0, sizeof(unsigned long long) * 64 * 8));
    }
    if (!out) return RVL_OK;
    if (iters < 1 || iters > 64) return fail(ctx, RVL_ERR_ARG, "iters must be in [1, 64]");
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, ctx->d_phase_dbg, sizeof(unsigned long long) * 8 * (size_t)iters, cudaMemcpyDeviceToHost));
    return RVL_OK;
}

extern "C" int rvl_set_fused(rvl_ctx *ctx, int on)
{
    GUARD();
    ctx->fused = on ? 1 : 0;
    return RVL_OK;
}

extern "C" int rvl_search(rvl_ctx *ctx, int iters, rvl_result *out)
{
    GUARD();
    if (ctx->world != 1) return fail(ctx, RVL_ERR_STATE, "rvl_search is the single-shard driver; use rvl_search_run or the split-phase calls when world > 1");
    int rc = rvl_search_run(ctx, iters);
    if (rc) return rc;
    if (out) return rvl_search_fetch(ctx, out);
    return rvl_synchronize(ctx);
}

extern "C" int rvl_score_once(rvl_ctx *ctx, double *out_scores)
{
    GUARD();
    int rc = ready(ctx);
    if (rc) return rc;
    rc = ensure_results(ctx, 1);
    if (rc) return rc;
    RvlParams P = make_params(ctx);
    CK(cudaMemcpyAsync(ctx->d_seed, ctx->d_seed0, sizeof(uint64_t) * (size_t)ctx->T * ctx->words, cudaMemcpyDeviceToDevice, ctx->stream));
    rc = prep_iteration(ctx);
    if (rc) return rc;
    rc = launch_score(ctx, ctx->d_scores);
    if (rc) return rc;
    if (ctx->fill_pending && ctx->F > 0) {
        rvl_launch_fill_duplicates(ctx->d_scores, ctx->F, (int64_t)ctx->T, ctx->d_rep, ctx->stream);
        ctx->launches++;
        ctx->fill_pending = false;
    }
    if (out_scores && ctx->F > 0)
        CK(cudaMemcpyAsync(out_scores, ctx->d_scores, sizeof(double) * (size_t)ctx->T * ctx->F, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    drain_score_events(ctx);
    return RVL_OK;
}

extern "C" int rvl_assoc(rvl_ctx *ctx, int64_t target_index, const uint8_t *y, int64_t n, double *out_ic)
{
    GUARD();
    if (ctx->T <= 0) return fail(ctx, RVL_ERR_STATE, "targets not set");
    if (!y || !out_ic || n != ctx->n || target_index < 0 || target_index >= ctx->T) return fail(ctx, RVL_ERR_ARG, "bad rvl_assoc arguments");
    std::vector<uint64_t> packed(ctx->words, 0ull);
    for (int64_t i = 0; i < n; i++) {
        if (y[i] == 1) packed[i >> 6] |= (1ull << (i & 63));
        else if (y[i] != 0) return fail(ctx, RVL_ERR_NOT_BINARY, "y entry %lld is not 0 or 1", (long long)i);
    }
    uint64_t *d_y = nullptr;
    double *d_o = nullptr;
    int *d_t = nullptr;
    CK(dalloc(&d_y, (size_t)ctx->words));
    CK(dalloc(&d_o, 1));
    CK(dalloc(&d_t, 1));
    int ti = (int)target_index;
    RvlParams P = make_params(ctx);
    cudaMemcpyAsync(d_y, packed.data(), sizeof(uint64_t) * ctx->words, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_t, &ti, sizeof(int), cudaMemcpyHostToDevice, ctx->stream);
    rvl_launch_assoc(&P, 1, ctx->d_tg, d_t, ctx->d_u, ctx->d_xc, d_y, ctx->d_bcvtab, d_o, (size_t)ctx->words, 1, 0, ctx->stream);
    ctx->launches++;
    cudaMemcpyAsync(out_ic, d_o, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_y); cudaFree(d_o); cudaFree(d_t);
    if (e != cudaSuccess) return fail(ctx, RVL_ERR_CUDA, "rvl_assoc: %s", cudaGetErrorString(e));
    return RVL_OK;
}

struct rvl_not_nan {
    __host__ __device__ bool operator()(double v) const { return v == v; }
};

extern "C" int rvl_pvalues(rvl_ctx *ctx, const double *cmi, int64_t F, const double *cmi_rand, int64_t n_rand, double *out_pval)
{
    GUARD();
    if (!cmi || !cmi_rand || !out_pval || F < 1 || n_rand < 1) return fail(ctx, RVL_ERR_ARG, "bad rvl_pvalues arguments");
    double *d_c = nullptr, *d_r = nullptr, *d_p = nullptr;
    CK(dalloc(&d_c, (size_t)F));
    CK(dalloc(&d_r, (size_t)n_rand));
    CK(dalloc(&d_p, (size_t)F));
    cudaMemcpyAsync(d_c, cmi, sizeof(double) * F, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_r, cmi_rand, sizeof(double) * n_rand, cudaMemcpyHostToDevice, ctx->stream);
    int64_t n_valid = 0;
    try {
        auto pol = thrust::cuda::par.on(ctx->stream);
        thrust::device_ptr<double> r(d_r);
        auto mid = thrust::partition(pol, r, r + n_rand, rvl_not_nan());
        n_valid = mid - r;
        thrust::sort(pol, r, r + n_valid);
    } catch (...) {
        cudaFree(d_c); cudaFree(d_r); cudaFree(d_p);
        return fail(ctx, RVL_ERR_CUDA, "device sort failed");
    }
    rvl_launch_pvalues(d_c, F, d_r, n_valid, n_rand, d_p, ctx->stream);
    ctx->launches += 3;
    cudaMemcpyAsync(out_pval, d_p, sizeof(double) * F, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_c); cudaFree(d_r); cudaFree(d_p);
    if (e != cudaSuccess) return fail(ctx, RVL_ERR_CUDA, "rvl_pvalues: %s", cudaGetErrorString(e));
    return RVL_OK;
}

// p-values of one iteration of the last search from the scores ALREADY in HBM (LIB:1868-1869): target 0's scores are
// cmi[, iter], the permutation-null targets' scores are cmi.rand -- contiguous in the [iters][T][F] score buffer, so
// nothing is downloaded or uploaded except the F p-values.  The null scores are sorted in a scratch copy.
extern "C" int rvl_pvalues_iter(rvl_ctx *ctx, int iter, double *out_pval)
{
    GUARD();
    if (!out_pval) return fail(ctx, RVL_ERR_ARG, "out_pval is NULL");
    if (ctx->iters < 1 || iter < 0 || iter >= ctx->iters) return fail(ctx, RVL_ERR_STATE, "no search result for iteration %d", iter);
    if (ctx->first_null != 1 || ctx->T < 2) return fail(ctx, RVL_ERR_STATE, "the last search had no permutation-null targets (rvl_set_null_targets)");
    if (ctx->world != 1) return fail(ctx, RVL_ERR_STATE, "rvl_pvalues_iter needs the null scores of every row: single shard only (gather and use rvl_pvalues)");
    const int64_t F = ctx->F, n_rand = (int64_t)(ctx->T - 1) * F;
    if (F < 1) return RVL_OK;
    int jrc = join_side(ctx);
    if (jrc) return jrc;
    if (ctx->fill_pending) { // duplicate rows get their group's score first (they count in the null distribution too)
        rvl_launch_fill_duplicates(ctx->d_scores, F, (int64_t)ctx->iters * ctx->T, ctx->d_rep, ctx->stream);
        ctx->launches++;
        ctx->fill_pending = false;
    }
    const double *base = ctx->d_scores + (size_t)iter * ctx->T * F;
    double *d_r = nullptr, *d_p = nullptr;
    CK(dalloc(&d_r, (size_t)n_rand));
    CK(dalloc(&d_p, (size_t)F));
    int64_t n_valid = 0;
    cudaError_t e = cudaMemcpyAsync(d_r, base + F, sizeof(double) * (size_t)n_rand, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) {
        try {
            rvl_pool_alloc alloc{ctx};
            auto pol = thrust::cuda::par(alloc).on(ctx->stream);
            thrust::device_ptr<double> r(d_r);
            auto mid = thrust::partition(pol, r, r + n_rand, rvl_not_nan());
            n_valid = mid - r;
            thrust::sort(pol, r, r + n_valid);
        } catch (...) {
            cudaFree(d_r); cudaFree(d_p);
            return fail(ctx, RVL_ERR_CUDA, "device sort failed");
        }
        rvl_launch_pvalues(base, F, d_r, n_valid, n_rand, d_p, ctx->stream);
        ctx->launches += 3;
        e = cudaMemcpyAsync(out_pval, d_p, sizeof(double) * (size_t)F, cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_r); cudaFree(d_p);
    if (e != cudaSuccess || e2 != cudaSuccess)
        return fail(ctx, RVL_ERR_CUDA, "rvl_pvalues_iter: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return RVL_OK;
}

// R's order(cmi[, iter], decreasing = !negative) restricted to its first k entries (the n.markers rows REVEALER.v1
// displays, LIB:1858-1864), on the scores resident in HBM: scores become sort keys that are ascending in R's order
// (NaN last, -0 == +0), a stable radix sort keeps the lower row first among ties.
struct rvl_order_key {
    const double *v;
    int negative;
    __host__ __device__ unsigned long long operator()(int64_t i) const
    {
        double s = v[i];
        if (s != s) return ~0ull;                      // NaN / NA: last
        if (s == 0.0) s = 0.0;                         // -0 and +0 tie
        unsigned long long u;
        memcpy(&u, &s, sizeof u);
        u = (u >> 63) ? ~u : (u | 0x8000000000000000ull); // ascending in s
        return negative ? u : ~u - 1ull;               // decreasing unless target.match == "negative"; never the NaN key
    }
};

extern "C" int rvl_topk(rvl_ctx *ctx, int iter, int64_t target_index, int64_t k, int64_t *out_rows, double *out_scores)
{
    GUARD();
    if (ctx->iters < 1 || iter < 0 || iter >= ctx->iters) return fail(ctx, RVL_ERR_STATE, "no search result for iteration %d", iter);
    if (target_index < 0 || target_index >= ctx->T || k < 0 || (k > 0 && !out_rows)) return fail(ctx, RVL_ERR_ARG, "bad rvl_topk arguments");
    const int64_t F = ctx->F;
    if (k > F) k = F;
    if (k == 0) return RVL_OK;
    int jrc = join_side(ctx);
    if (jrc) return jrc;
    if (ctx->fill_pending) {
        rvl_launch_fill_duplicates(ctx->d_scores, F, (int64_t)ctx->iters * ctx->T, ctx->d_rep, ctx->stream);
        ctx->launches++;
        ctx->fill_pending = false;
    }
    const double *col = ctx->d_scores + ((size_t)iter * ctx->T + (size_t)target_index) * F;
    unsigned long long *d_key = nullptr;
    int64_t *d_idx = nullptr;
    double *d_top = nullptr;
    CK(dalloc(&d_key, (size_t)F));
    CK(dalloc(&d_idx, (size_t)F));
    CK(dalloc(&d_top, (size_t)k));
    int rc = RVL_OK;
    try {
        rvl_pool_alloc alloc{ctx};
        auto pol = thrust::cuda::par(alloc).on(ctx->stream);
        thrust::device_ptr<unsigned long long> key(d_key);
        thrust::device_ptr<int64_t> idx(d_idx);
        thrust::copy(pol, thrust::counting_iterator<int64_t>(0), thrust::counting_iterator<int64_t>(F), idx);
        thrust::transform(pol, idx, idx + F, key, rvl_order_key{col, ctx->opts.negative});
        thrust::stable_sort_by_key(pol, key, key + F, idx);
        thrust::gather(pol, idx, idx + k, thrust::device_ptr<const double>(col), thrust::device_ptr<double>(d_top));
        ctx->launches += 4;
    } catch (...) {
        rc = fail(ctx, RVL_ERR_CUDA, "device sort failed");
    }
    cudaError_t e = cudaSuccess;
    if (!rc) {
        e = cudaMemcpyAsync(out_rows, d_idx, sizeof(int64_t) * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && out_scores) e = cudaMemcpyAsync(out_scores, d_top, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream);
    }
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_key); cudaFree(d_idx); cudaFree(d_top);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess) return fail(ctx, RVL_ERR_CUDA, "rvl_topk: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    for (int64_t i = 0; i < k; i++) out_rows[i] += ctx->row_offset; // global rows, like rvl_result.top
    return RVL_OK;
}

// ---- introspection ------------------------------------------------------------------------------

extern "C" int rvl_get_target_bandwidth(rvl_ctx *ctx, int64_t t, double *out)
{
    GUARD();
    if (ctx->T <= 0) return fail(ctx, RVL_ERR_STATE, "targets not set");
    if (!out || t < 0 || t >= ctx->T) return fail(ctx, RVL_ERR_ARG, "bad target index");
    RvlTarget tg;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(&tg, ctx->d_tg + t, sizeof(RvlTarget), cudaMemcpyDeviceToHost));
    *out = tg.bcv;
    return RVL_OK;
}

extern "C" int rvl_get_binary_bandwidth_table(rvl_ctx *ctx, double *out)
{
    GUARD();
    if (ctx->T <= 0) return fail(ctx, RVL_ERR_STATE, "targets not set");
    if (!out) return fail(ctx, RVL_ERR_ARG, "out is NULL");
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(out, ctx->d_bcvtab, sizeof(double) * ((size_t)ctx->n + 1), cudaMemcpyDeviceToHost));
    return RVL_OK;
}

extern "C" int rvl_get_matrix_dims(rvl_ctx *ctx, int64_t *out_rows, int64_t *out_samples)
{
    GUARD();
    if (ctx->F < 0) return fail(ctx, RVL_ERR_STATE, "feature matrix not loaded");
    if (out_rows) *out_rows = ctx->F;
    if (out_samples) *out_samples = ctx->n;
    return RVL_OK;
}

extern "C" int rvl_get_row_counts(rvl_ctx *ctx, int32_t *out)
{
    GUARD();
    if (ctx->F < 0) return fail(ctx, RVL_ERR_STATE, "feature matrix not loaded");
    if (!out) return fail(ctx, RVL_ERR_ARG, "out is NULL");
    if (ctx->F == 0) return RVL_OK;
    int32_t *d = nullptr;
    CK(dalloc(&d, (size_t)ctx->F));
    rvl_launch_row_counts(ctx->d_bits, ctx->F, ctx->words, d, ctx->stream);
    ctx->launches++;
    cudaMemcpyAsync(out, d, sizeof(int32_t) * ctx->F, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, RVL_ERR_CUDA, "row counts: %s", cudaGetErrorString(e));
    return RVL_OK;
}

extern "C" int64_t rvl_launch_count(const rvl_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int rvl_score_kernel_time(rvl_ctx *ctx, double *out_ms, int64_t *out_launches, int reset)
{
    GUARD();
    CK(cudaStreamSynchronize(ctx->stream));
    drain_score_events(ctx);
    if (out_ms) *out_ms = ctx->score_ms;
    if (out_launches) *out_launches = ctx->score_launches;
    if (reset) { ctx->score_ms = 0.0; ctx->score_launches = 0; }
    return RVL_OK;
}

extern "C" int rvl_get_counters_ex(rvl_ctx *ctx, uint64_t *out, int n_out, int reset)
{
    GUARD();
    if (!out || n_out < 1 || n_out > 8) return fail(ctx, RVL_ERR_ARG, "bad counter request");
    CK(cudaStreamSynchronize(ctx->stream));
    unsigned long long h[8];
    CK(cudaMemcpy(h, ctx->d_ctr, sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_out; i++) out[i] = (uint64_t)h[i];
    if (reset) CK(cudaMemset(ctx->d_ctr, 0, sizeof h));
    return RVL_OK;
}

extern "C" int rvl_get_counters(rvl_ctx *ctx, uint64_t *out4, int reset) { return rvl_get_counters_ex(ctx, out4, 4, reset); }

extern "C" int rvl_measure_fp64_peak(rvl_ctx *ctx, double *out_tflops)
{
    GUARD();
    if (!out_tflops) return fail(ctx, RVL_ERR_ARG, "out is NULL");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    double *d = nullptr;
    CK(dalloc(&d, (size_t)blocks * threads));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        rvl_launch_fp64_peak(blocks, threads, iters, d, ctx->stream);
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        double tf = 2.0 * 8.0 * (double)iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
        ctx->launches++;
    }
    cudaFree(d);
    *out_tflops = best;
    return RVL_OK;
}

extern "C" int rvl_set_profiling(rvl_ctx *ctx, int on)
{
    GUARD();
    CK(cudaStreamSynchronize(ctx->stream));
    drain_score_events(ctx);
    ctx->profiling = on ? 1 : 0;
    return RVL_OK;
}

extern "C" int rvl_set_dense_x(rvl_ctx *ctx, int on)
{
    GUARD();
    ctx->dense_x = on ? 1 : 0;
    return RVL_OK;
}

extern "C" int rvl_set_prune_theta(rvl_ctx *ctx, double theta)
{
    GUARD();
    if (!(theta >= 0.0) || theta > 1e-9) return fail(ctx, RVL_ERR_ARG, "prune theta must be in [0, 1e-9]");
    ctx->prune_theta = theta;
    return RVL_OK;
}

extern "C" int rvl_set_dedup(rvl_ctx *ctx, int on)
{
    GUARD();
    ctx->dedup = on ? 1 : 0;
    return RVL_OK;
}

extern "C" int64_t rvl_unique_rows(const rvl_ctx *ctx) { return ctx ? ctx->U : 0; }

// ---- peer-memory exchange (CUDA IPC over NVLink) ---------------------------------------------------

extern "C" int rvl_p2p_export(rvl_ctx *ctx, unsigned char *handle64)
{
    GUARD();
    if (!handle64) return fail(ctx, RVL_ERR_ARG, "handle is NULL");
    if (ctx->T <= 0 || ctx->words <= 0) return fail(ctx, RVL_ERR_STATE, "set the matrix/targets before rvl_p2p_export");
    if (ctx->world < 2) return fail(ctx, RVL_ERR_STATE, "rvl_p2p_export needs rvl_set_shard with world > 1");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaStreamSynchronize(ctx->stream));
    p2p_release(ctx);
    size_t rec = sizeof(RvlCandHead) + sizeof(uint64_t) * (size_t)ctx->words;
    size_t recs = 2 * (size_t)ctx->world * ctx->T * rec;
    ctx->xchg_stamps_off = (recs + 255) & ~(size_t)255;
    ctx->xchg_bytes = ctx->xchg_stamps_off + 2 * (size_t)ctx->world * ctx->T * sizeof(unsigned int);
    CK(cudaMalloc((void **)&ctx->d_xchg, ctx->xchg_bytes));
    CK(cudaMemset(ctx->d_xchg, 0, ctx->xchg_bytes));
    ctx->xchg_world = ctx->world; ctx->xchg_T = ctx->T; ctx->xchg_words = ctx->words;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->d_xchg));
    memcpy(handle64, &h, 64);
    return RVL_OK;
}

extern "C" int rvl_p2p_import(rvl_ctx *ctx, const unsigned char *handles, int world)
{
    GUARD();
    if (!handles || world != ctx->world || !ctx->d_xchg) return fail(ctx, RVL_ERR_ARG, "rvl_p2p_import: export first; world must match rvl_set_shard");
    std::vector<unsigned char *> ptrs(world, nullptr);
    for (int r = 0; r < world; r++) {
        if (r == ctx->rank) { ptrs[r] = ctx->d_xchg; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (void *q : ctx->opened) cudaIpcCloseMemHandle(q);
            ctx->opened.clear();
            return fail(ctx, RVL_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
        ctx->opened.push_back(p);
        ptrs[r] = (unsigned char *)p;
    }
    dfree(ctx->d_peers);
    CK(dalloc(&ctx->d_peers, (size_t)world));
    CK(cudaMemcpy(ctx->d_peers, ptrs.data(), sizeof(unsigned char *) * world, cudaMemcpyHostToDevice));
    ctx->p2p = true;
    ctx->epoch = 0; // every rank (re)connects at the same point of its history: stamps and slot parity start over together,
    ctx->xseq = 0;  // whatever searches this context ran on its own before
    return RVL_OK;
}

extern "C" int rvl_p2p_disable(rvl_ctx *ctx)
{
    GUARD();
    CK(cudaStreamSynchronize(ctx->stream));
    p2p_release(ctx);
    return RVL_OK;
}

// ---- several shards driven by ONE host thread (the single-process form an R session uses) ------------
// ctxs[i] must be rank i of n (rvl_set_shard) on n distinct devices, with matrix, targets and seed set.
// Peer access replaces CUDA IPC: every context's exchange buffer is directly addressable by the others.

extern "C" int rvl_group_connect(rvl_ctx **ctxs, int n)
{
    if (!ctxs || n < 1) return RVL_ERR_ARG;
    for (int i = 0; i < n; i++) {
        rvl_ctx *ctx = ctxs[i];
        if (!ctx) return RVL_ERR_ARG;
        if (ctx->world != n || ctx->rank != i) return fail(ctx, RVL_ERR_ARG, "rvl_group_connect: context %d must be rank %d of %d (rvl_set_shard)", i, i, n);
        if (ctx->T <= 0 || ctx->words <= 0) return fail(ctx, RVL_ERR_STATE, "set the matrix/targets before rvl_group_connect");
        for (int j = 0; j < i; j++)
            if (ctxs[j]->device == ctx->device) return fail(ctx, RVL_ERR_ARG, "rvl_group_connect: contexts %d and %d share device %d", j, i, ctx->device);
        if (ctx->T != ctxs[0]->T || ctx->words != ctxs[0]->words) return fail(ctx, RVL_ERR_ARG, "rvl_group_connect: targets / sample count differ between contexts");
    }
    if (n == 1) return RVL_OK;
    std::vector<unsigned char *> bufs(n, nullptr);
    for (int i = 0; i < n; i++) {
        rvl_ctx *ctx = ctxs[i];
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamSynchronize(ctx->stream));
        p2p_release(ctx);
        for (int j = 0; j < n; j++) {
            if (j == i) continue;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, ctx->device, ctxs[j]->device));
            if (!can) return fail(ctx, RVL_ERR_CUDA, "device %d cannot access device %d (no NVLink/PCIe peer path)", ctx->device, ctxs[j]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[j]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, RVL_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError(); // clear "already enabled"
        }
        size_t rec = sizeof(RvlCandHead) + sizeof(uint64_t) * (size_t)ctx->words;
        size_t recs = 2 * (size_t)n * ctx->T * rec;
        ctx->xchg_stamps_off = (recs + 255) & ~(size_t)255;
        ctx->xchg_bytes = ctx->xchg_stamps_off + 2 * (size_t)n * ctx->T * sizeof(unsigned int);
        CK(cudaMalloc((void **)&ctx->d_xchg, ctx->xchg_bytes));
        CK(cudaMemset(ctx->d_xchg, 0, ctx->xchg_bytes));
        ctx->xchg_world = n; ctx->xchg_T = ctx->T; ctx->xchg_words = ctx->words;
        bufs[i] = ctx->d_xchg;
    }
    for (int i = 0; i < n; i++) {
        rvl_ctx *ctx = ctxs[i];
        CK(cudaSetDevice(ctx->device));
        CK(dalloc(&ctx->d_peers, (size_t)n));
        CK(cudaMemcpy(ctx->d_peers, bufs.data(), sizeof(unsigned char *) * n, cudaMemcpyHostToDevice));
        ctx->p2p = true;
        ctx->epoch = 0;
        ctx->xseq = 0;
    }
    return RVL_OK;
}

// The whole search on every shard: all launches are asynchronous, so one host thread keeps n GPUs busy;
// the shards meet only through the stamped peer exchange inside k_argmax / k_advance.
extern "C" int rvl_group_search(rvl_ctx **ctxs, int n, int iters)
{
    if (!ctxs || n < 1) return RVL_ERR_ARG;
    bool fused = true;
    for (int i = 0; i < n; i++) {
        if (n > 1 && !ctxs[i]->p2p) return fail(ctxs[i], RVL_ERR_STATE, "rvl_group_search: call rvl_group_connect first");
        cudaSetDevice(ctxs[i]->device);
        fused = fused && fused_possible(ctxs[i]);
    }
    if (fused) { // one persistent launch per device; they meet through the stamped peer exchange inside k_search
        for (int i = 0; i < n; i++) {
            int rc = rvl_search_run(ctxs[i], iters);
            if (rc) return rc;
        }
    } else {
        for (int i = 0; i < n; i++) {
            int rc = rvl_search_begin(ctxs[i], iters);
            if (rc) return rc;
        }
        for (int it = 0; it < iters; it++) {
            for (int i = 0; i < n; i++) {
                int rc = rvl_iter_score(ctxs[i], it);
                if (rc) return rc;
            }
            for (int i = 0; i < n; i++) {
                int rc = rvl_iter_merge(ctxs[i], it);
                if (rc) return rc;
            }
        }
    }
    for (int i = 0; i < n; i++) {
        int rc = rvl_synchronize(ctxs[i]);
        if (rc) return rc;
    }
    return RVL_OK;
}
