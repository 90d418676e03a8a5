/*
 * revealer_b200.h -- C ABI of the B200-native REVEALER conditional-mutual-information search.
 *
 * The reference (genepattern/REVEALER, REVEALER_library.v5C.R, "LIB" below; live copy-B line
 * numbers) has NO native interface: its hot path is three R-level seams inside REVEALER.v1.
 * Every entry point here states which seam it replaces, so that an R maintainer can bind it
 * with one .Call shim (revealer_b200/r/r_shim.c).
 *
 *   scoring loop      for (k in 1:nrow(m.2)) cmi[k,iter] <- cond.assoc(target, m.2[k,], seed, "IC")
 *                                                                              LIB:1849-1856
 *   rank              order(cmi[,iter], decreasing = (target.match != "negative"))[1]
 *                                                                              LIB:1858-1864
 *   seed update       seed <- apply(rbind(seed, m.2[top,]), 2, "max");
 *                     cmi.seed[iter] <- assoc(target, seed, "IC")              LIB:2167-2171
 *   seed baseline     cmi.seed.iter0 <- assoc(target, seed, "IC")              LIB:1833
 *   permutation null  cmi.rand[k,j] <- cond.assoc(target.rand[j,], y, seed, "IC")   LIB:1853-1855
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types; no exception crosses the ABI;
 *   - every function returns 0 on success, a negative rvl_status otherwise;
 *     rvl_last_error(ctx) returns a message owned by ctx (valid until the next call on ctx);
 *   - host input buffers are read-only and not retained after the call returns;
 *   - one rvl_ctx == one GPU == one feature-row shard; a ctx is not re-entrant;
 *   - there is NO CPU fallback: without a CUDA device rvl_create fails with RVL_ERR_CUDA.
 *
 * Row-index convention: rows are 0-based here; the R shim adds 1.
 */
#ifndef REVEALER_B200_H
#define REVEALER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVL_ABI_VERSION 1
#define RVL_MAX_GRID 32

typedef struct rvl_ctx rvl_ctx;

typedef enum {
    RVL_OK = 0,
    RVL_ERR_CUDA = -1,       /* no device / CUDA runtime error (message has the CUDA string) */
    RVL_ERR_ARG = -2,        /* bad argument: NULL, size mismatch, n < 2, n_grid out of range ... */
    RVL_ERR_NOT_BINARY = -3, /* a feature or seed entry is not exactly 0 or 1 (or is NA/NaN) */
    RVL_ERR_STATE = -4,      /* call order: matrix / targets / seed not set yet */
    RVL_ERR_NONFINITE = -5   /* NA/NaN/Inf in a target */
} rvl_status;

/* Numeric constants the reference buries in calls (SURVEY section 5, config tier 4). */
typedef struct {
    int32_t n_grid;           /* n.grid = 25                      LIB:2497,2511,2517 */
    int32_t negative;         /* target.match == "negative"       LIB:1858 */
    double bandwidth_mult;    /* 0.25 in cond.mutual.inf's delta  LIB:2564 */
    double bandwidth_shrink;  /* -0.75 in delta*(1 + (-0.75)*rho) LIB:2532,2573 */
} rvl_opts;

/* Outputs of rvl_search; every pointer is caller-allocated (R: allocVector) or NULL to skip.
 * These are the objects REVEALER.v1 allocates at LIB:1837-1842. */
typedef struct {
    double *cmi;        /* [n_targets][iters][F_local]  unsorted scores, local row order (cmi[,iter] before LIB:1863) */
    int64_t *top;       /* [n_targets][iters]           GLOBAL row index of each iteration's winner (cmi.names[1,iter]) */
    double *top_score;  /* [n_targets][iters]           the winner's score */
    uint8_t *seed_iter; /* [n_targets][iters][n]        seed after each merge (seed.iter[iter,])   LIB:2169 */
    double *cmi_seed;   /* [n_targets][iters]           assoc(target, seed) after each merge       LIB:2170 */
    double *ic_seed0;   /* [n_targets]                  assoc(target, initial seed)                LIB:1833 */
} rvl_result;

/* ---- lifecycle ------------------------------------------------------------------------- */

int rvl_abi_version(void);

/* Bind a context to CUDA device `device`.  Fails (RVL_ERR_CUDA) when no such device exists. */
int rvl_create(rvl_ctx **out, int device);
void rvl_destroy(rvl_ctx *ctx);
const char *rvl_last_error(const rvl_ctx *ctx);

/* Run all work of this ctx on an existing CUDA stream (a cudaStream_t passed as void*), e.g.
 * torch's current stream so that a following torch.distributed collective is ordered after it. */
int rvl_set_stream(rvl_ctx *ctx, void *cuda_stream);
int rvl_set_options(rvl_ctx *ctx, const rvl_opts *opts);
int rvl_default_options(rvl_opts *opts);

/* Feature-row sharding: this ctx holds global rows [row_offset, row_offset+F_local)
 * of a matrix of F_global rows, as rank `rank` of `world`.  Default: rank 0 of 1, offset 0. */
int rvl_set_shard(rvl_ctx *ctx, int rank, int world, int64_t row_offset, int64_t F_global);

/* ---- inputs (replace the R objects m.2 / target / seed handed to the loop at LIB:1849) --- */

/* m.2 as R holds it: F x n doubles, column-major (m2[k + ld*s]); every entry must be 0 or 1.
 * Becomes uint64[F][words] on the device, sample s -> bit s%64 of word s/64.  PCIe is the bottleneck of this
 * call (8 bytes per bit kept), so the upload is cooperative: the DMA engine ships 64-sample words as doubles
 * from the front (packed by k_pack_f64_colmajor) while host threads pack words from the back (1/64 of the
 * bytes to send); the two sides meet wherever their speeds put them.  Same bits, same 0/1 validation. */
int rvl_load_matrix_f64_colmajor(rvl_ctx *ctx, const double *m2, int64_t F, int64_t n, int64_t ld);
/* Packer threads of the cooperative upload: n < 0 automatic (cores available to the process / world, at most
 * 16; none for small matrices), 0 = off (DMA + device packing only), n > 0 = that many. */
int rvl_set_host_threads(rvl_ctx *ctx, int n);
/* Of the last rvl_load_matrix_f64_colmajor: bytes that actually crossed PCIe (doubles of the DMA-side words + the
 * packed words of the host side) and how many 64-sample words the host threads packed. */
int rvl_get_upload_stats(rvl_ctx *ctx, uint64_t *out_h2d_bytes, uint64_t *out_host_packed_words);
/* Same matrix as one byte per entry, row-major (m2[k*n + s]). */
int rvl_load_matrix_u8_rowmajor(rvl_ctx *ctx, const uint8_t *m2, int64_t F, int64_t n);
/* Already bit-packed rows, `words` uint64 per row (words >= ceil(n/64)); padding bits must be 0. */
int rvl_load_matrix_bits(rvl_ctx *ctx, const uint64_t *bits, int64_t F, int64_t n, int64_t words);
/* As rvl_load_matrix_bits, but `bits` is a DEVICE pointer on this ctx's device (no H2D). */
int rvl_load_matrix_bits_device(rvl_ctx *ctx, const void *dev_bits, int64_t F, int64_t n, int64_t words);

/* ---- GCT ingest (replaces MSIG.Gct2Frame + data.matrix for ds2: LIB:2616-2626, LIB:1582) ------------
 * A GCT v1.2 file is memory-mapped; the host reads only the header and, on request, the name columns.
 * rvl_load_matrix_gct ships the data lines as text (about 2 bytes per entry instead of the 8 of an R
 * double matrix) and tokenises them on the device straight into the bit-packed matrix: no F x n
 * array of doubles ever exists.  Works without a context (and without a GPU) up to the load call. */
typedef struct rvl_gct rvl_gct;
int rvl_gct_open(const char *path, rvl_gct **out);
void rvl_gct_close(rvl_gct *gct);
/* Message of the last failed rvl_gct_* call on this handle (gct == NULL: of the last failed rvl_gct_open
 * on this thread). */
const char *rvl_gct_last_error(const rvl_gct *gct);
int rvl_gct_dims(const rvl_gct *gct, int64_t *n_rows, int64_t *n_cols);
/* Names: copy at most cap-1 bytes + NUL into buf, return the full length (< 0: bad index / malformed row). */
int rvl_gct_sample_name(const rvl_gct *gct, int64_t col, char *buf, size_t cap);
int rvl_gct_row_name(rvl_gct *gct, int64_t row, char *buf, size_t cap);
int rvl_gct_row_description(rvl_gct *gct, int64_t row, char *buf, size_t cap);
/* All row names (field 0) or descriptions (field 1) joined by '\n' into buf; returns the bytes needed
 * (call with cap = 0 to size the buffer), < 0 on error. */
int64_t rvl_gct_row_names(rvl_gct *gct, int field, char *buf, size_t cap);
/* One row as doubles, out[n_cols] (the continuous target row of ds1, LIB:1586); empty / "NA" -> NaN. */
int rvl_gct_row_values(rvl_gct *gct, int64_t row, double *out);
/* Rows [first_row, first_row + n_rows) of the file become the matrix of this ctx (a shard loads its own
 * block); output sample s is file column cols[s] (0-based, no repeats) -- the sample intersection and the
 * sort-by-target permutation of LIB:1596-1633 applied while parsing.  Every value must be 0 or 1
 * ("0", "1", "0.0", "1.00" ...), else RVL_ERR_NOT_BINARY; wrong field counts -> RVL_ERR_ARG. */
int rvl_load_matrix_gct(rvl_ctx *ctx, rvl_gct *gct, const int32_t *cols, int64_t n, int64_t first_row, int64_t n_rows);

/* Row selection on the device: keep rows[0..k) (local indices) of the loaded matrix, in that order --
 * m.2 <- m.2[retain, ] of the count filter (LIB:1655-1671), of the seed rows (LIB:1704-1706) and of the
 * excluded features (LIB:1716-1720) without re-uploading.  rvl_get_row_counts supplies rowSums(m.2). */
int rvl_select_rows(rvl_ctx *ctx, const int64_t *rows, int64_t k);
/* consolidate.identical.features = "identical" (LIB:1727-1748) on the device: the rows are ordered like their
 * 0/1 strings (paste(m.2[k,], collapse=""), stable as R's order()), every group of identical rows keeps its
 * first row, and the matrix becomes those rows in that order.  out_rows[i] = original local index of kept row i,
 * out_counts[i] = size of its group (the " (count)" suffix the reference appends); both hold F entries. */
int rvl_consolidate_identical(rvl_ctx *ctx, int64_t *out_rows, int32_t *out_counts, int64_t *out_k);
/* Rows rows[0..k) as one byte per sample, out[k][n] (the seed rows m.2[seed.names, ], LIB:1693-1700). */
int rvl_get_rows_u8(rvl_ctx *ctx, const int64_t *rows, int64_t k, uint8_t *out);

/* n_targets target profiles, row-major [n_targets][n].  Computes, on the device, per target:
 * mean, range, bcv(target) (MASS::bcv; pair histogram + Brent), and the (n-1)-entry table of
 * bcv() for binary vectors.  n_targets > 1 = a batch of independent searches (config C5), or,
 * with rvl_set_null_targets, the reference's permuted targets (LIB:1843-1844). */
int rvl_set_targets(rvl_ctx *ctx, const double *targets, int64_t n, int64_t n_targets);
/* As rvl_set_targets, for the reference's permutation nulls (target.rand, LIB:1843-1844; the 10 000
 * sample(target) draws of REVEALER_assess_features.v1, LIB:2885): targets 1..n_targets-1 are declared to
 * be permutations of target 0.  bcv() is permutation-invariant, so its O(n^2) pair histogram and
 * Brent search run once instead of n_targets times.  The claim is verified with an order-independent
 * fingerprint of the values; RVL_ERR_ARG if it does not hold. */
int rvl_set_targets_permuted(rvl_ctx *ctx, const double *targets, int64_t n, int64_t n_targets);
/* Targets [first_null, n_targets) are permutation-null targets: scored against target 0's
 * seed, excluded from the winner selection.  first_null <= 0 disables. */
int rvl_set_null_targets(rvl_ctx *ctx, int64_t first_null);

/* Initial seed(s): one byte per sample, 0 or 1 (all zero == "NULLSEED", LIB:1691).
 * per_target == 0: one seed [n] shared by all targets; else [n_targets][n]. */
int rvl_set_seed(rvl_ctx *ctx, const uint8_t *seed, int64_t n, int per_target);

/* ---- the hot path ---------------------------------------------------------------------- */

/* One scoring pass with the current seed: out[t*F_local + k] = cond.assoc(target_t, m.2[k,], seed_t, "IC").
 * Replaces the loop body LIB:1849-1856 for all rows (and all targets) at once. */
int rvl_score_once(rvl_ctx *ctx, double *out_scores);

/* The whole search, single shard: iters x (score, rank, merge, seed IC).  LIB:1846-2171. */
int rvl_search(rvl_ctx *ctx, int iters, rvl_result *out);

/* Split-phase form for feature-sharded runs (one ctx per rank).  Per iteration:
 *   rvl_iter_score  : score local rows + local best per target -> candidate records in the send buffer
 *   (caller all-gathers send -> recv across ranks: NCCL via torch.distributed, stream-ordered)
 *   rvl_iter_merge  : every rank resolves the same global winner from `world` records
 *                     (max score, ties -> lowest GLOBAL row, NaN never wins), ORs its row into
 *                     the seed, recomputes assoc(target, seed) and the per-iteration tables.
 * rvl_search_begin resets the iteration state (and computes ic_seed0); rvl_search_fetch copies results out.
 * The exchange buffers are DEVICE memory owned by the caller (so a torch tensor can back them):
 * send: rvl_candidate_bytes(ctx) bytes; recv: world * that. */
/* Peer-memory exchange instead of the caller's all-gather (ranks = processes on one NVLink/NVSwitch
 * node): every rank exports a CUDA-IPC handle of its exchange buffer (after the matrix and targets are
 * set), the host passes all `world` handles (64 bytes each, rank order) to every rank, and from then on
 * rvl_iter_score stores this rank's candidate record directly into every peer's buffer over NVLink
 * and rvl_iter_merge waits (bounded) for the peers' stamps -- no collective call between the two.
 * A stamp that does not arrive within ~2 s is reported as RVL_ERR_CUDA by rvl_search_fetch /
 * rvl_synchronize.  All ranks must call rvl_search_begin the same number of times. */
int rvl_p2p_export(rvl_ctx *ctx, unsigned char *handle64);
int rvl_p2p_import(rvl_ctx *ctx, const unsigned char *handles, int world);
int rvl_p2p_disable(rvl_ctx *ctx);
/* Several shards driven by ONE host thread (what a single R session does with n.gpus > 1): ctxs[i] is
 * rank i of n on n distinct devices, each with its row block, the (same) targets and seed set.
 * rvl_group_connect enables peer access and wires the same stamped peer exchange without IPC;
 * rvl_group_search enqueues the whole search on every device (all launches asynchronous) and waits;
 * results are then read per context with rvl_search_fetch (scores for its own rows; winners, seeds and
 * seed ICs identical on every context). */
int rvl_group_connect(rvl_ctx **ctxs, int n);
int rvl_group_search(rvl_ctx **ctxs, int n, int iters);
/* The whole search of this shard enqueued without waiting (rvl_search_fetch / rvl_synchronize wait): the split-phase
 * launches back to back (k_score, k_argmax, k_advance per iteration; nothing synchronises with the host), or -- after
 * rvl_set_fused(ctx, 1) -- ONE persistent cooperative launch (k_search: score, rank, peer exchange over NVLink, merge
 * and table update of every iteration inside the kernel).  world > 1 needs the peer exchange (rvl_p2p_import or
 * rvl_group_connect) and every rank must call it.  Results are bit-identical either way. */
int rvl_search_run(rvl_ctx *ctx, int iters);
/* on = 1 makes rvl_search / rvl_search_run / rvl_group_search use the single persistent launch (k_search) when the
 * device can launch it cooperatively.  Default 0: on B200 the persistent form measured 2-3 % slower than the
 * split-phase launches at 1 and at 8 GPUs, so it is kept as an option, tested bit-identical. */
int rvl_set_fused(rvl_ctx *ctx, int on);
int rvl_search_begin(rvl_ctx *ctx, int iters);
size_t rvl_candidate_bytes(const rvl_ctx *ctx);
int rvl_set_exchange_buffers(rvl_ctx *ctx, void *dev_send, void *dev_recv);
int rvl_iter_score(rvl_ctx *ctx, int iter);
int rvl_iter_merge(rvl_ctx *ctx, int iter);
int rvl_search_fetch(rvl_ctx *ctx, rvl_result *out);
int rvl_synchronize(rvl_ctx *ctx);

/* assoc(x, y, "IC") for target t and an arbitrary 0/1 vector y (LIB:2491-2501): seed baselines
 * at LIB:1820-1833 and REVEALER_assess_features' per-feature IC. */
int rvl_assoc(rvl_ctx *ctx, int64_t target_index, const uint8_t *y, int64_t n, double *out_ic);

/* Permutation p-values, LIB:1868-1869: pval[i] = #(cmi_rand > cmi[i]) / n_rand, on the device
 * (sort + binary search instead of the O(F^2 n.perm) scan).  A NaN score gives a NaN p-value (R: NA). */
int rvl_pvalues(rvl_ctx *ctx, const double *cmi, int64_t F, const double *cmi_rand, int64_t n_rand,
                double *out_pval);
/* The same for iteration `iter` (0-based) of the last search, from the scores still resident in HBM: target 0's
 * scores against the scores of the permutation-null targets (rvl_set_null_targets) -- no score crosses PCIe, only
 * the F p-values come back.  A NaN score gives a NaN p-value (R: NA).  Single shard only. */
int rvl_pvalues_iter(rvl_ctx *ctx, int iter, double *out_pval);

/* The first k entries of R's order(cmi[, iter], decreasing = !negative) for one target of the last search (the
 * n.markers rows REVEALER.v1 displays, LIB:1858-1864), ranked on the device from the resident scores: stable (ties ->
 * lower row), NaN last.  out_rows: GLOBAL row indices (this shard's rows only: with world > 1 merge the shards'
 * lists with the same rule); out_scores may be NULL.  k is clipped to the shard's row count. */
int rvl_topk(rvl_ctx *ctx, int iter, int64_t target_index, int64_t k, int64_t *out_rows, double *out_scores);

/* ---- introspection (parity tests, INTEGRATION checks) ----------------------------------- */

/* bcv(target_t) as computed on the device. */
int rvl_get_target_bandwidth(rvl_ctx *ctx, int64_t target_index, double *out_bcv);
/* bcv() of a binary vector of length n with n1 ones, n1 = 0..n (entries 0 and n are 0). out[n+1]. */
int rvl_get_binary_bandwidth_table(rvl_ctx *ctx, double *out_table);
/* Shape of the loaded matrix: rows of this shard and samples (dim(m.2)). */
int rvl_get_matrix_dims(rvl_ctx *ctx, int64_t *out_rows, int64_t *out_samples);
/* Row popcounts (rowSums(m.2), LIB:1656) of the local shard. out[F_local]. */
int rvl_get_row_counts(rvl_ctx *ctx, int32_t *out_counts);
/* Kernel launches issued by this ctx since creation (bench.py's gpu_launches claim). */
int64_t rvl_launch_count(const rvl_ctx *ctx);
/* Instrumentation switch (default off): on != 0 makes every k_score launch count its work (rvl_get_counters*) and
 * brackets it with a CUDA-event pair (rvl_score_kernel_time).  Off, the product path carries neither the per-evaluation
 * atomics nor the event records. */
int rvl_set_profiling(rvl_ctx *ctx, int on);
/* Device time of the dominant score kernel, summed over launches since the last reset (ms),
 * measured with CUDA events on the ctx stream; and the number of launches it covers. */
int rvl_score_kernel_time(rvl_ctx *ctx, double *out_ms, int64_t *out_launches, int reset);

/* Validation switch: on != 0 makes the score kernel recompute the x-axis kernel sums of every
 * feature over all n samples (the literal form) instead of using the per-iteration
 * bandwidth-interpolation table + minority-class visit.  Same results to ~1e-12; ~5x slower. */
int rvl_set_dense_x(rvl_ctx *ctx, int on);
/* Density-column pruning of the 3-D entropy epilogue.  A (y,z)-grid column whose cells can hold at
 * most theta / n_grid^3 of the total probability mass is replaced by the reference's
 * .Machine$double.eps floor (LIB:2581); this changes CMI by at most theta * (|log eps'| + 1)
 * (about 4e-13 for the default theta = 1e-14, against the 1e-6 relative tolerance on scores).
 * theta = 0 keeps only the exact rule (columns that cannot change fl(density + eps)). */
int rvl_set_prune_theta(rvl_ctx *ctx, double theta);
/* Exact-duplicate rows (LIB:1727-1748 consolidates them by name pasting).  Identical
 * rows have identical scores and R's stable order() ranks the first of them first, so by default
 * only the first row of each group is scored and the others receive a copy of its score (grouping is
 * done on the device at load time: hash, stable sort, exact compare).  on = 0 scores every row. */
int rvl_set_dedup(rvl_ctx *ctx, int on);
/* Number of distinct rows in the loaded shard. */
int64_t rvl_unique_rows(const rvl_ctx *ctx);
/* Work counters of the score kernel since the last reset: out4[0] evaluations that ran the KDE path,
 * out4[1] live (j,k) density columns (n_grid log() each), out4[2] dense-pass exp() evaluations. */
int rvl_get_counters(rvl_ctx *ctx, uint64_t *out4, int reset);
/* The same with more slots (n_out <= 8): out[1] counts only the columns evaluated cell by cell with a double-precision
 * log (both y-classes visible), out[3] (j,k) columns in the O(1) factorised form, out[4] one-class columns (the other
 * y-class cannot change fl(Q) anywhere in the column) whose eps-floor correction is summed cell by cell in single
 * precision, out[5] log() evaluations, out[6] FP64 flops outside exp() and log() (an FMA counts 2), out[7] single-precision
 * "soft" cells (out[4] * n_grid).  Needs rvl_set_profiling. */
int rvl_get_counters_ex(rvl_ctx *ctx, uint64_t *out, int n_out, int reset);
/* Development aid: nanosecond (globaltimer) stamps of the phases of the last rvl_search_run's k_search launch, out[iters][8]:
 * [0] block 0 starts scoring, [1] block 0 done, [2] every block done, [3] candidate records + stamps published,
 * [4] block 0's merge tasks done, [5] grid barrier passed.  Call once with out == NULL to arm the recording. */
int rvl_debug_phase_times(rvl_ctx *ctx, unsigned long long *out, int iters);
/* FP64 FMA peak of this device by a DFMA-chain micro-benchmark (TFLOP/s); the denominator of the
 * roofline bench.py reports, since the path is FP64-compute-bound. */
int rvl_measure_fp64_peak(rvl_ctx *ctx, double *out_tflops);

#ifdef __cplusplus
}
#endif
#endif /* REVEALER_B200_H */
