"""Host-side mirror of the reference's R interface for the CMI-search path.

Names, argument meaning and error behaviour follow REVEALER_library.v5C.R ("LIB", copy-B lines):

    assoc(x, y, metric)                 LIB:2491-2501
    cond_assoc(x, y, z, metric)         LIB:2503-2522
    revealer_search(...)                the iteration section of REVEALER.v1, LIB:1837-1864 + 2167-2171
    REVEALER_v1(...)                    REVEALER.v1's formals LIB:1504-1516 and its preprocessing
                                        LIB:1575-1707 (host-side, numpy; no plots, no PDF)

Everything numerical is done by librevealer_b200.so through its C ABI (include/revealer_b200.h).
This module never computes a score itself and has no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import rvl_opts, rvl_result


class RevealerError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rvl error %d: %s" % (code, msg))
        self.code = code


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def pack_bits(m):
    """0/1 matrix (F x n, any integer/bool/float dtype) -> (uint64[F][words], words), sample s at
    bit s%64 of word s/64, rows padded to an even number of words (16-byte rows)."""
    m = np.asarray(m)
    if m.ndim == 1:
        m = m[None, :]
    F, n = m.shape
    if not np.all((m == 0) | (m == 1)):
        raise RevealerError(_lib.RVL_ERR_NOT_BINARY, "matrix has an entry that is not 0 or 1")
    words = ((n + 63) // 64 + 1) & ~1
    by = np.packbits(m.astype(np.uint8), axis=1, bitorder="little")
    out = np.zeros((F, words * 8), dtype=np.uint8)
    out[:, : by.shape[1]] = by
    return out.view("<u8").reshape(F, words), words


def shard_rows(F, world, rank):
    """Contiguous row block of rank `rank`: [rank*F/world, (rank+1)*F/world)  (SURVEY 8e)."""
    lo = (rank * F) // world
    hi = ((rank + 1) * F) // world
    return lo, hi - lo


def resolve_winner(scores, rows, negative=False):
    """The deterministic rule every rank applies to the gathered candidates: R's stable
    order(decreasing = !negative)[1] -- best score, ties -> lowest global row, NaN never ahead of a
    number (LIB:1858-1862).  rows < 0 mark empty shards.  Returns the index into the candidate list."""
    best = -1
    for i, (v, r) in enumerate(zip(scores, rows)):
        if r < 0:
            continue
        if best < 0:
            best = i
            continue
        bv, br = scores[best], rows[best]
        vn, bn = np.isnan(v), np.isnan(bv)
        if vn or bn:
            ahead = (r < br) if (vn and bn) else bool(bn)
        elif v != bv:
            ahead = (v < bv) if negative else (v > bv)
        else:
            ahead = r < br
        if ahead:
            best = i
    return best


class GctFile:
    """A memory-mapped GCT v1.2 file (rvl_gct): what MSIG.Gct2Frame (LIB:2616-2626) returns, without the
    data frame -- dims, sample names, row names / descriptions and single rows are read on the host; the
    matrix itself goes to the device as text through RevealerContext.load_gct.  Needs no GPU."""

    def __init__(self, path):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.rvl_gct_open(str(path).encode(), C.byref(self._h))
        if rc != 0:
            self._h = None
            raise RevealerError(rc, self._lib.rvl_gct_last_error(None).decode("utf-8", "replace"))
        r, c = C.c_int64(), C.c_int64()
        self._lib.rvl_gct_dims(self._h, C.byref(r), C.byref(c))
        self.n_rows, self.n_cols = r.value, c.value
        self.path = str(path)
        self._names = {}

    def _err(self, rc=_lib.RVL_ERR_ARG):
        return RevealerError(rc, self._lib.rvl_gct_last_error(self._h).decode("utf-8", "replace") or "bad GCT request")

    @property
    def names(self):
        """Sample (column) names, header line 3 without Name / Description."""
        if "s" not in self._names:
            buf = C.create_string_buffer(4096)
            out = []
            for j in range(self.n_cols):
                ln = self._lib.rvl_gct_sample_name(self._h, j, buf, len(buf))
                if ln >= len(buf):
                    buf = C.create_string_buffer(ln + 1)
                    self._lib.rvl_gct_sample_name(self._h, j, buf, len(buf))
                out.append(buf.value.decode("utf-8", "replace"))
            self._names["s"] = out
        return self._names["s"]

    def _rows(self, field):
        if field not in self._names:
            need = self._lib.rvl_gct_row_names(self._h, field, None, 0)
            if need < 0:
                raise self._err()
            buf = C.create_string_buffer(max(int(need), 1))
            self._lib.rvl_gct_row_names(self._h, field, buf, len(buf))
            self._names[field] = buf.raw[:max(int(need) - 1, 0)].decode("utf-8", "replace").split("\n") if self.n_rows else []
        return self._names[field]

    @property
    def row_names(self):
        return self._rows(0)

    @property
    def descs(self):
        return self._rows(1)

    def row_values(self, i):
        """Row i as float64[n_cols] (empty / NA cells -> NaN)."""
        out = np.empty(self.n_cols, dtype=np.float64)
        if self._lib.rvl_gct_row_values(self._h, int(i), _dp(out)) != 0:
            raise self._err()
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rvl_gct_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class RevealerContext:
    """One GPU, one feature-row shard.  Thin object wrapper over rvl_ctx."""

    def __init__(self, device=0, n_grid=25, bandwidth_mult=0.25, bandwidth_shrink=-0.75, negative=False):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.rvl_create(C.byref(self._h), int(device))
        if rc != 0:
            self._h = None
            raise RevealerError(rc, "rvl_create failed: no usable CUDA device %d (there is no CPU fallback)" % device)
        self.device = device
        self.F = None
        self.n = None
        self.T = 0
        self.world, self.rank, self.row_offset = 1, 0, 0
        self._keep = []
        self.set_options(n_grid=n_grid, bandwidth_mult=bandwidth_mult, bandwidth_shrink=bandwidth_shrink,
                         negative=negative)

    # -- plumbing
    def _ck(self, rc):
        if rc != 0:
            raise RevealerError(rc, self._lib.rvl_last_error(self._h).decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rvl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_options(self, n_grid=25, bandwidth_mult=0.25, bandwidth_shrink=-0.75, negative=False):
        o = rvl_opts(int(n_grid), int(bool(negative)), float(bandwidth_mult), float(bandwidth_shrink))
        self._ck(self._lib.rvl_set_options(self._h, C.byref(o)))
        self.opts = o

    def set_stream(self, cuda_stream_ptr):
        self._ck(self._lib.rvl_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr))))

    def set_shard(self, rank, world, row_offset, F_global):
        self._ck(self._lib.rvl_set_shard(self._h, rank, world, row_offset, F_global))
        self.rank, self.world, self.row_offset = rank, world, row_offset

    def set_host_threads(self, n):
        """Packer threads of the cooperative upload of float64 matrices (-1 automatic, 0 off)."""
        self._ck(self._lib.rvl_set_host_threads(self._h, int(n)))

    def upload_stats(self):
        """(bytes that crossed PCIe, 64-sample words packed by host threads) of the last float64 load_matrix."""
        a, b = C.c_uint64(), C.c_uint64()
        self._ck(self._lib.rvl_get_upload_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    # -- inputs
    def load_matrix(self, m2):
        """m2: F x n.  float64 -> the R layout path (column-major doubles, validated and packed on the
        device); uint8/bool -> row-major bytes; uint64 -> already bit-packed rows."""
        m2 = np.asarray(m2)
        if m2.ndim != 2:
            raise RevealerError(_lib.RVL_ERR_ARG, "feature matrix must be 2-D (F x n)")
        F, n = m2.shape
        if m2.dtype == np.float64:
            a = np.asfortranarray(m2)
            self._ck(self._lib.rvl_load_matrix_f64_colmajor(self._h, _dp(a), F, n, max(F, 1)))
        elif m2.dtype in (np.uint8, np.bool_):
            a = np.ascontiguousarray(m2).view(np.uint8)
            self._ck(self._lib.rvl_load_matrix_u8_rowmajor(self._h, a.ctypes.data_as(C.POINTER(C.c_uint8)), F, n))
        else:
            raise RevealerError(_lib.RVL_ERR_ARG, "feature matrix dtype must be float64, uint8 or bool")
        self.F, self.n = F, n

    def load_matrix_block(self, m2_fortran, row_lo, n_rows):
        """Rows [row_lo, row_lo + n_rows) of a column-major float64 F_total x n matrix, read in place
        (ld = F_total): what a shard of an R matrix looks like (rvlR_search_multi)."""
        a = m2_fortran
        if a.dtype != np.float64 or a.ndim != 2 or not a.flags.f_contiguous:
            raise RevealerError(_lib.RVL_ERR_ARG, "load_matrix_block needs a Fortran-contiguous float64 matrix")
        Ft, n = a.shape
        if row_lo < 0 or n_rows < 0 or row_lo + n_rows > Ft:
            raise RevealerError(_lib.RVL_ERR_ARG, "row block out of range")
        ptr = C.cast(C.c_void_p(a.ctypes.data + 8 * int(row_lo)), C.POINTER(C.c_double))
        self._ck(self._lib.rvl_load_matrix_f64_colmajor(self._h, ptr, int(n_rows), n, max(Ft, 1)))
        self.F, self.n = int(n_rows), n

    def load_gct(self, gct, cols=None, first_row=0, n_rows=None):
        """Rows [first_row, first_row + n_rows) of a GctFile, tokenised on the device; output sample s is file
        column cols[s] (default: all columns in file order)."""
        cols = np.arange(gct.n_cols, dtype=np.int32) if cols is None else np.ascontiguousarray(cols, dtype=np.int32)
        n_rows = gct.n_rows - first_row if n_rows is None else int(n_rows)
        self.F = None
        self._ck(self._lib.rvl_load_matrix_gct(self._h, gct._h, cols.ctypes.data_as(C.POINTER(C.c_int32)), len(cols),
                                               int(first_row), n_rows))
        self.F, self.n = n_rows, len(cols)

    def select_rows(self, rows):
        """m.2 <- m.2[rows, ] on the device (count filter, seed / excluded rows removed)."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        self._ck(self._lib.rvl_select_rows(self._h, rows.ctypes.data_as(C.POINTER(C.c_int64)), len(rows)))
        self.F = len(rows)

    def consolidate_identical(self):
        """consolidate.identical.features = "identical" (LIB:1727-1748): returns (rows, counts) -- the original
        index and the group size of every kept row, in the reference's (pattern-sorted) order; the resident
        matrix becomes those rows."""
        rows = np.empty(max(self.F, 1), dtype=np.int64)
        counts = np.empty(max(self.F, 1), dtype=np.int32)
        k = C.c_int64()
        self._ck(self._lib.rvl_consolidate_identical(self._h, rows.ctypes.data_as(C.POINTER(C.c_int64)),
                                                     counts.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(k)))
        self.F = int(k.value)
        return rows[:self.F].copy(), counts[:self.F].copy()

    def get_rows(self, rows):
        """Rows of the loaded matrix as uint8[k][n]."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        out = np.empty((len(rows), self.n), dtype=np.uint8)
        self._ck(self._lib.rvl_get_rows_u8(self._h, rows.ctypes.data_as(C.POINTER(C.c_int64)), len(rows),
                                           out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def load_matrix_bits(self, bits, n):
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        F, words = bits.shape
        self._ck(self._lib.rvl_load_matrix_bits(self._h, bits.ctypes.data_as(C.POINTER(C.c_uint64)), F, n, words))
        self.F, self.n = F, n

    def load_matrix_bits_device(self, dev_ptr, F, n, words):
        self._ck(self._lib.rvl_load_matrix_bits_device(self._h, C.c_void_p(int(dev_ptr)), F, n, words))
        self.F, self.n = F, n

    def set_targets(self, targets, first_null=0, permuted=False):
        """targets: [n] or [T][n].  permuted=True declares rows 1.. to be permutations of row 0 (verified
        on the device): they then share row 0's bcv bandwidth instead of recomputing it T times."""
        t = np.ascontiguousarray(np.atleast_2d(np.asarray(targets, dtype=np.float64)))
        T, n = t.shape
        fn = self._lib.rvl_set_targets_permuted if permuted else self._lib.rvl_set_targets
        self._ck(fn(self._h, _dp(t), n, T))
        self.T, self.n = T, n
        if first_null and first_null > 0:
            self._ck(self._lib.rvl_set_null_targets(self._h, first_null))

    def set_seed(self, seed):
        s = np.asarray(seed)
        if not np.all((s == 0) | (s == 1)):
            raise RevealerError(_lib.RVL_ERR_NOT_BINARY, "seed has an entry that is not 0 or 1")
        s = np.ascontiguousarray(s.astype(np.uint8))
        per_target = 1 if s.ndim == 2 else 0
        n = s.shape[-1]
        self._ck(self._lib.rvl_set_seed(self._h, s.ctypes.data_as(C.POINTER(C.c_uint8)), n, per_target))

    # -- hot path
    def score_once(self):
        out = np.empty((self.T or 0, self.F or 0), dtype=np.float64)
        self._ck(self._lib.rvl_score_once(self._h, _dp(out)))  # call order errors come from the core
        return out

    def _alloc_result(self, iters, want_cmi=True):
        T, F, n = self.T, self.F, self.n
        res = dict(
            cmi=np.empty((T, iters, F), dtype=np.float64) if want_cmi else None,
            top=np.empty((T, iters), dtype=np.int64),
            top_score=np.empty((T, iters), dtype=np.float64),
            seed_iter=np.empty((T, iters, n), dtype=np.uint8),
            cmi_seed=np.empty((T, iters), dtype=np.float64),
            ic_seed0=np.empty((T,), dtype=np.float64),
        )
        r = rvl_result()
        r.cmi = _dp(res["cmi"]) if want_cmi else None
        r.top = res["top"].ctypes.data_as(C.POINTER(C.c_int64))
        r.top_score = _dp(res["top_score"])
        r.seed_iter = res["seed_iter"].ctypes.data_as(C.POINTER(C.c_uint8))
        r.cmi_seed = _dp(res["cmi_seed"])
        r.ic_seed0 = _dp(res["ic_seed0"])
        return res, r

    def search(self, iters, want_cmi=True, fetch=True):
        """iters x (score all rows, rank, OR-merge the winner into the seed, assoc(target, seed))."""
        if not fetch:
            self._ck(self._lib.rvl_search(self._h, int(iters), None))
            return None
        res, r = self._alloc_result(iters, want_cmi)
        self._ck(self._lib.rvl_search(self._h, int(iters), C.byref(r)))
        return res

    def search_run(self, iters):
        """Enqueue the whole search of this shard (one persistent launch when possible); search_fetch waits."""
        self._ck(self._lib.rvl_search_run(self._h, int(iters)))

    def set_fused(self, on):
        """on=False keeps the search on the split-phase launches (k_score / k_argmax / k_advance per iteration)."""
        self._ck(self._lib.rvl_set_fused(self._h, int(bool(on))))

    # split-phase (sharded) form
    def search_begin(self, iters):
        self._ck(self._lib.rvl_search_begin(self._h, int(iters)))

    def candidate_bytes(self):
        return int(self._lib.rvl_candidate_bytes(self._h))

    def set_exchange_buffers(self, send_ptr, recv_ptr):
        self._ck(self._lib.rvl_set_exchange_buffers(self._h, C.c_void_p(int(send_ptr)), C.c_void_p(int(recv_ptr))))

    def p2p_export(self):
        """64-byte CUDA-IPC handle of this rank's exchange buffer (see rvl_p2p_export)."""
        h = (C.c_uint8 * 64)()
        self._ck(self._lib.rvl_p2p_export(self._h, h))
        return bytes(h)

    def p2p_import(self, handles):
        """handles: the `world` exported handles concatenated in rank order (world * 64 bytes)."""
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self._ck(self._lib.rvl_p2p_import(self._h, buf, len(handles) // 64))

    def p2p_disable(self):
        self._ck(self._lib.rvl_p2p_disable(self._h))

    def iter_score(self, it):
        self._ck(self._lib.rvl_iter_score(self._h, int(it)))

    def iter_merge(self, it):
        self._ck(self._lib.rvl_iter_merge(self._h, int(it)))

    def search_fetch(self, iters, want_cmi=True):
        res, r = self._alloc_result(iters, want_cmi)
        self._ck(self._lib.rvl_search_fetch(self._h, C.byref(r)))
        return res

    def synchronize(self):
        self._ck(self._lib.rvl_synchronize(self._h))

    def assoc(self, y, target_index=0):
        y = np.asarray(y)
        if not np.all((y == 0) | (y == 1)):
            raise RevealerError(_lib.RVL_ERR_NOT_BINARY, "y has an entry that is not 0 or 1")
        y = np.ascontiguousarray(y.astype(np.uint8))
        out = C.c_double()
        self._ck(self._lib.rvl_assoc(self._h, int(target_index), y.ctypes.data_as(C.POINTER(C.c_uint8)), len(y),
                                     C.byref(out)))
        return out.value

    def pvalues(self, cmi, cmi_rand):
        c = np.ascontiguousarray(cmi, dtype=np.float64)
        r = np.ascontiguousarray(np.ravel(cmi_rand), dtype=np.float64)
        out = np.empty(len(c), dtype=np.float64)
        self._ck(self._lib.rvl_pvalues(self._h, _dp(c), len(c), _dp(r), len(r), _dp(out)))
        return out

    def topk(self, it, k, target_index=0):
        """(rows, scores): the first k entries of R's order() of iteration `it`'s scores, ranked on the device."""
        k = int(min(k, self.F or 0))
        rows = np.empty(k, dtype=np.int64)
        sc = np.empty(k, dtype=np.float64)
        self._ck(self._lib.rvl_topk(self._h, int(it), int(target_index), k, rows.ctypes.data_as(C.POINTER(C.c_int64)), _dp(sc)))
        return rows, sc

    def pvalues_iter(self, it):
        """p-values of iteration `it` of the last search from the resident scores (rvl_pvalues_iter)."""
        out = np.empty(self.F, dtype=np.float64)
        self._ck(self._lib.rvl_pvalues_iter(self._h, int(it), _dp(out)))
        return out

    # -- introspection
    def target_bandwidth(self, t=0):
        out = C.c_double()
        self._ck(self._lib.rvl_get_target_bandwidth(self._h, int(t), C.byref(out)))
        return out.value

    def binary_bandwidth_table(self):
        out = np.empty(self.n + 1, dtype=np.float64)
        self._ck(self._lib.rvl_get_binary_bandwidth_table(self._h, _dp(out)))
        return out

    def row_counts(self):
        out = np.empty(self.F or 0, dtype=np.int32)   # no matrix: the core reports the state error
        self._ck(self._lib.rvl_get_row_counts(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def launch_count(self):
        return int(self._lib.rvl_launch_count(self._h))

    def set_dense_x(self, on):
        """Validation mode: literal per-feature x-axis sums over all samples (see rvl_set_dense_x)."""
        self._ck(self._lib.rvl_set_dense_x(self._h, int(bool(on))))

    def set_dedup(self, on):
        """Score only the first row of every group of identical rows (default on; see rvl_set_dedup)."""
        self._ck(self._lib.rvl_set_dedup(self._h, int(bool(on))))

    def unique_rows(self):
        return int(self._lib.rvl_unique_rows(self._h))

    def set_prune_theta(self, theta):
        """Density-column pruning threshold of the 3-D epilogue (see rvl_set_prune_theta); 0 = exact rule only."""
        self._ck(self._lib.rvl_set_prune_theta(self._h, float(theta)))

    def set_profiling(self, on):
        """Work counters + per-launch timing of the score kernel (off by default; see rvl_set_profiling)."""
        self._ck(self._lib.rvl_set_profiling(self._h, int(bool(on))))

    def counters(self, reset=False):
        out = (C.c_uint64 * 8)()
        self._ck(self._lib.rvl_get_counters_ex(self._h, out, 8, int(reset)))
        return dict(kde_evaluations=int(out[0]), live_columns=int(out[1]), dense_exps=int(out[2]),
                    factorised_columns=int(out[3]), two_term_columns=int(out[4]), soft_columns=int(out[4]), logs=int(out[5]),
                    plain_flops=int(out[6]), soft_cells=int(out[7]))

    def measure_fp64_peak(self):
        out = C.c_double()
        self._ck(self._lib.rvl_measure_fp64_peak(self._h, C.byref(out)))
        return out.value

    def score_kernel_time(self, reset=False):
        ms, nl = C.c_double(), C.c_int64()
        self._ck(self._lib.rvl_score_kernel_time(self._h, C.byref(ms), C.byref(nl), int(reset)))
        return ms.value, nl.value


# ---------------------------------------------------------------------------------------------
# Reference-shaped functions
# ---------------------------------------------------------------------------------------------

def _unique_is_one(v):
    v = np.asarray(v)
    return v.size > 0 and bool(np.all(v == v.flat[0]))


def assoc(x, y, metric="IC", device=0, n_grid=25):
    """assoc(x, y, metric), LIB:2491-2501.  y must be a 0/1 vector (binary features only)."""
    if metric != "IC":
        raise RevealerError(_lib.RVL_ERR_ARG, 'only metric="IC" is implemented (assoc.metric is hard-coded to "IC", LIB:1522)')
    with RevealerContext(device, n_grid=n_grid) as ctx:
        ctx.set_targets(x)
        return ctx.assoc(y, 0)


def cond_assoc(x, y, z, metric="IC", device=0, n_grid=25):
    """cond.assoc(x, y, z, metric), LIB:2503-2522: 0 for constant x or y; IC(x,y) for constant z;
    CIC(x,y|z) otherwise.  y and z must be 0/1 vectors."""
    if metric != "IC":
        raise RevealerError(_lib.RVL_ERR_ARG, 'only metric="IC" is implemented (assoc.metric is hard-coded to "IC", LIB:1522)')
    with RevealerContext(device, n_grid=n_grid) as ctx:
        ctx.load_matrix(np.asarray(y, dtype=np.float64)[None, :])
        ctx.set_targets(x)
        ctx.set_seed(z)
        return float(ctx.score_once()[0, 0])


def revealer_search(target, m2, seed=None, max_n_iter=5, target_match="positive", n_grid=25,
                    bandwidth_mult=0.25, bandwidth_shrink=-0.75, target_rand=None, n_markers=30, device=0,
                    ctx=None):
    """The iteration section of REVEALER.v1 (LIB:1837-1864, 2167-2171) on in-memory inputs.

    target        double[n]  (already ordered as REVEALER.v1 orders samples, LIB:1626-1633)
    m2            F x n 0/1 matrix (seed rows already removed, LIB:1703-1705)
    seed          0/1 [n] or None == "NULLSEED" (all zero, LIB:1691)
    target_rand   optional n.perm x n permuted targets (LIB:1843-1844); adds the null scores and the
                  per-iteration permutation p-values of LIB:1868-1869
    Returns the objects REVEALER.v1 builds: cmi (F x iters, UNSORTED), order (iters x F, R's order()),
    top (0-based winner per iteration), seed_iter, cmi_seed, ic_seed0, and with target_rand also
    cmi_rand (iters x F x n.perm) and p_val (iters x F, aligned with cmi rows).
    """
    if target_match not in ("positive", "negative"):
        raise RevealerError(_lib.RVL_ERR_ARG, 'target.match must be "positive" or "negative"')
    negative = target_match == "negative"
    target = np.asarray(target, dtype=np.float64)
    n = target.shape[-1]
    own = ctx is None
    if own:
        ctx = RevealerContext(device, n_grid=n_grid, bandwidth_mult=bandwidth_mult,
                              bandwidth_shrink=bandwidth_shrink, negative=negative)
    try:
        if own or m2 is not None:
            ctx.load_matrix(m2)
        targets = target[None, :]
        first_null = 0
        if target_rand is not None:
            targets = np.vstack([targets, np.asarray(target_rand, dtype=np.float64)])
            first_null = 1
        ctx.set_targets(targets, first_null=first_null, permuted=target_rand is not None)
        ctx.set_seed(np.zeros(n, dtype=np.uint8) if seed is None else seed)
        res = ctx.search(max_n_iter)
        cmi = res["cmi"][0].T.copy()  # F x iters like R
        out = dict(cmi=cmi, top=res["top"][0], top_score=res["top_score"][0], seed_iter=res["seed_iter"][0],
                   cmi_seed=res["cmi_seed"][0], ic_seed0=float(res["ic_seed0"][0]))
        # display-side ranking (R's order(); stable, NaN last)
        key = np.where(np.isnan(cmi), np.inf, cmi if negative else -cmi)
        out["order"] = np.stack([np.argsort(key[:, it], kind="stable") for it in range(max_n_iter)])
        # the n.markers rows REVEALER.v1 displays (LIB:1858-1864): ranked on the device from the resident scores
        out["top_markers"] = np.stack([ctx.topk(it, n_markers)[0] for it in range(max_n_iter)])
        if target_rand is not None:
            out["cmi_rand"] = np.transpose(res["cmi"][1:], (1, 2, 0)).copy()  # iters x F x n.perm
            # p-values on the scores still resident in HBM (LIB:1868-1869); NaN score -> NaN p-value (R: NA)
            out["p_val"] = np.stack([ctx.pvalues_iter(it) for it in range(max_n_iter)])
            # FDR as REVEALER.v1 reports it (LIB:1870-1878): BH on p and on 1-p; rows with a negative score
            # take the lower-tail pair.  Display-side bookkeeping on F numbers: host numpy, like R's p.adjust.
            fdr = np.stack([p_adjust_fdr(p) for p in out["p_val"]])
            fdr_lower = np.stack([p_adjust_fdr(1.0 - p) for p in out["p_val"]])
            neg = (cmi.T < 0)
            out["FDR"] = np.where(neg, fdr_lower, fdr)
            out["p_val_signed"] = np.where(neg, 1.0 - out["p_val"], out["p_val"])
        return out
    finally:
        if own:
            ctx.close()


def assess_features(target, features, n_perm=10000, r_seed=5209761, n_grid=25, device=0, target_rand=None):
    """The numerical core of REVEALER_assess_features.v1 (LIB:2672-2933): for every feature row,
    IC = mutual.inf.v2(target, feature)$IC (LIB:2875) and its permutation p-value against n.perm
    shuffled targets (LIB:2885-2890: sum(null >= IC)/n.perm for IC >= 0, sum(null <= IC)/n.perm
    otherwise).  features: F x n 0/1 matrix.  target_rand: optional n.perm x n permuted targets
    (R's sample() stream is not reproduced; default: numpy PCG64(r_seed)).
    Returns dict(IC [F], p_val [F], null_IC [n_perm][F]).  Plots, ROC and squared error stay in R."""
    target = np.asarray(target, dtype=np.float64)
    n = target.shape[0]
    if target_rand is None:
        rng = np.random.Generator(np.random.PCG64(r_seed))
        target_rand = np.stack([rng.permutation(target) for _ in range(n_perm)]) if n_perm > 0 else np.empty((0, n))
    target_rand = np.asarray(target_rand, dtype=np.float64)
    ic = None
    null = []
    chunk = 60000                                   # targets per upload (the ABI takes <= 65535)
    with RevealerContext(device, n_grid=n_grid) as ctx:
        ctx.load_matrix(features)
        for lo in range(0, max(1, len(target_rand)), chunk):
            part = target_rand[lo:lo + chunk]
            ctx.set_targets(np.vstack([target[None, :], part]), permuted=True)
            ctx.set_seed(np.zeros(n, dtype=np.uint8))  # constant z: cond.assoc -> mutual.inf.v2 (LIB:2509)
            sc = ctx.score_once()
            ic = sc[0]
            null.append(sc[1:])
    null = np.concatenate(null, axis=0) if null else np.empty((0, len(ic)))
    n_perm = max(1, null.shape[0])
    p = np.where(ic >= 0, (null >= ic[None, :]).sum(axis=0), (null <= ic[None, :]).sum(axis=0)) / n_perm
    return dict(IC=ic, p_val=p, null_IC=null)


def revealer_search_multi_gpu(target, m2, seed=None, max_n_iter=5, target_match="positive", devices=(0, 1),
                              n_grid=25, bandwidth_mult=0.25, bandwidth_shrink=-0.75):
    """revealer_search with the feature rows sharded over several GPUs of THIS process (one host thread,
    rvl_group_connect / rvl_group_search): the form an R session uses with n.gpus > 1.  Returns the same
    dict as revealer_search (without the permutation extras)."""
    if target_match not in ("positive", "negative"):
        raise RevealerError(_lib.RVL_ERR_ARG, 'target.match must be "positive" or "negative"')
    negative = target_match == "negative"
    target = np.asarray(target, dtype=np.float64)
    m2 = np.asarray(m2)
    F, n = m2.shape
    world = len(devices)
    ctxs = []
    try:
        for r, dev in enumerate(devices):
            lo, cnt = shard_rows(F, world, r)
            c = RevealerContext(dev, n_grid=n_grid, bandwidth_mult=bandwidth_mult, bandwidth_shrink=bandwidth_shrink,
                                negative=negative)
            ctxs.append(c)
            c.set_shard(r, world, lo, F)
            if np.asarray(m2).dtype == np.float64:     # shards are loaded one after the other: each load may use all the cores
                c.set_host_threads(min(32, len(os.sched_getaffinity(0))))
            c.load_matrix(m2[lo:lo + cnt])
            c.set_targets(target)
            c.set_seed(np.zeros(n, dtype=np.uint8) if seed is None else seed)
        lib = ctxs[0]._lib
        arr = (C.c_void_p * world)(*[c._h for c in ctxs])
        ctxs[0]._ck(lib.rvl_group_connect(arr, world))
        rc = lib.rvl_group_search(arr, world, int(max_n_iter))
        if rc != 0:
            msgs = [lib.rvl_last_error(c._h).decode("utf-8", "replace") for c in ctxs]
            raise RevealerError(rc, "; ".join(m for m in msgs if m))
        parts = [c.search_fetch(max_n_iter) for c in ctxs]
        cmi = np.concatenate([p["cmi"][0] for p in parts], axis=1).T.copy()   # F x iters like R
        r0 = parts[0]
        out = dict(cmi=cmi, top=r0["top"][0], top_score=r0["top_score"][0], seed_iter=r0["seed_iter"][0],
                   cmi_seed=r0["cmi_seed"][0], ic_seed0=float(r0["ic_seed0"][0]))
        key = np.where(np.isnan(cmi), np.inf, cmi if negative else -cmi)
        out["order"] = np.stack([np.argsort(key[:, it], kind="stable") for it in range(max_n_iter)])
        return out
    finally:
        for c in ctxs:
            c.close()


def revealer_search_target_batch_multi_gpu(targets, m2, seed=None, max_n_iter=1, target_match="positive", devices=(0, 1),
                                           n_grid=25, bandwidth_mult=0.25, bandwidth_shrink=-0.75):
    """A batch of independent searches (one per target row of `targets`, e.g. 256 drug-response profiles against
    one matrix: BASELINE config C5) sharded by TARGET over several GPUs of this process: every device holds the
    whole bit matrix and its own contiguous block of targets, so nothing is exchanged at all ("prefer
    target-sharding when n_targets >= P").  All devices are enqueued before the first fetch.  Returns dict(cmi
    [T][iters][F], top, top_score, seed_iter, cmi_seed, ic_seed0) in target order -- equal to the T single runs."""
    if target_match not in ("positive", "negative"):
        raise RevealerError(_lib.RVL_ERR_ARG, 'target.match must be "positive" or "negative"')
    targets = np.ascontiguousarray(np.atleast_2d(np.asarray(targets, dtype=np.float64)))
    T, n = targets.shape
    world = len(devices)
    if T < world:
        raise RevealerError(_lib.RVL_ERR_ARG, "target sharding needs at least one target per device; shard rows instead")
    ctxs, spans = [], []
    try:
        for r, dev in enumerate(devices):
            lo, cnt = shard_rows(T, world, r)
            c = RevealerContext(dev, n_grid=n_grid, bandwidth_mult=bandwidth_mult, bandwidth_shrink=bandwidth_shrink,
                                negative=(target_match == "negative"))
            ctxs.append(c)
            spans.append((lo, cnt))
            c.load_matrix(m2)
            c.set_targets(targets[lo:lo + cnt])
            c.set_seed(np.zeros(n, dtype=np.uint8) if seed is None else seed)
            c.search_run(max_n_iter)                      # asynchronous: the devices run concurrently
        parts = [c.search_fetch(max_n_iter) for c in ctxs]
        return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}
    finally:
        for c in ctxs:
            c.close()


def p_adjust_fdr(p):
    """R's p.adjust(p, method = "fdr") (Benjamini-Hochberg): pmin(1, cummin(n / i * p[o]))[ro], o = decreasing order."""
    p = np.asarray(p, dtype=np.float64)
    n = len(p)
    o = np.argsort(-p, kind="stable")
    q = np.minimum(1.0, np.minimum.accumulate(n / np.arange(n, 0, -1) * p[o]))
    out = np.empty(n)
    out[o] = q
    return out


def read_gct(path):
    """MSIG.Gct2Frame, LIB:2616-2626: skip 2 lines, first column row names, second descriptions."""
    with open(path) as fh:
        fh.readline()
        fh.readline()
        header = fh.readline().rstrip("\n").split("\t")
        names = header[2:]
        rows, descs, data = [], [], []
        for line in fh:
            if not line.strip():
                continue
            p = line.rstrip("\n").split("\t")
            rows.append(p[0])
            descs.append(p[1])
            data.append([float(v) if v not in ("", "NA") else np.nan for v in p[2:]])
    return dict(ds=np.asarray(data, dtype=np.float64), row_names=rows, descs=descs, names=names)


def consolidate_identical_host(m2, rows2):
    """consolidate.identical.features = "identical", LIB:1727-1748, as the reference does it (numpy): order the
    rows by their pasted 0/1 strings (stable), keep the first row of every run of equal strings, append
    " (count)" to its name.  Returns (m2, rows2) in the reference's new row order."""
    m2 = np.asarray(m2)
    keys = [np.asarray(r != 0, dtype=np.uint8).tobytes() for r in m2]      # b"\x00\x01.." sorts like "01.."
    ind = sorted(range(len(keys)), key=lambda i: keys[i])                  # stable, like R's order()
    kept, counts = [], []
    for i in ind:
        if kept and keys[kept[-1]] == keys[i]:
            counts[-1] += 1
        else:
            kept.append(i)
            counts.append(1)
    return m2[kept], ["%s (%d)" % (rows2[i], c) for i, c in zip(kept, counts)]


def prepare_inputs(ds1, target_name, target_match="positive", ds2=None, seed_names=None, exclude_features=None,
                   count_thres_low=None, count_thres_high=None, consolidate_identical_features=False):
    """REVEALER.v1's preprocessing, LIB:1575-1720 (host side, numpy): drop NA-target samples, intersect the
    samples of ds1 and ds2, sort samples by target (decreasing unless target.match == "negative"), drop the
    target row from the features, count filter (seed rows always kept), seed = column-wise max of the seed
    rows, seed rows and excluded features removed.  ds1 / ds2: GCT path or dict(ds, row_names, names).
    Returns dict(target, m2, seed, seed_vectors, feature_names, sample_names, negative)."""
    d1 = read_gct(ds1) if isinstance(ds1, str) else ds1
    d2 = read_gct(ds2) if isinstance(ds2, str) else ds2
    m1 = np.asarray(d1["ds"], dtype=np.float64)
    names1 = list(d1["names"])
    rows1 = list(d1["row_names"])
    if target_name not in rows1:
        raise RevealerError(_lib.RVL_ERR_ARG, "target %r not found in ds1" % target_name)
    tgt = m1[rows1.index(target_name)]
    keep = ~np.isnan(tgt)                                   # LIB:1589-1594
    m1, names1 = m1[:, keep], [s for s, k in zip(names1, keep) if k]
    m2 = np.asarray(d2["ds"], dtype=np.float64)
    names2 = list(d2["names"])
    rows2 = list(d2["row_names"])
    pos2 = {s: i for i, s in enumerate(names2)}             # LIB:1596-1601 intersect
    overlap = [s for s in names1 if s in pos2]
    i1 = [names1.index(s) for s in overlap]
    i2 = [pos2[s] for s in overlap]
    m1, m2 = m1[:, i1], m2[:, i2]
    target = m1[rows1.index(target_name)]
    negative = target_match == "negative"                   # LIB:1626-1633
    ind = np.argsort(target if negative else -target, kind="stable")
    target, m2 = target[ind], m2[:, ind]
    sample_names = [overlap[i] for i in ind]
    if target_name in rows2:                                # LIB:1635-1638
        loc = rows2.index(target_name)
        m2 = np.delete(m2, loc, axis=0)
        rows2 = rows2[:loc] + rows2[loc + 1:]
    if seed_names is None or (isinstance(seed_names, str) and seed_names in ("NULL", "NULLSEED")):
        seed_list = []
    elif isinstance(seed_names, str):
        seed_list = [seed_names]
    else:
        seed_list = list(seed_names)
    if count_thres_low is not None and count_thres_high is not None:   # LIB:1655-1671
        sums = m2.sum(axis=1)
        retain = (sums >= count_thres_low) & (sums <= count_thres_high)
        for s in seed_list:
            if s in rows2:
                retain[rows2.index(s)] = True
        m2 = m2[retain]
        rows2 = [r for r, k in zip(rows2, retain) if k]
    if seed_list:                                           # LIB:1690-1707
        missing = [s for s in seed_list if s not in rows2]
        if missing:
            raise RevealerError(_lib.RVL_ERR_ARG, "seed(s) not found in ds2: %s" % ", ".join(missing))
        locs = [rows2.index(s) for s in seed_list]
        seed_vectors = m2[locs]
        seed = seed_vectors.max(axis=0)
        m2 = np.delete(m2, locs, axis=0)
        rows2 = [r for i, r in enumerate(rows2) if i not in set(locs)]
    else:
        seed = np.zeros(m2.shape[1])
        seed_vectors = seed[None, :]
    if exclude_features is not None:                        # LIB:1716-1720
        ex = {rows2.index(s) for s in exclude_features if s in rows2}
        m2 = np.delete(m2, sorted(ex), axis=0)
        rows2 = [r for i, r in enumerate(rows2) if i not in ex]
    if consolidate_identical_features == "identical":       # LIB:1727-1748
        m2, rows2 = consolidate_identical_host(m2, rows2)
    elif consolidate_identical_features:
        raise RevealerError(_lib.RVL_ERR_ARG, 'consolidate.identical.features: only F or "identical" (the "similar" '
                            'branch calls e1071::hamming.distance, which the reference never loads)')
    return dict(target=target, m2=m2, seed=seed, seed_vectors=seed_vectors, seed_list=seed_list, feature_names=rows2,
                sample_names=sample_names, negative=negative)


def prepare_inputs_device(ctx, ds1, target_name, target_match="positive", ds2=None, seed_names=None,
                          exclude_features=None, count_thres_low=None, count_thres_high=None,
                          consolidate_identical_features=False):
    """prepare_inputs with ds2 read by the GCT ingest (GctFile + rvl_load_matrix_gct): the feature matrix goes
    from the file to the bit-packed device matrix of `ctx` without an F x n host array.  The sample
    intersection and the sort-by-target permutation become the column map of the device tokeniser; the
    count filter uses the device row counts; seed / target / excluded rows are dropped with one device-side
    row selection.  Same return dict as prepare_inputs, with m2 = None (the matrix is resident in ctx)."""
    d1 = GctFile(ds1) if isinstance(ds1, str) else ds1
    g2 = ds2 if isinstance(ds2, GctFile) else GctFile(ds2)
    if isinstance(d1, GctFile):
        rows1, names1 = d1.row_names, list(d1.names)
        if target_name not in rows1:
            raise RevealerError(_lib.RVL_ERR_ARG, "target %r not found in ds1" % target_name)
        tgt = d1.row_values(rows1.index(target_name))
    else:
        rows1, names1 = list(d1["row_names"]), list(d1["names"])
        if target_name not in rows1:
            raise RevealerError(_lib.RVL_ERR_ARG, "target %r not found in ds1" % target_name)
        tgt = np.asarray(d1["ds"], dtype=np.float64)[rows1.index(target_name)]
    keep = ~np.isnan(tgt)                                   # LIB:1589-1594
    names1k = [s for s, k in zip(names1, keep) if k]
    tgt = tgt[keep]
    names2 = g2.names
    pos2 = {s: i for i, s in enumerate(names2)}             # LIB:1596-1601 intersect
    pos1 = {}
    for i, s in enumerate(names1k):
        pos1.setdefault(s, i)
    overlap = [s for s in names1k if s in pos2]
    target = tgt[[pos1[s] for s in overlap]]
    i2 = np.array([pos2[s] for s in overlap], dtype=np.int32)
    negative = target_match == "negative"                   # LIB:1626-1633
    ind = np.argsort(target if negative else -target, kind="stable")
    target = target[ind]
    sample_names = [overlap[i] for i in ind]
    ctx.load_gct(g2, cols=i2[ind])                          # every row of ds2, samples selected and ordered
    rows2 = list(g2.row_names)
    alive = np.ones(len(rows2), dtype=bool)
    if target_name in rows2:                                # LIB:1635-1638
        alive[rows2.index(target_name)] = False
    if seed_names is None or (isinstance(seed_names, str) and seed_names in ("NULL", "NULLSEED")):
        seed_list = []
    elif isinstance(seed_names, str):
        seed_list = [seed_names]
    else:
        seed_list = list(seed_names)

    def find(name):  # first live row with that name (R's match())
        for i, r in enumerate(rows2):
            if alive[i] and r == name:
                return i
        return -1

    if count_thres_low is not None and count_thres_high is not None:   # LIB:1655-1671
        sums = ctx.row_counts()
        retain = (sums >= count_thres_low) & (sums <= count_thres_high)
        for s in seed_list:
            i = find(s)
            if i >= 0:
                retain[i] = True
        alive &= retain
    if seed_list:                                           # LIB:1690-1707
        locs = [find(s) for s in seed_list]
        missing = [s for s, i in zip(seed_list, locs) if i < 0]
        if missing:
            raise RevealerError(_lib.RVL_ERR_ARG, "seed(s) not found in ds2: %s" % ", ".join(missing))
        seed_vectors = ctx.get_rows(locs).astype(np.float64)
        seed = seed_vectors.max(axis=0)
        alive[locs] = False
    else:
        seed = np.zeros(len(target))
        seed_vectors = seed[None, :]
    if exclude_features is not None:                        # LIB:1716-1720
        for s in exclude_features:
            i = find(s)
            if i >= 0:
                alive[i] = False
    idx = np.flatnonzero(alive)
    ctx.select_rows(idx)
    names = [rows2[i] for i in idx]
    if consolidate_identical_features == "identical":       # LIB:1727-1748, on the device
        kept, counts = ctx.consolidate_identical()
        names = ["%s (%d)" % (names[i], c) for i, c in zip(kept, counts)]
    elif consolidate_identical_features:
        raise RevealerError(_lib.RVL_ERR_ARG, 'consolidate.identical.features: only F or "identical"')
    return dict(target=target, m2=None, seed=seed, seed_vectors=seed_vectors, seed_list=seed_list,
                feature_names=names, sample_names=sample_names, negative=negative)


def REVEALER_v1(ds1, target_name, target_match="positive", ds2=None, seed_names=None, exclude_features=None,
                max_n_iter=5, pdf_output_file=None, count_thres_low=None, count_thres_high=None, n_markers=30,
                locs_table_file=None, n_grid=25, bandwidth_mult=0.25, bandwidth_shrink=-0.75, n_perm=10,
                r_seed=34578, device=0, ingest="host", consolidate_identical_features=False):
    """REVEALER.v1 with its 12 formals (LIB:1504-1516) plus the optional grid/bandwidth arguments.
    ds1 / ds2: GCT path or dict(ds, row_names, names) as MSIG.Gct2Frame returns.  Preprocessing
    mirrors LIB:1575-1707: drop NA-target samples, intersect samples, sort by target, drop the target
    row from the features, count filter (seeds always kept), build the seed as the column-wise max of
    the seed rows and remove those rows.  No plots/PDF (pdf_output_file and locs_table_file are accepted
    and ignored).  ingest: "host" (default, the drop-in behaviour) reads ds2 like MSIG.Gct2Frame and builds the F x n
    host matrix first; "device" opts into the device GCT tokeniser, which is faster but accepts only 0/1 spelled with
    an optional ".000" (read.delim also takes NA, exponents, padded fields) -- a file it rejects raises an error.  consolidate_identical_features: the
    reference's internal switch (LIB:1532, default F); "identical" runs LIB:1727-1748 (on the device with
    ingest="device").  n_perm (default 10, the reference's internal n.perm, LIB:1519) adds the permutation null with numpy's
    PCG64(r_seed) -- R's sample() stream is not reproduced, so p-values match R's in distribution, not draw by draw."""
    ctx = None
    if ingest not in ("host", "device"):
        raise RevealerError(_lib.RVL_ERR_ARG, 'ingest must be "host" or "device"')
    if ingest == "device":   # ds2: file -> text on the device -> bit matrix (no F x n host array)
        ctx = RevealerContext(device, n_grid=n_grid, bandwidth_mult=bandwidth_mult, bandwidth_shrink=bandwidth_shrink,
                              negative=(target_match == "negative"))
        try:
            p = prepare_inputs_device(ctx, ds1, target_name, target_match, ds2, seed_names, exclude_features,
                                      count_thres_low, count_thres_high, consolidate_identical_features)
        except Exception:
            ctx.close()
            raise
    else:
        p = prepare_inputs(ds1, target_name, target_match, ds2, seed_names, exclude_features, count_thres_low,
                           count_thres_high, consolidate_identical_features)
    target, m2, seed, seed_vectors, seed_list = p["target"], p["m2"], p["seed"], p["seed_vectors"], p["seed_list"]
    rows2, sample_names, negative = p["feature_names"], p["sample_names"], p["negative"]
    target_rand = None
    if n_perm > 0:                                          # LIB:1843-1844
        rng = np.random.Generator(np.random.PCG64(r_seed))
        target_rand = np.stack([rng.permutation(target) for _ in range(n_perm)])
    try:
        out = revealer_search(target, m2, seed, max_n_iter=max_n_iter, target_match=target_match, n_grid=n_grid,
                              bandwidth_mult=bandwidth_mult, bandwidth_shrink=bandwidth_shrink,
                              target_rand=target_rand, n_markers=n_markers, device=device, ctx=ctx)
    finally:
        if ctx is not None:
            ctx.close()
    out["feature_names"] = rows2
    out["sample_names"] = sample_names
    out["target"] = target
    out["seed_names_iter"] = [rows2[i] for i in out["top"]]       # seed.names.iter  LIB:2167
    med = np.median(target)                                         # LIB:1810-1815
    locs_t = (target <= med) if negative else (target > med)
    out["pct_explained_seed"] = out["seed_iter"][:, locs_t].sum(axis=1) / max(1, locs_t.sum())  # LIB:2171
    # seed baselines LIB:1817-1833 (one assoc per seed row and per cumulative OR)
    if seed_list:
        with RevealerContext(device, n_grid=n_grid, bandwidth_mult=bandwidth_mult,
                             bandwidth_shrink=bandwidth_shrink) as c2:
            c2.set_targets(target)
            cum = np.zeros_like(seed)
            out["cmi_orig_seed"], out["cmi_orig_seed_cum"] = [], []
            for v in seed_vectors:
                out["cmi_orig_seed"].append(c2.assoc(v))
                cum = np.maximum(cum, v)
                out["cmi_orig_seed_cum"].append(c2.assoc(cum))
    return out
