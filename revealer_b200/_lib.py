"""ctypes binding of librevealer_b200.so (include/revealer_b200.h).

The library is the product; this module only loads it.  If the .so is missing or a symbol is
absent the import of the hot path fails loudly -- there is no Python/CPU fallback.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("RVL_LIB") or os.path.join(HERE, "librevealer_b200.so")  # RVL_LIB: tuning builds only

RVL_OK = 0
RVL_ERR_CUDA = -1
RVL_ERR_ARG = -2
RVL_ERR_NOT_BINARY = -3
RVL_ERR_STATE = -4
RVL_ERR_NONFINITE = -5


class rvl_opts(C.Structure):
    _fields_ = [("n_grid", C.c_int32), ("negative", C.c_int32),
                ("bandwidth_mult", C.c_double), ("bandwidth_shrink", C.c_double)]


class rvl_result(C.Structure):
    _fields_ = [("cmi", C.POINTER(C.c_double)), ("top", C.POINTER(C.c_int64)),
                ("top_score", C.POINTER(C.c_double)), ("seed_iter", C.POINTER(C.c_uint8)),
                ("cmi_seed", C.POINTER(C.c_double)), ("ic_seed0", C.POINTER(C.c_double))]


_vp, _i, _i64, _dp = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_double)
_u8p, _u64p, _i32p, _i64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_int64)

# name -> (restype, argtypes); must list every symbol include/revealer_b200.h declares
SIGNATURES = {
    "rvl_abi_version": (_i, []),
    "rvl_create": (_i, [C.POINTER(_vp), _i]),
    "rvl_destroy": (None, [_vp]),
    "rvl_last_error": (C.c_char_p, [_vp]),
    "rvl_set_stream": (_i, [_vp, _vp]),
    "rvl_set_options": (_i, [_vp, C.POINTER(rvl_opts)]),
    "rvl_default_options": (_i, [C.POINTER(rvl_opts)]),
    "rvl_set_shard": (_i, [_vp, _i, _i, _i64, _i64]),
    "rvl_load_matrix_f64_colmajor": (_i, [_vp, _dp, _i64, _i64, _i64]),
    "rvl_set_host_threads": (_i, [_vp, _i]),
    "rvl_get_upload_stats": (_i, [_vp, _u64p, _u64p]),
    "rvl_load_matrix_u8_rowmajor": (_i, [_vp, _u8p, _i64, _i64]),
    "rvl_load_matrix_bits": (_i, [_vp, _u64p, _i64, _i64, _i64]),
    "rvl_load_matrix_bits_device": (_i, [_vp, _vp, _i64, _i64, _i64]),
    "rvl_gct_open": (_i, [C.c_char_p, C.POINTER(_vp)]),
    "rvl_gct_close": (None, [_vp]),
    "rvl_gct_last_error": (C.c_char_p, [_vp]),
    "rvl_gct_dims": (_i, [_vp, _i64p, _i64p]),
    "rvl_gct_sample_name": (_i, [_vp, _i64, C.c_char_p, C.c_size_t]),
    "rvl_gct_row_name": (_i, [_vp, _i64, C.c_char_p, C.c_size_t]),
    "rvl_gct_row_description": (_i, [_vp, _i64, C.c_char_p, C.c_size_t]),
    "rvl_gct_row_names": (_i64, [_vp, _i, C.c_char_p, C.c_size_t]),
    "rvl_gct_row_values": (_i, [_vp, _i64, _dp]),
    "rvl_load_matrix_gct": (_i, [_vp, _vp, _i32p, _i64, _i64, _i64]),
    "rvl_select_rows": (_i, [_vp, _i64p, _i64]),
    "rvl_consolidate_identical": (_i, [_vp, _i64p, _i32p, _i64p]),
    "rvl_get_rows_u8": (_i, [_vp, _i64p, _i64, _u8p]),
    "rvl_set_targets": (_i, [_vp, _dp, _i64, _i64]),
    "rvl_set_targets_permuted": (_i, [_vp, _dp, _i64, _i64]),
    "rvl_set_null_targets": (_i, [_vp, _i64]),
    "rvl_set_seed": (_i, [_vp, _u8p, _i64, _i]),
    "rvl_score_once": (_i, [_vp, _dp]),
    "rvl_search": (_i, [_vp, _i, C.POINTER(rvl_result)]),
    "rvl_p2p_export": (_i, [_vp, _u8p]),
    "rvl_p2p_import": (_i, [_vp, _u8p, _i]),
    "rvl_p2p_disable": (_i, [_vp]),
    "rvl_group_connect": (_i, [C.POINTER(_vp), _i]),
    "rvl_group_search": (_i, [C.POINTER(_vp), _i, _i]),
    "rvl_search_run": (_i, [_vp, _i]),
    "rvl_set_fused": (_i, [_vp, _i]),
    "rvl_search_begin": (_i, [_vp, _i]),
    "rvl_candidate_bytes": (C.c_size_t, [_vp]),
    "rvl_set_exchange_buffers": (_i, [_vp, _vp, _vp]),
    "rvl_iter_score": (_i, [_vp, _i]),
    "rvl_iter_merge": (_i, [_vp, _i]),
    "rvl_search_fetch": (_i, [_vp, C.POINTER(rvl_result)]),
    "rvl_synchronize": (_i, [_vp]),
    "rvl_assoc": (_i, [_vp, _i64, _u8p, _i64, _dp]),
    "rvl_pvalues": (_i, [_vp, _dp, _i64, _dp, _i64, _dp]),
    "rvl_pvalues_iter": (_i, [_vp, _i, _dp]),
    "rvl_topk": (_i, [_vp, _i, _i64, _i64, _i64p, _dp]),
    "rvl_get_target_bandwidth": (_i, [_vp, _i64, _dp]),
    "rvl_get_binary_bandwidth_table": (_i, [_vp, _dp]),
    "rvl_get_matrix_dims": (_i, [_vp, _i64p, _i64p]),
    "rvl_get_row_counts": (_i, [_vp, _i32p]),
    "rvl_launch_count": (_i64, [_vp]),
    "rvl_set_profiling": (_i, [_vp, _i]),
    "rvl_score_kernel_time": (_i, [_vp, _dp, _i64p, _i]),
    "rvl_set_dense_x": (_i, [_vp, _i]),
    "rvl_set_prune_theta": (_i, [_vp, C.c_double]),
    "rvl_set_dedup": (_i, [_vp, _i]),
    "rvl_unique_rows": (_i64, [_vp]),
    "rvl_get_counters": (_i, [_vp, _u64p, _i]),
    "rvl_get_counters_ex": (_i, [_vp, _u64p, _i, _i]),
    "rvl_measure_fp64_peak": (_i, [_vp, _dp]),
    "rvl_debug_phase_times": (_i, [_vp, _u64p, _i]),
}

_LIB = None


class RevealerLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library; raises RevealerLibraryError when it is missing (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise RevealerLibraryError(
            "librevealer_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`"
            " -- the CUDA extension is the only implementation of the hot path; there is no CPU fallback." % SO_PATH)
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RevealerLibraryError("librevealer_b200.so lacks symbol %s" % name) from e
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
