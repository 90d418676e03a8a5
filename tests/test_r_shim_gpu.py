"""The .Call shim (revealer_b200/r/r_shim.c) EXECUTED on the GPU against a miniature stand-in for the R runtime.

R is not installed here or on the GPU box, so the shim cannot be loaded into R.  tests/rstub/rstub_runtime.c implements
the ~30 R API entry points the shim uses (vectors, column-major matrices, named lists, dim attributes, external
pointers with finalizers, R_alloc, Rf_error as a longjmp); this test links the shim's real object code with it and
librevealer_b200.so and drives every entry point the way REVEALER_b200.R does: layouts (column-major m.2, n.perm x n
target.rand, F x n.perm x iter cmi.rand), 1-based indices, result list names, error conversion (no crash, message
from rvl_last_error), PROTECT balance and the finalizers.  Results must equal the Python host's (same C ABI).
It is still not R: the wrapper REVEALER_b200.R itself has never been run.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    from revealer_b200 import build as rb_build
    rb_build.build()
    out = str(tmp_path_factory.mktemp("rshim") / "revealer_b200_r_stub.so")
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    libdir = os.path.join(ROOT, "revealer_b200")
    cmd = [gcc, "-O1", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", os.path.join(libdir, "r", "r_shim.c"),
           os.path.join(ROOT, "tests", "rstub", "rstub_runtime.c"), "-I", os.path.join(ROOT, "tests", "rstub"),
           "-I", os.path.join(ROOT, "include"), "-L", libdir, "-lrevealer_b200", "-Wl,-rpath," + libdir, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    L = C.CDLL(out)
    vp = C.c_void_p
    for name, res, args in [("rstub_nil", vp, []), ("rstub_real", vp, [C.POINTER(C.c_double), C.c_long]),
                            ("rstub_real_matrix", vp, [C.POINTER(C.c_double), C.c_int, C.c_int]),
                            ("rstub_int", vp, [C.POINTER(C.c_int), C.c_long]), ("rstub_logical", vp, [C.c_int]),
                            ("rstub_string", vp, [C.c_char_p]), ("rstub_type", C.c_int, [vp]), ("rstub_length", C.c_long, [vp]),
                            ("rstub_nrow", C.c_int, [vp]), ("rstub_ncol", C.c_int, [vp]), ("rstub_dim", C.c_long, [vp, C.c_int]),
                            ("rstub_real_ptr", C.POINTER(C.c_double), [vp]), ("rstub_int_ptr", C.POINTER(C.c_int), [vp]),
                            ("rstub_elt", vp, [vp, C.c_long]), ("rstub_chars", C.c_char_p, [vp, C.c_long]),
                            ("rstub_name", C.c_char_p, [vp, C.c_long]), ("rstub_last_error", C.c_char_p, []),
                            ("rstub_protect_balance", C.c_int, []), ("rstub_call", vp, [vp, C.c_int, C.POINTER(vp)]),
                            ("rstub_reset", C.c_int, [])]:
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    yield R(L)
    L.rstub_reset()


class R:
    """The few things an R session does with the shim: build arguments, .Call, read the result."""

    def __init__(self, L):
        self.L = L

    def real(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64).ravel()
        return self.L.rstub_real(v.ctypes.data_as(C.POINTER(C.c_double)), len(v))

    def matrix(self, m):       # R stores matrices column-major
        m = np.asfortranarray(m, dtype=np.float64)
        return self.L.rstub_real_matrix(m.ctypes.data_as(C.POINTER(C.c_double)), m.shape[0], m.shape[1])

    def integer(self, v):
        v = np.ascontiguousarray(np.atleast_1d(v), dtype=np.int32)
        return self.L.rstub_int(v.ctypes.data_as(C.POINTER(C.c_int)), len(v))

    def call(self, name, *args):
        fn = C.cast(getattr(self.L, name), C.c_void_p)
        arr = (C.c_void_p * max(len(args), 1))(*args)
        out = self.L.rstub_call(fn, len(args), arr)
        assert self.L.rstub_protect_balance() == 0, "PROTECT / UNPROTECT out of balance in " + name
        if not out:
            raise RuntimeError(self.L.rstub_last_error().decode())
        return out

    def as_real(self, s):
        n = self.L.rstub_length(s)
        a = np.ctypeslib.as_array(self.L.rstub_real_ptr(s), shape=(n,)).copy()
        if self.L.rstub_dim(s, 2) >= 0:
            return a.reshape([self.L.rstub_dim(s, i) for i in range(3)], order="F")
        if self.L.rstub_nrow(s) * self.L.rstub_ncol(s) == n and n:
            return a.reshape((self.L.rstub_nrow(s), self.L.rstub_ncol(s)), order="F")
        return a

    def as_int(self, s):
        return np.ctypeslib.as_array(self.L.rstub_int_ptr(s), shape=(self.L.rstub_length(s),)).copy()

    def as_list(self, s):
        n = self.L.rstub_length(s)
        return {self.L.rstub_name(s, i).decode(): self.L.rstub_elt(s, i) for i in range(n)}


def _workload(n, F):
    from revealer_b200 import synth
    return synth.make_workload(n=n, F=F)


def test_rvlr_search_matches_the_python_host(shim):
    import revealer_b200 as rb
    w = _workload(90, 400)
    rng = np.random.default_rng(2)
    perm = np.stack([rng.permutation(w["target"]) for _ in range(3)])
    iters = 3
    ref = rb.revealer_search(w["target"], w["m2"], w["seed"], max_n_iter=iters, target_rand=perm, n_markers=25)
    ctx = shim.call("rvlR_create", shim.integer(0))
    res = shim.as_list(shim.call("rvlR_search", ctx, shim.real(w["target"]), shim.matrix(w["m2"]), shim.real(w["seed"]),
                                 shim.integer(iters), shim.L.rstub_logical(0), shim.integer(25), shim.real([0.25]),
                                 shim.real([-0.75]), shim.matrix(perm)))
    assert list(res) == ["cmi", "top", "top.score", "seed.iter", "cmi.seed", "cmi.seed.iter0", "cmi.rand"]
    cmi = shim.as_real(res["cmi"])
    assert cmi.shape == (400, iters) and np.array_equal(cmi, ref["cmi"], equal_nan=True)       # F x iter, rows of m.2
    assert list(shim.as_int(res["top"])) == [t + 1 for t in ref["top"]]                          # 1-based
    assert np.array_equal(shim.as_real(res["top.score"]), ref["top_score"])
    assert np.array_equal(shim.as_real(res["seed.iter"]), ref["seed_iter"].astype(float))        # iter x n
    assert np.array_equal(shim.as_real(res["cmi.seed"]), ref["cmi_seed"])
    assert shim.as_real(res["cmi.seed.iter0"])[0] == ref["ic_seed0"]
    cr = shim.as_real(res["cmi.rand"])
    assert cr.shape == (400, 3, iters)                                                           # F x n.perm x iter
    assert np.array_equal(np.transpose(cr, (2, 0, 1)), ref["cmi_rand"], equal_nan=True)
    for it in range(iters):                 # p-values and display ranking from the scores still resident on the device
        p = shim.as_real(shim.call("rvlR_pvalues_iter", ctx, shim.integer(it + 1)))
        assert np.array_equal(p, ref["p_val"][it], equal_nan=True)
        tk = shim.as_int(shim.call("rvlR_topk", ctx, shim.integer(it + 1), shim.integer(25)))
        assert list(tk) == [r + 1 for r in ref["top_markers"][it]]
        p2 = shim.as_real(shim.call("rvlR_pvalues", ctx, shim.real(cmi[:, it]), shim.real(cr[:, :, it].ravel(order="F"))))
        assert np.array_equal(p2, p, equal_nan=True)
    ic = shim.as_real(shim.call("rvlR_assoc", ctx, shim.real(w["target"]), shim.real(w["m2"][7])))[0]
    assert ic == rb.assoc(w["target"], w["m2"][7])


def test_rvlr_assess_matches_the_python_host(shim):
    """rvlR_assess (REVEALER_assess_features.v1's IC + permutation scores, LIB:2875 / 2885): F x n.perm layout, the target
    in row 0 of every chunk, a NULL target.rand, and the resident-matrix form."""
    import revealer_b200 as rb
    w = _workload(80, 250)
    rng = np.random.default_rng(5)
    perm = np.stack([rng.permutation(w["target"]) for _ in range(7)])
    ref = rb.assess_features(w["target"], w["m2"], target_rand=perm)
    ctx = shim.call("rvlR_create", shim.integer(0))
    res = shim.as_list(shim.call("rvlR_assess", ctx, shim.real(w["target"]), shim.matrix(w["m2"]), shim.matrix(perm),
                                 shim.integer(25)))
    assert list(res) == ["IC", "null.IC"]
    ic, null = shim.as_real(res["IC"]), shim.as_real(res["null.IC"])
    assert null.shape == (250, 7)
    assert np.array_equal(ic, ref["IC"], equal_nan=True) and np.array_equal(null.T, ref["null_IC"], equal_nan=True)
    res0 = shim.as_list(shim.call("rvlR_assess", ctx, shim.real(w["target"]), shim.L.rstub_nil(), shim.L.rstub_nil(),
                                  shim.integer(25)))                  # resident matrix, no permutations
    assert np.array_equal(shim.as_real(res0["IC"]), ic, equal_nan=True) and shim.L.rstub_type(res0["null.IC"]) == 0
    with pytest.raises(RuntimeError, match="n.perm x n"):
        shim.call("rvlR_assess", ctx, shim.real(w["target"]), shim.L.rstub_nil(), shim.matrix(perm[:, :-1]), shim.integer(25))


def test_rvlr_errors_become_r_errors_not_crashes(shim):
    w = _workload(60, 50)
    ctx = shim.call("rvlR_create", shim.integer(0))
    bad = w["m2"].astype(float)
    bad[3, 5] = 0.5                                                   # not binary: RVL_ERR_NOT_BINARY -> Rf_error
    args = lambda m, seed=w["seed"], it=2: (ctx, shim.real(w["target"]), shim.matrix(m), shim.real(seed), shim.integer(it),
                                            shim.L.rstub_logical(0), shim.integer(25), shim.real([0.25]), shim.real([-0.75]),
                                            shim.L.rstub_nil())
    with pytest.raises(RuntimeError, match="not 0 or 1"):
        shim.call("rvlR_search", *args(bad))
    with pytest.raises(RuntimeError, match="seed must be 0/1"):
        shim.call("rvlR_search", *args(w["m2"], seed=w["seed"] * 2.0))
    with pytest.raises(RuntimeError, match="dimension mismatch"):
        shim.call("rvlR_search", *args(w["m2"][:, :-1]))
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        shim.call("rvlR_create", shim.integer(4096))
    with pytest.raises(RuntimeError):                                 # iteration out of range -> status, not a fault
        shim.call("rvlR_pvalues_iter", ctx, shim.integer(99))
    ok = shim.as_list(shim.call("rvlR_search", *args(w["m2"])))       # the context is still usable after the errors
    assert shim.L.rstub_length(ok["top"]) == 2
    assert shim.L.rstub_reset() >= 1                                  # finalizers ran: rvl_destroy through R's "GC"


def test_rvlr_search_multi_shards_the_r_matrix_in_place(shim):
    """rvlR_search_multi (REVEALER.v1's n.gpus): row blocks of the column-major matrix passed in place (ld = F).  With one
    visible GPU it still runs the group path with a world of one; with two, the real peer exchange."""
    import torch
    import revealer_b200 as rb
    w = _workload(120, 1501)
    rng = np.random.default_rng(3)
    perm = np.stack([rng.permutation(w["target"]) for _ in range(2)])
    iters = 3
    ref = rb.revealer_search(w["target"], w["m2"], w["seed"], max_n_iter=iters, target_rand=perm)
    for devices in ([0], [0, 1]):
        if len(devices) > torch.cuda.device_count():
            continue
        res = shim.as_list(shim.call("rvlR_search_multi", shim.integer(devices), shim.real(w["target"]), shim.matrix(w["m2"]),
                                     shim.real(w["seed"]), shim.integer(iters), shim.L.rstub_logical(0), shim.integer(25),
                                     shim.real([0.25]), shim.real([-0.75]), shim.matrix(perm)))
        assert np.array_equal(shim.as_real(res["cmi"]), ref["cmi"], equal_nan=True)
        assert list(shim.as_int(res["top"])) == [t + 1 for t in ref["top"]]
        assert np.array_equal(shim.as_real(res["seed.iter"]), ref["seed_iter"].astype(float))
        assert np.array_equal(shim.as_real(res["cmi.seed"]), ref["cmi_seed"])
        assert np.array_equal(np.transpose(shim.as_real(res["cmi.rand"]), (2, 0, 1)), ref["cmi_rand"], equal_nan=True)


def test_rvlr_gct_ingest_path(shim, tmp_path):
    """The opt-in GCT path of the wrapper: open, info, target row, load with a column map, seed rows, row selection,
    identical-row consolidation, then the search on the resident matrix (m.2 = NULL)."""
    import revealer_b200 as rb
    w = _workload(70, 300)
    m2 = w["m2"].copy()
    m2[11] = m2[10]                                                   # a pair of identical rows
    path = str(tmp_path / "ds2.gct")
    with open(path, "w") as fh:
        fh.write("#1.2\n%d\t%d\nName\tDescription\t%s\n" % (300, 70, "\t".join("S%d" % j for j in range(70))))
        for i in range(300):
            fh.write("f%d\tna\t%s\n" % (i, "\t".join(str(int(v)) for v in m2[i])))
    g = shim.call("rvlR_gct_open", shim.L.rstub_string(path.encode()))
    info = shim.as_list(shim.call("rvlR_gct_info", g))
    assert shim.L.rstub_chars(info["row.names"], 299) == b"f299" and shim.L.rstub_chars(info["names"], 69) == b"S69"
    row = shim.as_real(shim.call("rvlR_gct_row", g, shim.integer(8)))
    assert np.array_equal(row, m2[7].astype(float))                  # 1-based row
    ctx = shim.call("rvlR_create", shim.integer(0))
    cols = np.arange(70, 0, -1)                                       # 1-based column map: samples reversed
    counts = shim.as_int(shim.call("rvlR_load_gct", ctx, g, shim.integer(cols)))
    assert np.array_equal(counts, m2.sum(axis=1))
    rows = shim.as_real(shim.call("rvlR_get_rows", ctx, shim.integer([3, 11])))
    assert np.array_equal(rows, m2[[2, 10]][:, ::-1].astype(float))
    keep = np.arange(2, 301)                                          # drop the first row
    shim.call("rvlR_select_rows", ctx, shim.integer(keep))
    cons = shim.as_list(shim.call("rvlR_consolidate_identical", ctx))
    kept, cnt = shim.as_int(cons["rows"]), shim.as_int(cons["counts"])
    uniq = len({r.tobytes() for r in m2[1:]})
    assert len(kept) == uniq and cnt.sum() == 299 and cnt.max() >= 2
    target, seed = w["target"][::-1].copy(), w["seed"][::-1].copy()
    res = shim.as_list(shim.call("rvlR_search", ctx, shim.real(target), shim.L.rstub_nil(), shim.real(seed), shim.integer(2),
                                 shim.L.rstub_logical(0), shim.integer(25), shim.real([0.25]), shim.real([-0.75]),
                                 shim.L.rstub_nil()))
    sub = m2[1:][:, ::-1][kept - 1]                                   # the matrix the device now holds
    ref = rb.revealer_search(target, np.ascontiguousarray(sub), seed, max_n_iter=2)
    assert np.array_equal(shim.as_real(res["cmi"]), ref["cmi"], equal_nan=True)
    assert list(shim.as_int(res["top"])) == [t + 1 for t in ref["top"]]
