"""Device GCT ingest (rvl_load_matrix_gct: text -> device tokeniser -> bit matrix), row selection and the
device form of REVEALER.v1's preprocessing, against the host path (numpy matrix -> rvl_load_matrix_*)."""
import numpy as np
import pytest

import revealer_b200 as rb
from revealer_b200 import _lib
from test_gct_cpu import write_gct

pytestmark = pytest.mark.gpu


def fast_gct(path, m, names=None, row_prefix="f"):
    """Write a 0/1 matrix as GCT quickly (numpy byte ops; one '0'/'1' character per value)."""
    F, n = m.shape
    names = names or ["S%d" % j for j in range(n)]
    body = np.empty((F, 2 * n), dtype=np.uint8)
    body[:, 0::2] = ord("\t")
    body[:, 1::2] = m.astype(np.uint8) + ord("0")
    with open(path, "wb") as fh:
        fh.write(("#1.2\n%d\t%d\nName\tDescription\t%s\n" % (F, n, "\t".join(names))).encode())
        for i in range(F):
            fh.write(("%s%d\td%d" % (row_prefix, i, i)).encode() + body[i].tobytes() + b"\n")


def matrix_of(ctx):
    return ctx.get_rows(np.arange(ctx.F))


@pytest.mark.parametrize("F,n", [(1, 2), (37, 64), (300, 65), (1000, 333)])
def test_device_ingest_equals_host_matrix(tmp_path, F, n):
    rng = np.random.default_rng(F * 1000 + n)
    m = (rng.random((F, n)) < 0.2).astype(np.uint8)
    p = str(tmp_path / "m.gct")
    fast_gct(p, m)
    with rb.GctFile(p) as g, rb.RevealerContext(0) as ctx:
        ctx.load_gct(g)
        assert np.array_equal(matrix_of(ctx), m)
        assert np.array_equal(ctx.row_counts(), m.sum(axis=1))
        cols = rng.permutation(n)[: max(2, n // 2)].astype(np.int32)      # sample subset, permuted
        ctx.load_gct(g, cols=cols)
        assert np.array_equal(matrix_of(ctx), m[:, cols])
        if F > 10:
            ctx.load_gct(g, cols=cols, first_row=7, n_rows=F - 10)         # a shard's row block
            assert np.array_equal(matrix_of(ctx), m[7:F - 3][:, cols])


def test_device_ingest_token_forms_line_endings_and_errors(tmp_path):
    rng = np.random.default_rng(11)
    F, n = 40, 21
    m = (rng.random((F, n)) < 0.3).astype(int)
    rows, descs, names = ["r%d" % i for i in range(F)], ["d"] * F, ["c%d" % j for j in range(n)]
    p = str(tmp_path / "t.gct")
    fmts = [lambda v: "%d" % v, lambda v: "%.1f" % v, lambda v: "%.3f" % v, lambda v: "%d." % v]
    for k, (eol, trailing, blank) in enumerate([("\n", True, 0), ("\r\n", True, 0), ("\n", False, 0), ("\n", True, 4)]):
        write_gct(p, rows, descs, names, m.tolist(), eol=eol, fmt=fmts[k], trailing_newline=trailing, blank_every=blank)
        with rb.GctFile(p) as g, rb.RevealerContext(0) as ctx:
            ctx.load_gct(g)
            assert np.array_equal(matrix_of(ctx), m)
    for badval in ("2", "0.5", "NA", "", "1e0", "-1", " 1"):
        bad = [[str(v) for v in row] for row in m.tolist()]
        bad[17][n - 1 if badval == "" else 5] = badval
        write_gct(p, rows, descs, names, bad, fmt=str)
        with rb.GctFile(p) as g, rb.RevealerContext(0) as ctx:
            with pytest.raises(rb.RevealerError) as e:
                ctx.load_gct(g)
            assert e.value.code == _lib.RVL_ERR_NOT_BINARY, badval
            with pytest.raises(rb.RevealerError):                         # nothing half-loaded is left behind
                ctx.row_counts()
    short = [[str(v) for v in row] for row in m.tolist()]
    short[3] = short[3][:-1]
    write_gct(p, rows, descs, names, short, fmt=str)
    with rb.GctFile(p) as g, rb.RevealerContext(0) as ctx:
        with pytest.raises(rb.RevealerError, match="fields"):
            ctx.load_gct(g)
    with rb.GctFile(p) as g, rb.RevealerContext(0) as ctx:
        with pytest.raises(rb.RevealerError):
            ctx.load_gct(g, cols=np.array([0, 0], dtype=np.int32))          # repeated column


def test_device_ingest_spans_several_text_chunks(tmp_path):
    """> 64 MB of text: the file is cut into line-aligned chunks; row numbering must carry across them."""
    rng = np.random.default_rng(2)
    F, n = 60000, 700                                   # ~84 MB of text
    m = (rng.random((F, n)) < 0.03).astype(np.uint8)
    p = str(tmp_path / "big.gct")
    fast_gct(p, m)
    with rb.GctFile(p) as g, rb.RevealerContext(0) as ctx:
        ctx.load_gct(g)
        assert np.array_equal(ctx.row_counts(), m.sum(axis=1))
        pick = np.array([0, 1, 29999, 30000, 47000, 47001, F - 1])
        assert np.array_equal(ctx.get_rows(pick), m[pick])
        ctx.load_gct(g, first_row=46990, n_rows=100)
        assert np.array_equal(matrix_of(ctx), m[46990:47090])


def test_select_rows_matches_numpy_indexing():
    rng = np.random.default_rng(9)
    m = (rng.random((500, 130)) < 0.1).astype(np.uint8)
    with rb.RevealerContext(0) as ctx:
        ctx.load_matrix(m)
        idx = np.flatnonzero(rng.random(500) < 0.4)
        ctx.select_rows(idx)
        assert ctx.F == len(idx) and np.array_equal(matrix_of(ctx), m[idx])
        assert ctx.unique_rows() == len(np.unique(m[idx], axis=0))
        ctx.select_rows(np.array([3, 0, 2]))
        assert np.array_equal(matrix_of(ctx), m[idx][[3, 0, 2]])
        with pytest.raises(rb.RevealerError):
            ctx.select_rows(np.array([5]))


def test_revealer_v1_device_ingest_equals_host_ingest(tmp_path):
    """REVEALER_v1 on GCT files: ds2 through the device tokeniser (no F x n host matrix) gives the same
    preprocessing and the same search as the host reader (read_gct + prepare_inputs)."""
    from revealer_b200 import synth
    w = synth.make_workload(n=150, F=1200)
    n = 150
    rng = np.random.default_rng(4)
    names2 = ["P%03d" % i for i in range(n)]
    perm = rng.permutation(n)
    names1 = [names2[i] for i in perm[:140]] + ["only_in_ds1_a", "only_in_ds1_b"]     # partial overlap, other order
    tvals = np.concatenate([w["target"][perm[:140]], [0.3, np.nan]])
    tvals[7] = np.nan                                                                  # an NA target sample
    f1, f2 = str(tmp_path / "ds1.gct"), str(tmp_path / "ds2.gct")
    write_gct(f1, ["OTHER", "TGT"], ["x", "y"], names1, [np.arange(len(names1), dtype=float).tolist(), tvals.tolist()])
    m2 = np.vstack([w["m2"], w["seed"][None, :]]).astype(np.uint8)
    fast_gct(f2, m2, names=names2, row_prefix="feat")
    seed_name = "feat%d" % (m2.shape[0] - 1)
    kw = dict(seed_names=[seed_name, "feat3"], exclude_features=["feat10", "feat11"], count_thres_low=3, count_thres_high=50)
    a = rb.REVEALER_v1(f1, "TGT", "positive", f2, max_n_iter=3, ingest="host", **kw)
    b = rb.REVEALER_v1(f1, "TGT", "positive", f2, max_n_iter=3, ingest="device", **kw)
    assert a["feature_names"] == b["feature_names"] and a["sample_names"] == b["sample_names"]
    assert np.array_equal(a["target"], b["target"])
    assert list(a["top"]) == list(b["top"]) and np.array_equal(a["seed_iter"], b["seed_iter"])
    assert np.array_equal(a["cmi"], b["cmi"], equal_nan=True)          # same kernels on the same bits: identical
    assert a["cmi_orig_seed"] == b["cmi_orig_seed"]


def test_identical_consolidation_matches_the_reference_procedure():
    """consolidate.identical.features = "identical" (LIB:1727-1748) on the device against the reference's own
    procedure restated in numpy (paste -> stable order -> first of each run, counts in the names)."""
    from revealer_b200.api import consolidate_identical_host
    rng = np.random.default_rng(12)
    for F, n in [(1, 5), (50, 3), (400, 70), (3000, 130)]:
        base = (rng.random((max(1, F // 3), n)) < 0.15).astype(np.uint8)
        m = base[rng.integers(0, len(base), F)]                    # many exact duplicates, scattered
        names = ["r%d" % i for i in range(F)]
        want_m, want_names = consolidate_identical_host(m, names)
        with rb.RevealerContext(0) as ctx:
            ctx.load_matrix(m)
            kept, counts = ctx.consolidate_identical()
            got = matrix_of(ctx)
            assert ctx.unique_rows() == ctx.F == len(kept)
        assert ["%s (%d)" % (names[i], c) for i, c in zip(kept, counts)] == want_names
        assert np.array_equal(got, want_m)
        assert counts.sum() == F


def test_revealer_v1_with_identical_consolidation_device_equals_host(tmp_path):
    from revealer_b200 import synth
    w = synth.make_workload(n=90, F=700)
    names2 = ["P%03d" % i for i in range(90)]
    f1, f2 = str(tmp_path / "ds1.gct"), str(tmp_path / "ds2.gct")
    write_gct(f1, ["TGT"], ["y"], names2, [w["target"].tolist()])
    fast_gct(f2, w["m2"].astype(np.uint8), names=names2, row_prefix="feat")
    kw = dict(max_n_iter=2, consolidate_identical_features="identical", count_thres_low=3, count_thres_high=40)
    a = rb.REVEALER_v1(f1, "TGT", "positive", f2, ingest="host", **kw)
    b = rb.REVEALER_v1(f1, "TGT", "positive", f2, ingest="device", **kw)
    assert a["feature_names"] == b["feature_names"] and any("(2)" in s or "(3)" in s for s in a["feature_names"])
    assert list(a["top"]) == list(b["top"]) and np.array_equal(a["cmi"], b["cmi"], equal_nan=True)


@pytest.mark.parametrize("F,n,threads", [(3000, 130, 1), (5000, 1000, 3), (20000, 257, 8)])
def test_cooperative_upload_equals_device_only_upload(F, n, threads):
    """rvl_load_matrix_f64_colmajor with host packer threads (words claimed from the back while the DMA engine
    ships words from the front) leaves exactly the bits the device-only path leaves, and rejects the same input."""
    rng = np.random.default_rng(F + n)
    m = np.asfortranarray((rng.random((F, n)) < 0.07).astype(np.float64))
    with rb.RevealerContext(0) as ctx:
        ctx.set_host_threads(0)
        ctx.load_matrix(m)
        want = matrix_of(ctx)
        assert np.array_equal(want, m.astype(np.uint8))
        ctx.set_host_threads(threads)
        for _ in range(3):                                   # the meeting point of the two sides varies run to run
            ctx.load_matrix(m)
            assert np.array_equal(matrix_of(ctx), want)
        for s in (0, n // 2, n - 1):                         # a bad entry in a DMA-side word, in the middle, in a packer-side word
            bad = m.copy(order="F")
            bad[F // 3, s] = 0.5
            with pytest.raises(rb.RevealerError) as e:
                ctx.load_matrix(bad)
            assert e.value.code == _lib.RVL_ERR_NOT_BINARY
        bad = m.copy(order="F")
        bad[5, n - 1] = np.nan
        with pytest.raises(rb.RevealerError):
            ctx.load_matrix(bad)
        ctx.load_matrix(m)                                   # and the context is still usable
        assert np.array_equal(ctx.row_counts(), want.sum(axis=1))
