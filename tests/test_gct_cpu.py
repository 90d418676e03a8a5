"""Host side of the GCT ingest (rvl_gct_*: memory-mapped file, header, names, single rows) against the
plain-Python MSIG.Gct2Frame mirror (api.read_gct, LIB:2616-2626).  No GPU needed: the handle works
without a context."""
import numpy as np
import pytest

import revealer_b200 as rb


def write_gct(path, row_names, descs, names, ds, eol="\n", fmt=repr, trailing_newline=True, blank_every=0):
    lines = ["#1.2", "%d\t%d" % (len(row_names), len(names)), "Name\tDescription\t" + "\t".join(names)]
    for i, (r, d, row) in enumerate(zip(row_names, descs, ds)):
        lines.append("%s\t%s\t%s" % (r, d, "\t".join("" if (isinstance(v, float) and np.isnan(v)) else fmt(v) for v in row)))
        if blank_every and (i + 1) % blank_every == 0:
            lines.append("")
    text = eol.join(lines) + (eol if trailing_newline else "")
    with open(path, "wb") as fh:
        fh.write(text.encode())


@pytest.mark.parametrize("eol,trailing,blank", [("\n", True, 0), ("\r\n", True, 0), ("\n", False, 0), ("\n", True, 3)])
def test_gct_handle_matches_python_reader(tmp_path, eol, trailing, blank):
    rng = np.random.default_rng(5)
    F, n = 17, 9
    ds = rng.standard_normal((F, n)).round(3)
    ds[2, 4] = np.nan
    rows = ["gene_%d" % i for i in range(F)]
    rows[5] = "name with spaces"
    descs = ["desc %d" % i for i in range(F)]
    names = ["S%d" % i for i in range(n)]
    p = str(tmp_path / "x.gct")
    write_gct(p, rows, descs, names, ds.tolist(), eol=eol, trailing_newline=trailing, blank_every=blank)
    with rb.GctFile(p) as g:
        assert (g.n_rows, g.n_cols) == (F, n)
        assert g.names == names and g.row_names == rows and g.descs == descs
        for i in (0, 2, F - 1):
            v = g.row_values(i)
            assert np.array_equal(np.isnan(v), np.isnan(ds[i])) and np.allclose(v[~np.isnan(v)], ds[i][~np.isnan(ds[i])])
        with pytest.raises(rb.RevealerError):
            g.row_values(F)
    if eol == "\n":
        ref = rb.read_gct(p)
        assert ref["row_names"] == rows and ref["names"] == names


def test_gct_open_errors(tmp_path):
    p = str(tmp_path / "bad.gct")
    open(p, "w").write("#1.3\n1\t1\nName\tDescription\ta\nr\td\t1\n")
    with pytest.raises(rb.RevealerError, match="GCT v1.2"):
        rb.GctFile(p)
    open(p, "w").write("#1.2\n1\t2\nName\tDescription\ta\nr\td\t1\n")
    with pytest.raises(rb.RevealerError, match="column-name line"):
        rb.GctFile(p)
    with pytest.raises(rb.RevealerError, match="cannot open"):
        rb.GctFile(str(tmp_path / "missing.gct"))
    open(p, "w").write("#1.2\n3\t1\nName\tDescription\ta\nr\td\t1\n")     # header promises 3 rows, file has 1
    with rb.GctFile(p) as g:
        with pytest.raises(rb.RevealerError, match="data lines"):
            g.row_names


def test_gct_index_of_a_large_file_uses_several_threads(tmp_path):
    """> 8 MB of data lines: the line index is built by four threads on byte ranges cut at newlines; the
    order of rows, blank-line skipping and single-row access must not depend on where the cuts fall."""
    F, n = 30000, 200
    m = (np.random.default_rng(0).random((F, n)) < 0.1).astype(np.uint8)
    body = np.empty((F, 2 * n), dtype=np.uint8)
    body[:, 0::2] = 9
    body[:, 1::2] = m + 48
    p = str(tmp_path / "big.gct")
    with open(p, "wb") as fh:
        fh.write(("#1.2\n%d\t%d\nName\tDescription\t%s\n" % (F, n, "\t".join("S%d" % j for j in range(n)))).encode())
        for i in range(F):
            fh.write(b"row_%d\tdesc %d" % (i, i) + body[i].tobytes() + b"\n")
            if i % 1000 == 7:
                fh.write(b"\n")
    with rb.GctFile(p) as g:
        assert g.row_names == ["row_%d" % i for i in range(F)]
        assert g.descs[12345] == "desc 12345"
        for i in (0, 7, 8, 14999, 15000, F - 1):
            assert np.array_equal(g.row_values(i), m[i].astype(float))
