"""GCT ingest timing at config C3's shape: 100 000 rows x 1 000 samples as a GCT v1.2 text file (~200 MB)
through rvl_load_matrix_gct, next to the host reader (api.read_gct) on a slice and the R-layout upload."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import revealer_b200 as rb
from revealer_b200 import synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
w = synth.make_workload("C3_1kx100k", F=F)
m = w["m2"].astype(np.uint8)
n = m.shape[1]
d = tempfile.mkdtemp()
p = os.path.join(d, "c3.gct")
body = np.empty((F, 2 * n), dtype=np.uint8)
body[:, 0::2] = 9
body[:, 1::2] = m + 48
t0 = time.perf_counter()
with open(p, "wb") as fh:
    fh.write(("#1.2\n%d\t%d\nName\tDescription\t%s\n" % (F, n, "\t".join("S%d" % j for j in range(n)))).encode())
    for i in range(F):
        fh.write(b"feature_%d\tna" % i + body[i].tobytes() + b"\n")
print("wrote %s: %.1f MB in %.1f s" % (p, os.path.getsize(p) / 1e6, time.perf_counter() - t0))
with rb.RevealerContext(0) as ctx:
    for rep in range(4):
        t0 = time.perf_counter()
        g = rb.GctFile(p)
        t1 = time.perf_counter()
        ctx.load_gct(g)
        t2 = time.perf_counter()
        names = g.row_names
        t3 = time.perf_counter()
        g.close()
        print("open %.2f ms  load_gct %.2f ms (%.2f GB/s of text)  row names %.2f ms" %
              ((t1 - t0) * 1e3, (t2 - t1) * 1e3, os.path.getsize(p) / (t2 - t1) / 1e9, (t3 - t2) * 1e3))
    cnt = ctx.row_counts()
    assert np.array_equal(cnt, m.sum(axis=1)), "ingest mismatch"
    md = np.asfortranarray(m.astype(np.float64))
    for rep in range(2):
        t0 = time.perf_counter()
        ctx.load_matrix(md)
        print("R-layout upload of the same matrix (pageable host memory): %.2f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter()
sub = os.path.join(d, "sub.gct")
with open(p, "rb") as fi, open(sub, "wb") as fo:
    for i, line in enumerate(fi):
        if i == 1:
            line = b"%d\t%d\n" % (2000, n)
        if i >= 2003:
            break
        fo.write(line)
t0 = time.perf_counter()
r = rb.read_gct(sub)
dt = time.perf_counter() - t0
print("host reader (python mirror of MSIG.Gct2Frame) on 2000 rows: %.1f ms -> %.1f s for %d rows" % (dt * 1e3, dt * F / 2000, F))
