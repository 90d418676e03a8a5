"""Scratch GPU check: CUDA path vs oracle on a small synthetic case (not a test)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import revealer_b200 as rb
from revealer_b200 import synth
from oracle import oracle as O

n, F = int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 300
w = synth.make_workload(n=n, F=F)
x, m2, seed = w["target"], w["m2"], w["seed"]
ctx = rb.RevealerContext(0)
ctx.load_matrix(m2)
ctx.set_targets(x)
print("bcv target gpu/oracle", ctx.target_bandwidth(0), O.bcv(x))
tab = ctx.binary_bandwidth_table()
for n1 in (1, 3, 10, n // 2, n - 1):
    y = np.zeros(n); y[:n1] = 1
    print(" bcv binary n1=%d gpu %.17g oracle %.17g" % (n1, tab[n1], O.bcv(y)))
for sd, nm in ((seed, "seeded"), (np.zeros(n, np.uint8), "nullseed")):
    ctx.set_seed(sd)
    t = time.time(); g = ctx.score_once()[0]; tg = time.time() - t
    t = time.time(); o = O.score_rows(x, m2.astype(float), sd.astype(float)); to = time.time() - t
    rel = np.abs(g - o) / np.maximum(np.abs(o), 1e-300)
    rel[(o == 0) & (g == 0)] = 0
    print(nm, "max rel err", np.nanmax(rel), "argmax", np.nanargmax(rel), "gpu s", tg, "oracle s", to)
    bad = np.argsort(-rel)[:5]
    for b in bad:
        print("   row", b, "n1", m2[b].sum(), "gpu", g[b], "oracle", o[b])
ctx.set_seed(seed)
res = ctx.search(3)
ores = O.search(x, m2.astype(float), seed.astype(float), 3)
print("top gpu", res["top"][0], "oracle", ores["top"])
print("cmi_seed gpu", res["cmi_seed"][0], "oracle", ores["cmi_seed"])
print("ic0", res["ic_seed0"][0], ores["ic_seed0"])
print("seed_iter equal", np.array_equal(res["seed_iter"][0], ores["seed_iter"].astype(np.uint8)))
print("cmi max abs diff", np.nanmax(np.abs(res["cmi"][0] - ores["cmi"])))
print("assoc", ctx.assoc(seed), O.assoc(x, seed.astype(float)))
print("launches", ctx.launch_count(), "score ms", ctx.score_kernel_time())
