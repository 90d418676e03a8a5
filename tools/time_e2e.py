"""Where the end-to-end (host-buffer) call spends its time, per API call, for a 12.5k-row shard."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import revealer_b200 as rb
from revealer_b200 import synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
w = synth.make_workload("C3_1kx100k", F=F)
m = torch.empty((1000, F), dtype=torch.float64).pin_memory().numpy()
m[...] = w["m2"].T
mr = m.T
ctx = rb.RevealerContext(0)
def t(f, *a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(*a); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
for rep in range(3):
    a, _ = t(ctx.load_matrix, mr)
    b, _ = t(ctx.set_targets, w["target"])
    c, _ = t(ctx.set_seed, w["seed"])
    d, _ = t(ctx.search, 10, True, False)
    e, _ = t(ctx.search_fetch, 10)
    print("F=%d load %.2f ms  targets %.2f  seed %.2f  search %.2f  fetch %.2f" % (F, a, b, c, d, e))
# PCIe floor for the same bytes: one pinned -> device copy of the whole matrix
src = torch.from_numpy(m)
dst = torch.empty_like(src, device="cuda")
for rep in range(3):
    a, _ = t(lambda: dst.copy_(src, non_blocking=True))
    print("raw H2D of %.0f MB: %.2f ms = %.1f GB/s" % (src.numel() * 8 / 1e6, a, src.numel() * 8 / a / 1e6))
