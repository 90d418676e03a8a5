"""Phase timing of the fused search kernel (k_search) on one GPU: where an iteration's time goes.
  python tools/time_phases.py [workload] [rows]"""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import revealer_b200 as rb
from revealer_b200 import synth

wl = sys.argv[1] if len(sys.argv) > 1 else "C3_1kx100k"
n, F, iters, _, T = synth.CONFIGS[wl]
F = int(sys.argv[2]) if len(sys.argv) > 2 else F
w = synth.make_workload(wl, F=F)
with rb.RevealerContext(0) as ctx:
    ctx.load_matrix(w["m2"]); ctx.set_targets(np.atleast_2d(w["target"])); ctx.set_seed(w["seed"])
    for _ in range(3):
        ctx.search(iters, fetch=False)
    ctx._ck(ctx._lib.rvl_debug_phase_times(ctx._h, None, 0))
    import torch
    ctx.set_profiling(True)
    ctx.score_kernel_time(reset=True)
    ctx.search(iters, fetch=False)
    kms, kn = ctx.score_kernel_time(reset=True)
    out = (C.c_uint64 * (8 * iters))()
    ctx._ck(ctx._lib.rvl_debug_phase_times(ctx._h, out, iters))
    t = np.array(out, dtype=np.int64).reshape(iters, 8)
    t0 = t[0, 0]
    print("k_search by events: %.1f us (%d launch); entry -> first scoring %.1f us; entry -> last barrier-stamp %.1f us" %
          (kms * 1e3, kn, (t[0, 0] - t[0, 6]) / 1e3, (t[iters - 1, 5] - t[0, 6]) / 1e3))
    print("it  score(b0)  all_done-b0  publish  merge(b0)-after-publish  barrier-wait  iteration")
    for it in range(iters):
        s, b0, alld, pub, mrg, bar = t[it, :6]
        nxt = t[it + 1, 0] if it + 1 < iters else bar
        print("%2d %9.1f %11.1f %8.1f %12.1f %14.1f %10.1f   us" % (it, (b0 - s) / 1e3, (alld - b0) / 1e3, (pub - alld) / 1e3,
              (mrg - pub) / 1e3, (bar - mrg) / 1e3, (nxt - s) / 1e3))
