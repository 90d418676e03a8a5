"""Single-process multi-GPU check: revealer_search_multi_gpu (one host thread, peer access, stamped
exchange) must reproduce the single-GPU search exactly."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import revealer_b200 as rb
from revealer_b200 import synth

ng = torch.cuda.device_count()
w = synth.make_workload(n=300, F=6001)
single = rb.revealer_search(w["target"], w["m2"], w["seed"], max_n_iter=5)
ok = True
for devs in ([0, 1], list(range(ng))):
    t0 = time.time()
    multi = rb.revealer_search_multi_gpu(w["target"], w["m2"], w["seed"], max_n_iter=5, devices=devs)
    same = all(np.array_equal(single[k], multi[k]) for k in ("cmi", "top", "seed_iter", "cmi_seed", "top_score"))
    print("devices", devs, "identical to single GPU:", same, "winners", multi["top"], "%.2f s" % (time.time() - t0))
    ok = ok and same
sys.exit(0 if ok else 1)
