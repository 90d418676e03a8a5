#!/usr/bin/env python
"""bench.py -- REVEALER CMI-search throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3_1kx100k] [--impl ours|reference]
  torchrun --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

A *step* is one complete search over the resident matrix: `iters` x (score every feature row against
the target given the current seed, rank, OR-merge the winner into the seed, assoc(target, seed)).
metric value = feature-CMI evaluations / s = F_global * n_targets * iters / (device time per step),
timed with CUDA events on the launching stream, max over ranks.  Feature rows are sharded across
ranks (strong scaling on the named matrix); the only exchange is one candidate record per rank and
iteration (peer-memory stores over NVLink by default, `--exchange nccl` for the all-gather).

Extra keys of the JSON line: `e2e` = the same search through the reference-facing calls with HOST buffers
(R's column-major doubles in pinned memory: upload + bit-packing + targets + seed + search + D2H of the
scores, wall clock, max over ranks; `phases_ms_rank0_last_step` splits it), `e2e_gct` = the same fed from a
GCT v1.2 text file through the device tokeniser (single rank), `roofline` (FP64 work executed / measured DFMA
peak, plus the HBM view), `cpu_baseline` (restated oracle on the host threads, rank 0, N=1).  Every rank is
bound to the cores NVML reports as local to its GPU before the pinned buffers are allocated.

`--impl reference` times the reference's own CPU algorithm for the same path.  R, MASS and misc3d do
not exist in this image, so it runs the restated C oracle in *reference mode* (three bcv() bandwidth
searches per cond.assoc call, as LIB:2564 does) on all host threads, on a bounded sample of the same
workload; it is labelled "port", never "R".
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "feature-CMI evaluations/sec"
UNIT = "evaluations/s"

# FP64 work convention for the roofline numerator (DESIGN.md "Roofline"): FP64-pipe instructions
# the kernel issues per transcendental, counted in the SASS of this build (profiles/), 2 flop each.
FLOP_PER_EXP = 2 * 18   # rvl_exp_neg: 14 DFMA + 2 DMUL + 2 DADD in the SASS of score.o
FLOP_PER_LOG = 2 * 8    # rvl_log_pos (table-driven): 6 DFMA + 1 DMUL + 1 DADD (rvl_common.cuh)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C3_1kx100k")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=0, help="override the workload's iteration count")
    ap.add_argument("--features", type=int, default=0, help="override the workload's row count (not the headline config)")
    ap.add_argument("--cpu-sample-rows", type=int, default=0, help="rows in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="candidate exchange between ranks: peer-memory stores over NVLink (CUDA IPC) or NCCL all-gather")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # NVML missing: report that instead of clocks
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def cpu_reference_rate(wl_name, sample_rows, threads=0, budget_s=15.0):
    """evaluations/s of the restated reference algorithm (oracle, reference mode) on host cores."""
    from oracle import oracle as O
    from revealer_b200 import synth
    n, F, iters, n_seed, T = synth.CONFIGS[wl_name]
    threads = threads or O.max_threads()
    if not sample_rows:
        # ~1 us per sample^2 pair-loop + kde: calibrate on a few rows, then size for ~budget_s
        w = synth.make_workload(wl_name, row_lo=0, row_hi=2 * threads)
        m = w["m2"].astype(np.float64)
        t0 = time.perf_counter()
        O.score_rows(np.atleast_2d(w["target"])[0], m, w["seed"].astype(np.float64), hoist=False, nthreads=threads)
        per = (time.perf_counter() - t0) / len(m)
        sample_rows = int(max(4 * threads, min(F, budget_s / max(per, 1e-6))))
    w = synth.make_workload(wl_name, row_lo=0, row_hi=sample_rows)
    m = w["m2"].astype(np.float64)
    x = np.atleast_2d(w["target"])[0]
    t0 = time.perf_counter()
    O.score_rows(x, m, w["seed"].astype(np.float64), hoist=False, nthreads=threads)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.score_rows(x, m[: max(threads, sample_rows // 4)], w["seed"].astype(np.float64), hoist=True, nthreads=threads)
    dth = (time.perf_counter() - t0) / max(threads, sample_rows // 4) * sample_rows
    return dict(value=sample_rows / dt, unit=UNIT, cores=threads, kind="port",
                sample="first %d feature rows of %s, 1 scoring pass, restated C oracle in reference mode (3 bcv "
                       "searches per call as LIB:2564), %d threads, %.1f s; with bandwidths hoisted: %.0f eval/s; "
                       "R/MASS/misc3d absent -> not R" % (sample_rows, wl_name, threads, dt, sample_rows / dth),
                seconds=dt)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from revealer_b200 import synth
    n, F, iters, n_seed, T = synth.CONFIGS[args.workload]
    t_all = time.perf_counter()
    vals = []
    rows = args.cpu_sample_rows
    for i in range(args.warmup + args.steps):
        r = cpu_reference_rate(args.workload, rows, budget_s=8.0)
        rows = rows or int(round(r["value"] * r["seconds"]))
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    sec = float(np.mean([r["seconds"] for r in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "samples": n, "features": F, "iterations": iters,
                   "targets": T, "step": "one scoring pass over a bounded sample of %d rows" % rows},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": vals[-1]["cores"], "kind": "port",
                         "sample": vals[-1]["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


def bind_to_gpu_numa(index):
    """One process per GPU: run on the cores NVML reports as local to that GPU, so that the pinned host
    buffers of the e2e path (first touch) sit on the NUMA node the GPU's PCIe link hangs off."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import revealer_b200 as rb
    from revealer_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    numa_cores = bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, F, iters, n_seed, T = synth.CONFIGS[args.workload]
    if args.iters:
        iters = args.iters
    if args.features:
        F = args.features
    lo, cnt = rb.shard_rows(F, world, rank)
    w = synth.make_workload(args.workload, F=F, row_lo=lo, row_hi=lo + cnt)
    targets = np.atleast_2d(w["target"])
    m2_u8 = w["m2"]

    stream = torch.cuda.current_stream()
    ctx = rb.RevealerContext(local)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_shard(rank, world, lo, F)

    # ---- resident inputs (outside the timed region for `value`)
    ctx.load_matrix(m2_u8)
    ctx.set_targets(targets)
    ctx.set_seed(w["seed"])
    nb = ctx.candidate_bytes()
    send = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    recv = torch.zeros(nb * world, dtype=torch.uint8, device="cuda") if world > 1 else send
    use_p2p = world > 1 and args.exchange == "p2p"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    if use_p2p:   # every rank maps every other rank's exchange buffer (CUDA IPC); NCCL only carries the handles
        ok = 1
        try:
            mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).cuda()
        except rb.RevealerError:
            ok, mine = 0, torch.zeros(64, dtype=torch.uint8, device="cuda")
        allh = torch.empty(64 * world, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allh, mine)
        if ok:
            try:
                ctx.p2p_import(allh.cpu().numpy().tobytes())
            except rb.RevealerError:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:      # peer mapping unavailable on some rank: every rank falls back together
            ctx.p2p_disable()
            use_p2p = False
    use_nccl = world > 1 and not use_p2p
    if use_nccl:
        ctx.set_exchange_buffers(send.data_ptr(), recv.data_ptr())

    def one_search():
        ctx.search_begin(iters)
        for it in range(iters):
            ctx.iter_score(it)
            if use_nccl:
                dist.all_gather_into_tensor(recv, send)
            ctx.iter_merge(it)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_search()
    barrier()
    ctx.score_kernel_time(reset=True)
    ctx.counters(reset=True)
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.zero_()                     # L2 flush between timed steps (inputs are smaller than L2)
        a.record(stream)
        one_search()
        b.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    gpu_launches = ctx.launch_count() - launches0      # this rank's own kernels (NCCL's are not counted)
    k_ms, k_n = ctx.score_kernel_time(reset=True)
    ctr = ctx.counters(reset=True)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = F * T * iters / (ms * 1e-3)

    # ---- e2e: through the reference-facing call with HOST buffers (R layout: column-major doubles),
    # H2D + bit-packing + search + D2H of the score matrix and winners inside the timed region.
    e2e = None
    if not args.no_e2e:
        m2_f64 = torch.empty((n, max(cnt, 1)), dtype=torch.float64).pin_memory().numpy()[:, :cnt]
        m2_f64[...] = m2_u8.T                     # column-major F x n  == C-order n x F
        m2_r = m2_f64.T                           # F x n view, Fortran-contiguous, pinned
        tgt_h, seed_h = targets.copy(), w["seed"].copy()

        phases = {}

        def one_e2e():
            t0 = time.perf_counter()
            ctx.load_matrix(m2_r)
            t1 = time.perf_counter()
            ctx.set_targets(tgt_h)
            ctx.set_seed(seed_h)
            t2 = time.perf_counter()
            one_search()
            r = ctx.search_fetch(iters)
            t3 = time.perf_counter()
            phases.update(load_matrix=(t1 - t0) * 1e3, targets_seed=(t2 - t1) * 1e3, search_fetch=(t3 - t2) * 1e3)
            return r

        one_e2e()
        barrier()
        t0 = time.perf_counter()
        reps = max(1, min(args.steps, 3))
        for _ in range(reps):
            res = one_e2e()
        barrier()
        dt = (time.perf_counter() - t0) / reps
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        pcie_bytes, host_words = ctx.upload_stats()      # of the last load: what actually crossed PCIe
        host_in = cnt * n * 8 + targets.nbytes + seed_h.nbytes
        h2d = pcie_bytes + targets.nbytes + seed_h.nbytes
        d2h = sum(v.nbytes for v in res.values() if v is not None)
        e2e = {"value": F * T * iters / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3,
               "host_input_bytes_per_step": int(host_in),
               "upload": "cooperative: %d of %d 64-sample words packed by host threads, the rest sent as doubles and "
                         "packed on the device (rank 0, last step)" % (host_words, (n + 63) // 64),
               "phases_ms_rank0_last_step": {k: round(v, 3) for k, v in phases.items()},
               "call": "rvl_load_matrix_f64_colmajor + rvl_set_targets + rvl_set_seed + search + rvl_search_fetch"}

    # ---- the same search fed from a GCT v1.2 text file on local disk (rvl_load_matrix_gct: text over PCIe,
    # tokenised on the device) instead of the R double matrix -- informational, single rank only
    e2e_gct = None
    if not args.no_e2e and world == 1 and cnt * n <= 200_000_000:
        import tempfile
        import revealer_b200 as rb
        with tempfile.TemporaryDirectory() as td:
            gpath = os.path.join(td, "m2.gct")
            body = np.empty((cnt, 2 * n), dtype=np.uint8)
            body[:, 0::2] = 9
            body[:, 1::2] = m2_u8 + 48
            with open(gpath, "wb") as fh:
                fh.write(("#1.2\n%d\t%d\nName\tDescription\t%s\n" % (cnt, n, "\t".join("S%d" % j for j in range(n)))).encode())
                for i in range(cnt):
                    fh.write(b"f%d\tna" % i + body[i].tobytes() + b"\n")
            del body
            gbytes = os.path.getsize(gpath)

            def one_gct():
                with rb.GctFile(gpath) as g:
                    ctx.load_gct(g)
                ctx.set_targets(tgt_h)
                ctx.set_seed(seed_h)
                one_search()
                return ctx.search_fetch(iters)

            one_gct()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                one_gct()
            barrier()
            dtg = (time.perf_counter() - t0) / reps
            e2e_gct = {"value": F * T * iters / dtg, "unit": UNIT, "ms_per_step": dtg * 1e3, "file_bytes": int(gbytes),
                       "call": "rvl_gct_open + rvl_load_matrix_gct (page-cached file) + rvl_set_targets + rvl_set_seed + "
                               "search + rvl_search_fetch"}

    # ---- roofline of the dominant kernel (k_score)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fp64_peak = ctx.measure_fp64_peak()
    roofline = None
    if k_n:
        dur = k_ms / k_n * 1e-3                                   # average launch duration (s)
        evals = ctr["kde_evaluations"] / k_n
        G = 25
        direct = ctr["live_columns"] / k_n            # (j,k) columns evaluated cell by cell
        fact = ctr["factorised_columns"] / k_n        # (j,k) columns in the O(1) factorised form
        # per evaluation (score.cu, rvl_cmi_epilogue_f): pass A 2 G^2 logs, pass B one log per non-dead
        # column, pass C G logs per direct column, G for the z marginal, 1 in the table lookup
        logs = evals * (2 * G * G + G + 1) + direct * (G + 1) + fact
        exps = ctr["dense_exps"] / k_n + evals * (2 * G + 1)
        two = ctr.get("two_term_columns", 0) / k_n    # direct columns whose cells take 2 FMAs instead of 4
        plain = evals * (G * G * 18 + G * G * 24 + 80) + direct * G * 10 - two * G * 4 + ctr["dense_exps"] / k_n * 6
        flops = exps * FLOP_PER_EXP + logs * FLOP_PER_LOG + plain
        words = ((n + 63) // 64 + 1) & ~1
        bytes_alg = cnt * T * (words * 8 + 8)
        traffic = None
        try:  # DRAM bytes per launch from the committed ncu capture of this workload (1 GPU)
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01c_k_score_traffic.json")))
            if tj["workload"] == args.workload and world == 1 and not args.features:
                traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
        except Exception:
            pass
        roofline = {
            "kernel": "k_score", "bound": "fp64", "achieved": flops / dur / 1e12, "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": flops / dur / 1e12 / fp64_peak if fp64_peak else None, "traffic": traffic,
            "peak_source": "measured live: rvl_measure_fp64_peak (DFMA chain); MEASURED_PEAKS.json has no FP64 figure",
            "launch_ms": dur * 1e3, "launches": int(k_n), "share_of_step": k_ms / (ms * args.steps),
            "convention": "executed work from kernel counters: exp=%d flop, log=%d flop (FP64-pipe SASS "
                          "instructions x2), + plain FMA/ADD" % (FLOP_PER_EXP, FLOP_PER_LOG),
            "per_launch": {"evaluations": evals, "exp": exps, "log": logs, "flop": flops},
            "hbm": {"bound": "hbm", "achieved": bytes_alg / dur / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_alg / dur / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": bytes_alg,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_rate(args.workload, args.cpu_sample_rows)
        cpu.pop("seconds", None)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "samples": n, "features": F, "iterations": iters, "targets": T,
                       "sharding": "feature rows, contiguous blocks, %d rank(s)" % world,
                       "exchange": ("peer-memory stores (CUDA IPC over NVLink)" if use_p2p else
                                    ("NCCL all_gather" if world > 1 else "none")),
                       "l2": "flushed between timed steps (256 MiB write); bit matrix %.1f MB < L2" % (cnt * words * 8 / 1e6),
                       "step": "one full search: iterations x (score + rank + merge + seed IC)",
                       "host_affinity": ("%d NUMA-local cores of the GPU (NVML)" % numa_cores) if numa_cores else "unbound"},
            "e2e": e2e, "e2e_gct": e2e_gct, "gpu_launches": int(gpu_launches), "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
