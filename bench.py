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
`value` counts every row of the matrix (each receives its score); rows that are exact copies of another row
are scored once (R's stable order() ranks the first of them first) -- `evaluations_executed` and
`value_executed` state what the kernels actually evaluated.

Extra keys of the JSON line:
  check          winners, CRC of the winners' scores and a position-weighted checksum of every score bit pattern,
                 identical at every N by construction; at N > 1 rank 0 also runs the whole workload on one GPU
                 and asserts that the sharded run selected the same features with the same scores.  Parity of
                 these numbers with the reference algorithm is what tests/ checks -- against the RESTATED
                 oracle (R is not installed: "parity unpinned").
  e2e            the same search through the reference-facing calls with HOST buffers as R holds them
                 (column-major doubles in ordinary pageable memory): upload + bit-packing + targets + seed +
                 search + D2H of the scores, wall clock, max over ranks; `pinned` = the same from pinned memory.
  e2e_gct        the same fed from a GCT v1.2 text file through the device tokeniser (single rank).
  roofline       the dominant kernel (k_score) against the resource that binds it -- instruction issue: warp instructions
                 per evaluation of the committed ncu capture x evaluations executed / launch time, over 4 schedulers x SMs
                 x the SM clock (`frac`); `fp64` = FP64 work executed / measured DFMA peak (the earlier rounds' headline),
                 `algorithmic` = the same launch counted by SURVEY 8(d)'s literal per-evaluation work, `hbm` the HBM view,
                 `ncu` the committed capture's own counters.
  cpu_baseline   restated oracle on the host cores (rank 0, N = 1): all cores, and one core (the reference is
                 serial R).
  configs        short runs of the other named shapes (C2; C5; C4 at 8 GPUs) with their own check blocks.

`--impl reference` times the reference's own CPU algorithm for the same path.  R, MASS and misc3d do
not exist in this image, so it runs the restated C oracle in *reference mode* (three bcv() bandwidth
searches per cond.assoc call, as LIB:2564 does) on an explicit number of host threads (all cores the process
may use, whatever OMP_NUM_THREADS the launcher exports), on a bounded sample of the same workload; it is
labelled "port", never "R".
"""
import argparse
import json
import os
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "feature-CMI evaluations/sec"
UNIT = "evaluations/s"
HOST_CORES = len(os.sched_getaffinity(0))   # taken before any NUMA binding: the cores this process may use

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
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU legs (0 = every core the process may use)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other named configs")
    ap.add_argument("--no-single-rank-check", action="store_true")
    ap.add_argument("--fused", action="store_true",
                    help="the one persistent k_search launch per search instead of the per-iteration launches (k_score / k_argmax / "
                         "k_advance); bit-identical results, measured 2-3 %% slower on B200")
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


# ------------------------------------------------------------------------------------------------ CPU legs

def cpu_reference_rate(wl_name, sample_rows, threads, budget_s=15.0, single_core=True):
    """evaluations/s of the restated reference algorithm (oracle, reference mode) on `threads` host threads
    (explicit: omp_set_num_threads, so the launcher's OMP_NUM_THREADS does not matter), and on one thread."""
    from oracle import oracle as O
    from revealer_b200 import synth
    n, F, iters, n_seed, T = synth.CONFIGS[wl_name]
    threads = max(1, int(threads))
    if not sample_rows:
        # calibrate on a few rows, then size the sample for ~budget_s
        w = synth.make_workload(wl_name, row_lo=0, row_hi=2 * threads)
        m = w["m2"].astype(np.float64)
        t0 = time.perf_counter()
        O.score_rows(np.atleast_2d(w["target"])[0], m, w["seed"].astype(np.float64), hoist=False, nthreads=threads)
        per = (time.perf_counter() - t0) / len(m)
        sample_rows = int(max(4 * threads, min(F, budget_s / max(per, 1e-6))))
    w = synth.make_workload(wl_name, row_lo=0, row_hi=sample_rows)
    m = w["m2"].astype(np.float64)
    x = np.atleast_2d(w["target"])[0]
    z = w["seed"].astype(np.float64)
    t0 = time.perf_counter()
    O.score_rows(x, m, z, hoist=False, nthreads=threads)
    dt = time.perf_counter() - t0
    nh = max(threads, sample_rows // 4)
    t0 = time.perf_counter()
    O.score_rows(x, m[:nh], z, hoist=True, nthreads=threads)
    hoisted = nh / (time.perf_counter() - t0)
    out = dict(value=sample_rows / dt, unit=UNIT, cores=threads, kind="port",
               sample="first %d feature rows of %s, 1 scoring pass, restated C oracle in reference mode (3 bcv "
                      "searches per call as LIB:2564), %d threads (explicit), %.1f s; with bandwidths hoisted: %.0f eval/s; "
                      "R/MASS/misc3d absent -> not R" % (sample_rows, wl_name, threads, dt, hoisted),
               seconds=dt, rows=sample_rows)
    if single_core:   # the reference itself is a serial R loop: the same port on ONE core
        n1 = max(8, min(sample_rows, int(sample_rows / threads / 4) or 8))
        t0 = time.perf_counter()
        O.score_rows(x, m[:n1], z, hoist=False, nthreads=1)
        d1 = time.perf_counter() - t0
        out["single_core"] = dict(value=n1 / d1, unit=UNIT, cores=1, sample="first %d rows, 1 thread, %.1f s" % (n1, d1))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from revealer_b200 import synth
    n, F, iters, n_seed, T = synth.CONFIGS[args.workload]
    threads = args.cpu_threads or HOST_CORES
    t_all = time.perf_counter()
    vals = []
    rows = args.cpu_sample_rows
    for i in range(args.warmup + args.steps):
        r = cpu_reference_rate(args.workload, rows, threads, budget_s=8.0, single_core=(i == args.warmup + args.steps - 1))
        rows = rows or r["rows"]
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
                         "sample": vals[-1]["sample"], "single_core": vals[-1].get("single_core"),
                         "host_cores_available": HOST_CORES},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line))


def bind_to_gpu_numa(index):
    """One process per GPU: run on the cores NVML reports as local to that GPU, so that the host
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


# ------------------------------------------------------------------------------------------------ result checks

def score_checksum(cmi, row_offset, F_global, t_offset=0):
    """Position-weighted sum (mod 2^64) of the bit patterns of cmi[t][it][local row]: independent of how rows or
    targets are sharded, so the sum over ranks is the same number at every N iff every score is bit-identical."""
    T, iters, Fl = cmi.shape
    bits = np.ascontiguousarray(cmi).view(np.uint64)
    rows = (np.arange(Fl, dtype=np.uint64) + np.uint64(row_offset + 1))
    tot = np.uint64(0)
    with np.errstate(over="ignore"):
        for t in range(T):
            for it in range(iters):
                salt = np.uint64(((t + t_offset) * iters + it) * F_global)
                tot += np.sum(bits[t, it] * (rows + salt) * np.uint64(0x9E3779B97F4A7C15), dtype=np.uint64)
    return int(tot)


def result_check(res, row_offset, F_global, dist, world, t_offset=0, gather_targets=False):
    """The `check` block: winners (global rows), CRC32 of the winners' scores and seed ICs, checksum of all scores."""
    import torch
    cs = score_checksum(res["cmi"], row_offset, F_global, t_offset)
    top, top_score, cmi_seed = res["top"], res["top_score"], res["cmi_seed"]
    if world > 1:
        t = torch.tensor([np.int64(np.uint64(cs).astype(np.int64))], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)                      # wraps mod 2^64
        cs = int(np.uint64(np.int64(t.item())))
        if gather_targets:   # target-sharded: every rank holds the winners of its own targets
            parts = [None] * world
            dist.all_gather_object(parts, (top, top_score, cmi_seed))
            top = np.concatenate([p[0] for p in parts])
            top_score = np.concatenate([p[1] for p in parts])
            cmi_seed = np.concatenate([p[2] for p in parts])
    return {"winners": [int(v) for v in np.ravel(top)],
            "top_score_crc": "%08x" % zlib.crc32(np.ascontiguousarray(top_score).tobytes()),
            "cmi_seed_crc": "%08x" % zlib.crc32(np.ascontiguousarray(cmi_seed).tobytes()),
            "cmi_checksum": "%016x" % cs,
            "parity": "tests/ (-m gpu) hold these scores to the RESTATED oracle (R absent: parity unpinned)"}


# ------------------------------------------------------------------------------------------------ one workload

class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.numa_cores = bind_to_gpu_numa(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.current_stream()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def connect(self, ctx):
        """Candidate exchange between the ranks' contexts: peer-memory stores (CUDA IPC) or NCCL all-gather."""
        import revealer_b200 as rb
        torch, dist, world = self.torch, self.dist, self.world
        nb = ctx.candidate_bytes()
        send = torch.zeros(nb, dtype=torch.uint8, device="cuda")
        recv = torch.zeros(nb * world, dtype=torch.uint8, device="cuda") if world > 1 else send
        use_p2p = world > 1 and self.args.exchange == "p2p"
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
        return use_p2p, use_nccl, send, recv

    def run(self, wl_name, steps, warmup, main):
        """One named workload.  main: the headline run (e2e, roofline, cpu baseline, profiling counters)."""
        import revealer_b200 as rb
        from revealer_b200 import synth
        torch, dist, args, world, rank, local = self.torch, self.dist, self.args, self.world, self.rank, self.local
        n, F, iters, n_seed, T = synth.CONFIGS[wl_name]
        if main and args.iters:
            iters = args.iters
        if main and args.features:
            F = args.features
        # C5-style batches shard by TARGET when there are at least as many targets as ranks: every rank
        # holds the whole matrix and its own block of targets, and nothing is exchanged at all
        by_target = world > 1 and T >= world
        if by_target:
            lo, cnt = 0, F
            t_lo, t_cnt = rb.shard_rows(T, world, rank)
        else:
            lo, cnt = rb.shard_rows(F, world, rank)
            t_lo, t_cnt = 0, T
        w = synth.make_workload(wl_name, F=F, row_lo=lo, row_hi=lo + cnt)
        targets = np.atleast_2d(w["target"])[t_lo:t_lo + t_cnt]
        m2_u8 = w["m2"]

        ctx = rb.RevealerContext(local)
        ctx.set_stream(self.stream.cuda_stream)
        if not by_target:
            ctx.set_shard(rank, world, lo, F)
        ctx.set_fused(args.fused)
        # ---- resident inputs (outside the timed region for `value`)
        ctx.load_matrix(m2_u8)
        ctx.set_targets(targets)
        ctx.set_seed(w["seed"])
        if by_target:
            use_p2p = use_nccl = False
        else:
            use_p2p, use_nccl, send, recv = self.connect(ctx)

        def one_search():
            if not use_nccl:       # every launch of the search enqueued by one call; ranks meet through peer memory on the device
                ctx.search_run(iters)
                return
            ctx.search_begin(iters)
            for it in range(iters):
                ctx.iter_score(it)
                dist.all_gather_into_tensor(recv, send)
                ctx.iter_merge(it)

        for _ in range(warmup):
            one_search()
        self.barrier()
        launches0 = ctx.launch_count()
        sampler = ClockSampler(local) if main else None
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for a, b in ev:
            self.flush.zero_()                     # L2 flush between timed steps (inputs are smaller than L2)
            a.record(self.stream)
            one_search()
            b.record(self.stream)
        self.barrier()
        clocks = sampler.stop() if sampler else None
        ms = self.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev) / steps)
        gpu_launches = ctx.launch_count() - launches0      # this rank's own kernels (NCCL's are not counted)
        value = F * T * iters / (ms * 1e-3)
        res = ctx.search_fetch(iters)
        # the timed region above runs the product path as shipped (no instrumentation); the dominant kernel's own duration
        # and work counters come from a few more searches with the profiling switch on (CUDA events around every launch)
        k_ms, k_n, ctr, ms_prof = 0.0, 0, None, None
        if main:
            ctx.set_profiling(True)
            ctx.score_kernel_time(reset=True)
            ctx.counters(reset=True)
            nprof = max(1, min(steps, 3))
            self.barrier()
            pe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nprof)]
            for a, b in pe:
                self.flush.zero_()
                a.record(self.stream)
                one_search()
                b.record(self.stream)
            self.barrier()
            ms_prof = sum(a.elapsed_time(b) for a, b in pe) / nprof
            k_ms, k_n = ctx.score_kernel_time(reset=True)
            ctr = ctx.counters(reset=True)
            ctx.set_profiling(False)
        check = result_check(res, lo, F, dist, world, t_offset=t_lo, gather_targets=by_target)
        uniq = ctx.unique_rows()
        executed = self.max_over_ranks(0.0)  # placeholder overwritten below
        ex_t = torch.tensor([float(uniq) * t_cnt * iters], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ex_t, op=dist.ReduceOp.SUM)
        executed = float(ex_t.item())

        # ---- at N > 1: the same workload on ONE GPU (rank 0), asserted equal
        if world > 1 and rank == 0 and not args.no_single_rank_check and F * n <= 400_000_000:
            wf = synth.make_workload(wl_name, F=F)
            with rb.RevealerContext(local) as c1:
                c1.set_stream(self.stream.cuda_stream)
                c1.load_matrix(wf["m2"])
                c1.set_targets(np.atleast_2d(wf["target"]))
                c1.set_seed(wf["seed"])
                r1 = c1.search(iters)
            chk1 = result_check(r1, 0, F, dist, 1)
            same = all(chk1[k] == check[k] for k in ("winners", "top_score_crc", "cmi_seed_crc", "cmi_checksum"))
            check["single_rank_equal"] = bool(same)
            if not same:
                raise SystemExit("bench.py: the %d-rank run differs from the 1-rank run of %s: %s vs %s" % (world, wl_name, check, chk1))
            del wf, r1
        elif world > 1:
            check["single_rank_equal"] = None if rank == 0 else "rank 0 only"
        if world > 1:
            self.barrier()

        out = {"value": value, "ms_per_step": ms, "steps": steps, "warmup": warmup,
               "evaluations_per_step": F * T * iters, "evaluations_executed_per_step": executed,
               "value_executed": executed / (ms * 1e-3),
               "config": {"workload": wl_name, "samples": n, "features": F, "iterations": iters, "targets": T,
                          "sharding": ("targets, contiguous blocks, %d rank(s); whole matrix on every rank, no exchange" % world)
                          if by_target else ("feature rows, contiguous blocks, %d rank(s)" % world),
                          "exchange": ("peer-memory stores (CUDA IPC over NVLink)" if use_p2p else
                                       ("NCCL all_gather" if use_nccl else "none")),
                          "dedup": "rows identical to an earlier row are scored once and receive its score (%d of %d "
                                   "rows of this rank are distinct); `value` counts all rows" % (uniq, cnt)},
               "check": check, "gpu_launches": int(gpu_launches)}
        if not main:
            if k_n:
                out["k_score_ms_per_launch"] = k_ms / k_n
            ctx.close()
            return out

        words = ((n + 63) // 64 + 1) & ~1
        out["config"].update({
            "l2": "flushed between timed steps (256 MiB write); bit matrix %.1f MB < L2" % (cnt * words * 8 / 1e6),
            "step": "one full search: iterations x (score + rank + merge + seed IC)",
            "host_affinity": ("%d NUMA-local cores of the GPU (NVML)" % self.numa_cores) if self.numa_cores else "unbound"})
        out["clocks"] = clocks

        # ---- e2e: through the reference-facing call with HOST buffers (R layout: column-major doubles),
        # H2D + bit-packing + search + D2H of the score matrix and winners inside the timed region.
        tgt_h, seed_h = targets.copy(), w["seed"].copy()
        reps = max(1, min(steps, 3))
        if not args.no_e2e and cnt * n * 8 > (4 << 30):
            out["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                          "skipped": "the R-layout copy of this rank's shard would take %.1f GB of host memory (twice: pageable and "
                                     "pinned); run with fewer rows per rank" % (cnt * n * 8 / 1e9)}
        elif not args.no_e2e:
            def e2e_leg(m2_r):
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

                self.barrier()
                t0 = time.perf_counter()
                one_e2e()
                first = self.max_over_ranks(time.perf_counter() - t0)
                self.barrier()
                t0 = time.perf_counter()
                for _ in range(reps):
                    r = one_e2e()
                self.barrier()
                dt = self.max_over_ranks((time.perf_counter() - t0) / reps)
                pcie_bytes, host_words = ctx.upload_stats()      # of the last load: what actually crossed PCIe
                return dt, first, pcie_bytes, host_words, phases, r

            # R's allocVector memory is ordinary pageable memory: that is the headline; pinned for comparison
            m2_page = np.empty((n, max(cnt, 1)), dtype=np.float64)[:, :cnt]
            m2_page[...] = m2_u8.T                    # column-major F x n  == C-order n x F
            dt, first, pcie_bytes, host_words, phases, r = e2e_leg(m2_page.T)
            host_in = cnt * n * 8 + targets.nbytes + seed_h.nbytes
            h2d = pcie_bytes + targets.nbytes + seed_h.nbytes
            d2h = sum(v.nbytes for v in r.values() if v is not None)
            e2e = {"value": F * T * iters / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3,
                   "host_memory": "pageable (numpy / malloc, as R's allocVector)",
                   "first_call_ms": first * 1e3,
                   "host_input_bytes_per_step": int(host_in),
                   "upload": "cooperative: %d of %d 64-sample words packed by host threads, the rest sent as doubles and "
                             "packed on the device (rank 0, last step)" % (host_words, (n + 63) // 64),
                   "phases_ms_rank0_last_step": {k: round(v, 3) for k, v in phases.items()},
                   "call": "rvl_load_matrix_f64_colmajor + rvl_set_targets + rvl_set_seed + search + rvl_search_fetch"}
            del m2_page
            m2_pin = torch.empty((n, max(cnt, 1)), dtype=torch.float64).pin_memory().numpy()[:, :cnt]
            m2_pin[...] = m2_u8.T
            dtp, firstp, pb, hw, ph, _ = e2e_leg(m2_pin.T)
            e2e["pinned"] = {"value": F * T * iters / dtp, "ms_per_step": dtp * 1e3, "h2d_bytes_per_step":
                             int(pb + targets.nbytes + seed_h.nbytes), "host_packed_words": int(hw),
                             "phases_ms_rank0_last_step": {k: round(v, 3) for k, v in ph.items()}}
            del m2_pin
            out["e2e"] = e2e

        # ---- the same search fed from a GCT v1.2 text file on local disk (rvl_load_matrix_gct: text over PCIe,
        # tokenised on the device) instead of the R double matrix -- informational, single rank only
        if not args.no_e2e and world == 1 and cnt * n <= 200_000_000:
            import tempfile
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
                self.barrier()
                t0 = time.perf_counter()
                for _ in range(reps):
                    one_gct()
                self.barrier()
                dtg = (time.perf_counter() - t0) / reps
                out["e2e_gct"] = {"value": F * T * iters / dtg, "unit": UNIT, "ms_per_step": dtg * 1e3, "file_bytes": int(gbytes),
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
        if k_n:
            dur = k_ms / k_n * 1e-3                                   # average launch duration (s)
            evals = ctr["kde_evaluations"] / k_n
            G = 25
            logs = ctr["logs"] / k_n                      # log() evaluations the epilogue executed (counted by the kernel)
            exps = ctr["dense_exps"] / k_n + evals * (2 * G + 1)
            plain = ctr["plain_flops"] / k_n              # FMA / ADD / MUL outside exp and log, 2 flop per FMA (kernel-counted)
            flops = exps * FLOP_PER_EXP + logs * FLOP_PER_LOG + plain
            bytes_alg = cnt * t_cnt * (words * 8 + 8)
            # SURVEY 8(d)'s literal per-evaluation work, for EVERY row of the shard (the reference evaluates duplicates too):
            # exp = 25 n + 100, log = 16 900, plain flop = 50 n + 1.8e5  (n = 1000: 25 100 exp, 16 900 log, 2.3e5 flop)
            alg_eval = cnt * t_cnt
            alg_flops = alg_eval * ((G * n + 4 * G) * FLOP_PER_EXP + (G ** 3 + 2 * G * G + G) * FLOP_PER_LOG + 2 * G * n + 1.8e5)
            traffic, inst_per_eval, ncu_view = None, None, None
            try:  # DRAM bytes and warp instructions per launch from the committed ncu capture of this workload (1 GPU)
                tj = json.load(open(os.path.join(ROOT, "profiles", "k_score_traffic.json")))
                if tj["workload"] == wl_name and not args.features:
                    inst_per_eval = tj["warp_instructions_per_launch"] / tj["evaluations_per_launch"]
                    if world == 1:
                        traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
                        ncu_view = dict(tj.get("ncu") or {}, source=tj.get("source"))  # the committed capture's own counters
            except Exception:
                pass
            fp64_view = {
                "bound": "fp64", "achieved": flops / dur / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": flops / dur / 1e12 / fp64_peak if fp64_peak else None,
                "peak_source": "measured live: rvl_measure_fp64_peak (DFMA chain); MEASURED_PEAKS.json has no FP64 figure",
                "convention": "EXECUTED FP64 work from kernel counters: exp=%d flop, log=%d flop (FP64-pipe SASS instructions x2), "
                              "+ plain FMA/ADD; the soft cells (%d per launch, 9 FP32 instructions + 1 MUFU each) are not FP64 "
                              "work" % (FLOP_PER_EXP, FLOP_PER_LOG, int(ctr.get("soft_cells", 0) / k_n)),
            }
            # Since the one-class density columns moved to single precision the FP64 pipe is ~30 % busy and the kernel is bound by
            # INSTRUCTION ISSUE (one warp instruction per clock and scheduler): that is the ceiling the headline frac is quoted
            # against.  Numerator: warp instructions per evaluation of the committed ncu capture (same workload) x the evaluations
            # this launch executed; denominator: 4 schedulers x SMs x the SM clock sampled during the timed region.
            sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz")
            sms = torch.cuda.get_device_properties(local).multi_processor_count
            if inst_per_eval and sm_mhz:
                issue_peak = 4.0 * sms * sm_mhz * 1e6 / 1e9           # G warp-instructions/s
                issue_ach = inst_per_eval * evals / dur / 1e9
                top = {"bound": "issue", "achieved": issue_ach, "peak": issue_peak, "unit": "G warp-inst/s",
                       "frac": issue_ach / issue_peak,
                       "peak_source": "4 warp schedulers x %d SMs x %.0f MHz (SM clock sampled by NVML during the timed region)" % (sms, sm_mhz),
                       "convention": "%.0f warp instructions per evaluation (ncu smsp__inst_executed.sum of the committed capture, "
                                     "profiles/k_score_traffic.json) x %.0f evaluations executed per launch (kernel counter) / launch "
                                     "duration; the FP64-pipe view of earlier rounds is kept under `fp64`" % (inst_per_eval, evals)}
            else:
                top = dict(fp64_view)
            out["roofline"] = {
                "kernel": "k_search" if args.fused else "k_score", **top, "traffic": traffic, "fp64": fp64_view, "ncu": ncu_view,
                "launch_ms": dur * 1e3, "launches": int(k_n), "share_of_step": k_ms / (ms_prof * nprof),
                "measured": "CUDA events around every launch of the kernel on its stream, in %d extra searches with the "
                            "profiling switch on (%.3f ms per search there, %.3f ms in the uninstrumented timed region)"
                            % (nprof, ms_prof, ms),
                "per_launch": {"evaluations": evals, "exp": exps, "log": logs, "flop": flops,
                               "soft_cells": ctr.get("soft_cells", 0) / k_n},
                "algorithmic": {"convention": "SURVEY 8(d) literal work per evaluation (25n+100 exp, 16 900 log, 50n+1.8e5 flop) "
                                              "x every row of the shard, same flop weights; > 1 means work the kernel avoids "
                                              "(x-axis table, dead / factorised density columns, duplicate rows), not a fuller pipe",
                                "evaluations": alg_eval, "flop": alg_flops, "achieved": alg_flops / dur / 1e12,
                                "frac": alg_flops / dur / 1e12 / fp64_peak if fp64_peak else None},
                "hbm": {"bound": "hbm", "achieved": bytes_alg / dur / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_alg / dur / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": bytes_alg,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            }
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference_rate(wl_name, args.cpu_sample_rows, args.cpu_threads or HOST_CORES)
            cpu.pop("seconds", None)
            cpu.pop("rows", None)
            cpu["host_cores_available"] = HOST_CORES
            out["cpu_baseline"] = cpu
        ctx.close()
        return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    from revealer_b200 import synth
    B = Bench(args)
    res = B.run(args.workload, args.steps, args.warmup, main=True)
    extra = {}
    if not args.no_extra and not args.features and not args.iters:
        names = [k for k in ("C2_500x20k", "C5_256targets_1kx100k") if k != args.workload]
        if B.world == 8 and args.workload != "C4_10kx500k":
            names.append("C4_10kx500k")          # BASELINE: "10k samples x 500k features ... 8 GPU sharded"
        for k in names:
            extra[k] = B.run(k, 3, 3, main=False)
    if B.rank == 0:
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": B.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": res["config"],
            "evaluations_per_step": res["evaluations_per_step"],
            "evaluations_executed_per_step": res["evaluations_executed_per_step"], "value_executed": res["value_executed"],
            "check": res["check"], "e2e": res.get("e2e"), "e2e_gct": res.get("e2e_gct"),
            "gpu_launches": res["gpu_launches"], "clocks": res.get("clocks"),
            "roofline": res.get("roofline"), "cpu_baseline": res.get("cpu_baseline"), "configs": extra,
        }
        print(json.dumps(line))
    if B.world > 1:
        B.dist.destroy_process_group()


if __name__ == "__main__":
    main()
