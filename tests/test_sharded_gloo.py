"""world_size-2 gloo test (CPU) of the sharded search's HOST logic: row partition, candidate records,
the all-gather, and the deterministic winner rule every rank applies.  Scores come from the CPU oracle
(allowed in tests); the CUDA kernels' counterpart of this exchange is test_two_shards_equal_single_shard
(gpu) and bench.py --gpus N."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from revealer_b200 import synth
    from revealer_b200.api import resolve_winner, shard_rows
    n, F, iters = 60, 91, 3
    lo, cnt = shard_rows(F, world, rank)
    w = synth.make_workload(n=n, F=F, row_lo=lo, row_hi=lo + cnt)
    m2 = w["m2"].astype(float)
    seed = w["seed"].astype(float)
    tops = []
    for it in range(iters):
        col = O.score_rows(w["target"], m2, seed)
        k = O.top1(col) if cnt else -1
        rec = torch.zeros(2 + n, dtype=torch.float64)          # [score][global row][winner row]
        rec[0] = col[k] if cnt else float("nan")
        rec[1] = lo + k if cnt else -1
        if cnt:
            rec[2:] = torch.from_numpy(m2[k])
        out = [torch.zeros_like(rec) for _ in range(world)]
        dist.all_gather(out, rec)
        sc = [float(o[0]) for o in out]
        rows = [int(o[1]) for o in out]
        win = resolve_winner(sc, rows)
        seed = np.maximum(seed, out[win][2:].numpy())
        tops.append(rows[win])
    q.put((rank, tops, seed.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_search_matches_single_process(oracle):
    from revealer_b200 import synth
    world, port = 2, 29731 + os.getpid() % 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = synth.make_workload(n=60, F=91)
    ref = oracle.search(w["target"], w["m2"].astype(float), w["seed"].astype(float), 3)
    for rank, tops, seed in got:
        assert tops == list(ref["top"]), (rank, tops, ref["top"])
        assert np.array_equal(np.array(seed), ref["seed_iter"][-1])
