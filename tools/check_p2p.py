"""2+-rank check (torchrun): the peer-memory exchange must select the same features as the NCCL all-gather
and as a single-rank run of the whole matrix."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import revealer_b200 as rb
from revealer_b200 import synth

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, F, iters = 300, 5000, 6
lo, cnt = rb.shard_rows(F, world, rank)
w = synth.make_workload(n=n, F=F, row_lo=lo, row_hi=lo + cnt)
res = {}
for mode in ("nccl", "p2p"):
    ctx = rb.RevealerContext(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.set_shard(rank, world, lo, F)
    ctx.load_matrix(w["m2"]); ctx.set_targets(w["target"]); ctx.set_seed(w["seed"])
    nb = ctx.candidate_bytes()
    send = torch.zeros(nb, dtype=torch.uint8, device="cuda"); recv = torch.zeros(nb * world, dtype=torch.uint8, device="cuda")
    ctx.set_exchange_buffers(send.data_ptr(), recv.data_ptr())
    if mode == "p2p":
        mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).cuda()
        allh = torch.empty(64 * world, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allh, mine)
        ctx.p2p_import(allh.cpu().numpy().tobytes())
    for rep in range(2):                      # two searches: the stamp epoch must separate them
        ctx.search_begin(iters)
        for it in range(iters):
            ctx.iter_score(it)
            if mode == "nccl":
                dist.all_gather_into_tensor(recv, send)
            ctx.iter_merge(it)
        res[mode] = ctx.search_fetch(iters)
    dist.barrier()
    ctx.close()
ok = all(np.array_equal(res["nccl"][k], res["p2p"][k]) for k in ("top", "seed_iter", "cmi", "cmi_seed", "top_score"))
if rank == 0:
    full = synth.make_workload(n=n, F=F)
    with rb.RevealerContext(local) as c1:
        c1.load_matrix(full["m2"]); c1.set_targets(full["target"]); c1.set_seed(full["seed"])
        single = c1.search(iters)
    ok1 = np.array_equal(single["top"], res["p2p"]["top"]) and np.array_equal(single["seed_iter"], res["p2p"]["seed_iter"])
    print("p2p == nccl:", ok, "| p2p == single rank:", ok1, "| winners", res["p2p"]["top"][0])
dist.destroy_process_group()
sys.exit(0 if ok else 1)
