import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import hdpde_b200 as H
from hdpde_b200 import dist as HD
import bench
local_rank = int(os.environ.get("LOCAL_RANK", 0)); rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl")
hc = H.Handle(local_rank)
HD.attach_nccl(hc)
names = sys.argv[1:] or ["c4_nonlocal_fisher_d10_L5_one_solve"]
W = bench.strong_workloads(H)
for name in names:
    kw = {}
    if ":" in name:
        name, g = name.split(":"); kw["group_size"] = int(g)
    prob, alg = W[name]
    for r in range(4):
        dist.barrier(); torch.cuda.synchronize()
        if r == 3 and rank in (0, 5): os.environ["HDP_PHASE_TRACE"] = "1"
        u, st = H._ml_picard_batch(prob, alg, n_solves=1, seed=31, handle=hc, **kw)
        os.environ.pop("HDP_PHASE_TRACE", None)
    if rank == 0: print(name, kw, float(u[0]), st["device_ms"], st["phase_ms"], st["n_units"], st["n_launches"], flush=True)
dist.destroy_process_group()
