"""Sharded single solves under torchrun: the library's exchange kernel over peer windows against the ncclAllReduce
it replaces, same box, same build (best and median of REPS, max over ranks).
  torchrun ... profiles/tools/strong_ab.py [REPS]"""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
import hdpde_b200 as H
from hdpde_b200 import dist as HD
import bench
local_rank = int(os.environ.get("LOCAL_RANK", 0)); rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 9
hp = H.Handle(local_rank); HD.attach_nccl(hp)
hn = H.Handle(local_rank); HD.attach_nccl(hn, peer_windows=False)
for name, (prob, alg) in bench.strong_workloads(H).items():
    for tag, hc in (("peer", hp), ("nccl", hn)):
        rows = []
        for r in range(reps + 2):
            dist.barrier(); torch.cuda.synchronize()
            u, st = H._ml_picard_batch(prob, alg, n_solves=1, seed=31, handle=hc)
            ms = HD.max_over_ranks(st["device_ms"])
            if r >= 2: rows.append((ms, st["phase_ms"]))
        rows.sort(key=lambda t: t[0])
        if rank == 0:
            print(name, tag, st["exchange_kind"], repr(float(u[0])), "best %.4f median %.4f ms" % (rows[0][0], rows[len(rows) // 2][0]),
                  {k: round(v, 4) for k, v in rows[0][1].items()}, "launches", st["n_launches"], "G", st["group_size"], flush=True)
hp.close(); hn.close()
dist.destroy_process_group()
