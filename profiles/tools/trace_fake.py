"""Per-launch view of the chain of ONE rank of an N-rank sharded solve on a single GPU (trace build of the library,
HDP_TRACE_FAKE_RANKS=r/N: sharding built for N ranks, exchange left out — timings only, the value is rank r's share).
  HDPDE_B200_LIB=$PWD/highdimpde.jl_b200/libhdpde_b200_trace.so HDP_TRACE_FAKE_RANKS=0/8 python profiles/tools/trace_fake.py [names...]
Under `ncu --metrics gpu__time_duration.sum` the same script gives the kernels' own durations."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, hdpde_b200 as H
import bench
h = H.Handle(0)
W = bench.strong_workloads(H)
names = sys.argv[1:] or list(W)
reps = int(os.environ.get("TRACE_REPS", "4"))
for name in names:
    prob, alg = W[name]
    for r in range(reps):
        if r == reps - 1: os.environ["HDP_PHASE_TRACE"] = "1"
        u, st = H._ml_picard_batch(prob, alg, n_solves=1, seed=31, handle=h)
        os.environ.pop("HDP_PHASE_TRACE", None)
    print(name, float(u[0]), round(st['device_ms'], 3), {k: round(v, 3) for k, v in st['phase_ms'].items()}, st['n_units'], 'G', st['group_size'],
          'launches', st['n_launches'], flush=True)
