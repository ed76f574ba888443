"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: work sharding of
independent solves, NCCL unique-id distribution, result gathering (hdpde_b200.dist)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import hdpde_b200 as H
    from hdpde_b200 import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. sharding covers every solve exactly once, contiguous, balanced
        n = 1001
        lo, hi = D.shard_range(n, rank, world)
        sizes = D.all_gather_object((lo, hi))
        assert sizes[0][0] == 0 and sizes[-1][1] == n
        for a, b in zip(sizes[:-1], sizes[1:]):
            assert a[1] == b[0]
        assert max(h_ - l_ for l_, h_ in sizes) - min(h_ - l_ for l_, h_ in sizes) <= 1
        # 2. unique-id distribution: every rank ends up with rank 0's 128 bytes
        uid = D.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
        assert uid == bytes(range(128))
        # 3. gathering per-rank results in solve order; seeds follow the global solve index
        local = np.arange(lo, hi, dtype=np.float64) * 2.0
        full = D.gather_solves(local, n)
        assert np.array_equal(full, np.arange(n) * 2.0)
        assert D.local_seed(100, rank, world, n) == 100 + lo
        # 4. max-over-ranks timing reduction
        assert D.max_over_ranks(float(rank + 1)) == float(world)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_range_edges():
    import hdpde_b200  # noqa
    from hdpde_b200 import dist as D
    assert D.shard_range(0, 0, 4) == (0, 0)
    assert [D.shard_range(3, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert D.shard_range(10, 0, 1) == (0, 10)
