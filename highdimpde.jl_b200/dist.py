"""Multi-GPU plumbing: one process per GPU (torchrun), `torch.distributed` only for
rendezvous-level exchanges (NCCL unique id, result gathering, timing reductions).

How the path shards:
 * independent solves (different seeds / start points): contiguous shards of the solve
   index, NO data-path collective — `shard_range`, `gather_solves`;
 * one large solve: every rank evaluates a slice of the tree's work units and the
   per-unit partial sums are all-reduced by NCCL inside libhdpde_b200
   (`attach_nccl` → hdp_comm_init_rank; the reference's analogue is the
   `Threads.@spawn`/`fetch`/`sum` of src/MLP.jl:235-259).
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of n items for `rank` of `world`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def local_seed(seed: int, rank: int, world: int, n: int) -> int:
    """Solve i always uses seed+i, whatever the number of ranks."""
    return seed + shard_range(n, rank, world)[0]


def _dist():
    import torch.distributed as dist
    return dist


def all_gather_object(obj):
    dist = _dist()
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def broadcast_bytes(data: bytes | None, n: int, src: int = 0) -> bytes:
    import torch
    dist = _dist()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(n, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.tensor(list(data), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())


def broadcast_int(v: int, src: int = 0) -> int:
    """The same 64-bit integer on every rank (rank `src`'s) — e.g. the seed of a sharded solve."""
    data = int(v).to_bytes(8, "little") if _dist().get_rank() == src else None
    return int.from_bytes(broadcast_bytes(data, 8, src), "little")


def gather_solves(local: np.ndarray, n: int) -> np.ndarray:
    """All ranks' per-solve results, concatenated in solve order."""
    parts = all_gather_object(np.asarray(local))
    full = np.concatenate(parts) if parts else np.zeros(0)
    assert full.shape[0] == n
    return full


def max_over_ranks(x: float) -> float:
    import torch
    dist = _dist()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def min_over_ranks(x: float) -> float:
    import torch
    dist = _dist()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())


def attach_nccl(handle, peer_windows: bool = True):
    """Join this rank's handle to an NCCL communicator over all ranks and — ranks of one node with a peer path —
    attach the peer exchange windows (`attach_peer_windows`), so that the exchange of a sharded solve is the
    library's own kernel over NVLink peer memory; NCCL stays for exchanges larger than a window slot."""
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = handle.nccl_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128)
    handle.comm_init_rank(uid, world, rank)
    if peer_windows and world > 1:
        attach_peer_windows(handle)
    return rank, world


def attach_peer_windows(handle) -> bool:
    """hdp_peer_window_export on every rank, the 64-byte IPC handles gathered in rank order, hdp_peer_window_attach.
    All or nothing: if any rank cannot map a peer (different nodes, no P2P path, more than 16 ranks) every rank
    detaches and sharded solves keep the ncclAllReduce.  Returns whether the windows are attached."""
    from . import _capi
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    ok, mine = True, b""
    if world > _capi.HDP_PEER_MAX_RANKS:
        ok = False
    else:
        try:
            mine = handle.peer_window_export()
        except _capi.HdpError:
            ok = False
    parts = all_gather_object((ok, mine))
    ok = all(p[0] for p in parts)
    if ok:
        try:
            handle.peer_window_attach(b"".join(p[1] for p in parts), world, rank)
        except _capi.HdpError:
            ok = False
    ok = all(all_gather_object(ok))
    if not ok:
        handle.peer_window_detach()
    return ok
