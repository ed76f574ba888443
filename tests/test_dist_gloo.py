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


class _FakeHandle:
    """Stands in for hdpde_b200.Handle in the handshake test: records the calls, fails where told to."""

    def __init__(self, H, rank, fail_export=False, fail_attach=False):
        self.H, self.rank, self.fail_export, self.fail_attach = H, rank, fail_export, fail_attach
        self.attached, self.detached, self.seen = None, 0, None

    def peer_window_export(self):
        if self.fail_export:
            raise self.H.HdpError(self.H.capi.HDP_ERR_CUDA, "no window")
        return bytes([self.rank + 1]) * self.H.capi.HDP_PEER_HANDLE_BYTES

    def peer_window_attach(self, handles, nranks, rank):
        self.seen = (handles, nranks, rank)
        if self.fail_attach:
            raise self.H.HdpError(self.H.capi.HDP_ERR_CUDA, "no peer path")
        self.attached = nranks

    def peer_window_detach(self):
        self.detached += 1
        self.attached = None


def _peer_window_handshake(rank, world, H, D):
    nb = H.capi.HDP_PEER_HANDLE_BYTES
    # every rank fine: attached everywhere, each rank was given all handles in rank order
    h = _FakeHandle(H, rank)
    assert D.attach_peer_windows(h) is True
    assert h.attached == world and h.detached == 0
    assert h.seen == (b"".join(bytes([r + 1]) * nb for r in range(world)), world, rank)
    # the last rank cannot map its peers: everyone detaches, including the ranks whose own attach succeeded
    h = _FakeHandle(H, rank, fail_attach=(rank == world - 1))
    assert D.attach_peer_windows(h) is False
    assert h.attached is None and h.detached == 1
    # rank 0 cannot even export: nobody tries to attach, everyone detaches
    h = _FakeHandle(H, rank, fail_export=(rank == 0))
    assert D.attach_peer_windows(h) is False
    assert h.seen is None and h.attached is None and h.detached == 1


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
        # 4. max-over-ranks timing reduction; the seed of a sharded solve is rank 0's
        assert D.max_over_ranks(float(rank + 1)) == float(world) and D.min_over_ranks(float(rank + 1)) == 1.0
        assert D.broadcast_int(0xFEDCBA9876543210 + rank) == 0xFEDCBA9876543210
        # 5. ONE solve sharded over the two ranks — the library's N > 1 data path with the product's device code
        #    compiled for the host (tests/hostemu): each process runs its rank's geometry, units and local combine,
        #    the exchange buffers are summed by a real all-reduce, each rank finishes the crown.
        _sharded_solve(rank, world, H, D)
        # 6. the peer-window handshake (dist.attach_peer_windows) is all or nothing: handles gathered in rank order,
        #    and one rank that cannot export or cannot map a peer makes EVERY rank detach (they then all keep NCCL)
        _peer_window_handshake(rank, world, H, D)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _sharded_solve(rank, world, H, D):
    import ctypes as C
    import subprocess
    sys_path = os.path.dirname(os.path.abspath(__file__))
    import sys
    sys.path.insert(0, sys_path)
    from helpers import zoo, random_tapes
    from oracle import oracle as O
    O.build()
    emu_dir = os.path.join(sys_path, "hostemu")
    if rank == 0:
        subprocess.run(["make", "-C", emu_dir], check=True, capture_output=True)
    dist.barrier()
    emu = C.CDLL(os.path.join(emu_dir, "_build", "libhdp_hostemu.so"))
    emu.hdp_emu_shard_begin.restype = C.c_void_p
    for name, target in (("nonlocal_ac_neumann_d5", 400), ("allen_cahn_neumann_d5", 60)):
        prob, alg, dt = zoo(H)[name]
        tapes = random_tapes(H, prob, alg, dt, seed=29)
        pb, keep = prob.to_c()
        sm, k2 = alg.mc_sample._descriptor(prob.d)
        uo, info = O.ml_picard(pb, sm, alg.M, alg.L, alg.K, np.asarray(prob.x, dtype=dt), O.Rng.tape(dt, *tapes))
        cfg = H.capi.hdp_mlp_cfg()
        cfg.M, cfg.L, cfg.K, cfg.n_solves, cfg.n_x0, cfg.seed, cfg.target_units = alg.M, alg.L, alg.K, 1, 1, 0, target
        cfg.sampler = sm
        ts = [np.ascontiguousarray(t, dtype=dt) if t is not None else None for t in tapes]
        cfg.r_tape, cfg.n_tape = ts[0].ctypes.data, ts[1].ctypes.data
        if ts[2] is not None:
            cfg.s_tape = ts[2].ctypes.data
        x = np.ascontiguousarray(prob.x, dtype=dt)
        n = C.c_int64(0)
        e = emu.hdp_emu_shard_begin(C.byref(pb), C.byref(cfg), x.ctypes.data_as(C.c_void_p), world, rank, C.byref(n))
        assert e
        xb = np.zeros(n.value)
        emu.hdp_emu_shard_xbuf(C.c_void_p(e), rank, xb.ctypes.data_as(C.c_void_p))
        t = torch.from_numpy(xb)
        dist.all_reduce(t)  # the one collective of the path (ncclAllReduce(double, sum) on the GPUs)
        u = np.zeros(1)
        st = H.capi.hdp_mlp_stats()
        rc = emu.hdp_emu_shard_end(C.c_void_p(e), rank, C.byref(cfg), xb.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p),
                                   C.byref(st))
        assert rc == 0
        assert abs(u[0] - uo) <= 1e-12 * abs(uo), (name, u[0], uo)
        assert st.bounces == info["bounces"]          # every wall hit counted once across the ranks
        both = D.all_gather_object(float(u[0]))
        assert both[0] == both[1]                     # the ranks finish with identical bits
        assert n.value * 8 < 64 * 1024                # the exchange is a double per crown node, not per unit


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
