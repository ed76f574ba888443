"""Peer exchange windows (hdp_peer_window_*, include/hdpde.h): argument checking without a GPU, export / attach
rules on one GPU.  The exchange itself needs two GPUs: tests/test_gpu_multi.py (profiles/multi_gpu_check.py runs
every sharded solve through the peer windows AND through the ncclAllReduce and compares)."""
import ctypes as C

import numpy as np
import pytest

import hdpde_b200 as H
from hdpde_b200 import _capi as capi


def test_peer_window_null_arguments():
    lib = capi.load_library()
    buf = (C.c_uint8 * capi.HDP_PEER_HANDLE_BYTES)()
    assert lib.hdp_peer_window_export(None, buf) == capi.HDP_ERR_BAD_ARG
    assert lib.hdp_peer_window_attach(None, buf, 2, 0) == capi.HDP_ERR_BAD_ARG
    assert lib.hdp_peer_window_detach(None) == capi.HDP_ERR_BAD_ARG


def test_peer_window_constants_match_header():
    import os, re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "hdpde.h")).read()
    for name in ("HDP_PEER_HANDLE_BYTES", "HDP_PEER_MAX_RANKS", "HDP_PEER_SLOT_DOUBLES"):
        assert int(re.search(r"#define %s (\d+)" % name, hdr).group(1)) == getattr(capi, name)
    assert C.sizeof(capi.hdp_mlp_stats) % 8 == 0 and hasattr(capi.hdp_mlp_stats, "exchange_kind")


@pytest.mark.gpu
def test_peer_window_export_and_attach_rules(handle):
    h = handle
    ipc = h.peer_window_export()
    assert len(ipc) == capi.HDP_PEER_HANDLE_BYTES and any(ipc)
    # windows follow the communicator: without hdp_comm_init_rank there is nothing to attach to
    with pytest.raises(H.HdpError) as e:
        h.peer_window_attach(ipc * 2, 2, 0)
    assert "communicator" in str(e.value)
    with pytest.raises(H.HdpError):
        h.peer_window_attach(ipc * 17, 17, 0)
    h.peer_window_detach()
    # the handle still solves, on one GPU, with no exchange
    d = 10
    prob = H.ParabolicPDEProblem(0.0, 1.0, np.zeros(d), (0.0, 0.5), g=H.ExpSumSq(0.25), f=H.FAllenCahn(1.0),
                                 neumann_bc=(np.full(d, -0.5), np.full(d, 0.5)))
    u, st = H._ml_picard_batch(prob, H.MLP(M=3, L=3, K=4), n_solves=1, seed=5, handle=h)
    assert st["exchange_kind"] == "none" and st["exchange_bytes"] == 0 and np.isfinite(u[0])
