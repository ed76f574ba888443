import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def _ensure_built():
    # `make` is a no-op when the in-tree library is up to date; it prevents testing a stale .so
    import __graft_entry__ as g

    g.build()


_ensure_built()


@pytest.fixture(scope="session")
def H():
    import hdpde_b200

    return hdpde_b200


@pytest.fixture(scope="session")
def O():
    from oracle import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def emu():
    """Host emulation (G = 1) of the product's MLP device code — test infrastructure."""
    import ctypes as C

    d = os.path.join(ROOT, "tests", "hostemu")
    subprocess.run(["make", "-C", d], check=True, capture_output=True)
    return C.CDLL(os.path.join(d, "_build", "libhdp_hostemu.so"))


@pytest.fixture(scope="session")
def handle(H):
    try:
        h = H.Handle(0)
    except H.HdpError as e:
        if e.code == H.capi.HDP_ERR_NO_DEVICE:  # plain `pytest tests` on a machine without a B200: skip, don't error
            pytest.skip("no sm_100 device: the hot path has no CPU fallback (run with -m gpu on a B200)")
        raise
    yield h
    h.close()
