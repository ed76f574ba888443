"""CPU-only: the C-ABI library loads, exports every symbol include/hdpde.h declares, the
ctypes structs have the C layout, the host-side mirror validates its inputs, and the
product fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "hdpde.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hdp_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(H):
    lib = H.load_library()
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hdpde.h but not exported"
        assert n in H.capi.SYMBOLS, f"{n} has no ctypes binding"
    assert set(H.capi.SYMBOLS) == set(names)
    assert lib.hdp_version() == 100


def test_struct_layouts_match_c(H, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hdpde.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(hdp_sampler),sizeof(hdp_problem),sizeof(hdp_mlp_cfg),sizeof(hdp_mlp_counts),sizeof(hdp_mlp_stats),"
                   "sizeof(hdp_devinfo),offsetof(hdp_problem,lo),offsetof(hdp_mlp_cfg,sampler),offsetof(hdp_mlp_cfg,r_tape));"
                   'printf("%zu %zu\\n", offsetof(hdp_mlp_cfg,user), sizeof(hdp_user_exprs));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    c = H.capi
    want = [C.sizeof(c.hdp_sampler), C.sizeof(c.hdp_problem), C.sizeof(c.hdp_mlp_cfg), C.sizeof(c.hdp_mlp_counts),
            C.sizeof(c.hdp_mlp_stats), C.sizeof(c.hdp_devinfo), c.hdp_problem.lo.offset, c.hdp_mlp_cfg.sampler.offset,
            c.hdp_mlp_cfg.r_tape.offset, c.hdp_mlp_cfg.user.offset, C.sizeof(c.hdp_user_exprs)]
    assert [int(v) for v in out] == want


def test_no_cpu_fallback(H):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(H.HdpError) as e:
        H.Handle(0)
    assert e.value.code == H.capi.HDP_ERR_NO_DEVICE
    prob = H.ParabolicPDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), g=H.SumSq(), f=H.FZero())
    with pytest.raises(H.HdpError):
        H.solve(prob, H.MLP())


def test_product_never_imports_oracle():
    """The oracle and the host emulation are test infrastructure: nothing under the product
    package may import, link or execute them."""
    pkg = os.path.join(ROOT, "highdimpde.jl_b200")
    banned = ("import oracle", "from oracle", "hdp_oracle", "oracle/", "hostemu", "libhdp_hostemu")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                for b in banned:
                    assert b not in txt, f"{os.path.join(dirpath, f)} references {b!r}"
    # and the shared library does not link them
    out = subprocess.run(["ldd", os.path.join(pkg, "libhdpde_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "hostemu" not in out


def test_host_mirror_validation(H):
    with pytest.raises(TypeError):
        H.ParabolicPDEProblem(lambda x, p, t: 0.0, 0.1, np.zeros(3), (0.0, 0.5), g=H.SumSq(), f=H.FZero())
    with pytest.raises(TypeError):
        H.ParabolicPDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), g=lambda x: 0.0, f=H.FZero())
    with pytest.raises(TypeError):
        H.ParabolicPDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), g=H.SumSq(), f=H.FNLFisher())
    with pytest.raises(ValueError):
        H.ParabolicPDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), g=H.SumSq(), f=H.FZero(),
                              neumann_bc=(np.zeros(2), np.ones(2)))
    # MLP requires x0_sample = NoSampling (src/MLP.jl:65-66) — checked before any device work
    prob = H.ParabolicPDEProblem(0.0, 0.1, np.zeros(3), (0.0, 0.5), g=H.SumSq(), f=H.FZero(),
                                 x0_sample=H.UniformSampling(-1.0, 1.0))
    with pytest.raises(H.HdpError) as e:
        H._ml_picard_batch(prob, H.MLP(), handle=object())
    assert e.value.code == H.capi.HDP_ERR_X0_SAMPLE_NOT_SUPPORTED
    # kinds are ordinary callables too (g(x) in PIDESolution, src/MLP.jl:99)
    assert H.SumSq(2.0)(np.array([1.0, 2.0])) == 7.0
    assert abs(H.ExpSumSq(0.25)(np.array([2.0])) - np.exp(-1.0)) < 1e-15
    assert H.NoSampling()(np.zeros(3)) is None and not H._integrate(H.NoSampling())
    pb, keep = H.PIDEProblem(0.0, 0.1, np.zeros(4), (0.0, 0.5), H.GaussDensity(0.01), H.FNLReplicatorMutator(1.0),
                             neumann_bc=(-1.0, 1.0)).to_c()
    assert pb.d == 4 and pb.has_bc == 1 and pb.bc_is_scalar == 1 and pb.f_kind == 18 and pb.g_kind == 4


def test_counts_entry_point_needs_no_gpu(H):
    assert H.mlp_counts(4, 4, 10)["path_steps"] == 4916
    with pytest.raises(H.HdpError):
        H.mlp_counts(4, 9, 10)
