"""bench.py's roofline block reads DRAM traffic and executed-instruction counts from the committed ncu export of the
units kernel — with the export's units applied (round 1 read 2.24 Gbyte as 2.24 MB).  CPU-only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles", "tools"))


def test_units_profile_is_read_with_units():
    import bench
    import ncu_metrics as NM
    prof = bench.units_profile()
    assert prof["source"].startswith("profiles/r2") and prof["source"].endswith("_mlp_units_full_metrics.csv")
    m = NM.load(os.path.join(ROOT, prof["source"]))
    assert m["dram__bytes_read.sum#unit"] in ("Gbyte", "Mbyte", "Kbyte", "byte")
    # one launch of 4096 C2 solves: gigabytes of scratch-row traffic, a few milliseconds, ~20 fp64 instr per coordinate-step
    assert 1e9 < prof["traffic"] < 2e10
    assert prof["traffic"] == prof["dram_read"] + prof["dram_write"]
    assert 3.0 < prof["kernel_ms_under_ncu"] < 10.0
    assert 15.0 < prof["fp64_instr_per_coord_step"] < 39.0
    assert 100.0 < prof["warp_instr_per_vec_group"] < 170.0
    assert 0.5 < prof["dispatch_port_busy"] < 1.0
    # the capture is of the kernel the bench line names
    assert "mlp_units_kernel<double, 1, 4, 2>" in m["Kernel Name"]
