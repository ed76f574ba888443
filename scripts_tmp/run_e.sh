python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sde or mlt or reflect or sample" 2>&1 | tail -2
echo "== unroll 2"; python scripts_tmp/k1.py
echo "== unroll 1"; HDPDE_B200_LIB=$PWD/highdimpde.jl_b200/libhdpde_b200_k1u1.so python scripts_tmp/k1.py
