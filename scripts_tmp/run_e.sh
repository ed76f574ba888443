echo "== unroll 2"; python scripts_tmp/k1.py
echo "== unroll 3"; HDPDE_B200_LIB=$PWD/highdimpde.jl_b200/libhdpde_b200_k1u3.so python scripts_tmp/k1.py
echo "== unroll 4"; HDPDE_B200_LIB=$PWD/highdimpde.jl_b200/libhdpde_b200_k1u4.so python scripts_tmp/k1.py
