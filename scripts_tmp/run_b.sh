for v in "" _nopf _inl _nopfinl; do
  echo "=== variant $v"; SWEEP_QUICK=1 HDPDE_B200_LIB=$PWD/highdimpde.jl_b200/libhdpde_b200$v.so python profiles/sweep.py > gpurun_out/sweepb$v.txt 2>&1; grep -E "^C" gpurun_out/sweepb$v.txt | cut -c1-50,80-100
done
