for v in _u2mb3 _u3mb3 _u3mb4 _u4mb3 _u4mb2; do
  echo "=== variant $v"; SWEEP_QUICK=1 HDPDE_B200_LIB=$PWD/highdimpde.jl_b200/libhdpde_b200$v.so python profiles/sweep.py > gpurun_out/sweeph$v.txt 2>&1; grep -E "^C(1|2|4 |5 )" gpurun_out/sweeph$v.txt | cut -c1-50,80-100
done
