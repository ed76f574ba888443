python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in "" _rec; do
  echo "=== variant $v"; SWEEP_QUICK=1 HDPDE_B200_LIB=$PWD/highdimpde.jl_b200/libhdpde_b200$v.so python profiles/sweep.py > gpurun_out/sweepc$v.txt 2>&1; grep -E "^C" gpurun_out/sweepc$v.txt | cut -c1-50,80-100
done
