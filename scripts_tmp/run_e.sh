for v in "" _pf; do
  echo "=== variant $v"; SWEEP_QUICK=1 HDPDE_B200_LIB=$PWD/highdimpde.jl_b200/libhdpde_b200$v.so python profiles/sweep.py > gpurun_out/sweepe$v.txt 2>&1; grep -E "^C" gpurun_out/sweepe$v.txt | cut -c1-50,80-100
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
