set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
for v in "" _mb5 _mb4; do
  echo "=== variant $v"; HDPDE_B200_LIB=$PWD/highdimpde.jl_b200/libhdpde_b200$v.so python profiles/sweep.py > gpurun_out/sweep$v.txt 2>&1; grep -E "C1.*n=  4096|C2.*n=  (512|4096)|C4 |C5 .*n=    64" gpurun_out/sweep$v.txt
done
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; cat gpurun_out/bench_a.json | head -c 1500
