python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
SWEEP_QUICK=1 python profiles/sweep.py > gpurun_out/sweepd.txt 2>&1; grep -E "^C" gpurun_out/sweepd.txt | cut -c1-50,80-100
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_d.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value']);print(d['extra'])"
