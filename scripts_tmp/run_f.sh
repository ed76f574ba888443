python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_r1j.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']);
for k,v in d['extra'].items(): print(k, {a:b for a,b in v.items() if a in ('kernel_ms','device_ms','solves_per_sec','coord_steps_per_sec')})"
python profiles/sweep.py > gpurun_out/sweep_r1j.txt 2>&1; cat gpurun_out/sweep_r1j.txt | cut -c1-50,80-100
