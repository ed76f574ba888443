python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_r1j.json 2> gpurun_out/bench_8gpu_r1j.err; echo "bench8 rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_8gpu_r1j.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 profiles/multi_gpu_check.py 5 2>&1 | grep "\[multi\]"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -2
