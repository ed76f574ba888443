set -x
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python profiles/prof_workload.py mlp 4096 > gpurun_out/plain_mlp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mlp_units -s 1 -c 1 -o gpurun_out/prof_mlp_units_r1k -f python profiles/prof_workload.py mlp 4096 > gpurun_out/ncu_mlp.log 2>&1
tail -1 gpurun_out/plain_mlp.log | cut -c1-200
