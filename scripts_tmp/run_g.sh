set -x
python bench.py --steps 2 --warmup 3 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python profiles/prof_workload.py mlp 4096 > gpurun_out/plain_mlp.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mlp_units -s 1 -c 1 -o gpurun_out/prof_mlp_units_r1j -f python profiles/prof_workload.py mlp 4096 > gpurun_out/ncu_mlp.log 2>&1
python profiles/prof_workload.py sde > gpurun_out/plain_sde.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sde_loop -s 1 -c 1 -o gpurun_out/prof_sde_loop_r1j -f python profiles/prof_workload.py sde > gpurun_out/ncu_sde.log 2>&1
tail -2 gpurun_out/plain_sde.log
