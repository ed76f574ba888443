#!/bin/bash
# final captures of the round (one GPU).  Plain runs first, ncu after each has exited 0.
set -x
R=r2z
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${R}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${R}_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference_arm.json 2>> gpurun_out/${R}_bench.err; echo "ref rc=$?"
python profiles/sweep.py > gpurun_out/${R}_sweep.txt 2>&1
python profiles/sde_sweep.py > gpurun_out/${R}_sde_sweep.txt 2>&1
python profiles/tools/single_solve_ab.py 15 > gpurun_out/${R}_single_solves.txt 2>&1
for w in "mlp 4096" "c5 64" "c1 4096" "sde" "ens"; do
  tag=$(echo $w | cut -d' ' -f1)
  python profiles/prof_workload.py $w > gpurun_out/${R}_plain_$tag.log 2>&1 || { echo "plain $tag failed"; continue; }
  case $tag in mlp|c5|c1) k=mlp_units;; sde) k=sde_loop;; ens) k=em_ensemble;; esac
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_${tag}_${R} python profiles/prof_workload.py $w > gpurun_out/${R}_ncu_$tag.log 2>&1
  ncu -i gpurun_out/prof_${tag}_${R}.ncu-rep --page raw --csv > gpurun_out/${R}_${tag}_full_metrics.csv 2>/dev/null
done
python bench.py --steps 2 --warmup 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${R}_ncu_bench.log 2>&1
ls -la gpurun_out | grep $R
