python -m pytest tests -m gpu -x -q 2>&1 | tail -2
SWEEP_QUICK=1 python profiles/sweep.py > gpurun_out/sweepg.txt 2>&1; grep -E "^C" gpurun_out/sweepg.txt | cut -c1-50,80-100
