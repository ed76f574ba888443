python -m pytest tests -m gpu -x -q 2>&1 | tail -2
SWEEP_QUICK=1 python profiles/sweep.py > gpurun_out/sweepm.txt 2>&1; cat gpurun_out/sweepm.txt | cut -c1-50,80-100
