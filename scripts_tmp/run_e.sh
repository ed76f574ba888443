python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python profiles/sweep.py > gpurun_out/sweepk.txt 2>&1; cat gpurun_out/sweepk.txt | cut -c1-50,80-100
