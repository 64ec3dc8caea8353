#!/bin/bash
# One GPU-box call of the development loop: parity tests, the bench line (A/B on
# the gradient-only items), kernel times at 10k.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
( time python -m pytest tests -m gpu -q -x --timeout 900 ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
cat gpurun_out/bench_split.json | cut -c1-600
IID_GRAD_SPLIT=0 python bench.py --no-extras --steps 3 > gpurun_out/bench_nosplit.json 2> gpurun_out/bench_nosplit.err
cat gpurun_out/bench_nosplit.json | cut -c1-300
python scripts/gpu_time.py > gpurun_out/times10k.txt 2>&1
cat gpurun_out/times10k.txt
