#!/bin/bash
# One GPU-box call of the development loop: smoke, parity tests, the bench line
# of both arms, kernel times at 10k.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( time python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "import json;d=json.load(open('gpurun_out/bench.json'));print(d['ms_per_step'],d['roofline']['frac'],d['e2e']['ms_per_step'],d['gpu_launches'],d['clocks']);print(d['extras']['hmc_au561'])"; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
python scripts/gpu_time.py > gpurun_out/times10k.txt 2>&1
cat gpurun_out/times10k.txt
