#!/bin/bash
# One GPU-box call of the development loop: parity tests, the bench line, A/B
# runs of the kernel options.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
( time python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python -c "import json;d=json.load(open('gpurun_out/bench.json'));print(d['roofline']['frac'],d['e2e']['ms_per_step'],d['extras']['hmc_au561'])"
for v in $AB; do
  env $v python bench.py --no-extras --steps 3 > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python -c "import json,sys;d=json.load(open('gpurun_out/bench_$v.json'));print('$v',d['ms_per_step'],d['roofline']['frac'])"
done
python scripts/gpu_time.py > gpurun_out/times10k.txt 2>&1
cat gpurun_out/times10k.txt
