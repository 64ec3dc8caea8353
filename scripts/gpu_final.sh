#!/bin/bash
# The records of a round's last state (one GPU): parity tests, ncu capture of the
# histogram F(Q) pass, configs[4] on one GPU, the bench line of both arms.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
bash scripts/gpu_ncu_fqhist.sh
python scripts/gpu_cfg5.py > gpurun_out/r2_cfg5_1gpu.txt 2>&1; cat gpurun_out/r2_cfg5_1gpu.txt
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python -c "import json;d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['roofline']['frac'],d['e2e'],d['gpu_launches'],d['clocks']);print(d['extras']['hmc_au561'])"
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
