#!/bin/bash
# ncu evidence for the bench workload: launch list and one full capture of the
# dominant kernel (numbers printed under ncu are never bench values).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches_bench50k.csv python bench.py --steps 2 --warmup 1 --no-extras \
    > gpurun_out/ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:debye2 -s 2 -c 1 \
    -o gpurun_out/grad50k -f python bench.py --steps 1 --warmup 1 --no-extras \
    > gpurun_out/ncu_full_run.log 2>&1
ncu -i gpurun_out/grad50k.ncu-rep --page raw --csv > gpurun_out/grad50k_raw.csv 2>/dev/null
ncu -i gpurun_out/grad50k.ncu-rep --page source --csv > gpurun_out/grad50k_source.csv 2>/dev/null
ls -la gpurun_out/
