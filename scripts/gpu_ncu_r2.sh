#!/bin/bash
# Round-2 ncu evidence for the bench workload (row-ownership gradient kernel):
# launch list of a bench run and one full capture of the dominant kernel
# (numbers printed under ncu are never bench values).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/r2_launches_bench50k.csv python bench.py --steps 2 --warmup 3 --no-extras \
    > gpurun_out/r2_ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:debye2 -s 3 -c 1 \
    -o gpurun_out/r2_grad50k -f python bench.py --steps 1 --warmup 3 --no-extras \
    > gpurun_out/r2_ncu_full_run.log 2>&1
ncu -i gpurun_out/r2_grad50k.ncu-rep --page raw --csv > gpurun_out/r2_grad50k_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_grad50k.ncu-rep --page source --csv > gpurun_out/r2_grad50k_source.csv 2>/dev/null
ls -la gpurun_out/ | grep r2_
