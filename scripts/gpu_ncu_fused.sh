#!/bin/bash
# One full ncu capture (with source) of the fused evaluation kernel at Au561.
mkdir -p gpurun_out
sed -i 's/range(401)/range(30)/' scripts/fused_phase_times.py
ncu --set full --clock-control none --import-source on -k regex:fused_eval -s 12 -c 1 \
    -o gpurun_out/r2_fused561 -f python scripts/fused_phase_times.py > gpurun_out/r2_ncu_fused_run.log 2>&1
ncu -i gpurun_out/r2_fused561.ncu-rep --page raw --csv > gpurun_out/r2_fused561_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fused561.ncu-rep --page source --csv > gpurun_out/r2_fused561_source.csv 2>/dev/null
ncu -i gpurun_out/r2_fused561.ncu-rep --page source --print-source cuda --csv > gpurun_out/r2_fused561_cuda.csv 2>/dev/null
tail -3 gpurun_out/r2_ncu_fused_run.log
ls -la gpurun_out/ | grep fused561
