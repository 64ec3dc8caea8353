#!/bin/bash
# ncu evidence for the fused evaluation kernel at Au561: the launch list of a
# run of 16-step leapfrog chains (one launch per chain) and one full capture
# (with source) of one launch that walks 4 steps.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/r2_launches_leapfrog_chain_au561.csv env LF_REPS=12 python scripts/lf_phase_times.py 16 \
    > gpurun_out/r2_ncu_chain_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_eval -s 8 -c 1 \
    -o gpurun_out/r2_fused561 -f env LF_REPS=12 python scripts/lf_phase_times.py 4 > gpurun_out/r2_ncu_fused_run.log 2>&1
ncu -i gpurun_out/r2_fused561.ncu-rep --page raw --csv > gpurun_out/r2_fused561_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fused561.ncu-rep --page source --csv > gpurun_out/r2_fused561_source.csv 2>/dev/null
tail -3 gpurun_out/r2_ncu_fused_run.log
ls -la gpurun_out/ | grep -i "fused561\|chain_au561"
