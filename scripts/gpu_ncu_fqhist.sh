#!/bin/bash
# ncu evidence for the histogram F(Q) pass at Au 10k: launch list of one get_fq
# and one full capture of fq_hist_kernel.
mkdir -p gpurun_out
cat > /tmp/fqhist_target.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from pyiid_b200 import ElasticScatter, structures
atoms = structures.fcc_sphere('Au', 10000)
scat = ElasticScatter(precision='fp32', device=0)
for _ in range(4):
    fq = scat.get_fq(atoms)
print(float(abs(fq).max()))
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_fq_hist_au10k.csv python /tmp/fqhist_target.py > gpurun_out/r2_ncu_fqhist_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fq_hist_kernel -s 2 -c 1 \
    -o gpurun_out/r2_fqhist10k -f python /tmp/fqhist_target.py > gpurun_out/r2_ncu_fqhist_run.log 2>&1
ncu -i gpurun_out/r2_fqhist10k.ncu-rep --page raw --csv > gpurun_out/r2_fqhist10k_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fqhist10k.ncu-rep --page source --csv > gpurun_out/r2_fqhist10k_source.csv 2>/dev/null
tail -2 gpurun_out/r2_ncu_fqhist_run.log
tail -8 gpurun_out/r2_launches_fq_hist_au10k.csv | cut -d, -f5,15 
