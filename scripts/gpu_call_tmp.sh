mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python scripts/hmc_breakdown.py > gpurun_out/hmc_breakdown.txt 2>&1; head -12 gpurun_out/hmc_breakdown.txt; grep "NUTS device" gpurun_out/hmc_breakdown.txt
python scripts/small_n_breakdown.py > gpurun_out/small_n.txt 2>&1; tail -15 gpurun_out/small_n.txt
