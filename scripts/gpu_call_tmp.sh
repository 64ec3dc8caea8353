mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python scripts/gpu_time.py 2>&1 | grep fp64
