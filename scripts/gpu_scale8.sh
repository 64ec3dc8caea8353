#!/bin/bash
# 8-GPU bench line (one process per GPU, NCCL), as the driver launches it.
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
python -c "import json;d=json.load(open('gpurun_out/bench_n8.json'));print(d['n_gpus'],d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['roofline']['frac'],d['check'])"; tail -2 gpurun_out/bench_n8.err
