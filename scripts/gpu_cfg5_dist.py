"""configs[4] under torchrun: 100k-atom Au/Pt alloy, F(Q) -> G(r) + Rw, pair-tile
sharding over the ranks + NCCL all-reduce of the F(Q) partial sums.

    python -m torch.distributed.run --nproc-per-node 8 scripts/gpu_cfg5_dist.py
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank = int(os.environ.get('RANK', 0)); local = int(os.environ.get('LOCAL_RANK', 0))
world = int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from pyiid_b200 import ElasticScatter, structures
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
atoms = structures.alloy_sphere(n)
ideal = structures.alloy_sphere(n, sigma=0.0)
scat = ElasticScatter()
target = scat.get_pdf(ideal)
scat._ensure_wrapped(atoms)
be = scat.pdf_backend
pos = atoms.get_positions()
for _ in range(2):
    e, scale, _, _ = be.energy_forces(pos, target, 'rw', 1.0, want_forces=False)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
K = 5
t = time.perf_counter()
for _ in range(K):
    e, scale, _, _ = be.energy_forces(pos, target, 'rw', 1.0, want_forces=False)
dt = (time.perf_counter() - t) / K
if world > 1:
    tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
pairq = n * (n - 1) // 2 * be.nq
if rank == 0:
    print(json.dumps({'workload': 'Au/Pt %d atoms, F(Q)->G(r)+Rw, 330-bin PDF grid, host positions in' % n,
                      'n_gpus': world, 'ms_per_evaluation': dt * 1e3, 'pairq_per_s': pairq / dt,
                      'rw': e, 'scale': scale}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
