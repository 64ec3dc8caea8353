"""End-to-end time of grad_fq (host buffers) at the bench workload, for the
download tunables IID_DL_THREADS / IID_DL_CHUNK_MB (read at library load)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
atoms, scat = bench.build_workload(50000)
be = scat._load(atoms, scat.exp['qbin'], 'fq')
pos = atoms.get_positions()
for _ in range(2):
    be.grad_fq(pos, with_fq=True)
ts = []
for _ in range(5):
    t = time.perf_counter(); g, f = be.grad_fq(pos, with_fq=True); ts.append(time.perf_counter() - t)
print(os.environ.get('IID_DL_THREADS'), os.environ.get('IID_DL_CHUNK_MB'), 'e2e ms', [round(1e3 * x, 1) for x in ts], flush=True)
