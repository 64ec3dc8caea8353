"""Device / wall time of the spring restraint kernels and of the fused
Calc1D + Spring evaluation (run on the GPU box)."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyiid_b200 import ElasticScatter, Calc1D, Spring, MultiCalc, structures
from pyiid_b200.backend import Backend
from oracle import spring as osp

for prec in ('fp32', 'fp64'):
    be = Backend.get(prec, None, 'fq')
    dev = 'cuda:%d' % be.device
    for n in (561, 10000, 50000):
        pos = structures.fcc_sphere('Au', n).get_positions()
        with be._on_stream():
            p = torch.from_numpy(pos).to(dev)
            e = torch.zeros(1, dtype=torch.float64, device=dev)
            f = torch.zeros((n, 3), dtype=torch.float64, device=dev)
            for t, k, rt in (('rep', 10., 3.0), ('att', 1e-3, 30.)):
                ty = {'rep': 0, 'att': 2}[t]
                fn = lambda: be.lib.iid_spring_partial(be.h, p.data_ptr(), n, ty, k, rt, None,
                                                       e.data_ptr(), f.data_ptr(), None, None)
                for _ in range(3): assert fn() == 0
                K = 20
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                for _ in range(K): fn()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / K
                print('%s %s n=%6d  %.4f ms  %.3e ordered pairs/s' % (prec, t, n, ms, n * (n - 1.) / ms * 1e3))
# CPU restatement of the reference (N x N numpy arrays) for scale
pos = structures.fcc_sphere('Au', 2000).get_positions()
t0 = time.perf_counter(); osp.pair_energy(pos, 10., 3.0, 'rep'); osp.pair_force(pos, 10., 3.0, 'rep')
dt = time.perf_counter() - t0
print('numpy restatement n=2000 energy+force %.3f s  %.3e ordered pairs/s' % (dt, 2000 * 1999. / dt))

# fused Calc1D + Spring at Au561
scat = ElasticScatter()
ideal = structures.icosahedron('Au', 5)
target = scat.get_pdf(ideal)
atoms = structures.icosahedron('Au', 5)
atoms.positions *= 1.05
scat._ensure_wrapped(atoms)
be = scat._load(atoms, scat.pdf_qbin, 'PDF')
be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), 0.0)
pos = atoms.get_positions()
for springs in ([], [('rep', 200., 2.5)], [('rep', 200., 2.5), ('att', .01, 20.)]):
    be.set_restraints(springs)
    for _ in range(5): be.energy_forces(pos, target, 'rw', 100.)
    t0 = time.perf_counter()
    for i in range(2000): be.energy_forces(pos + 1e-6 * i, target, 'rw', 100.)
    print('Au561 fused evaluation with %d springs: %.1f us' % (len(springs), (time.perf_counter() - t0) / 2000 * 1e6))
sp = Spring(k=200., rt=2.5, sp_type='rep')
t0 = time.perf_counter()
for i in range(500): Backend.get('fp32', None, 'fq').spring(pos, 'rep', 200., 2.5, None, True, True)
print('Au561 stand-alone spring call (host in/out): %.1f us' % ((time.perf_counter() - t0) / 500 * 1e6))
