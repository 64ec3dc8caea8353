"""configs[4]: 100k-atom Au/Pt alloy, F(Q) -> G(r) + Rw (single GPU here)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, structures
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
atoms = structures.alloy_sphere(n)
ideal = structures.alloy_sphere(n, sigma=0.0)
res = {}
for prec in ('fp32', 'fp64'):
    scat = ElasticScatter(precision=prec)
    t = time.perf_counter(); target = scat.get_pdf(ideal); t1 = time.perf_counter() - t
    scat._ensure_wrapped(atoms)
    be = scat.pdf_backend
    be.set_timing(True)
    t = time.perf_counter()
    e, scale, f, pdf = be.energy_forces(atoms.get_positions(), target, 'rw', 1.0, want_forces=(prec == 'fp32'), want_pdf=True)
    t2 = time.perf_counter() - t
    res[prec] = (e, scale, pdf)
    print(prec, 'first pdf s %.3f' % t1, 'energy(+forces) s %.3f' % t2, 'Rw', e, 'scale', scale, 'last kernel ms', be.last_kernel_ms()[0], flush=True)
    if f is not None:
        print('  force max', np.abs(f).max(), 'sum', np.abs(f.sum(0)).max())
d = np.abs(res['fp32'][2] - res['fp64'][2]).max() / np.abs(res['fp64'][2]).max()
print('G(r) fp32 vs fp64 normalised max err', d, 'Rw diff', abs(res['fp32'][0] - res['fp64'][0]))
