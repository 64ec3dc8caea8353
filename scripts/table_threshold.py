"""Fused energy+forces wall time with the direct and the tabulated force pass
over structure size (picks force_table_min_n)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, structures
scat = ElasticScatter()
for n in (150, 309, 561, 800, 1000, 1250, 1500, 2000):
    atoms = structures.fcc_sphere('Au', n)
    target = scat.get_pdf(structures.fcc_sphere('Au', n, sigma=0.0))
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.pdf_qbin, 'PDF')
    be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), 0.0)
    pos = atoms.get_positions()
    out = []
    for min_n in (10 ** 9, 2):
        be.set_option('force_table_min_n', min_n)
        for _ in range(6):
            be.energy_forces(pos, target, 'rw', 100.)
        t = time.perf_counter()
        K = 500
        for i in range(K):
            be.energy_forces(pos + 1e-6 * i, target, 'rw', 100.)
        out.append((time.perf_counter() - t) / K * 1e6)
    be.set_option('force_table_min_n', 600)
    print('n=%5d  direct %.1f us   table %.1f us' % (n, out[0], out[1]))
