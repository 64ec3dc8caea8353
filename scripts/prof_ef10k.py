"""Energy + forces at Au 10k (GPU box) for an ncu launch list: the kernels of one
evaluation and their times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, structures

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
atoms = structures.fcc_sphere('Au', n)
scat = ElasticScatter(precision='fp32', device=0)
scat._ensure_wrapped(atoms)
be = scat._load(atoms, scat.pdf_qbin, 'PDF')
be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), scat.exp['qmin'])
pos = atoms.get_positions()
target = be.pdf(structures.fcc_sphere('Au', n, sigma=0.0).get_positions())
be.set_option('graph', int(os.environ.get('EF_GRAPH', '1')))
for _ in range(4):
    be.energy_forces(pos, target, 'rw', 100.)
t = time.perf_counter()
reps = 20
for i in range(reps):
    be.energy_forces(pos + 1e-6 * i, target, 'rw', 100.)
print('%d atoms: %.3f ms per evaluation' % (n, (time.perf_counter() - t) / reps * 1e3))
