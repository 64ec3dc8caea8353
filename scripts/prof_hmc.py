"""ncu target: a few fused energy+force evaluations at Au561 (configs[1])."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, structures
scat = ElasticScatter()
ideal = structures.icosahedron('Au', 5)
target = scat.get_pdf(ideal)
atoms = structures.icosahedron('Au', 5)
atoms.positions *= 1.05
scat._ensure_wrapped(atoms)
be = scat._load(atoms, scat.pdf_qbin, 'PDF')
be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), 0.0)
pos = atoms.get_positions()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(3):
    be.energy_forces(pos, target, 'rw', 100.)
t = time.perf_counter()
for i in range(n):
    be.energy_forces(pos + 1e-6 * i, target, 'rw', 100.)
print('us per evaluation', (time.perf_counter() - t) / n * 1e6, be.sizes())
