import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from pyiid_b200 import ElasticScatter, structures
ideal = structures.icosahedron('Au', 5)
scat = ElasticScatter(precision='fp32')
target = scat.get_pdf(ideal)
atoms = structures.icosahedron('Au', 5); atoms.positions *= 1.05
scat._ensure_wrapped(atoms)
be = scat._load(atoms, scat.pdf_qbin, 'PDF')
be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), 0.0)
be.set_option('graph', 0)
pos = atoms.get_positions()
for i in range(401):
    be.energy_forces(pos, target, 'rw', 100.)
print(be.sizes())
