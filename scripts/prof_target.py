"""Tiny target for ncu: one warm-up and one timed call of the chosen pair sum."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyiid_b200 import ElasticScatter, structures
what = sys.argv[1] if len(sys.argv) > 1 else 'grad'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
prec = sys.argv[3] if len(sys.argv) > 3 else 'fp32'
atoms = structures.fcc_sphere('Au', n)
scat = ElasticScatter(precision=prec)
scat._ensure_wrapped(atoms)
if what == 'force':
    be = scat._load(atoms, scat.pdf_qbin, 'PDF')
    be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), 0.0)
    target = be.pdf(structures.fcc_sphere('Au', n, sigma=0.0).get_positions())
    for _ in range(2):
        be.energy_forces(atoms.get_positions(), target, 'rw', 100.)
else:
    be = scat._load(atoms, scat.exp['qbin'], 'fq')
    for _ in range(2):
        (be.fq if what == 'fq' else be.grad_fq)(atoms.get_positions())
print('done', what, n, prec)
