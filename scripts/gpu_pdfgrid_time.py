"""Full gradient on the 330-bin PDF grid (get_grad_pdf's pair sum) at 10k atoms:
one 11-warp block per item against two 6-warp blocks."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyiid_b200 import ElasticScatter, structures
atoms = structures.fcc_sphere('Au', 10000)
res = {}
for prec in ('fp32',):
    scat = ElasticScatter(precision=prec)
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.pdf_qbin, 'PDF')
    be.set_timing(True)
    pos = atoms.get_positions()
    for nw in (8, 12, 8):
        be.set_option('grad_nw_max', nw)
        ts = []
        for it in range(4):
            g = be.grad_fq(pos)
            ts.append(round(be.last_kernel_ms()[0], 3))
        res[nw] = g
        print(prec, 'grad on the PDF grid (330 bins), grad_nw_max', nw, 'kernel ms', ts, flush=True)
    print('max |diff| / max', float(np.abs(res[8] - res[12]).max() / np.abs(res[8]).max()))
