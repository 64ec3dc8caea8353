"""Energy + forces per evaluation over structure sizes (GPU box): fused launch on
/ off, forces against each other."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, structures


def nerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


for n in (55, 147, 309, 561, 700, 923, 1100, 1415):
    atoms = structures.fcc_sphere('Au', n)
    ideal = atoms.copy()
    atoms.positions = atoms.positions * 1.03
    scat = ElasticScatter(precision='fp32', device=0)
    target = scat.get_pdf(ideal)
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.pdf_qbin, 'PDF')
    be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), scat.exp['qmin'])
    pos = atoms.get_positions()
    out = {}
    for fused in (0, 1):
        be.set_option('fused', fused)
        n0 = None
        for _ in range(5):
            r = be.energy_forces(pos, target, 'rw', 100.)
        n0 = be.launch_count()
        t = time.perf_counter()
        reps = 200
        for i in range(reps):
            r = be.energy_forces(pos, target, 'rw', 100.)
        dt = (time.perf_counter() - t) / reps * 1e6
        out[fused] = (dt, (be.launch_count() - n0) / reps, r)
    be.set_option('fused', 1)
    print('n = %4d: launch sequence %6.1f us (%d launches), fused %6.1f us (%d launches); '
          'energy %.1e forces %.1e' % (len(atoms), out[0][0], out[0][1], out[1][0], out[1][1],
                                       abs(out[1][2][0] - out[0][2][0]) / abs(out[0][2][0]),
                                       nerr(out[1][2][2], out[0][2][2])))
