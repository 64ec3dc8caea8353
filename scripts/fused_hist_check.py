"""The fused launch with its F(Q) phase through the pair histogram against the
direct phase and the FP64 handle; time per evaluation (GPU box)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, structures


def nerr(a, b):
    return np.abs(np.asarray(a, float) - np.asarray(b, float)).max() / np.abs(b).max()


for name, atoms, ideal in (
        ('Au55', structures.icosahedron('Au', 2), structures.icosahedron('Au', 2)),
        ('Au561', structures.icosahedron('Au', 5), structures.icosahedron('Au', 5)),
        ('AuPt400', structures.alloy_sphere(400, seed=3), structures.alloy_sphere(400, seed=3)),
        ('Au923', structures.fcc_sphere('Au', 923), structures.fcc_sphere('Au', 923, sigma=0.0))):
    atoms.positions = atoms.positions * 1.03 + np.random.RandomState(1).normal(0, 0.03, atoms.positions.shape)
    res = {}
    for prec in ('fp64', 'fp32'):
        scat = ElasticScatter(precision=prec, device=0)
        target = scat.get_pdf(ideal)
        scat._ensure_wrapped(atoms)
        be = scat._load(atoms, scat.pdf_qbin, 'PDF')
        be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), scat.exp['qmin'])
        pos = atoms.get_positions()
        if prec == 'fp64':
            res = be.energy_forces(pos, target, 'rw', 1.)
            continue
        for hist in (0, 1):
            be.set_option('fused_hist', hist)
            for _ in range(5):
                e, sc, f = be.energy_forces(pos, target, 'rw', 1.)[:3]
            e2, sc2, f2 = be.energy_forces(pos, target, 'rw', 1.)[:3]
            t = time.perf_counter()
            for i in range(200):
                be.energy_forces(pos, target, 'rw', 1.)
            us = (time.perf_counter() - t) / 200 * 1e6
            print('%-8s hist=%d  %.1f us  energy %.2e  scale %.2e  forces %.2e  reproducible %s' % (
                name, hist, us, abs(e - res[0]) / abs(res[0]), abs(sc - res[1]) / abs(res[1]),
                nerr(f, res[2]), e == e2 and np.array_equal(f, f2)))
