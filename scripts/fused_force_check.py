"""Forces of the fused evaluation (table and direct force pass) against the
FP64 handle, Au55 / Au561 / a two-element sphere (GPU box)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, structures


def nerr(a, b):
    return np.abs(np.asarray(a, float) - np.asarray(b, float)).max() / np.abs(b).max()


for name, atoms, ideal in (
        ('Au55', structures.icosahedron('Au', 2), structures.icosahedron('Au', 2)),
        ('Au561', structures.icosahedron('Au', 5), structures.icosahedron('Au', 5)),
        ('AuPt400', structures.alloy_sphere(400, seed=3), structures.alloy_sphere(400, seed=3))):
    atoms.positions = atoms.positions * 1.03 + np.random.RandomState(1).normal(0, 0.03, atoms.positions.shape)
    res = {}
    for prec in ('fp64', 'fp32'):
        scat = ElasticScatter(precision=prec, device=0)
        target = scat.get_pdf(ideal)
        scat._ensure_wrapped(atoms)
        be = scat._load(atoms, scat.pdf_qbin, 'PDF')
        be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), scat.exp['qmin'])
        pos = atoms.get_positions()
        for pot in ('rw', 'chi_sq'):
            if prec == 'fp64':
                res[pot] = be.energy_forces(pos, target, pot, 1.)
                continue
            for table in (0, 1):
                be.set_option('fused_table', table)
                e, sc, f = be.energy_forces(pos, target, pot, 1.)[:3]
                e0, s0, f0 = res[pot][:3]
                print('%-8s %-6s table=%d  energy %.2e  scale %.2e  forces %.2e' % (
                    name, pot, table, abs(e - e0) / abs(e0), abs(sc - s0) / abs(s0), nerr(f, f0)))
