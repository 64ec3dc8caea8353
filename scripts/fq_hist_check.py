"""F(Q) through the radial pair histogram against the direct pass and the FP64
handle, and their kernel times (GPU box)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, structures


def nerr(a, b):
    return np.abs(np.asarray(a, float) - np.asarray(b, float)).max() / np.abs(b).max()


cases = [('Au2000', structures.fcc_sphere('Au', 2000)),
         ('AuPt3000', structures.alloy_sphere(3000, seed=3)),
         ('Au10000', structures.fcc_sphere('Au', 10000))]
if len(sys.argv) > 1:
    cases.append(('AuPt%s' % sys.argv[1], structures.alloy_sphere(int(sys.argv[1]), seed=5)))
for name, atoms in cases:
    pos = atoms.get_positions()
    s64 = ElasticScatter(precision='fp64', device=0)
    s64._ensure_wrapped(atoms)
    f64 = s64._load(atoms, s64.exp['qbin'], 'fq').fq(pos)
    s32 = ElasticScatter(precision='fp32', device=0)
    s32._ensure_wrapped(atoms)
    for kind, qb in (('fq', s32.exp['qbin']), ('PDF', s32.pdf_qbin)):
        be = s32._load(atoms, qb, kind)
        ref = f64 if kind == 'fq' else None
        out = {}
        for hist in (0, 1):
            be.set_option('fq_hist', hist)
            be.set_option('fq_hist_min_n', 2)
            f = be.fq(pos)
            f2 = be.fq(pos)
            be.set_timing(True)
            be.fq(pos)
            ms = be.last_kernel_ms()[0]
            be.set_timing(False)
            t = time.perf_counter()
            for _ in range(5):
                be.fq(pos)
            wall = (time.perf_counter() - t) / 5 * 1e3
            out[hist] = f
            print('%-9s %-3s grid hist=%d: kernel %.3f ms, call %.3f ms, reproducible %s%s' % (
                name, kind, hist, ms, wall, np.array_equal(f, f2),
                ', vs fp64 %.2e' % nerr(f, ref) if ref is not None else ''))
        print('%-9s %-3s grid: histogram vs direct %.2e' % (name, kind, nerr(out[1], out[0])))
