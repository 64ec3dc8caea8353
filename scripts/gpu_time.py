"""Developer timing run on the GPU box (kernel-only, CUDA events)."""
import os, sys, time, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyiid_b200 import ElasticScatter, structures

def smi():
    try:
        return subprocess.check_output(['nvidia-smi', '--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active', '--format=csv,noheader'], text=True).strip()
    except Exception as e:
        return str(e)

cases = [(10000, 'fp32'), (10000, 'fp64')]
if len(sys.argv) > 1 and sys.argv[1] == 'big':
    cases.append((50000, 'fp32'))
for n, prec in cases:
    atoms = structures.fcc_sphere('Au' if n == 10000 else 'Pt', n)
    scat = ElasticScatter(precision=prec)
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.exp['qbin'], 'fq')
    be.set_timing(True)
    pos = atoms.get_positions()
    for what in ('fq', 'grad'):
        ts = []
        for it in range(4):
            if what == 'fq': be.fq(pos)
            else: be.grad_fq(pos)
            ms, pq = be.last_kernel_ms()
            ts.append(round(ms, 3))
        print(n, prec, what, 'kernel ms', ts, 'pairQ/s %.4g' % (pq / (min(ts) * 1e-3)), '|', smi(), flush=True)
    be2 = scat._load(atoms, scat.pdf_qbin, 'PDF')
    be2.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), 0.0)
    be2.set_timing(True)
    target = be2.pdf(structures.fcc_sphere('Au', n, sigma=0.0).get_positions())
    for it in range(3):
        t = time.time(); e, s, f, _ = be2.energy_forces(pos, target, 'rw', 100.); wall = time.time() - t
        ms, pq = be2.last_kernel_ms()
        print(n, prec, 'energy_forces wall ms %.3f' % (wall * 1e3), 'force kernel ms %.3f' % ms, 'E', e, flush=True)
