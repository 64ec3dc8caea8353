"""Developer check run on the GPU box: parity vs the oracle at several sizes
and kernel timings.  Not part of the test-suite."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from pyiid_b200 import ElasticScatter, Calc1D, structures
from pyiid_b200.backend import Backend

def nerr(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    return float(np.abs(a - b).max() / np.abs(b).max())

out = {}
exp = dict(oracle.DEFAULT_EXP)
for name, atoms in [('Au10', structures.random_atoms(10, 1)),
                    ('Au100', structures.random_atoms(100, 2)),
                    ('Au561', structures.icosahedron('Au', 5)),
                    ('AuPt300', structures.alloy_sphere(300)),
                    ('Au1000', structures.fcc_sphere('Au', 1000))]:
    for prec in (('fp32',) if 'quick' in sys.argv else ('fp32', 'fp64')):
        scat = ElasticScatter(precision=prec)
        t = time.time()
        fq = scat.get_fq(atoms); g = scat.get_grad_fq(atoms); pdf = scat.get_pdf(atoms)
        dt = time.time() - t
        pos = atoms.get_positions()
        sf = atoms.get_array('F(Q) scatter'); sp = atoms.get_array('PDF scatter')
        rec = {'t_first': dt}
        for oprec in ('fp32', 'fp64'):
            if len(atoms) > 600 and oprec == 'fp32' and prec == 'fp64':
                continue
            ofq = oracle.experiment_fq(pos if oprec == 'fp64' and prec == 'fp64' else pos.astype(np.float32), sf, exp, oprec)
            og = oracle.experiment_grad_fq(pos if oprec == 'fp64' and prec == 'fp64' else pos.astype(np.float32), sf, exp, oprec)
            opdf = oracle.experiment_pdf(pos if oprec == 'fp64' and prec == 'fp64' else pos.astype(np.float32), sp, exp, oprec)
            rec['vs_' + oprec] = {'fq': nerr(fq, ofq), 'grad': nerr(g, og), 'pdf': nerr(pdf, opdf)}
        out[name + '_' + prec] = rec
        print(name, prec, json.dumps(rec), flush=True)

# energy/forces
a1, a2 = structures.atomic_square()
for prec in ('fp32', 'fp64'):
    scat = ElasticScatter(precision=prec)
    for pot in ('rw', 'chi_sq'):
        target = scat.get_pdf(a1)
        calc = Calc1D(target_data=target, exp_function=scat.get_pdf, exp_grad_function=scat.get_grad_pdf, potential=pot, conv=1.)
        a2c = a2.copy(); a2c.set_calculator(calc); scat._ensure_wrapped(a2c)
        e = a2c.get_potential_energy(); f = a2c.get_forces()
        oe, of, _ = oracle.calc1d_energy_forces(a2.get_positions(), a2c.get_array('PDF scatter'), exp, target, pot, 1., 'fp64')
        # generic (unfused) route through get_grad_pdf
        gp = scat.get_grad_pdf(a2c)
        from pyiid_b200.calc import wrap_grad_rw, wrap_grad_chi_sq
        f2 = (wrap_grad_rw if pot == 'rw' else wrap_grad_chi_sq)(gp, scat.get_pdf(a2c), target)
        print('square', prec, pot, 'E', e, oe, 'force err', nerr(f, of), 'generic err', nerr(f2, of), flush=True)

# timings
if 'quick' in sys.argv:
    sys.exit(0)
for n, prec in [(10000, 'fp32'), (10000, 'fp64'), (50000, 'fp32')]:
    atoms = structures.fcc_sphere('Au' if n == 10000 else 'Pt', n)
    scat = ElasticScatter(precision=prec)
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.exp['qbin'], 'fq')
    be.set_timing(True)
    pos = atoms.get_positions()
    res = {}
    for what in ('fq', 'grad'):
        ts = []
        for it in range(3):
            t = time.time()
            if what == 'fq': be.fq(pos)
            else: be.grad_fq(pos)
            wall = time.time() - t
            ms, pq = be.last_kernel_ms()
            ts.append((ms, wall))
        res[what] = ts
        print(n, prec, what, 'kernel ms, wall s:', ts, 'pairQ/s', pq / (ts[-1][0] * 1e-3), flush=True)
    if n == 10000:
        be2 = scat._load(atoms, scat.pdf_qbin, 'PDF')
        be2.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), 0.0)
        target = be2.pdf(structures.fcc_sphere('Au', n, sigma=0.0).get_positions())
        for it in range(3):
            t = time.time(); e, s, f, _ = be2.energy_forces(pos, target, 'rw', 100.); wall = time.time() - t
            ms, pq = be2.last_kernel_ms()
            print(n, prec, 'energy_forces wall', wall, 'force kernel ms', ms, 'E', e, flush=True)
