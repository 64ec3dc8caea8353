"""Generate the golden fixtures under tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference):

    PYTHONPATH=. python tests/golden/make_golden.py

Every stored output comes from the reference's own numba kernels
(pyiid/experiments/elasticscatter/kernels/cpu_flat.py, cpu_experimental.py)
and its float64 host stage (kernels/master_kernel.py), loaded through
oracle/ref_shim.py and driven like cpu_wrappers/flat_serial_cpu_wrap.py:13-133
with the float64 normaliser (flat_multi_cpu_wrap.py:55).  The "f64" entries
come from the same source with f4 -> f8.  Inputs (positions, scatter-factor
arrays) are stored next to the outputs, so the tests do not depend on this
package's structure builders or form-factor table.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
import oracle  # noqa: E402
from pyiid_b200 import structures, formfactors  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
EXP = dict(oracle.DEFAULT_EXP)
QB = EXP['qbin']
PQB = float(oracle.pdf_qbin(EXP))
NQ_F = int(np.floor(EXP['qmax'] / QB))
NQ_P = int(np.floor(EXP['qmax'] / PQB))
RGRID = oracle.r_grid(EXP)


def scatter_arrays(numbers):
    out = []
    for qbin, nq in ((QB, NQ_F), (PQB, NQ_P)):
        arr = np.zeros((len(numbers), nq), np.float32)
        formfactors.get_scatter_array(arr, numbers, qbin)
        out.append(arr)
    return out


def ref_pdf(pos, scat_p, prec):
    mk = ref_shim.master()
    fq = ref_shim.ref_fq(pos, scat_p, PQB, prec)
    return mk.get_pdf_at_qmin(np.array(fq, dtype=np.float64), EXP['rstep'],
                              PQB, RGRID, EXP['qmin'])


def ref_grad_pdf(pos, scat_p, prec):
    mk = ref_shim.master()
    g = np.array(ref_shim.ref_grad_fq(pos, scat_p, PQB, prec))
    out = np.zeros((g.shape[0], 3, len(RGRID)))
    for i in range(g.shape[0]):
        for w in range(3):
            out[i, w] = mk.get_pdf_at_qmin(np.array(g[i, w], dtype=np.float64),
                                           EXP['rstep'], PQB, RGRID, EXP['qmin'])
    return out


def ref_energy_forces(gpdf, gcalc, target):
    mk = ref_shim.master()
    res = {}
    rw, scale = mk.get_rw(target, gcalc, weight=None)
    f = np.zeros((gpdf.shape[0], 3))
    mk.get_grad_rw(f, gpdf, gcalc, target, rw, scale)
    res['rw'] = (rw, scale, f)
    chi, scale = mk.get_chi_sq(target, gcalc)
    f = np.zeros((gpdf.shape[0], 3))
    mk.get_grad_chi_sq(f, gpdf, gcalc, target, scale)
    res['chi_sq'] = (chi, scale, f)
    return res


def case(name, numbers, pos, target_pos, store_grad_pdf=False):
    numbers = np.asarray(numbers)
    sf, sp = scatter_arrays(numbers)
    d = {'numbers': numbers, 'positions': pos, 'target_positions': target_pos,
         'scatter_fq': sf, 'scatter_pdf': sp}
    for prec, tag in (('fp32', 'f32'), ('fp64', 'f64')):
        d['fq_' + tag] = np.asarray(ref_shim.ref_fq(pos, sf, QB, prec))
        d['grad_fq_' + tag] = np.asarray(ref_shim.ref_grad_fq(pos, sf, QB, prec))
        gcalc = ref_pdf(pos, sp, prec)
        target = ref_pdf(target_pos, sp, prec)
        d['pdf_' + tag] = gcalc
        d['target_pdf_' + tag] = target
        gp = ref_grad_pdf(pos, sp, prec)
        if store_grad_pdf:
            d['grad_pdf_' + tag] = gp
        for pot, (val, scale, f) in ref_energy_forces(gp, gcalc, target).items():
            d['%s_%s' % (pot, tag)] = np.array([val, scale], dtype=np.float64)
            d['%s_forces_%s' % (pot, tag)] = f
    d['fq_f32_asis_na'] = np.asarray(ref_shim.ref_fq(pos, sf, QB, 'fp32', True))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **d)
    print(name, 'N =', len(numbers), {k: np.asarray(v).shape for k, v in d.items()
                                      if k.endswith('f32')})


def main():
    if not ref_shim.available():
        raise SystemExit('reference mount not found')
    rs = np.random.RandomState(20161017)
    # Au4 square and its 0.75-scaled copy (reference tests/__init__.py:131-141)
    a1, a2 = structures.atomic_square()
    case('au4_square', a2.numbers, a2.get_positions(), a1.get_positions(), True)
    # random Au10 in a 10 A box (reference tests/__init__.py:83-95)
    a = structures.random_atoms(10, 1)
    b = structures.random_atoms(10, 2)
    case('au10_random', a.numbers, a.get_positions(), b.get_positions())
    # config 1: Au55 Mackay icosahedron, perturbed vs ideal
    ico = structures.icosahedron('Au', 2)
    pos = ico.get_positions() + rs.normal(0, 0.05, (55, 3))
    case('au55_ico', ico.numbers, pos, ico.get_positions())
    # two elements, ragged size (not a multiple of the 32-atom tile)
    al = structures.alloy_sphere(37, seed=3)
    pos = al.get_positions()
    case('aupt37_alloy', al.numbers, pos,
         structures.alloy_sphere(37, sigma=0.0, seed=3).get_positions())
    # known answers the reference's own tests hold
    mk = ref_shim.master()
    x = np.arange(0, 2 * np.pi, .1)
    c60 = np.loadtxt(os.path.join(ref_shim.REF_ROOT,
                                  'pyiid/tests/test_master/c60_scat.txt'),
                     dtype=np.float32)
    ffq = rs.normal(size=NQ_P)
    exp2 = dict(EXP, qmin=1.3, rmin=1.5, rmax=32.0, rstep=0.013)
    rg2 = oracle.r_grid(exp2)
    pq2 = float(oracle.pdf_qbin(exp2))
    nq2 = int(np.floor(exp2['qmax'] / pq2))
    ffq2 = rs.normal(size=nq2)
    np.savez_compressed(
        os.path.join(HERE, 'known_answers.npz'),
        x=x,
        rw_sin_cos=np.array(mk.get_rw(np.sin(x), np.cos(x)), dtype=float),
        rw_sin_sin=np.array(mk.get_rw(np.sin(x), np.sin(x)), dtype=float),
        chi_sin_cos=np.array(mk.get_chi_sq(np.sin(x), np.cos(x)), dtype=float),
        chi_sin_sin=np.array(mk.get_chi_sq(np.sin(x), np.sin(x)), dtype=float),
        c60_row=c60[0], c60_all_rows_equal=np.array(bool(np.all(c60 == c60[0]))),
        random_fq=ffq,
        random_fq_pdf=mk.get_pdf_at_qmin(ffq.copy(), EXP['rstep'], PQB, RGRID, 0.0),
        random_fq2=ffq2,
        random_fq2_pdf=mk.get_pdf_at_qmin(ffq2.copy(), exp2['rstep'], pq2, rg2,
                                          exp2['qmin']),
        exp2_vals=np.array([exp2['qmin'], exp2['rmin'], exp2['rmax'],
                            exp2['rstep'], pq2]),
        k_to_ij=np.array([ref_shim.kernels()[0].k_to_ij(k) for k in range(12)]),
    )
    print('known answers written')


if __name__ == '__main__':
    main()
