"""Isotropic atomic displacement parameters (Debye-Waller factors).

The reference holds two kernels nothing calls, get_adp_fq
(kernels/cpu_nxn.py:114-121: fq = norm * omega * tau) and get_adp_grad_fq
(kernels/cpu_flat.py:156-174: grad = norm * (tau * grad_omega + omega *
grad_tau)), and builds no tau (its wrappers pass adps = None,
cpu_wrappers/flat_multi_cpu_wrap.py:18-19).  tests/golden/adp_aupt24.npz comes
from THOSE kernels with tau = exp(-(u_i^2 + u_j^2) Q^2 / 2), grad_tau = 0
(tests/golden/make_golden_adp.py).  CPU tests: the oracle's restatement against
that fixture and the host-side table logic; GPU tests: the CUDA path (one
form-factor row f t per (element, displacement) class, normaliser from f) against
fixture and oracle, 1e-5 in FP32 mode, 1e-10 in FP64 mode."""
import numpy as np
import pytest

import oracle
from conftest import golden, nerr, TOL32, TOL64
from pyiid_b200 import ElasticScatter, Calc1D, ase_shim, backend

EXP = oracle.DEFAULT_EXP


# ---- CPU: oracle and host logic -------------------------------------------------
@pytest.mark.parametrize('prec,tag,tol', [('fp32', 'f32', 1e-6), ('fp64', 'f64', 1e-13)])
def test_oracle_restates_the_reference_adp_kernels(prec, tag, tol):
    g = golden('adp_aupt24')
    qb = float(g['qbin'])
    fq = oracle.wrap_adp_fq(g['positions'], g['scatter'], g['adps'], qb, prec)
    grad = oracle.wrap_adp_grad_fq(g['positions'], g['scatter'], g['adps'], qb, prec)
    assert nerr(fq, g['fq_' + tag]) < tol and nerr(grad, g['grad_' + tag]) < tol
    # and the displacements matter at this size: 30 % of F(Q)
    assert nerr(oracle.wrap_fq(g['positions'], g['scatter'], qb, prec), g['fq_' + tag]) > 0.1
    # zero displacements: tau = 1, the plain kernels
    z = np.zeros(len(g['adps']))
    assert np.array_equal(oracle.wrap_adp_fq(g['positions'], g['scatter'], z, qb, prec),
                          oracle.wrap_fq(g['positions'], g['scatter'], qb, prec))


def test_adp_tables_factorise_tau():
    """tau_ij = t_i t_j: the rows the pair sums get reproduce the reference-shaped
    tau array pair by pair; the normaliser rows stay the plain form factors."""
    g = golden('adp_aupt24')
    table, idx = backend.element_table(g['scatter'], g['numbers'])
    pair, norm, cls = backend.adp_tables(table, idx, g['adps'], float(g['qbin']))
    assert len(pair) == 4 and pair.shape == norm.shape  # (Z, u2) classes of the fixture
    assert np.array_equal(norm[cls], np.asarray(g['scatter'], np.float64))
    i, j = oracle.pair_indices(len(cls))
    tau = oracle.adp_tau(g['adps'], table.shape[1], float(g['qbin']), 'fp64')
    f = np.asarray(g['scatter'], np.float64)
    assert np.allclose(pair[cls][i] * pair[cls][j], f[i] * f[j] * tau, rtol=1e-12, atol=0)
    with pytest.raises(ValueError):
        # more classes than the tiling by class is meant for
        backend.adp_tables(table, np.zeros(40, np.int32), np.linspace(0, 0.01, 40),
                           float(g['qbin']))


# ---- GPU ------------------------------------------------------------------------
def _atoms(g):
    a = ase_shim.Atoms(numbers=g['numbers'], positions=g['positions'])
    a.set_array('adps', g['adps'])
    return a


@pytest.mark.gpu
@pytest.mark.parametrize('prec,tag,tol', [('fp32', 'f32', TOL32), ('fp64', 'f64', TOL64)])
def test_fq_and_gradient_with_adps_match_the_reference_kernels(prec, tag, tol):
    g = golden('adp_aupt24')
    scat = ElasticScatter(precision=prec)
    atoms = _atoms(g)
    fq, grad = scat.get_fq(atoms), scat.get_grad_fq(atoms)
    # (same form factors on both sides: the fixture stores the package's table)
    assert np.array_equal(atoms.get_array('F(Q) scatter'), g['scatter'])
    tol_g = tol if prec == 'fp64' else 5e-5  # float32 reference: its own noise floor
    assert nerr(fq, g['fq_' + tag]) < tol and nerr(grad, g['grad_' + tag]) < tol_g
    if prec == 'fp32':  # the FP32 mode against the float64 reference arithmetic
        assert nerr(grad, g['grad_f64']) < TOL32
    # G(r) on the PDF grid against the oracle's re-drive
    qb = float(oracle.pdf_qbin(EXP))
    sp = atoms.get_array('PDF scatter')
    ofq = oracle.wrap_adp_fq(g['positions'], sp, g['adps'], qb, 'fp64')
    opdf = oracle.get_pdf_at_qmin(np.array(ofq, dtype=np.float64), EXP['rstep'], qb,
                                  oracle.r_grid(EXP), EXP['qmin'])
    assert nerr(scat.get_pdf(atoms), opdf) < max(tol, 1e-9)
    # a change of the displacements is seen; zeros are the plain structure, bit for bit
    plain = ase_shim.Atoms(numbers=g['numbers'], positions=g['positions'])
    fq0 = scat.get_fq(plain)
    assert nerr(fq0, fq) > 0.1
    atoms.set_array('adps', None)
    atoms.set_array('adps', np.zeros(len(atoms)))
    assert np.array_equal(scat.get_fq(atoms), fq0)
    atoms.set_array('adps', None)
    atoms.set_array('adps', g['adps'])
    assert np.array_equal(scat.get_fq(atoms), fq)
    atoms.set_array('adps', None)
    atoms.set_array('adps', np.zeros((len(atoms), 3)))
    with pytest.raises(ValueError):
        scat.get_fq(atoms)


@pytest.mark.gpu
def test_forces_with_adps_are_minus_half_the_energy_gradient():
    """Calc1D on a structure with displacements: tau does not depend on the
    positions, so the forces stay -1/2 dE/dq (the reference's convention,
    SURVEY.md 8a note 1) -- central differences in FP64 mode."""
    g = golden('adp_aupt24')
    scat = ElasticScatter(precision='fp64')
    ideal = _atoms(g)
    target = scat.get_pdf(ideal)
    atoms = _atoms(g)
    atoms.positions = atoms.positions * 1.02
    atoms.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf,
                                exp_grad_function=scat.get_grad_pdf, conv=1., potential='rw'))
    f = atoms.get_forces()
    h = 1e-5
    for (a, w) in ((0, 0), (5, 1), (11, 2)):
        e = []
        for s in (1, -1):
            b = atoms.copy()
            b.positions[a, w] += s * h
            b.set_calculator(atoms.calc)
            e.append(b.get_potential_energy())
        fd = (e[0] - e[1]) / (2 * h)
        assert abs(f[a, w] + 0.5 * fd) < 1e-6 * np.abs(f).max(), (a, w, f[a, w], fd)
    # and the fused FP32 evaluation agrees with the FP64 mode
    s32 = ElasticScatter(precision='fp32')
    e32, _, f32 = s32.get_pdf_energy_forces(atoms, target, 'rw', 1.)
    assert abs(e32 - atoms.get_potential_energy()) < TOL32 * abs(e32) and nerr(f32, f) < TOL32
