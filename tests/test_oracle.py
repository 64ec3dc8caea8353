"""CPU tests of the oracle itself: against the golden vectors generated from
the reference (tests/golden/make_golden.py) and against the known answers the
reference's own tests hold."""
import numpy as np
import pytest

import oracle
from conftest import golden, nerr

CASES = ['au4_square', 'au10_random', 'au55_ico', 'aupt37_alloy']
EXP = oracle.DEFAULT_EXP


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('prec,tag', [('fp32', 'f32'), ('fp64', 'f64')])
def test_debye_sums_match_reference_bit_for_bit(name, prec, tag):
    g = golden(name)
    fq = oracle.experiment_fq(g['positions'], g['scatter_fq'], EXP, prec)
    grad = oracle.experiment_grad_fq(g['positions'], g['scatter_fq'], EXP, prec)
    # same libm, same operation order: identical bits
    assert np.array_equal(fq, g['fq_' + tag])
    assert np.array_equal(grad, g['grad_fq_' + tag])


@pytest.mark.parametrize('name', CASES)
def test_threaded_oracle_equals_serial(name):
    g = golden(name)
    a = oracle.wrap_fq(g['positions'], g['scatter_fq'], .1, nthreads=4, chunk=7)
    assert nerr(a, g['fq_f32']) < 1e-7
    b = oracle.wrap_fq_grad(g['positions'], g['scatter_fq'], .1, nthreads=4, chunk=5)
    assert nerr(b, g['grad_fq_f32']) < 1e-6


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('prec,tag', [('fp32', 'f32'), ('fp64', 'f64')])
def test_pdf_and_potentials_match_reference(name, prec, tag):
    g = golden(name)
    pdf = oracle.experiment_pdf(g['positions'], g['scatter_pdf'], EXP, prec)
    assert nerr(pdf, g['pdf_' + tag]) < 1e-13
    target = g['target_pdf_' + tag]
    for pot in ('rw', 'chi_sq'):
        e, f, scale = oracle.calc1d_energy_forces(
            g['positions'], g['scatter_pdf'], EXP, target, pot, 1., prec)
        val, sc = g['%s_%s' % (pot, tag)]
        assert abs(e - val) <= 1e-12 * max(1., abs(val))
        assert abs(scale - sc) <= 1e-12 * max(1., abs(sc))
        assert nerr(f, g['%s_forces_%s' % (pot, tag)]) < 1e-11


def test_grad_pdf_matches_reference():
    g = golden('au4_square')
    gp = oracle.experiment_grad_pdf(g['positions'], g['scatter_pdf'], EXP, 'fp32')
    assert nerr(gp, g['grad_pdf_f32']) < 1e-13


def test_as_is_float32_normaliser_is_reproduced():
    g = golden('au55_ico')
    asis = oracle.wrap_fq(g['positions'], g['scatter_fq'], .1, 'fp32', na_mode=1)
    assert nerr(asis, g['fq_f32_asis_na']) < 1e-6


def test_reference_known_answers():
    """pyiid/tests/test_master/test_master_kernel.py:19-48."""
    k = golden('known_answers')
    x = np.arange(0, 2 * np.pi, .1)
    assert oracle.get_rw(np.sin(x), np.cos(x))[0] == 1
    assert abs(oracle.get_rw(np.sin(x), np.sin(x))[0]) < 1e-15
    assert abs(oracle.get_chi_sq(np.sin(x), np.cos(x))[0] - 63.01399) < 1e-5
    assert abs(oracle.get_chi_sq(np.sin(x), np.sin(x))[0]) < 1e-15
    assert np.allclose(k['rw_sin_cos'], oracle.get_rw(np.sin(x), np.cos(x)))
    assert np.allclose(k['chi_sin_cos'], oracle.get_chi_sq(np.sin(x), np.cos(x)))


def test_k_to_ij_map():
    """kernels/__init__.py:15-19: (1,0),(2,0),(2,1),(3,0),(3,1),(3,2),..."""
    k = golden('known_answers')
    pos = np.arange(15, dtype=np.float64).reshape(5, 3) ** 1.5
    scat = np.ones((5, 4), np.float32)
    d, r, norm, om = oracle.pair_internals(pos, scat, .1)
    q = pos.astype(np.float32)
    for kk, (i, j) in enumerate(k['k_to_ij'][:10]):
        assert np.array_equal(d[kk], q[i] - q[j])


def test_pdf_transform_known_vectors():
    k = golden('known_answers')
    out = oracle.get_pdf_at_qmin(k['random_fq'].copy(), EXP['rstep'],
                                 float(oracle.pdf_qbin(EXP)), oracle.r_grid(EXP), 0.0)
    assert nerr(out, k['random_fq_pdf']) < 1e-13
    qmin, rmin, rmax, rstep, pq = k['exp2_vals']
    rg = np.arange(rmin, rmax, rstep)
    out = oracle.get_pdf_at_qmin(k['random_fq2'].copy(), rstep, pq, rg, qmin)
    assert nerr(out, k['random_fq2_pdf']) < 1e-13


def test_gradient_convention_is_minus_half_true_derivative():
    """SURVEY.md section 8a note 1: reference grad = -1/2 dF/dq."""
    g = golden('au10_random')
    pos = g['positions'].astype(np.float32).astype(np.float64)
    scat = g['scatter_fq']
    grad = oracle.wrap_fq_grad(pos, scat, .1, 'fp64')
    h = 1e-5
    for (i, w) in [(0, 0), (3, 1), (7, 2)]:
        p1, p2 = pos.copy(), pos.copy()
        p1[i, w] += h
        p2[i, w] -= h
        fd = (oracle.wrap_fq(p1, scat, .1, 'fp64') -
              oracle.wrap_fq(p2, scat, .1, 'fp64')) / (2 * h)
        assert nerr(-0.5 * fd, grad[i, w]) < 1e-6


def test_rw_gradient_is_linear_in_grad_pdf():
    """master_kernel.get_grad_rw :293-347 equals sum_r c_r dG[i,w,r] with the
    closed-form c used by the CUDA potential kernel."""
    g = golden('au4_square')
    gc, go, gp = g['pdf_f64'], g['target_pdf_f64'], g['grad_pdf_f64']
    for pot in ('rw', 'chi_sq'):
        a = np.dot(gc, go) / np.dot(gc, gc)
        scale = a if a > 0 else 1.0
        d = go - scale * gc
        if pot == 'rw':
            val = np.sqrt(np.dot(d, d) / np.dot(go, go)) if a > 0 else 1.0
            pref = -val / np.dot(d, d)
            ref = oracle.wrap_grad_rw(gp, gc, go)
        else:
            pref = -2.0
            ref = oracle.wrap_grad_chi_sq(gp, gc, go)
        c = pref * (scale * d + np.dot(gc, d) / np.dot(gc, gc) * (go - 2 * a * gc))
        assert nerr(np.tensordot(gp, c, axes=([2], [0])), ref) < 1e-12
