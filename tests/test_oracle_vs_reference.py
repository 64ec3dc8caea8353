"""Pin the oracle to the real reference where it is mounted (build container
only; skipped on the GPU box, which has no /root/reference)."""
import numpy as np
import pytest

import oracle
from oracle import ref_shim
from conftest import nerr

pytestmark = pytest.mark.skipif(not ref_shim.available(),
                                reason='reference mount not present')


@pytest.mark.parametrize('n', [2, 3, 17, 60])
@pytest.mark.parametrize('prec', ['fp32', 'fp64'])
def test_c_oracle_is_bit_identical_to_reference_kernels(n, prec):
    rs = np.random.RandomState(n)
    pos = rs.random_sample((n, 3)) * 10
    f = (79 * np.exp(-0.01 * np.arange(250))).astype(np.float32)
    scat = np.tile(f, (n, 1))
    scat[::3] *= np.float32(0.9)
    assert np.array_equal(ref_shim.ref_fq(pos, scat, .1, prec),
                          oracle.wrap_fq(pos, scat, .1, prec))
    assert np.array_equal(ref_shim.ref_grad_fq(pos, scat, .1, prec),
                          oracle.wrap_fq_grad(pos, scat, .1, prec))
    q, s, d, r, norm, om = ref_shim.ref_pair_arrays(pos, scat, .1, prec)
    d2, r2, n2, o2 = oracle.pair_internals(pos, scat, .1, prec)
    for a, b in ((d, d2), (r, r2), (norm, n2), (om, o2)):
        assert np.array_equal(a, b)


def test_flat_equals_nxn_reference_variant():
    """reference tests/test_scatter_internals.py:39-94 (nxn vs flat)."""
    base, flat, exp, mk = ref_shim.kernels()
    nxn = ref_shim.nxn_kernels()
    n, nq = 9, 50
    rs = np.random.RandomState(0)
    q = (rs.random_sample((n, 3)) * 10).astype(np.float32)
    scat = (rs.random_sample((n, nq)) + 1).astype(np.float32)
    d = np.zeros((n, n, 3), np.float32)
    nxn.get_d_array(d, q)
    r = np.zeros((n, n), np.float32)
    nxn.get_r_array(r, d)
    dk, rk, normk, omk = oracle.pair_internals(q, scat, .1)
    assert np.array_equal(base.antisymmetric_reshape(dk), d.astype(np.float64))
    assert np.array_equal(base.symmetric_reshape(rk), r.astype(np.float64))


def test_master_kernel_restatement():
    mk = ref_shim.master()
    rs = np.random.RandomState(1)
    exp = oracle.DEFAULT_EXP
    qb = float(oracle.pdf_qbin(exp))
    f = rs.normal(size=330)
    for qmin in (0.0, 1.3):
        a = mk.get_pdf_at_qmin(f.copy(), exp['rstep'], qb, oracle.r_grid(exp), qmin)
        b = oracle.get_pdf_at_qmin(f.copy(), exp['rstep'], qb, oracle.r_grid(exp), qmin)
        assert np.array_equal(a, b)
    gc, go = rs.normal(size=400), rs.normal(size=400)
    gp = rs.normal(size=(5, 3, 400))
    for sign in (1, -1):
        assert np.allclose(mk.get_rw(go, sign * gc), oracle.get_rw(go, sign * gc))
        assert np.allclose(mk.get_chi_sq(go, sign * gc), oracle.get_chi_sq(go, sign * gc))
        rw, scale = mk.get_rw(go, sign * gc)
        ref = np.zeros((5, 3))
        mk.get_grad_rw(ref, gp, sign * gc, go, rw, scale)
        assert nerr(oracle.get_grad_rw(gp, sign * gc, go, rw, scale), ref) < 1e-14
        chi, scale = mk.get_chi_sq(go, sign * gc)
        ref = np.zeros((5, 3))
        mk.get_grad_chi_sq(ref, gp, sign * gc, go, scale)
        assert nerr(oracle.get_grad_chi_sq(gp, sign * gc, go, scale), ref) < 1e-14
