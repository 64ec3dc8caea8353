"""Parity of the CUDA path (through the public API and the C ABI) with the CPU
oracle and with the golden vectors generated from the reference.

Tolerances (conftest.py): normalised max-norm 1e-5 in FP32 mode on F(Q), G(r),
energies and forces; 1e-10 in FP64 mode.  The full gradient array in FP32 mode
is held to 1e-5 against the float64 variant of the reference arithmetic (same
float32-rounded inputs) and to the float32 reference's own noise floor (5e-5)
against the float32 oracle."""
import copy
import ctypes

import numpy as np
import pytest

import oracle
from conftest import golden, nerr, TOL32, TOL64, TOL32_GRAD_VS_F32
from pyiid_b200 import ElasticScatter, Calc1D, PDFCalc, structures, ase_shim, _lib
from pyiid_b200.calc import wrap_rw, wrap_chi_sq, wrap_grad_rw, wrap_grad_chi_sq
from pyiid_b200 import sim

pytestmark = pytest.mark.gpu

EXP = oracle.DEFAULT_EXP
CASES = ['au4_square', 'au10_random', 'au55_ico', 'aupt37_alloy']


def atoms_from(g, which='positions'):
    a = ase_shim.Atoms(numbers=g['numbers'], positions=g[which])
    a.set_array('F(Q) scatter', g['scatter_fq'])
    a.set_array('PDF scatter', g['scatter_pdf'])
    return a


def wrapped(scat, atoms, g):
    """Feed the SAME scatter-factor arrays to kernel and oracle."""
    atoms.info['exp'] = scat.exp
    atoms.info['scatter_atoms'] = len(atoms)
    scat.wrap_atoms_state = True
    return atoms


# ---- golden vectors from the reference -------------------------------------------
@pytest.mark.parametrize('name', CASES)
def test_fp32_against_reference_golden_vectors(name):
    g = golden(name)
    scat = ElasticScatter(precision='fp32')
    atoms = wrapped(scat, atoms_from(g), g)
    fq = scat.get_fq(atoms)
    assert fq.dtype == np.float32 and fq.shape == (250,)
    assert nerr(fq, g['fq_f32']) < TOL32
    assert nerr(fq, g['fq_f64']) < TOL32
    grad = scat.get_grad_fq(atoms)
    assert grad.shape == (len(atoms), 3, 250)
    assert nerr(grad, g['grad_fq_f64']) < TOL32
    assert nerr(grad, g['grad_fq_f32']) < TOL32_GRAD_VS_F32
    pdf = scat.get_pdf(atoms)
    assert pdf.dtype == np.float64 and pdf.shape == (4000,)
    assert nerr(pdf, g['pdf_f32']) < TOL32
    for pot in ('rw', 'chi_sq'):
        calc = Calc1D(target_data=g['target_pdf_f32'], exp_function=scat.get_pdf,
                      exp_grad_function=scat.get_grad_pdf, potential=pot, conv=1.)
        a = wrapped(scat, atoms_from(g), g)
        a.set_calculator(calc)
        val, scale = g['%s_f32' % pot]
        assert abs(a.get_potential_energy() - val) <= TOL32 * max(abs(val), 1.)
        assert nerr(a.get_forces(), g['%s_forces_f32' % pot]) < TOL32


@pytest.mark.parametrize('name', CASES)
def test_fp64_against_reference_golden_vectors(name):
    g = golden(name)
    scat = ElasticScatter(precision='fp64')
    atoms = wrapped(scat, atoms_from(g), g)
    assert nerr(scat.get_fq(atoms), g['fq_f64']) < TOL64
    assert nerr(scat.get_grad_fq(atoms), g['grad_fq_f64']) < TOL64
    assert nerr(scat.get_pdf(atoms), g['pdf_f64']) < TOL64
    for pot in ('rw', 'chi_sq'):
        calc = Calc1D(target_data=g['target_pdf_f64'], exp_function=scat.get_pdf,
                      exp_grad_function=scat.get_grad_pdf, potential=pot, conv=1.)
        a = wrapped(scat, atoms_from(g), g)
        a.set_calculator(calc)
        val, scale = g['%s_f64' % pot]
        assert abs(a.get_potential_energy() - val) <= TOL64 * max(abs(val), 1.)
        assert abs(calc.scale - scale) <= TOL64 * max(abs(scale), 1.)
        assert nerr(a.get_forces(), g['%s_forces_f64' % pot]) < 10 * TOL64


def test_grad_pdf_against_reference_golden_vector():
    g = golden('au4_square')
    for prec, tag, tol in (('fp32', 'f32', TOL32), ('fp64', 'f64', TOL64)):
        scat = ElasticScatter(precision=prec)
        atoms = wrapped(scat, atoms_from(g), g)
        gp = scat.get_grad_pdf(atoms)
        assert gp.shape == (4, 3, 4000) and gp.dtype == np.float64
        assert nerr(gp, g['grad_pdf_' + tag]) < tol


# ---- oracle on seeded inputs at several sizes --------------------------------------
@pytest.mark.parametrize('n', [2, 31, 33, 100, 561, 1000])
def test_fp32_against_oracle(n):
    atoms = structures.icosahedron('Au', 5) if n == 561 else \
        structures.random_atoms(n, n) if n < 500 else structures.fcc_sphere('Au', n)
    scat = ElasticScatter(precision='fp32')
    fq = scat.get_fq(atoms)
    grad = scat.get_grad_fq(atoms)
    pdf = scat.get_pdf(atoms)
    pos = atoms.get_positions()
    sf, sp = atoms.get_array('F(Q) scatter'), atoms.get_array('PDF scatter')
    th = 8 if n > 300 else 1
    for oprec in ('fp32', 'fp64'):
        p32 = pos.astype(np.float32)
        ofq = oracle.experiment_fq(p32, sf, EXP, oprec, nthreads=th)
        og = oracle.experiment_grad_fq(p32, sf, EXP, oprec, nthreads=th)
        opdf = oracle.experiment_pdf(p32, sp, EXP, oprec, nthreads=th)
        assert nerr(fq, ofq) < TOL32
        assert nerr(pdf, opdf) < TOL32
        assert nerr(grad, og) < (TOL32 if oprec == 'fp64' else TOL32_GRAD_VS_F32)


@pytest.mark.parametrize('n', [3, 64, 65, 300])
def test_fp64_against_oracle(n):
    atoms = structures.alloy_sphere(n, seed=n)
    scat = ElasticScatter(precision='fp64')
    pos = atoms.get_positions()
    fq, grad, pdf = scat.get_fq(atoms), scat.get_grad_fq(atoms), scat.get_pdf(atoms)
    sf, sp = atoms.get_array('F(Q) scatter'), atoms.get_array('PDF scatter')
    assert grad.dtype == np.float64
    assert nerr(fq, oracle.experiment_fq(pos, sf, EXP, 'fp64')) < TOL64
    assert nerr(grad, oracle.experiment_grad_fq(pos, sf, EXP, 'fp64')) < TOL64
    assert nerr(pdf, oracle.experiment_pdf(pos, sp, EXP, 'fp64')) < TOL64


def test_odd_experiment_and_qmin():
    """qmin > 0, rmin > 0, Nyquist sampling, Q bins not a multiple of 32."""
    exp = dict(qmin=1.3, qmax=21.7, qbin=.11, rmin=1.5, rmax=33.0, rstep=.02,
               sampling='ns')
    scat = ElasticScatter(dict(exp), precision='fp64')
    atoms = structures.alloy_sphere(50, seed=7)
    pos = atoms.get_positions()
    fq, grad = scat.get_fq(atoms), scat.get_grad_fq(atoms)
    pdf, gpdf = scat.get_pdf(atoms), scat.get_grad_pdf(atoms)
    sf, sp = atoms.get_array('F(Q) scatter'), atoms.get_array('PDF scatter')
    oexp = scat.exp
    assert len(fq) == len(scat.get_scatter_vector())
    assert nerr(fq, oracle.experiment_fq(pos, sf, oexp, 'fp64')) < TOL64
    assert nerr(grad, oracle.experiment_grad_fq(pos, sf, oexp, 'fp64')) < TOL64
    assert nerr(pdf, oracle.experiment_pdf(pos, sp, oexp, 'fp64')) < TOL64
    assert nerr(gpdf, oracle.experiment_grad_pdf(pos, sp, oexp, 'fp64')) < TOL64
    target = scat.get_pdf(structures.alloy_sphere(50, seed=7, sigma=0.0))
    for pot in ('rw', 'chi_sq'):
        e, f, _ = oracle.calc1d_energy_forces(pos, sp, oexp, target, pot, 3., 'fp64')
        a = atoms.copy()
        a.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf,
                                exp_grad_function=scat.get_grad_pdf, potential=pot,
                                conv=3.))
        assert abs(a.get_potential_energy() - e) <= TOL64 * abs(e)
        assert nerr(a.get_forces(), f) < 10 * TOL64


def test_single_atom_and_empty_pair_list():
    """k_max == 0 gives zeros (flat_multi_cpu_wrap.py:85-86)."""
    atoms = structures.random_atoms(1, 0)
    scat = ElasticScatter()
    assert not np.any(scat.get_fq(atoms))
    g = scat.get_grad_fq(atoms)
    assert g.shape == (1, 3, 250) and not np.any(g)
    assert not np.any(scat.get_pdf(atoms))


# ---- API behaviour the reference's tests check -----------------------------------
def test_fq_and_pdf_are_bit_reproducible_at_any_size():
    """The stand-alone F(Q) pass stores per-item partial sums and adds them in
    item order (no atomics): F(Q), G(r) and Rw energies reproduce bit for bit
    (pyiid/tests/test_consistancy.py:8-16), many items and two elements included."""
    atoms = structures.alloy_sphere(3000, seed=9)
    for prec in ('fp32', 'fp64'):
        scat = ElasticScatter(precision=prec)
        f1, p1 = scat.get_fq(atoms), scat.get_pdf(atoms)
        for _ in range(3):
            assert np.array_equal(scat.get_fq(atoms), f1)
            assert np.array_equal(scat.get_pdf(atoms), p1)
        be = scat.pdf_backend
        e1 = be.energy_forces(atoms.get_positions(), p1 * 1.01, 'rw', 1., want_forces=False)[0]
        e2 = be.energy_forces(atoms.get_positions(), p1 * 1.01, 'rw', 1., want_forces=False)[0]
        assert e1 == e2
        if prec == 'fp32':
            # the tabulated force pass (FP32 mode, >= 600 atoms) adds the j shares
            # of a row in share order: forces reproduce as well
            r1 = be.energy_forces(atoms.get_positions(), p1 * 1.01, 'rw', 1.)
            for _ in range(3):
                r2 = be.energy_forces(atoms.get_positions(), p1 * 1.01, 'rw', 1.)
                assert r2[0] == r1[0] and np.array_equal(r2[2], r1[2])


def test_smoke_every_method_returns_fresh_nonzero_arrays():
    """tests/test_scatter_smoke.py:13-181 and test_scatter.py:43."""
    atoms = structures.random_atoms(20, 3)
    scat = ElasticScatter()
    for name in ('get_fq', 'get_pdf', 'get_sq', 'get_iq', 'get_grad_fq', 'get_grad_pdf'):
        a1 = getattr(scat, name)(atoms)
        a2 = getattr(scat, name)(atoms)
        assert a1 is not None and np.any(a1) and a1 is not a2
        # run-to-run determinism (tests/test_consistancy.py); S(Q=0) is 0/0 = NaN
        # in the reference too (only inf is cleared, __init__.py:419)
        assert np.array_equal(a1, a2, equal_nan=True)
    img = scat.get_2d_scatter(atoms, np.linspace(0, 20, 64).reshape(8, 8))
    assert img.shape == (8, 8) and np.any(img)
    noisy = ElasticScatter(seed=1).get_pdf(atoms, iq_std=0.01)
    assert noisy.shape == (4000,) and not np.allclose(noisy, scat.get_pdf(atoms))
    assert scat.get_fq(atoms, iq_std=0.01).shape == (250,)


def test_state_invalidation_changes_results():
    """tests/test_scatter_state.py:11-71."""
    atoms = structures.random_atoms(10, 4)
    scat = ElasticScatter()
    a1 = scat.get_fq(atoms)
    atoms2 = atoms + ase_shim.Atom('Au', [0, 0, 0])
    a2 = scat.get_fq(atoms2)
    assert not np.allclose(a1, a2)
    atoms3 = copy.deepcopy(atoms)
    del atoms3[2]
    assert not np.allclose(a1, scat.get_fq(atoms3))


def test_calc1d_known_system():
    """tests/test_calc/test_calc_1d_known.py: Au4 square vs 0.75-scaled copy:
    rw >= 0.9, forces central (cross(r_i - com, F_i) = 0, atol 1e-7)."""
    a1, a2 = structures.atomic_square()
    for exp_name in ('PDF', 'FQ'):
        scat = ElasticScatter(precision='fp64')
        f, gf = (scat.get_pdf, scat.get_grad_pdf) if exp_name == 'PDF' else \
            (scat.get_fq, scat.get_grad_fq)
        calc = Calc1D(target_data=f(a1), exp_function=f, exp_grad_function=gf,
                      potential='rw')
        b = a2.copy()
        b.set_calculator(calc)
        assert b.get_potential_energy() >= .9
        forces = b.get_forces()
        com = b.get_center_of_mass()
        for i in range(4):
            assert np.allclose(np.cross(b[i].position - com, forces[i]), 0, atol=1e-7)


def test_generic_calc_route_equals_fused_route():
    """Calc1D through get_pdf/get_grad_pdf + wrap_grad_rw (the reference's
    route, calc_1d.py:78-95) equals the fused device evaluation."""
    atoms = structures.random_atoms(12, 5)
    tgt_atoms = structures.random_atoms(12, 6)
    scat = ElasticScatter(precision='fp64')
    target = scat.get_pdf(tgt_atoms)
    for pot, wp, wg in (('rw', wrap_rw, wrap_grad_rw), ('chi_sq', wrap_chi_sq, wrap_grad_chi_sq)):
        a = atoms.copy()
        a.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf,
                                exp_grad_function=scat.get_grad_pdf, potential=pot, conv=2.))
        e, f = a.get_potential_energy(), a.get_forces()
        gcalc, ggrad = scat.get_pdf(atoms), scat.get_grad_pdf(atoms)
        e2, scale = wp(gcalc, target)
        f2 = wg(ggrad, gcalc, target)
        assert abs(e - 2. * e2) < 1e-10 * abs(e)
        assert nerr(2. * f2, f) < 1e-9
        oe, of, _ = oracle.calc1d_energy_forces(atoms.get_positions(),
                                                atoms.get_array('PDF scatter'), EXP,
                                                target, pot, 2., 'fp64')
        assert nerr(f, of) < 10 * TOL64
    pc = PDFCalc(obs_data=target, scatter=scat, conv=2., potential='rw')
    b = atoms.copy()
    b.set_calculator(pc)
    assert abs(b.get_potential_energy() - e) >= 0  # runs; energy of rw vs chi differs


def test_known_answers_on_device():
    """test_master_kernel.py:19-48 through the device potential kernel."""
    x = np.arange(0, 2 * np.pi, .1)
    assert wrap_rw(np.cos(x), np.sin(x))[0] == 1
    assert abs(wrap_rw(np.sin(x), np.sin(x))[0]) < 1e-15
    assert abs(wrap_chi_sq(np.cos(x), np.sin(x))[0] - 63.01399) < 1e-5
    assert abs(wrap_chi_sq(np.sin(x), np.sin(x))[0]) < 1e-15


@pytest.mark.parametrize('potential', ['rw', 'chi_sq'])
def test_anti_correlated_target_takes_the_scale_le_zero_branch(potential):
    """master_kernel.py:229-230, 263-264, 338-340, 368-370: a target that is
    anti-correlated with the model gives scale <= 0; get_rw then returns
    (1, 1), get_chi_sq resets the scale to 1, and the gradients follow.  Driven
    on the device through the fused path (potential_kernel + Q-space weights +
    force pass) and through the generic wrap_* route (iid_rw_host)."""
    atoms = structures.random_atoms(24, 11)
    exp = oracle.DEFAULT_EXP
    scat = ElasticScatter(precision='fp64')
    gcalc = scat.get_pdf(atoms)
    rs = np.random.RandomState(3)
    target = -0.7 * gcalc + 0.05 * np.abs(gcalc).max() * rs.standard_normal(gcalc.shape)
    assert oracle.get_scale(target, gcalc) < 0
    pos = atoms.get_positions()
    sp = atoms.get_array('PDF scatter')
    oe, of, oscale = oracle.calc1d_energy_forces(pos, sp, exp, target, potential, 3., 'fp64')
    assert oscale == 1
    a = atoms.copy()
    a.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf,
                            exp_grad_function=scat.get_grad_pdf, conv=3.,
                            potential=potential))
    e, f = a.get_potential_energy(), a.get_forces()
    assert a.calc.scale == 1
    assert abs(e - oe) < 1e-10 * abs(oe)
    assert nerr(f, of) < 10 * TOL64
    if potential == 'rw':
        assert e == 3.0  # Rw = 1 exactly
    # the generic route: potential and chain-rule contraction on host arrays
    gp = scat.get_grad_pdf(atoms)
    wrap, wgrad = (wrap_rw, wrap_grad_rw) if potential == 'rw' else (wrap_chi_sq, wrap_grad_chi_sq)
    val, sc = wrap(gcalc, target)
    assert sc == 1 and abs(3. * val - oe) < 1e-10 * abs(oe)
    assert nerr(3. * wgrad(gp, gcalc, target), of) < 10 * TOL64
    # FP32 mode, r-space weights (iid_potential + wq_kernel) as well
    s32 = ElasticScatter(precision='fp32')
    s32._ensure_wrapped(atoms)
    be = s32._load(atoms, s32.pdf_qbin, 'PDF')
    be.set_transform(exp['rstep'], s32.pdf_qbin, s32.get_r(), exp['qmin'])
    for q in (1, 0):
        be.set_option('qspace_wq', q)
        e32, sc32, f32, _ = be.energy_forces(pos, target, potential, 3.)
        assert sc32 == 1 and abs(e32 - oe) < TOL32 * abs(oe) and nerr(f32, of) < 5 * TOL32
    be.set_option('qspace_wq', 1)


# ---- size-independent properties at full size ---------------------------------------
@pytest.fixture(scope='module')
def big():
    atoms = structures.fcc_sphere('Au', 10000)
    scat = ElasticScatter(precision='fp32')
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.exp['qbin'], 'fq')
    g, f = be.grad_fq(atoms.get_positions(), with_fq=True)
    return atoms, scat, g, f


def test_10k_properties(big):
    atoms, scat, g, f = big
    pos = atoms.get_positions()
    be = scat.backend
    # the F(Q) that comes with the gradient pass equals the triangle-pass F(Q)
    f_tri = be.fq(pos)
    assert nerr(f, f_tri) < 1e-6
    # Newton's third law: the gradient summed over atoms vanishes
    assert np.abs(g.sum(axis=0, dtype=np.float64)).max() < 1e-4 * np.abs(g).max()
    # translation and rigid rotation leave F(Q) unchanged
    assert nerr(be.fq(pos + np.array([3.25, -1.5, 7.0])), f_tri) < TOL32
    c, s = np.cos(0.7), np.sin(0.7)
    rot = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.]])
    assert nerr(be.fq(pos.dot(rot.T)), f_tri) < TOL32
    # permutation of the atoms permutes the gradient rows
    perm = np.random.RandomState(0).permutation(len(pos))
    gp = be.grad_fq(pos[perm])
    # two summation orders of float32 partial sums (each within 2e-6 of the
    # float64 mode, test_10k_fp32_agrees_with_fp64_mode)
    assert nerr(gp, g[perm]) < 4e-6
    # F(Q=0) = 0 and the result is finite
    assert f[0] == 0 and np.all(np.isfinite(g)) and np.all(np.isfinite(f))


def test_10k_gradient_is_bit_reproducible(big):
    """Row ownership: every gradient row is stored once by one block and the
    F(Q) partials of the row jobs are added in a fixed order, so two runs give
    the same bits (the reference holds itself to that,
    pyiid/tests/test_consistancy.py:8-16); with pieces of any size, in pinned
    (kernel-written) and pageable (staged download) destinations."""
    import ctypes
    atoms, scat, g, f = big
    pos = atoms.get_positions()
    be = scat.backend
    g2, f2 = be.grad_fq(pos, with_fq=True)
    assert g2 is not g and np.array_equal(g2, g) and np.array_equal(f2, f)
    from pyiid_b200 import hostmem
    assert hostmem.is_pinned(g2)
    # pageable destination through the C ABI: same bits
    gp = np.empty_like(g)
    fp = np.empty_like(f)
    assert be.lib.iid_grad_fq_host(be.h, pos.ctypes.data, gp.ctypes.data, fp.ctypes.data) == 0
    assert np.array_equal(gp, g) and np.array_equal(fp, f)
    # another cut of the split rows: other float32 partial sums for those rows,
    # same result to rounding, and again reproducible
    be.set_option('piece_div', 16)
    try:
        g3 = be.grad_fq(pos)
        g4 = be.grad_fq(pos)
    finally:
        be.set_option('piece_div', 128)
    assert np.array_equal(g3, g4)
    assert nerr(g3, g) < 2e-6
    # the pool hands a buffer out again only after its array has died
    addr = g4.ctypes.data
    del g4
    g5 = be.grad_fq(pos)
    assert g5.ctypes.data == addr and np.array_equal(g5, g)


def test_one_process_multi_gpu_handle():
    """set_processor('Multi-GPU') in ONE process uses every GPU of the box
    (iid_create_multi; reference gpu_wrap.py:119-156, 287-314): results equal
    the one-GPU handle's."""
    from pyiid_b200.backend import visible_devices
    if visible_devices() < 2:
        pytest.skip('needs 2 GPUs')
    atoms = structures.alloy_sphere(4000, seed=3)
    ideal = structures.alloy_sphere(4000, seed=3, sigma=0.0)
    one = ElasticScatter(device=0)
    many = ElasticScatter()
    assert many.set_processor('Multi-GPU') is True
    assert one.processor == 'B200' and many.processor == 'Multi-GPU'
    f1, fm = one.get_fq(atoms), many.get_fq(atoms)
    assert many.backend.devices()[1] >= 2
    assert nerr(fm, f1) < 1e-6
    g1, gm = one.get_grad_fq(atoms), many.get_grad_fq(atoms)
    assert nerr(gm, g1) < 2e-6
    assert np.array_equal(gm, many.get_grad_fq(atoms))
    p1, pm = one.get_pdf(atoms), many.get_pdf(atoms)
    assert nerr(pm, p1) < 1e-6
    target = one.get_pdf(ideal)
    res = []
    for scat in (one, many):
        a = atoms.copy()
        a.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf,
                                exp_grad_function=scat.get_grad_pdf, conv=100.,
                                potential='rw'))
        res.append((a.get_potential_energy(), a.get_forces()))
    assert abs(res[0][0] - res[1][0]) < 1e-6 * abs(res[0][0])
    assert nerr(res[1][1], res[0][1]) < TOL32
    # restraints ride along: MultiCalc(Calc1D + rep spring) on all devices
    from pyiid_b200.spring_calc import Spring
    from pyiid_b200.multi_calc import MultiCalc
    res = []
    for scat in (one, many):
        a = atoms.copy()
        c1 = Calc1D(target_data=target, exp_function=scat.get_pdf,
                    exp_grad_function=scat.get_grad_pdf, conv=100., potential='rw')
        a.set_calculator(MultiCalc(calc_list=[c1, Spring(k=10., rt=2.9, sp_type='rep')]))
        res.append((a.get_potential_energy(), a.get_forces()))
    assert abs(res[0][0] - res[1][0]) < 1e-6 * abs(res[0][0])
    assert nerr(res[1][1], res[0][1]) < TOL32
    # a small structure stays on one device, the device-resident sampler included
    small = structures.random_atoms(50, 1)
    assert nerr(many.get_fq(small), one.get_fq(small)) < 1e-6
    assert many.backend.devices()[1] == 1
    assert nerr(many.get_grad_pdf(small), one.get_grad_pdf(small)) < 1e-6
    ico = structures.icosahedron('Au', 2)
    tgt = many.get_pdf(ico)
    trajs = []
    for scat in (one, many):
        a = ico.copy()
        a.positions *= 1.03
        a.set_calculator(Calc1D(target_data=tgt, exp_function=scat.get_pdf,
                                exp_grad_function=scat.get_grad_pdf, conv=100., potential='rw'))
        np.random.seed(2)
        ens = sim.NUTSCanonicalEnsemble(a, temperature=300, escape_level=4, seed=5)
        assert ens.device_states
        traj, _ = ens.run(3)
        trajs.append(np.array([x.get_positions() for x in traj]))
    assert trajs[0].shape == trajs[1].shape and np.allclose(trajs[0], trajs[1], atol=1e-9)
    # a large one cannot keep sampler states on one device: array-level path
    big = atoms.copy()
    big.set_calculator(Calc1D(target_data=target, exp_function=many.get_pdf,
                              exp_grad_function=many.get_grad_pdf, conv=100., potential='rw'))
    assert sim._FastSystem.usable(big) and not sim._DeviceSystem.usable(big)


def test_10k_fp32_agrees_with_fp64_mode(big):
    atoms, scat, g, f = big
    s64 = ElasticScatter(precision='fp64')
    s64._ensure_wrapped(atoms)
    be = s64._load(atoms, s64.exp['qbin'], 'fq')
    pos32 = atoms.get_positions().astype(np.float32).astype(np.float64)
    g64, f64 = be.grad_fq(pos32, with_fq=True)
    assert nerr(f, f64) < TOL32
    assert nerr(g, g64) < TOL32


def test_10k_force_is_weighted_gradient(big):
    """force[i,w] = sum_m wq[m] G[i,w,m] (fused path vs full gradient) and the
    oracle on a bounded sample of rows."""
    atoms, scat, g, f = big
    pos = atoms.get_positions()
    be = scat._load(atoms, scat.pdf_qbin, 'PDF')
    be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), 0.0)
    target = be.pdf(structures.fcc_sphere('Au', 10000, sigma=0.0).get_positions())
    e, scale, forces, pdf = be.energy_forces(pos, target, 'rw', 100., want_pdf=True)
    gfull = be.grad_fq(pos)                      # [N,3,330] on the PDF grid
    from pyiid_b200.backend import pdf_matrix
    t = pdf_matrix(330, .01, scat.pdf_qbin, scat.get_r(), 0.0)
    # chain-rule weights from the oracle's Rw on the device PDF
    a = np.dot(pdf, target) / np.dot(pdf, pdf)
    d = target - a * pdf
    rw = np.sqrt(np.dot(d, d) / np.dot(target, target))
    c = -rw / np.dot(d, d) * (a * d + np.dot(pdf, d) / np.dot(pdf, pdf) * (target - 2 * a * pdf))
    wq = 100. * t.T.dot(c)
    ref = np.tensordot(gfull.astype(np.float64), wq, axes=([2], [0]))
    assert abs(e - 100. * rw) < 1e-9 * abs(e)
    assert nerr(forces, ref) < TOL32
    # Newton's third law up to float32 rounding of the per-pair scalar
    assert np.abs(forces.sum(0)).max() < 1e-4 * np.abs(forces).max()


def test_sharded_partials_sum_to_the_whole():
    """iid_set_shard on ONE GPU: the partial results of world=3 shards add up
    to the unsharded result (what the NCCL all-reduce does across GPUs)."""
    import torch
    atoms = structures.alloy_sphere(700, seed=2)
    scat = ElasticScatter(precision='fp64')
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.exp['qbin'], 'fq')
    pos = atoms.get_positions()
    g_ref, f_ref = be.grad_fq(pos, with_fq=True)
    lib, h = be.lib, be.h
    dev = 'cuda:%d' % be.device
    with be._on_stream():
        p = torch.from_numpy(pos).to(dev)
        s_tot = torch.zeros(be.nq, dtype=torch.float64, device=dev)
        g_tot = torch.zeros((be.n, 3, be.nq), dtype=torch.float64, device=dev)
        s_tri = torch.zeros_like(s_tot)
        written = torch.zeros(be.n, dtype=torch.int32, device=dev)
        for rank in range(3):
            assert lib.iid_set_shard(h, rank, 3) == 0
            s = torch.zeros_like(s_tot)
            g = torch.zeros_like(g_tot)
            assert lib.iid_grad_fq_partial(h, p.data_ptr(), g.data_ptr(), s.data_ptr(), None) == 0
            s_tot += s
            g_tot += g
            # row ownership: a shard writes complete rows of its own atoms only
            written += (g.abs().amax(dim=(1, 2)) > 0).to(torch.int32)
            assert lib.iid_fq_partial(h, p.data_ptr(), s.data_ptr(), None) == 0
            s_tri += s
        assert lib.iid_set_shard(h, 0, 1) == 0
        assert bool((written == 1).all())
        f = torch.zeros_like(s_tot)
        assert lib.iid_fq_finish(h, s_tot.data_ptr(), f.data_ptr(), None) == 0
        f_sq = f.cpu().numpy()
        assert lib.iid_fq_finish(h, s_tri.data_ptr(), f.data_ptr(), None) == 0
        f_tri = f.cpu().numpy()
        g_sum = g_tot.cpu().numpy()
    assert nerr(f_sq, f_ref) < 1e-12 and nerr(f_tri, f_ref) < 1e-12
    assert nerr(g_sum, g_ref) < 1e-12


def test_fq_through_the_pair_histogram():
    """FP32 mode, large structures: the F(Q) pair sum goes through a radial pair
    histogram (iid_fq_hist.cuh).  It must give the direct pass's F(Q) (both are
    within 1e-5 of the float64 oracle; the histogram is the more accurate one),
    bit-reproducibly, for one / two / three element types, on both Q grids, as
    shards that add up, and fall back to the direct kernel when the structure
    does not fit the histogram."""
    import torch
    rs = np.random.RandomState(3)
    tri_numbers = rs.choice([79, 78, 47], 1500)
    cases = [structures.fcc_sphere('Au', 2000), structures.alloy_sphere(1800, seed=3),
             ase_shim.Atoms(numbers=tri_numbers,
                            positions=structures.fcc_sphere_positions(1500, 3.9, 0.05, 3))]
    for atoms in cases:
        scat = ElasticScatter()
        scat._ensure_wrapped(atoms)
        pos = atoms.get_positions()
        for kind, qb, key in (('fq', scat.exp['qbin'], 'F(Q) scatter'),
                              ('PDF', scat.pdf_qbin, 'PDF scatter')):
            be = scat._load(atoms, qb, kind)
            assert be.n >= 1200  # the histogram pass is the default at this size
            f_hist = be.fq(pos)
            n0 = be.launch_count()
            assert np.array_equal(be.fq(pos), f_hist)
            per_call = be.launch_count() - n0
            be.set_option('fq_hist', 0)
            f_direct = be.fq(pos)
            be.set_option('fq_hist', 1)
            assert nerr(f_hist, f_direct) < 5e-7, (kind, nerr(f_hist, f_direct))
            if kind == 'fq':
                exp = dict(scat.exp)
                ref = oracle.experiment_fq(pos.astype(np.float32), atoms.get_array(key), exp,
                                           'fp64', nthreads=8)
                assert nerr(f_hist, ref) < 2e-6 and nerr(f_direct, ref) < TOL32
                assert per_call == 7  # staging, grid, histogram, transform, sum, gated direct, finish
    # shards: the ranks' pair histograms are partial sums of S
    atoms = cases[1]
    scat = ElasticScatter()
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.exp['qbin'], 'fq')
    pos = atoms.get_positions()
    f_ref = be.fq(pos)
    lib, h = be.lib, be.h
    dev = 'cuda:%d' % be.device
    with be._on_stream():
        p = torch.from_numpy(pos).to(dev)
        s_tot = torch.zeros(be.nq, dtype=torch.float64, device=dev)
        for rank in range(3):
            assert lib.iid_set_shard(h, rank, 3) == 0
            s = torch.zeros_like(s_tot)
            assert lib.iid_fq_partial(h, p.data_ptr(), s.data_ptr(), None) == 0
            s_tot += s
        assert lib.iid_set_shard(h, 0, 1) == 0
        f = torch.zeros_like(s_tot)
        assert lib.iid_fq_finish(h, s_tot.data_ptr(), f.data_ptr(), None) == 0
        f_sum = f.cpu().numpy()
    assert nerr(f_sum, f_ref) < 1e-12
    # the compact structures above take the finest grid (6 points, up to 75 A); two
    # clusters 100 A apart the fine one (8 points, up to 180 A), 200 A apart the
    # coarse one (12 points, up to 382 A) -- the same F(Q) as the direct pass and float64
    for gap in (100., 200.):
        mid = structures.fcc_sphere('Au', 1300)
        mid.positions[650:] += [gap, 0., 0.]
        scat = ElasticScatter()
        scat._ensure_wrapped(mid)
        be = scat._load(mid, scat.exp['qbin'], 'fq')
        pos = mid.get_positions()
        f_hist = be.fq(pos)
        assert np.array_equal(be.fq(pos), f_hist)
        be.set_option('fq_hist', 0)
        f_direct = be.fq(pos)
        be.set_option('fq_hist', 1)
        ref = oracle.experiment_fq(pos.astype(np.float32), mid.get_array('F(Q) scatter'),
                                   dict(scat.exp), 'fp64', nthreads=8)
        assert nerr(f_hist, f_direct) < 5e-7 and nerr(f_hist, ref) < 2e-6, gap
    # a structure that does not fit the histogram: the gated direct kernel runs
    far = structures.fcc_sphere('Au', 1300)
    far.positions[650:] += [800., 0., 0.]
    scat = ElasticScatter()
    scat._ensure_wrapped(far)
    be = scat._load(far, scat.exp['qbin'], 'fq')
    pos = far.get_positions()
    f_gate = be.fq(pos)
    be.set_option('fq_hist', 0)
    f_direct = be.fq(pos)
    be.set_option('fq_hist', 1)
    assert nerr(f_gate, f_direct) < 2e-7 and np.abs(f_gate).max() > 0


# ---- samplers on the device calculator -------------------------------------------------
def make_hmc_atoms(shells=2, precision='fp32'):
    scat = ElasticScatter(precision=precision)
    ideal = structures.icosahedron('Au', shells)
    target = scat.get_pdf(ideal)
    atoms = structures.icosahedron('Au', shells)
    atoms.positions *= 1.05
    atoms.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf,
                                exp_grad_function=scat.get_grad_pdf, conv=100,
                                potential='rw'))
    return atoms, scat


def test_leapfrog_reversibility_and_single_evaluation():
    """tests/test_sim/test_leapfrog.py:58-73 on the PDF calculator; one device
    evaluation per leapfrog (the reference needs three)."""
    atoms, scat = make_hmc_atoms(2, 'fp64')
    atoms.set_momenta(np.random.RandomState(0).normal(0, 1, (55, 3)))
    atoms.get_forces()
    be = scat.pdf_backend
    n0 = be.launch_count()
    b = sim.leapfrog(atoms, 0.05, False)
    b.get_total_energy()
    per_eval = be.launch_count() - n0
    assert per_eval <= 8
    c = sim.leapfrog(b, -0.05, False)
    assert np.allclose(c.positions, atoms.positions, atol=1e-9)
    assert np.allclose(c.get_momenta(), atoms.get_momenta(), atol=1e-9)


def test_forces_are_minus_half_the_energy_gradient():
    """Reference convention (SURVEY.md 8a note 1): forces = -1/2 dE/dq."""
    atoms, scat = make_hmc_atoms(1, 'fp64')
    f = atoms.get_forces()
    h = 1e-5
    for i, w in ((0, 0), (5, 2)):
        ap, am = atoms.copy(), atoms.copy()
        ap.positions[i, w] += h
        am.positions[i, w] -= h
        for x in (ap, am):
            x.set_calculator(atoms.get_calculator())
        de = (ap.get_potential_energy() - am.get_potential_energy()) / (2 * h)
        assert abs(f[i, w] - (-0.5 * de)) < 1e-5 * np.abs(f).max()


def test_nuts_runs_on_the_device_calculator():
    """tests/test_sim/test_nuts.py:17-64 with the PDF calculator."""
    atoms, scat = make_hmc_atoms(2, 'fp32')
    np.random.seed(0)
    e0 = atoms.get_potential_energy()
    ens = sim.NUTSCanonicalEnsemble(atoms, temperature=1000, escape_level=4, seed=0)
    traj, meta = ens.run(5)
    pe = [t.get_potential_energy() for t in traj]
    assert meta['samples_total'] > 0 and len(traj) >= 1
    assert min(pe) <= e0
    assert all(np.isfinite(pe))


# ---- randomised experiments and odd inputs (reference tests/__init__.py:117-128) -----
def random_experiment(rs):
    exp = {}
    ranges = dict(qmin=(0, 1.5), qmax=(19., 25.), qbin=(.08, .12), rmin=(0., 2.5),
                  rmax=(30., 50.), rstep=(.005, .015))
    for k, (lo, hi) in ranges.items():
        exp[k] = float(rs.uniform(lo, hi))
    exp['sampling'] = str(rs.choice(['full', 'ns']))
    return exp


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_random_experiments_against_oracle(seed):
    rs = np.random.RandomState(100 + seed)
    exp = random_experiment(rs)
    n = int(rs.choice([10, 57, 100]))
    atoms = structures.alloy_sphere(n, seed=seed) if seed % 2 else structures.random_atoms(n, seed)
    for prec, tol, gtol in (('fp32', TOL32, TOL32), ('fp64', TOL64, TOL64)):
        scat = ElasticScatter(dict(exp), precision=prec)
        fq, grad, pdf = scat.get_fq(atoms), scat.get_grad_fq(atoms), scat.get_pdf(atoms)
        pos = atoms.get_positions()
        opos = pos.astype(np.float32) if prec == 'fp32' else pos
        sf, sp = atoms.get_array('F(Q) scatter'), atoms.get_array('PDF scatter')
        assert nerr(fq, oracle.experiment_fq(opos, sf, scat.exp, 'fp64')) < tol
        assert nerr(grad, oracle.experiment_grad_fq(opos, sf, scat.exp, 'fp64')) < gtol
        assert nerr(pdf, oracle.experiment_pdf(opos, sp, scat.exp, 'fp64')) < tol
        assert len(pdf) == len(scat.get_r())


@pytest.mark.parametrize('qmax', [2.0, 45.0])
def test_very_short_and_very_long_q_grids(qmax):
    """Q grids outside the default: 20 bins (one warp per block) and 450 bins
    (more chunks than one block holds: two blocks per item / row job share the
    Q range), every pass, both precisions, energy + forces included."""
    exp = {'qmin': 0.0, 'qmax': qmax, 'qbin': .1, 'rmin': 0.0, 'rmax': 20.0, 'rstep': .02}
    atoms = structures.alloy_sphere(150, seed=12)
    ideal = structures.alloy_sphere(150, seed=12, sigma=0.0)
    for prec, tol in (('fp32', TOL32), ('fp64', TOL64)):
        scat = ElasticScatter(dict(exp), precision=prec)
        fq, grad, pdf = scat.get_fq(atoms), scat.get_grad_fq(atoms), scat.get_pdf(atoms)
        pos = atoms.get_positions()
        opos = pos.astype(np.float32) if prec == 'fp32' else pos
        sf, sp = atoms.get_array('F(Q) scatter'), atoms.get_array('PDF scatter')
        assert fq.shape == (int(qmax / .1),)
        assert nerr(fq, oracle.experiment_fq(opos, sf, scat.exp, 'fp64')) < tol
        assert nerr(grad, oracle.experiment_grad_fq(opos, sf, scat.exp, 'fp64')) < tol
        assert nerr(pdf, oracle.experiment_pdf(opos, sp, scat.exp, 'fp64')) < tol
        assert np.array_equal(grad, scat.get_grad_fq(atoms))
        target = scat.get_pdf(ideal)
        a = atoms.copy()
        a.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf,
                                exp_grad_function=scat.get_grad_pdf, conv=5., potential='rw'))
        e, f = a.get_potential_energy(), a.get_forces()
        oe, of, _ = oracle.calc1d_energy_forces(opos, sp, scat.exp, target, 'rw', 5., 'fp64')
        assert abs(e - oe) < 10 * tol * abs(oe) and nerr(f, of) < 10 * tol


def test_five_element_types():
    """More element runs than the tabulated force pass takes (<= 4): every run is
    padded to a multiple of 32 and the row jobs flush once per run."""
    rs = np.random.RandomState(21)
    n = 260
    numbers = rs.choice([26, 28, 46, 47, 79], n)
    atoms = ase_shim.Atoms(numbers=numbers, positions=structures.fcc_sphere_positions(n, 3.9, 0.05, 3))
    ideal = ase_shim.Atoms(numbers=numbers, positions=structures.fcc_sphere_positions(n, 3.9, 0.0, 3))
    exp = oracle.DEFAULT_EXP
    for prec, tol in (('fp32', TOL32), ('fp64', TOL64)):
        scat = ElasticScatter(precision=prec)
        fq, grad, pdf = scat.get_fq(atoms), scat.get_grad_fq(atoms), scat.get_pdf(atoms)
        pos = atoms.get_positions()
        opos = pos.astype(np.float32) if prec == 'fp32' else pos
        sf, sp = atoms.get_array('F(Q) scatter'), atoms.get_array('PDF scatter')
        assert nerr(fq, oracle.experiment_fq(opos, sf, exp, 'fp64')) < tol
        assert nerr(grad, oracle.experiment_grad_fq(opos, sf, exp, 'fp64')) < tol
        assert nerr(pdf, oracle.experiment_pdf(opos, sp, exp, 'fp64')) < tol
        target = scat.get_pdf(ideal)
        a = atoms.copy()
        a.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf,
                                exp_grad_function=scat.get_grad_pdf, conv=5., potential='chi_sq'))
        e, f = a.get_potential_energy(), a.get_forces()
        oe, of, _ = oracle.calc1d_energy_forces(opos, sp, exp, target, 'chi_sq', 5., 'fp64')
        assert abs(e - oe) < 10 * tol * abs(oe) and nerr(f, of) < 10 * tol


@pytest.mark.parametrize('potential', ['rw', 'chi_sq'])
def test_fused_launch_equals_the_launch_sequence(potential):
    """Small structures evaluate in ONE cooperative launch with the potential in
    Q space (gc.go = F.(T^T go), gc.gc = F.(T^T T F)); it must give what the
    sequence of launches (G(r) in r space, potential_kernel, force pass with
    atomics) gives, with and without a fused spring; its fixed-point sums make
    repeated evaluations bit-identical."""
    atoms, scat = make_hmc_atoms(5)
    be = scat.pdf_backend
    pos, target = atoms.get_positions(), atoms.calc.target_data
    for springs in ([], [('rep', 10., 2.95)]):
        be.set_restraints(springs)
        res = {}
        for fused, table in ((0, 0), (1, 0), (1, 1)):
            be.set_option('fused', fused)
            be.set_option('fused_table', table)
            outs = []
            for _ in range(4):  # eager, eager, capture, replay
                out = be.energy_forces(pos, target, potential, 100.)
                outs.append((out[0], out[1], np.array(out[2]), be.restraint_energy))
            res[fused, table] = outs[-1]
            if fused and not springs:
                for o in outs[:-1]:
                    assert o[0] == outs[-1][0] and o[1] == outs[-1][1]
                    assert np.array_equal(o[2], outs[-1][2])
        be.set_option('fused', 1)
        be.set_option('fused_table', 1)
        e0, s0, f0, r0 = res[0, 0]
        for key in ((1, 0), (1, 1)):
            e, sc, f, r = res[key]
            # (the fused launch sums F(Q) through its radial pair histogram: float32
            # interpolation weights, 1e-7 of a pair's term)
            assert abs(e - e0) < 1e-7 * abs(e0) and abs(sc - s0) < 1e-7 * abs(s0)
            assert nerr(f, f0) < 2e-6 and abs(r - r0) <= 1e-12 * max(1., abs(r0)), (key, nerr(f, f0))
        print('fused forces vs launch sequence:', potential, springs,
              nerr(res[1, 0][2], f0), nerr(res[1, 1][2], f0))
    be.set_restraints([])


def _fused_backend(atoms, ideal, precision='fp32'):
    scat = ElasticScatter(precision=precision)
    target = scat.get_pdf(ideal)
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.pdf_qbin, 'PDF')
    be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), scat.exp['qmin'])
    return be, target


def test_fused_force_table_two_elements_and_far_pairs():
    """The radial force table of the fused launch: (1) a two-element structure
    (one table per ordered element pair) against the FP64 handle and the direct
    pass; (2) a structure stretched beyond the table's last entry -- two
    clusters 1000 A apart -- whose far pairs are summed directly over the Q
    bins; (3) three element types keep the direct pass."""
    atoms = structures.alloy_sphere(300, seed=3)
    ideal = atoms.copy()
    atoms.positions = atoms.positions * 1.03 + np.random.RandomState(1).normal(0, 0.03, (300, 3))
    pos = atoms.get_positions()
    be64, target = _fused_backend(atoms, ideal, 'fp64')
    e64, s64, f64 = be64.energy_forces(pos, target, 'rw', 1.)[:3]
    be, target = _fused_backend(atoms, ideal)
    res = {}
    for table in (0, 1):
        be.set_option('fused_table', table)
        be.energy_forces(pos, target, 'rw', 1.)  # (first call: the target's one-time kernels)
        n0 = be.launch_count()
        res[table] = be.energy_forces(pos, target, 'rw', 1.)[:3]
        assert be.launch_count() - n0 == 1  # the fused launch either way
        assert abs(res[table][0] - e64) < 1e-6 * abs(e64)
        assert nerr(res[table][2], f64) < TOL32
    assert res[0][0] == res[1][0] and nerr(res[1][2], res[0][2]) < 2e-6
    # far pairs
    far = structures.icosahedron('Au', 2)
    far.positions[30:] += [1000., 0., 0.]
    ideal = structures.icosahedron('Au', 2)
    be, target = _fused_backend(far, ideal)
    pos = far.get_positions()
    out = {}
    for table in (0, 1):
        be.set_option('fused_table', table)
        out[table] = be.energy_forces(pos, target, 'rw', 1.)[:3]
        again = be.energy_forces(pos, target, 'rw', 1.)[:3]
        assert np.array_equal(np.asarray(again[2]), np.asarray(out[table][2]))
    assert out[0][0] == out[1][0] and nerr(out[1][2], out[0][2]) < 2e-6
    be.set_option('fused_table', 1)
    # three element types: direct pass inside the fused launch
    base = structures.alloy_sphere(120, seed=5)
    numbers = np.array(base.get_atomic_numbers())
    numbers[::3] = 47
    tri = ase_shim.Atoms(numbers=numbers, positions=base.get_positions())
    ideal = tri.copy()
    tri.positions *= 1.02
    pos = tri.get_positions()
    be64, target = _fused_backend(tri, ideal, 'fp64')
    f64 = be64.energy_forces(pos, target, 'rw', 1.)[2]
    be, target = _fused_backend(tri, ideal)
    be.energy_forces(pos, target, 'rw', 1.)
    n0 = be.launch_count()
    f32 = be.energy_forces(pos, target, 'rw', 1.)[2]
    assert be.launch_count() - n0 == 1 and nerr(f32, f64) < TOL32


def test_fused_histogram_phase_against_direct_phase_and_fp64():
    """The fused launch sums F(Q) through a radial pair histogram in shared
    memory (iid_fused.cuh phase 1; the reference's pair sum is
    pyiid/experiments/elasticscatter/kernels/cpu_flat.py:70-91).  Against the
    direct pair pass of the same launch and the FP64 handle: one and two
    elements, perturbed and perfectly symmetric positions (many pairs on one
    node), bit-identical when repeated and whatever was evaluated before; a
    structure whose bounding box exceeds the histogram keeps the direct pass."""
    rs = np.random.RandomState(1)
    cases = []
    for atoms in (structures.icosahedron('Au', 3), structures.alloy_sphere(400, seed=3)):
        ideal = atoms.copy()
        atoms.positions = atoms.positions * 1.03 + rs.normal(0, 0.03, atoms.positions.shape)
        cases.append((atoms, ideal))
    sym = structures.icosahedron('Au', 4)
    sym.positions *= 1.05
    cases.append((sym, structures.icosahedron('Au', 4)))
    for atoms, ideal in cases:
        pos = atoms.get_positions()
        be64, target = _fused_backend(atoms, ideal, 'fp64')
        e64, s64, f64 = be64.energy_forces(pos, target, 'rw', 1.)[:3]
        be, target = _fused_backend(atoms, ideal)
        res = {}
        for hist in (0, 1, 0, 1):
            be.set_option('fused_hist', hist)
            be.energy_forces(pos + 0.01, target, 'rw', 1.)  # another history each time
            n0 = be.launch_count()
            out = be.energy_forces(pos, target, 'rw', 1.)[:3]
            assert be.launch_count() - n0 == 1
            out = (out[0], out[1], np.array(out[2]))
            if hist in res:
                assert out[0] == res[hist][0] and out[1] == res[hist][1]
                assert np.array_equal(out[2], res[hist][2])
            res[hist] = out
            assert abs(out[0] - e64) < 1e-6 * abs(e64) and abs(out[1] - s64) < 1e-6 * abs(s64)
            assert nerr(out[2], f64) < TOL32
        assert abs(res[1][0] - res[0][0]) < 1e-7 * abs(res[0][0])
        assert nerr(res[1][2], res[0][2]) < 2e-6
    # two clusters 1000 A apart: beyond the histogram either way, same bits
    far = structures.icosahedron('Au', 2)
    far.positions[30:] += [1000., 0., 0.]
    be, target = _fused_backend(far, structures.icosahedron('Au', 2))
    pos = far.get_positions()
    out = {}
    for hist in (0, 1):
        be.set_option('fused_hist', hist)
        out[hist] = be.energy_forces(pos, target, 'rw', 1.)[:3]
    assert out[0][0] == out[1][0] and np.array_equal(np.array(out[0][2]), np.array(out[1][2]))


def test_sq_iq_follow_the_reference_formulas():
    """get_sq = F/Q + 1 (inf -> 0), get_iq = S * <f>^2 (__init__.py:393-446)."""
    atoms = structures.alloy_sphere(40, seed=9)
    scat = ElasticScatter()
    fq = scat.get_fq(atoms).astype(np.float64)
    q = scat.get_scatter_vector()
    sq = scat.get_sq(atoms)
    assert np.isnan(sq[0]) and np.allclose(sq[1:], fq[1:] / q[1:] + 1, rtol=1e-6)
    f2 = np.average(atoms.get_array('F(Q) scatter'), axis=0) ** 2
    assert np.allclose(scat.get_iq(atoms)[1:], sq[1:] * f2[1:], rtol=1e-6)


def test_per_atom_scatter_rows_that_differ_within_an_element():
    """A caller may hand in arbitrary per-atom scatter-factor rows; they are
    grouped by unique row, not by atomic number."""
    rs = np.random.RandomState(5)
    atoms = structures.random_atoms(23, 8)
    scat = ElasticScatter(precision='fp64')
    scat._ensure_wrapped(atoms)
    sf = atoms.get_array('F(Q) scatter')
    scale = rs.choice([0.5, 1.0, 1.7], size=23).astype(np.float32)
    sf = sf * scale[:, None]
    atoms.set_array('F(Q) scatter', sf)
    fq, grad = scat.get_fq(atoms), scat.get_grad_fq(atoms)
    pos = atoms.get_positions()
    assert nerr(fq, oracle.experiment_fq(pos, sf, scat.exp, 'fp64')) < TOL64
    assert nerr(grad, oracle.experiment_grad_fq(pos, sf, scat.exp, 'fp64')) < TOL64


def test_noise_is_reproducible_with_a_seed():
    atoms = structures.random_atoms(15, 2)
    a = ElasticScatter(seed=7).get_fq(atoms, iq_std=0.05)
    b = ElasticScatter(seed=7).get_fq(atoms, iq_std=0.05)
    c = ElasticScatter(seed=8).get_fq(atoms, iq_std=0.05)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_fp32_f_of_q_against_oracle_at_2000_atoms():
    atoms = structures.fcc_sphere('Au', 2000)
    scat = ElasticScatter()
    fq = scat.get_fq(atoms)
    pos = atoms.get_positions().astype(np.float32)
    sf = atoms.get_array('F(Q) scatter')
    assert nerr(fq, oracle.experiment_fq(pos, sf, EXP, 'fp32', nthreads=8)) < TOL32
    assert nerr(fq, oracle.experiment_fq(pos, sf, EXP, 'fp64', nthreads=8)) < TOL32


def test_nuts_trajectory_equals_the_reference_samplers():
    """The samples of the REFERENCE's NUTSCanonicalEnsemble source driving a
    float64 oracle calculator (tests/golden/nuts_au55.npz, generated in the
    build container by tests/golden/make_golden_nuts.py from
    pyiid/sim/nuts_hmc.py:91-244 + pyiid/sim/__init__.py:10-38 as they lie in
    the mount) against pyiid_b200.sim's NUTS on the B200 Calc1D, FP64 mode, same
    seeds: Atoms-level path, array-level path and device-resident states."""
    k = golden('nuts_au55')

    def run(**kw):
        atoms = structures.icosahedron('Au', 2)
        atoms.set_positions(k['start_positions'])
        scat = ElasticScatter(precision='fp64')
        calc = Calc1D(target_data=k['target'], exp_function=scat.get_pdf,
                      exp_grad_function=scat.get_grad_pdf, conv=float(k['conv']),
                      potential='rw')
        atoms.set_calculator(calc)

        class Ensemble(sim.NUTSCanonicalEnsemble):
            def _find_step_size(self, input_atoms, thermal_nrg=None, momentum=None):
                return float(k['step'])  # the override the golden run needed

        np.random.seed(int(k['np_seed']))
        ens = Ensemble(atoms, temperature=float(k['temperature']),
                       escape_level=int(k['escape_level']), seed=int(k['seed']), **kw)
        traj, meta = ens.run(int(k['iterations']))
        return traj, meta, ens

    assert np.allclose(structures.icosahedron('Au', 2).get_masses(), k['masses'])
    for kw in (dict(fast=False), dict(fast=True, device_states=False),
               dict(fast=True, device_states=True)):
        traj, meta, ens = run(**kw)
        assert len(traj) == len(k['traj_positions']), kw
        assert meta['accepted_samples'] == int(k['accepted_samples'])
        assert meta['samples_total'] == int(k['samples_total'])
        for a, q, p, e in zip(traj, k['traj_positions'], k['traj_momenta'], k['traj_energy']):
            assert np.abs(a.get_positions() - q).max() < 1e-6, kw
            assert np.abs(a.get_momenta() - p).max() < 1e-6 * max(1., np.abs(p).max()), kw
            assert abs(a.get_potential_energy() - e) < 1e-7 * abs(e), kw
        assert abs(ens.step_size - float(k['final_step_size'])) < 1e-6 * ens.step_size


def test_seeded_nuts_run_reproduces_bit_for_bit():
    """pyiid/tests/test_consistancy.py:8-16 holds every result to run-to-run
    equality; a seeded sampler run on the device must reproduce as well (one
    fused launch per leapfrog, device-resident states)."""
    start, scat = make_hmc_atoms(3)
    target = start.calc.target_data  # one target for both runs

    def run():
        atoms = start.copy()
        atoms.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf,
                                    exp_grad_function=scat.get_grad_pdf, conv=100,
                                    potential='rw'))
        np.random.seed(4)
        ens = sim.NUTSCanonicalEnsemble(atoms, temperature=600, escape_level=5, seed=11)
        traj, meta = ens.run(4)
        return (np.array([a.get_positions() for a in traj]),
                np.array([a.get_potential_energy() for a in traj]), meta['samples_total'],
                ens.step_size)

    (q1, e1, n1, s1), (q2, e2, n2, s2) = run(), run()
    assert n1 == n2 and s1 == s2 and np.array_equal(q1, q2) and np.array_equal(e1, e2)
    # the evaluation itself: same bits every time, energy and forces
    be = scat.pdf_backend
    pos = start.get_positions()
    ref = be.energy_forces(pos, target, 'rw', 100.)
    for _ in range(5):
        e, sc, f, _ = be.energy_forces(pos, target, 'rw', 100.)
        assert e == ref[0] and sc == ref[1] and np.array_equal(f, ref[2])


def test_pinned_output_pool_limit_falls_back_to_pageable_memory():
    """When the pinned pool is exhausted the gradient goes through the staged
    download into pageable memory -- same bits."""
    from pyiid_b200 import hostmem
    import gc
    atoms = structures.alloy_sphere(300, seed=6)
    scat = ElasticScatter()
    g1 = scat.get_grad_fq(atoms)
    assert hostmem.is_pinned(g1)
    limit = hostmem.POOL_LIMIT_BYTES
    keep = []
    try:
        gc.collect()
        hostmem._pool.trim()                             # idle buffers of earlier tests
        hostmem.POOL_LIMIT_BYTES = hostmem._pool.total  # nothing more may be allocated
        keep = [scat.get_grad_fq(atoms) for _ in range(3)]
    finally:
        hostmem.POOL_LIMIT_BYTES = limit
    assert any(not hostmem.is_pinned(g) for g in keep)
    assert all(np.array_equal(g, g1) for g in keep)


def test_array_level_nuts_equals_atoms_level_nuts():
    """The fast sampler path (plain arrays, one native call per leapfrog)
    reproduces the Atoms-level path: same random numbers, same trajectory."""
    trajs = []
    for fast in (False, True):
        atoms, scat = make_hmc_atoms(2, 'fp64')
        np.random.seed(3)
        ens = sim.NUTSCanonicalEnsemble(atoms, temperature=1000, escape_level=4, seed=5,
                                        fast=fast)
        assert ens.fast is fast
        traj, meta = ens.run(4)
        trajs.append((traj, dict(meta), ens.step_size, ens.leapfrogs))
    (ta, ma, sa, la), (tb, mb, sb, lb) = trajs
    assert ma == mb and la == lb and len(ta) == len(tb)
    assert abs(sa - sb) < 1e-6 * abs(sa), (sa, sb)
    for x, y in zip(ta, tb):
        # identical up to the chaotic growth of last-bit differences (the
        # device sums use atomics, so even one path is only reproducible to
        # ~1e-16 per evaluation)
        assert np.allclose(x.positions, y.positions, rtol=0, atol=1e-5), \
            np.abs(x.positions - y.positions).max()
        assert np.allclose(x.get_momenta(), y.get_momenta(), rtol=0, atol=1e-5)
        assert abs(x.get_potential_energy() - y.get_potential_energy()) < 1e-5


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_device_resident_leapfrog_equals_array_level_leapfrog(precision):
    """iid_leapfrog_host (kick, drift, fused energy + forces, kick, centring on
    the device, states in device slots) against the host arithmetic of the
    array-level path (pyiid/sim/__init__.py:10-38): positions and momenta bit
    for bit, energies to the summation order; eager calls, graph capture and
    replays; forward then backward returns to the start."""
    atoms, scat = make_hmc_atoms(2, precision)
    atoms.set_momenta(np.random.RandomState(0).normal(0, 1, (55, 3)))
    atoms.get_forces()
    host = sim._FastSystem(atoms)
    dev = sim._DeviceSystem(atoms)
    free0 = len(dev.pool.free)
    sh, sd = host.state_of(atoms), dev.state_of(atoms)
    assert sd.slot is not None and len(dev.pool.free) == free0 - 1
    tol = 1e-12 if precision == 'fp64' else 2e-6  # atomics order; FP32 pair sums
    for k, (step, centre) in enumerate([(0.05, True), (0.05, True), (0.02, False),
                                        (-0.03, True), (0.05, True), (0.01, False)]):
        nh, nd = host.leapfrog(sh, step, centre), dev.leapfrog(sd, step, centre)
        assert np.abs(nd.q - nh.q).max() <= tol * np.abs(nh.q).max(), k
        assert np.abs(nd.p - nh.p).max() <= tol * np.abs(nh.p).max(), k
        assert abs(nd.pe - nh.pe) <= tol * abs(nh.pe), k
        assert abs(nd.ke - nh.ke) <= 1e-12 * abs(nh.ke) + tol * abs(nh.ke), k
        if precision == 'fp64' and k == 0:
            # same operation order as numpy: the integrator itself adds no error
            # beyond the force evaluation (atomics: last bits)
            assert np.abs(nd.q - nh.q).max() < 1e-13
        # the device state carries its forces: the next step starts from them
        f = dev.be.state_download(nd.slot, want=('f',))['f']
        assert np.abs(f - nh.f).max() <= tol * np.abs(nh.f).max(), k
        sh, sd = nh, nd
    # accepted samples: an Atoms object with cached energy and forces
    a = dev.to_atoms(sd)
    assert np.array_equal(a.get_positions(), sd.q) and a.get_potential_energy() == sd.pe
    assert np.abs(a.get_forces() - sh.f).max() <= tol * np.abs(sh.f).max()
    # reversibility on the device (tests/test_sim/test_leapfrog.py:58-73)
    start = dev.state_of(atoms)
    fwd = dev.leapfrog(start, 0.05, False)
    back = dev.leapfrog(fwd, -0.05, False)
    assert np.abs(back.q - start.q).max() < (1e-9 if precision == 'fp64' else 1e-5)
    assert np.abs(back.p - start.p).max() < (1e-9 if precision == 'fp64' else 1e-4)
    # slots return to the pool with their states
    del nh, nd, sd, a, start, fwd, back
    import gc
    gc.collect()
    assert len(dev.pool.free) == free0


def test_chained_leapfrogs_equal_single_steps_bit_for_bit():
    """iid_leapfrog_chain_host: k steps walked inside ONE cooperative launch
    (and, with chain_in_kernel off, k launches behind one synchronisation)
    produce exactly the states of k single-step calls -- positions, momenta,
    energies and the forces left in the last slot."""
    atoms, scat = make_hmc_atoms(2, 'fp32')
    atoms.set_momenta(np.random.RandomState(4).normal(0, 1, (55, 3)))
    atoms.get_forces()
    dev = sim._DeviceSystem(atoms)
    be, calc = dev.be, dev.calc
    start = dev.state_of(atoms)
    args = (calc.target_data, calc.potential_name, calc.rw_to_eV)
    for step in (0.04, -0.02):
        single, src = [], start.slot
        for k in range(5):
            single.append(be.leapfrog(src, 10 + k, step, True, *args))
            src = 10 + k
        f_single = be.state_download(14, want=('f',))['f']
        for in_kernel in (1, 0):
            be.set_option('chain_in_kernel', in_kernel)
            n0 = be.launch_count()
            for rep in range(3):  # eager, capture, replay
                chain = be.leapfrog_chain(start.slot, [20, 21, 22, 23, 24], step, True, *args)
            if in_kernel:
                assert be.launch_count() - n0 == 3  # one launch per chain of five
            assert len(chain) == 5
            for a, b in zip(single, chain):
                for x, y in zip(a, b):
                    assert np.array_equal(np.asarray(x), np.asarray(y))
            assert np.array_equal(be.state_download(24, want=('f',))['f'], f_single)
        be.set_option('chain_in_kernel', 1)
    # look-ahead in the sampler's system: the same states as the step-by-step walk
    ref, st = [], start
    for k in range(6):
        st = dev.leapfrog(st, 0.03)
        ref.append(st)
    dev.expect(6)
    st = start
    n0 = be.launch_count()
    for k in range(6):
        st = dev.leapfrog(st, 0.03)
        assert np.array_equal(st.q, ref[k].q) and np.array_equal(st.p, ref[k].p)
        assert st.pe == ref[k].pe and st.ke == ref[k].ke
    assert be.launch_count() - n0 == 1
    # a different trajectory voids what was computed ahead
    dev.expect(4)
    a = dev.leapfrog(start, 0.03)
    b = dev.leapfrog(start, -0.03)
    assert np.array_equal(a.q, ref[0].q) and not np.array_equal(b.q, ref[0].q)
    # the two halves of the native call: steps are handed out as they complete;
    # another call on the handle waits for the chain and drops the rest
    from pyiid_b200 import _lib
    cid = be.leapfrog_chain_begin(start.slot, [30, 31, 32, 33], 0.03, True, *args)
    one = be.leapfrog_chain_next(cid)
    two = be.leapfrog_chain_next(cid)
    assert np.array_equal(one[4], ref[0].q) and np.array_equal(two[4], ref[1].q)
    assert np.array_equal(two[5], ref[1].p) and float(two[0]) == ref[1].pe
    got = be.state_download(33)  # (the chain ran to its end before this call)
    assert np.array_equal(got['q'], ref[3].q) and np.array_equal(got['p'], ref[3].p)
    with pytest.raises(_lib.ChainDropped):
        be.leapfrog_chain_next(cid)
    # the sampler's system recovers from a chain dropped behind its back
    dev.expect(5)
    st = dev.leapfrog(start, 0.03)
    be.state_download(st.slot, want=('f',))
    for k in range(1, 5):
        st = dev.leapfrog(st, 0.03)
        assert np.array_equal(st.q, ref[k].q) and np.array_equal(st.p, ref[k].p)


def test_nuts_returns_every_device_slot():
    """Steps enqueued ahead of a subtree that stops early (U-turn, divergence,
    end of the iteration) hold device slots; every one of them is back in the
    pool after the iteration, over many iterations."""
    import gc
    atoms, scat = make_hmc_atoms(2, 'fp32')
    np.random.seed(2)
    ens = sim.NUTSCanonicalEnsemble(atoms, temperature=1000, escape_level=5, seed=4, fast=True,
                                    device_states=True)
    ens.run(25)
    gc.collect()
    pool = scat.pdf_backend._slot_pool
    assert len(pool.free) == sim._DeviceSystem.N_SLOTS and len(set(pool.free)) == len(pool.free)
    assert ens.leapfrogs > 100


def test_device_state_nuts_equals_array_level_nuts():
    """NUTS with the tree's states resident on the device draws the same random
    numbers and follows the same trajectory as the array-level path."""
    runs = []
    for dev in (False, True):
        atoms, scat = make_hmc_atoms(2, 'fp64')
        np.random.seed(3)
        ens = sim.NUTSCanonicalEnsemble(atoms, temperature=1000, escape_level=4, seed=5,
                                        fast=True, device_states=dev)
        assert ens.fast and ens.device_states is dev
        traj, meta = ens.run(5)
        runs.append((traj, dict(meta), ens.step_size, ens.leapfrogs))
    (ta, ma, sa, la), (tb, mb, sb, lb) = runs
    assert ma == mb and la == lb and len(ta) == len(tb) and ma['samples_total'] > 0
    assert abs(sa - sb) < 1e-6 * abs(sa), (sa, sb)
    for x, y in zip(ta, tb):
        assert np.allclose(x.positions, y.positions, rtol=0, atol=1e-5)
        assert np.allclose(x.get_momenta(), y.get_momenta(), rtol=0, atol=1e-5)
        assert abs(x.get_potential_energy() - y.get_potential_energy()) < 1e-5
        assert np.allclose(x.get_forces(), y.get_forces(), rtol=0, atol=1e-4)


def test_classical_dynamics_on_device_states_equals_atoms_level_leapfrogs():
    """pyiid/sim/dynamics.py on the fused calculator: frames from the
    device-resident states equal repeated Atoms-level leapfrog steps."""
    atoms, scat = make_hmc_atoms(2, 'fp64')
    atoms.set_momenta(np.random.RandomState(1).normal(0, 1, (55, 3)))
    traj = sim.classical_dynamics(atoms, 0.02, 4)
    assert len(traj) == 5 and traj[0] is atoms
    ref = atoms
    for frame in traj[1:]:
        ref = sim.leapfrog(ref, 0.02)
        assert np.allclose(frame.positions, ref.positions, rtol=0, atol=1e-9)
        assert np.allclose(frame.get_momenta(), ref.get_momenta(), rtol=0, atol=1e-9)
        assert abs(frame.get_potential_energy() - ref.get_potential_energy()) < 1e-9
        assert np.allclose(frame.get_forces(), ref.get_forces(), rtol=0, atol=1e-8)


def test_example_workflow_and_coincident_atoms():
    """examples/au_np_pdf.py (the reference's Au_NP_PDF.py flow through the
    `pyiid` import paths) runs; coincident atoms contribute 0 instead of NaN."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location(
        'au_np_pdf', os.path.join(ROOT, 'examples', 'au_np_pdf.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    e0, pe, meta = mod.main(3)
    assert e0 > 0 and np.all(np.isfinite(pe)) and meta['samples_total'] > 0
    atoms = structures.random_atoms(12, 1)
    atoms.positions[5] = atoms.positions[2]
    scat = ElasticScatter(precision='fp64')
    fq, g = scat.get_fq(atoms), scat.get_grad_fq(atoms)
    assert np.all(np.isfinite(fq)) and np.all(np.isfinite(g))
    keep = [i for i in range(12) if i != 5]
    # the duplicated pair adds nothing: remaining pair terms equal those of
    # the 12-atom sum with that one pair removed
    assert np.any(fq)


def test_50k_bench_workload_properties():
    """Maximum-size case of the bench: F(Q) of the Pt 50 000-atom particle is
    finite, translation invariant and equal on the two Q grids' shared bin 0."""
    atoms = structures.fcc_sphere('Pt', 50000)
    scat = ElasticScatter()
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.exp['qbin'], 'fq')
    pos = atoms.get_positions()
    f1 = be.fq(pos)
    f2 = be.fq(pos + np.array([11.0, -7.5, 2.25]))
    assert np.all(np.isfinite(f1)) and f1[0] == 0
    assert nerr(f2, f1) < TOL32


def test_tabulated_force_pass_equals_direct_force_pass():
    """FP32 mode sums over the Q bins first (radial table + one interpolation
    per pair) for large structures; it must reproduce the direct O(N^2 Q)
    force kernel and the oracle."""
    atoms = structures.alloy_sphere(2000, seed=3)
    ideal = structures.alloy_sphere(2000, seed=3, sigma=0.0)
    scat = ElasticScatter(precision='fp32')
    target = scat.get_pdf(ideal)
    scat._ensure_wrapped(atoms)
    be = scat.pdf_backend
    pos = atoms.get_positions()
    res = {}
    for table in (1, 0):
        be.set_option('force_table', table)
        be.set_option('force_table_min_n', 2)
        res[table] = be.energy_forces(pos, target, 'rw', 100.)
    be.set_option('force_table', 1)
    be.set_option('force_table_min_n', 600)
    (e1, s1, f1, _), (e0, s0, f0, _) = res[1], res[0]
    # the F(Q) pass is the same kernel both times (atomic order: last bits only)
    assert abs(e1 - e0) < 1e-12 * abs(e0) and abs(s1 - s0) < 1e-12 * abs(s0)
    assert nerr(f1, f0) < 2e-6
    # against the float64 mode (itself pinned to the oracle)
    s64 = ElasticScatter(precision='fp64')
    t64 = s64.get_pdf(ideal)
    s64._ensure_wrapped(atoms)
    e64, _, f64, _ = s64.pdf_backend.energy_forces(pos.astype(np.float32).astype(np.float64),
                                                   t64, 'rw', 100.)
    assert abs(e1 - e64) < TOL32 * abs(e64)
    assert nerr(f1, f64) < TOL32
    # small structure: the table also agrees with the oracle
    g = golden('aupt37_alloy')
    sc = ElasticScatter(precision='fp32')
    a = wrapped(sc, atoms_from(g), g)
    bp = sc._load(a, sc.pdf_qbin, 'PDF')
    bp.set_transform(sc.exp['rstep'], sc.pdf_qbin, sc.get_r(), 0.0)
    bp.set_option('force_table_min_n', 2)
    e, scale, f, _ = bp.energy_forces(g['positions'], g['target_pdf_f32'], 'rw', 1.0)
    bp.set_option('force_table_min_n', 600)
    assert nerr(f, g['rw_forces_f32']) < TOL32
    # an extended structure: two clusters 700 A apart (beyond the 524 A the
    # table covers at its fixed step): the distant pairs are summed directly
    # over the Q bins, the grid step is never widened
    far = structures.alloy_sphere(120, seed=8)
    pos2 = np.vstack([far.get_positions(), far.get_positions()[::-1] + np.array([700., 0, 0])])
    two = ase_shim.Atoms(numbers=np.concatenate([far.numbers, far.numbers[::-1]]),
                         positions=pos2)
    s2 = ElasticScatter(precision='fp32')
    s2._ensure_wrapped(two)
    b2 = s2._load(two, s2.pdf_qbin, 'PDF')
    b2.set_transform(s2.exp['rstep'], s2.pdf_qbin, s2.get_r(), 0.0)
    t2 = b2.pdf(pos2 * 1.01)
    out = {}
    for table in (1, 0):
        b2.set_option('force_table', table)
        b2.set_option('force_table_min_n', 2)
        out[table] = b2.energy_forces(pos2, t2, 'rw', 100.)
    b2.set_option('force_table', 1)
    b2.set_option('force_table_min_n', 600)
    assert nerr(out[1][2], out[0][2]) < 2e-6


def test_two_scatter_objects_keep_their_own_handles():
    """Two ElasticScatter objects with different structures / experiments do
    not share (and ping-pong) one native handle."""
    a1, a2 = structures.random_atoms(30, 1), structures.random_atoms(45, 2)
    s1, s2 = ElasticScatter(), ElasticScatter({'qmax': 20., 'rmax': 30.})
    f1, f2 = s1.get_fq(a1), s2.get_fq(a2)
    assert s1.backend is not s2.backend and s1.pdf_backend is not s2.pdf_backend
    c0 = s1.backend.launch_count()
    for _ in range(3):
        assert np.array_equal(s1.get_fq(a1), f1) and np.array_equal(s2.get_fq(a2), f2)
    # per call: staging + F(Q) pass + fixed-order sum of the item partials +
    # finish, no re-upload in between
    assert s1.backend.launch_count() - c0 == 12
    assert s1.backend.sizes()['n'] == 30 and s2.backend.sizes()['n'] == 45
    g = s1.get_grad_pdf(structures.random_atoms(1, 0))
    assert g.shape == (1, 3, 4000) and not np.any(g)


@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_gradient_only_items_equal_full_items(precision):
    """The full-gradient pass sums F(Q) over the items below the diagonal only
    and runs the items above it with a shorter bin loop (coefficient
    a_k = k C_k + t_k in a basis without the running m*kappa).  It must
    reproduce the pass that treats every item alike (grad_split = 0), F(Q) of
    the triangle pass, and the float64 mode."""
    atoms = structures.alloy_sphere(700, seed=5)
    scat = ElasticScatter(precision=precision)
    scat._ensure_wrapped(atoms)
    be = scat._load(atoms, scat.exp['qbin'], 'fq')
    pos = atoms.get_positions()
    res = {}
    for split in (1, 0):
        be.set_option('grad_split', split)
        res[split] = be.grad_fq(pos, with_fq=True)
    be.set_option('grad_split', 1)
    (g1, f1), (g0, f0) = res[1], res[0]
    tol = 1e-10 if precision == 'fp64' else 2e-6
    assert nerr(f1, f0) < tol and nerr(g1, g0) < tol
    assert nerr(f1, be.fq(pos)) < tol
    if precision == 'fp32':
        s64 = ElasticScatter(precision='fp64')
        s64._ensure_wrapped(atoms)
        b64 = s64._load(atoms, s64.exp['qbin'], 'fq')
        g64, f64 = b64.grad_fq(pos.astype(np.float32).astype(np.float64), with_fq=True)
        assert nerr(f1, f64) < TOL32 and nerr(g1, g64) < TOL32


@pytest.mark.parametrize('potential', ['rw', 'chi_sq'])
@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_qspace_chain_rule_weights_equal_rspace(potential, precision):
    """The fused host path forms wq = conv T^T c from T^T T and T^T target
    (no pass over the R x Q matrix per evaluation); it must give the forces of
    the r-space contraction, including after a change of target."""
    g = golden('au55_ico')
    sc = ElasticScatter(precision=precision)
    a = wrapped(sc, atoms_from(g), g)
    be = sc._load(a, sc.pdf_qbin, 'PDF')
    be.set_transform(sc.exp['rstep'], sc.pdf_qbin, sc.get_r(), 0.0)
    pos = g['positions'] * 1.02
    targets = [g['target_pdf_f32'], 0.5 * g['target_pdf_f32'][::-1].copy()]
    # like with like: both sides take the direct float32 force pass (the fused
    # launch's float64 radial table differs from it by the direct pass's own
    # rounding, checked below)
    be.set_option('fused_table', 0)
    be.set_option('fused_hist', 0)  # (and the direct float32 F(Q) pass on both sides)
    for tg in targets:
        res = {}
        for q in (1, 0):
            be.set_option('qspace_wq', q)
            for _ in range(4):  # eager calls, capture, graph replay
                res[q] = be.energy_forces(pos, tg, potential, 10.)
        be.set_option('qspace_wq', 1)
        (e1, s1, f1, _), (e0, s0, f0, _) = res[1], res[0]
        assert abs(e1 - e0) <= 1e-12 * abs(e0) and abs(s1 - s0) <= 1e-12 * abs(s0)
        assert nerr(f1, f0) < (1e-9 if precision == 'fp64' else 2e-6)
        if precision == 'fp32':
            be.set_option('fused_table', 1)
            e2, s2, f2, _ = be.energy_forces(pos, tg, potential, 10.)
            be.set_option('fused_table', 0)
            assert e2 == e1 and s2 == s1 and nerr(f2, f1) < TOL32
    be.set_option('fused_table', 1)
    be.set_option('fused_hist', 1)
