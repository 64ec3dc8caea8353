"""Spring restraints (reference pyiid/calc/spring_calc.py, multi_calc.py).

CPU part: oracle/spring.py against the outputs of the reference's own
functions (tests/golden/springs.npz) and the reference tests' known answers.
GPU part: the CUDA pair kernels behind iid_spring_host / Spring / MultiCalc
against the oracle, the golden vectors and the reference tests' properties
(tests/test_calc/test_spring.py, test_spring_known_system.py,
test_multi_spring_known_system.py).
"""
import copy
import os

import numpy as np
import pytest

from oracle import spring as osp
from conftest import nerr

HERE = os.path.dirname(os.path.abspath(__file__))
SYSTEMS = ['au4_square', 'au4_small', 'au10_random', 'au55_ico', 'aupt37_alloy']
# reference tests/__init__.py:216-218
REF_KWARGS = [{'k': 100, 'rt': 5., 'sp_type': 'rep'},
              {'k': 100, 'rt': 1., 'sp_type': 'com'},
              {'k': 100, 'rt': 1., 'sp_type': 'att'}]


@pytest.fixture(scope='module')
def gold():
    g = np.load(os.path.join(HERE, 'golden', 'springs.npz'))
    kwargs = [(t, float(k), float(rt)) for t, k, rt in g['kwargs']]
    return g, kwargs


def oracle_all(pos, com, t, k, rt, precision='fp32'):
    if t == 'com':
        return (osp.com_energy(pos, com, k, rt, precision),
                osp.com_force(pos, com, k, rt, precision),
                osp.com_atomwise(pos, com, k, rt, precision))
    return (osp.pair_energy(pos, k, rt, t, precision),
            osp.pair_force(pos, k, rt, t, precision),
            osp.pair_atomwise(pos, k, rt, t, precision))


# ---------------------------------------------------------------- CPU: oracle
@pytest.mark.parametrize('name', SYSTEMS)
def test_oracle_equals_reference_outputs(gold, name):
    """Energy and forces bit for bit, the float32 atomwise sums too."""
    g, kwargs = gold
    pos, com = g[name + '/positions'], g[name + '/com']
    for i, (t, k, rt) in enumerate(kwargs):
        e, f, a = oracle_all(pos, com, t, k, rt)
        assert e == float(g['%s/%d/energy' % (name, i)])
        assert np.array_equal(f, g['%s/%d/forces' % (name, i)])
        assert np.array_equal(np.asarray(a, np.float64),
                              g['%s/%d/atomwise' % (name, i)])


def test_oracle_known_answers_of_the_reference_tests(gold):
    """test_spring_known_system.py: energy >= 100 on the Au4 square for every
    spring type, forces non-zero and central (no torque about the centre)."""
    g, _ = gold
    pos, com = g['au4_square/positions'], g['au4_square/com']
    for kw in REF_KWARGS:
        e, f, _ = oracle_all(pos, com, kw['sp_type'], kw['k'], kw['rt'])
        assert e >= 100
        for i in range(4):
            assert np.any(f[i])
            assert np.allclose(np.cross(pos[i] - com, f[i]), 0, atol=1e-4)


def test_oracle_voxel_energy_is_the_energy_of_adding_an_atom(gold):
    """test_spring.py:17-42 (rtol 2e-7 there): pins the voxel restatement,
    which cannot be run from the reference under this numpy."""
    g, _ = gold
    pos = g['au4_square/positions'] - g['au4_square/positions'].min(0) + 1.0
    shape = (5, 5, 2)
    for t, k, rt in (('rep', 100, 5.), ('att', 100, 1.)):
        vox = osp.voxel_energy(pos, k, rt, 1.0, shape, t)
        e0 = osp.pair_energy(pos, k, rt, t)
        want = np.zeros(shape)
        for i in range(shape[0]):
            for j in range(shape[1]):
                for l in range(shape[2]):
                    p2 = np.vstack([pos, [(i + .5), (j + .5), (l + .5)]])
                    want[i, j, l] = osp.pair_energy(p2, k, rt, t) - e0
        assert np.allclose(vox, want, rtol=2e-6, atol=1e-3)


def test_oracle_force_is_minus_half_the_energy_gradient(gold):
    """The ordered-pair energy counts every pair twice, the force once:
    force = -1/2 dE/dq (float64 restatement, central differences)."""
    g, _ = gold
    pos = g['au10_random/positions']
    for t, k, rt in (('rep', 10., 6.), ('att', 10., 2.)):
        f = osp.pair_force(pos, k, rt, t, 'fp64')
        h = 1e-6
        for i, w in ((0, 0), (3, 1), (7, 2)):
            p, m = pos.copy(), pos.copy()
            p[i, w] += h
            m[i, w] -= h
            de = (osp.pair_energy(p, k, rt, t, 'fp64') -
                  osp.pair_energy(m, k, rt, t, 'fp64')) / (2 * h)
            assert abs(f[i, w] + 0.5 * de) < 1e-6 * max(1., abs(de))


def test_calculators_import_under_the_reference_paths():
    from pyiid.calc.spring_calc import Spring, spring_nrg, att_spring_force  # noqa: F401
    from pyiid.calc.multi_calc import MultiCalc
    from pyiid_b200 import Calc1D, ElasticScatter
    s = Spring(k=100, rt=5., sp_type='rep')
    assert (s.k, s.rt, s.sp_type) == (100, 5., 'rep')
    scat = ElasticScatter()
    c = Calc1D(target_data=np.zeros(4000), exp_function=scat.get_pdf,
               exp_grad_function=scat.get_grad_pdf)
    # one fused Calc1D + rep/att springs is a single device sequence
    assert MultiCalc(calc_list=[c, s])._plan is not None
    assert MultiCalc(calc_list=[s, Spring(sp_type='att')])._plan is None
    assert MultiCalc(calc_list=[c, Spring(sp_type='com')])._plan is None
    m2 = copy.deepcopy(MultiCalc(calc_list=[c, s]))
    assert m2._plan is not None and m2.calc_list[0]._fused is scat


# ---------------------------------------------------------------- GPU: parity
def _atoms(pos, numbers=None):
    from pyiid_b200.ase_shim import Atoms
    return Atoms(numbers=[79] * len(pos) if numbers is None else numbers,
                 positions=pos)


@pytest.mark.gpu
@pytest.mark.parametrize('name', SYSTEMS)
def test_gpu_springs_match_reference_outputs(gold, name):
    """FP32 handles reproduce the reference's float32 pair arithmetic: each
    pair term is bit-exact, only the float64 summation order differs."""
    from pyiid_b200.backend import Backend
    g, kwargs = gold
    be = Backend.get('fp32', None, 'fq')
    pos, com = g[name + '/positions'], g[name + '/com']
    for i, (t, k, rt) in enumerate(kwargs):
        e, f, a = be.spring(pos, t, k, rt, com, True, True, True)
        eref = float(g['%s/%d/energy' % (name, i)])
        fref = g['%s/%d/forces' % (name, i)]
        aref = g['%s/%d/atomwise' % (name, i)]
        assert abs(e - eref) <= 1e-13 * max(1., abs(eref))
        assert np.abs(f - fref).max() <= 1e-12 * max(1., np.abs(fref).max())
        if t == 'com':
            assert abs(a.sum() - aref) <= 1e-12 * max(1., abs(aref))
        else:  # the reference sums this one in float32
            assert np.abs(a - aref).max() <= 2e-6 * max(1., np.abs(aref).max())


@pytest.mark.gpu
def test_gpu_springs_fp64_mode_and_ragged_sizes():
    from pyiid_b200.backend import Backend
    rs = np.random.RandomState(5)
    for n in (1, 2, 127, 128, 129, 700):
        pos = rs.random_sample((n, 3)) * (2.0 * n ** (1. / 3.))
        com = pos.mean(0)
        for prec in ('fp32', 'fp64'):
            be = Backend.get(prec, None, 'fq')
            for t, k, rt in (('rep', 10., 3.), ('att', 2.5, 2.), ('com', 4., 1.5)):
                e, f, a = be.spring(pos, t, k, rt, com, True, True, True)
                eo, fo, ao = oracle_all(pos, com, t, k, rt, prec)
                assert abs(e - eo) <= 1e-12 * max(1., abs(eo))
                assert np.abs(f - fo).max() <= 1e-11 * max(1., np.abs(fo).max())
                if t == 'com':
                    assert abs(a.sum() - ao) <= 1e-12 * max(1., abs(ao))
                else:
                    tol = 2e-6 if prec == 'fp32' else 1e-12
                    assert np.abs(a - ao).max() <= tol * max(1., np.abs(ao).max())
    # no atoms, coincident atoms (0/0 -> 0 as in the reference)
    be = Backend.get('fp32', None, 'fq')
    assert be.spring(np.zeros((0, 3)), 'rep', 1., 1.)[0] == 0.0
    pos = np.array([[0., 0, 0], [0, 0, 0], [1, 0, 0]])
    e, f, _ = be.spring(pos, 'rep', 10., 2., None, True, True)
    assert e == osp.pair_energy(pos, 10., 2., 'rep')
    assert np.array_equal(f, osp.pair_force(pos, 10., 2., 'rep'))


@pytest.mark.gpu
def test_gpu_large_structure_properties():
    """Full-size check through size-independent properties: total force of the
    pair springs vanishes, energy equals the oracle on a 10k structure."""
    from pyiid_b200 import structures
    from pyiid_b200.backend import Backend
    atoms = structures.fcc_sphere('Au', 10000)
    pos = atoms.get_positions()
    be = Backend.get('fp64', None, 'fq')
    for t, k, rt in (('rep', 10., 3.0), ('att', 1e-3, 30.0)):
        e, f, a = be.spring(pos, t, k, rt, None, True, True, True)
        assert np.abs(f.sum(0)).max() <= 1e-9 * np.abs(f).sum()
        # ordered-pair energy = -1/2 sum of the atomwise energies
        assert abs(e + 0.5 * a.sum()) <= 1e-11 * abs(e)
    e32, f32, _ = Backend.get('fp32', None, 'fq').spring(pos, 'rep', 10., 3.0, None,
                                                         True, True)
    e64, f64, _ = be.spring(pos, 'rep', 10., 3.0, None, True, True)
    assert abs(e32 - e64) < 1e-4 * abs(e64)
    assert nerr(f32, f64) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize('kw', REF_KWARGS)
def test_gpu_spring_known_system(kw):
    """test_spring_known_system.py:18-52 and
    test_multi_spring_known_system.py:21-49 on the product classes."""
    from pyiid_b200 import structures
    from pyiid.calc.spring_calc import Spring
    from pyiid.calc.multi_calc import MultiCalc
    atoms1, atoms2 = structures.atomic_square()
    atoms1.set_calculator(Spring(**kw))
    assert atoms1.get_potential_energy() >= 100
    forces = atoms1.get_forces()
    com = atoms1.get_center_of_mass()
    for i in range(len(atoms1)):
        assert np.any(forces[i])
        assert np.allclose(np.cross(atoms1[i].position - com, forces[i]), 0, atol=1e-4)
    calc = MultiCalc(calc_list=[Spring(**kw), Spring(**kw)])
    a = structures.atomic_square()[0]
    a.set_calculator(calc)
    assert a.get_potential_energy() >= 100
    assert np.allclose(a.get_potential_energy(), 2 * atoms1.get_potential_energy())
    assert np.allclose(a.get_forces(), 2 * forces)
    com = a.get_center_of_mass()
    for i in range(len(a)):
        assert np.allclose(np.cross(a[i].position - com, a.get_forces()[i]), 0, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('kw', [REF_KWARGS[0], REF_KWARGS[2]])
def test_gpu_voxel_and_atomwise_energy(kw):
    """test_spring.py:17-70: voxel energy = energy change on adding an atom at
    the voxel centre; atomwise energy = energy change on deleting the atom."""
    from pyiid_b200 import structures
    from pyiid_b200.ase_shim import Atom
    from pyiid.calc.spring_calc import Spring
    atoms = copy.deepcopy(structures.atomic_square()[0])
    resolution = 1.
    atoms.center(resolution)
    atoms.set_calculator(Spring(**kw))
    e0 = atoms.get_potential_energy()
    vox = atoms.calc.calculate_voxel_energy(atoms, resolution)
    want = np.zeros(vox.shape)
    im, jm, km = vox.shape
    assert (im, jm, km) == (5, 5, 2)
    for i in range(im):
        for j in range(jm):
            for k in range(km):
                a2 = copy.deepcopy(atoms)
                a2 += Atom('Au', ((i + .5) * resolution, (j + .5) * resolution,
                                  (k + .5) * resolution))
                want[i, j, k] = a2.get_potential_energy() - e0
    assert np.allclose(vox, want, rtol=2e-6, atol=1e-3)
    assert np.allclose(vox, osp.voxel_energy(atoms.get_positions(), kw['k'], kw['rt'],
                                             resolution, vox.shape, kw['sp_type']),
                       rtol=1e-12, atol=1e-9)
    aw = atoms.calc.calculate_atomwise_energy(atoms)
    want = np.zeros(len(atoms))
    for atom in atoms:
        a2 = copy.deepcopy(atoms)
        del a2[atom.index]
        want[atom.index] = a2.get_potential_energy() - e0
    assert np.allclose(aw, want, rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('precision', ['fp32', 'fp64'])
def test_gpu_multicalc_fused_equals_sum_of_calculators(precision):
    """Calc1D + rep/att springs evaluated as one device sequence (graph
    replay included) equals the calculators evaluated one by one."""
    from pyiid_b200 import ElasticScatter, Calc1D, structures
    from pyiid.calc.spring_calc import Spring
    from pyiid.calc.multi_calc import MultiCalc
    scat = ElasticScatter(precision=precision)
    ideal = structures.icosahedron('Au', 2)
    target = scat.get_pdf(ideal)
    atoms = structures.icosahedron('Au', 2)
    atoms.positions *= 1.04
    c1 = Calc1D(target_data=target, exp_function=scat.get_pdf,
                exp_grad_function=scat.get_grad_pdf, conv=30., potential='rw')
    s1 = Spring(k=20, rt=3.2, sp_type='rep', precision=precision)
    s2 = Spring(k=0.05, rt=8., sp_type='att', precision=precision)
    multi = MultiCalc(calc_list=[c1, s1, s2])
    assert multi._plan is not None
    e_parts, f_parts = 0.0, 0.0
    for c in (c1, s1, s2):
        a = copy.deepcopy(atoms)
        a.set_calculator(copy.deepcopy(c))
        e_parts += a.get_potential_energy()
        f_parts = f_parts + a.get_forces()
    for rep in range(4):   # eager, capture, replay
        a = copy.deepcopy(atoms)
        a.set_calculator(copy.deepcopy(multi))
        e, f = a.get_potential_energy(), a.get_forces()
        assert abs(e - e_parts) <= 1e-10 * abs(e_parts)
        assert nerr(f, f_parts) < 1e-9
    # the springs must be dropped again for a plain Calc1D evaluation
    a = copy.deepcopy(atoms)
    a.set_calculator(copy.deepcopy(c1))
    a2 = copy.deepcopy(atoms)
    a2.set_calculator(Spring(k=20, rt=3.2, sp_type='rep', precision=precision))
    assert abs(a.get_potential_energy() + a2.get_potential_energy() +
               0.0 - (e_parts - s2.get_potential_energy(atoms))) <= 1e-9 * abs(e_parts)
    # a list that cannot be fused takes the reference's loop
    loop = MultiCalc(calc_list=[c1, Spring(k=3, rt=4., sp_type='com',
                                           precision=precision)])
    assert loop._plan is None
    a = copy.deepcopy(atoms)
    a.set_calculator(loop)
    b = copy.deepcopy(atoms)
    b.set_calculator(Spring(k=3, rt=4., sp_type='com', precision=precision))
    c = copy.deepcopy(atoms)
    c.set_calculator(copy.deepcopy(c1))
    assert abs(a.get_potential_energy() - b.get_potential_energy() -
               c.get_potential_energy()) < 1e-9
    assert nerr(a.get_forces(), b.get_forces() + c.get_forces()) < 1e-9


@pytest.mark.gpu
def test_gpu_nuts_on_a_restrained_potential_fast_equals_atoms_level():
    """The sampler on Calc1D + repulsive spring (the refinement set-up of the
    reference's examples): the array-level path, which evaluates the fused
    MultiCalc sequence once per leapfrog, reproduces the Atoms-level path."""
    from pyiid_b200 import ElasticScatter, Calc1D, structures, sim
    from pyiid.calc.spring_calc import Spring
    from pyiid.calc.multi_calc import MultiCalc
    runs = []
    for fast in (False, True):
        scat = ElasticScatter(precision='fp64')
        ideal = structures.icosahedron('Au', 2)
        target = scat.get_pdf(ideal)
        atoms = structures.icosahedron('Au', 2)
        atoms.positions *= 1.05
        calc = MultiCalc(calc_list=[
            Calc1D(target_data=target, exp_function=scat.get_pdf,
                   exp_grad_function=scat.get_grad_pdf, conv=100., potential='rw'),
            Spring(k=200, rt=2.9, sp_type='rep', precision='fp64')])
        atoms.set_calculator(calc)
        np.random.seed(3)
        ens = sim.NUTSCanonicalEnsemble(atoms, temperature=1000, escape_level=4,
                                        seed=5, fast=fast)
        assert ens.fast is fast
        traj, meta = ens.run(4)
        runs.append((traj, dict(meta), ens.step_size, ens.leapfrogs))
    (ta, ma, sa, la), (tb, mb, sb, lb) = runs
    assert ma == mb and la == lb and len(ta) == len(tb)
    assert ma['samples_total'] > 0
    assert abs(sa - sb) < 1e-6 * abs(sa)
    for x, y in zip(ta, tb):
        assert np.allclose(x.positions, y.positions, rtol=0, atol=1e-5)
        assert abs(x.get_potential_energy() - y.get_potential_energy()) < 1e-5
