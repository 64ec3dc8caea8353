"""CPU tests of the host layer: experiment bookkeeping, scatter-factor cache,
the F(Q)->G(r) matrix, form factors, structure builders, ASE stand-ins and the
sampler logic (on a CPU toy calculator -- no elastic-scattering compute here)."""
import copy

import numpy as np
import pytest

import oracle
from conftest import golden, nerr
from pyiid_b200 import ElasticScatter, Calc1D, ase_shim, formfactors, structures
from pyiid_b200.backend import pdf_matrix, element_table
from pyiid_b200 import sim


def test_experiment_defaults_and_grids():
    """elasticscatter/__init__.py:86-89, 180-204, 526-558."""
    s = ElasticScatter()
    assert s.exp == dict(qmin=0.0, qmax=25, qbin=.1, rmin=0.0, rmax=40.0,
                         rstep=.01, sampling='full')
    assert abs(s.pdf_qbin - 0.0756865) < 1e-7
    assert len(s.get_scatter_vector()) == 250
    assert len(s.get_scatter_vector(pdf=True)) == 330
    assert len(s.get_r()) == 4000
    s.update_experiment({'sampling': 'ns', 'qmax': 20.})
    assert abs(s.exp['rstep'] - np.pi / 20.) < 1e-15
    with pytest.raises(ValueError):
        ElasticScatter(seed='a')
    assert s.set_processor() is True
    assert s.set_processor('Multi-GPU', 'flat') is True
    assert s.processor == 'B200' and s.alg == 'flat'
    assert s.set_processor('nonsense') is None


def test_wrap_atoms_state_follows_reference_semantics():
    """tests/test_scatter_state.py:11-71 (the cache checks need no device)."""
    s = ElasticScatter()
    atoms = structures.random_atoms(10, 0)
    assert s._check_wrap_atoms_state(atoms) is False
    assert s._check_wrap_atoms_state(atoms) is True
    assert atoms.arrays['F(Q) scatter'].shape == (10, 250)
    assert atoms.arrays['PDF scatter'].shape == (10, 330)
    assert atoms.arrays['F(Q) scatter'].dtype == np.float32
    assert atoms.info['exp'] == s.exp and atoms.info['scatter_atoms'] == 10
    atoms2 = atoms + ase_shim.Atom('Au', [0, 0, 0])
    assert len(atoms2) == 11
    assert s._check_wrap_atoms_state(atoms2) is False
    assert s._check_wrap_atoms_state(atoms2) is True
    atoms3 = copy.deepcopy(atoms)
    del atoms3[3]
    assert s._check_wrap_atoms_state(atoms3) is False
    s.update_experiment({'qmax': 20.})
    assert s.check_wrap_atoms_state(atoms) is False


def test_wrap_atoms_uses_one_row_per_element_on_the_right_grid():
    s = ElasticScatter()
    atoms = structures.alloy_sphere(40, seed=5)
    s._wrap_atoms(atoms)
    for name, qbin in (('F(Q) scatter', .1), ('PDF scatter', s.pdf_qbin)):
        arr = atoms.arrays[name]
        q = np.arange(arr.shape[1]) * qbin
        for z in (78, 79):
            rows = arr[atoms.numbers == z]
            assert np.all(rows == rows[0])
            assert np.allclose(rows[0], formfactors.form_factor(z, q), rtol=1e-6)


def test_deepcopy_shares_the_scatter_object():
    s = ElasticScatter()
    assert copy.deepcopy(s) is s


def test_carbon_form_factor_golden_vector():
    """tests/test_master/test_master_kernel.py:8-16 + c60_scat.txt."""
    k = golden('known_answers')
    assert bool(k['c60_all_rows_equal'])
    f = formfactors.form_factor(6, np.arange(250) * .1)
    assert np.allclose(f, k['c60_row'], rtol=1e-6)
    for z in formfactors.WK95:
        assert abs(formfactors.form_factor(z, np.zeros(1))[0] - z) < 0.045
    with pytest.raises(KeyError):
        formfactors.form_factor(120, np.zeros(1))


def test_pdf_matrix_reproduces_the_fft_path():
    """T @ F == get_pdf_at_qmin(F) for the reference-generated vectors and for
    odd experiments (qmin > 0, rmin > 0, 'ns' sampling)."""
    k = golden('known_answers')
    exp = oracle.DEFAULT_EXP
    t = pdf_matrix(330, exp['rstep'], float(oracle.pdf_qbin(exp)), oracle.r_grid(exp), 0.0)
    assert nerr(t.dot(k['random_fq']), k['random_fq_pdf']) < 1e-12
    qmin, rmin, rmax, rstep, pq = k['exp2_vals']
    rg = np.arange(rmin, rmax, rstep)
    t = pdf_matrix(len(k['random_fq2']), rstep, pq, rg, qmin)
    assert nerr(t.dot(k['random_fq2']), k['random_fq2_pdf']) < 1e-12
    rs = np.random.RandomState(3)
    for e in (dict(qmin=0.7, qmax=21.3, rmin=0.9, rmax=33.3, rstep=np.pi / 21.3),
              dict(qmin=0.0, qmax=19., rmin=2.5, rmax=50., rstep=.015)):
        pq = np.pi / (e['rmax'] + 12 * np.pi / e['qmax'])
        nq = int(np.floor(e['qmax'] / pq))
        rg = np.arange(e['rmin'], e['rmax'], e['rstep'])
        f = rs.normal(size=nq)
        ref = oracle.get_pdf_at_qmin(f.copy(), e['rstep'], pq, rg, e['qmin'])
        assert nerr(pdf_matrix(nq, e['rstep'], pq, rg, e['qmin']).dot(f), ref) < 1e-12


def test_element_table_grouping():
    rs = np.random.RandomState(0)
    numbers = np.array([79, 78, 79, 79, 78])
    rows = {78: rs.rand(7).astype(np.float32), 79: rs.rand(7).astype(np.float32)}
    scat = np.array([rows[z] for z in numbers])
    table, idx = element_table(scat, numbers)
    assert table.shape == (2, 7) and np.array_equal(table[idx], scat.astype(np.float64))
    scat[2, 3] += 1  # same Z, different row: falls back to unique rows
    table, idx = element_table(scat, numbers)
    assert table.shape == (3, 7) and np.array_equal(table[idx], scat.astype(np.float64))


def test_structures():
    assert [len(structures.icosahedron_positions(s)) for s in range(6)] == \
        [1, 13, 55, 147, 309, 561]
    p = structures.fcc_sphere_positions(500, structures.A_AU, sigma=0.0)
    d = np.linalg.norm(p[:, None] - p[None], axis=2)
    d[d == 0] = 9
    assert abs(d.min() - structures.A_AU / np.sqrt(2)) < 1e-9
    assert p.min() == 0.0
    a = structures.alloy_sphere(1000)
    assert 400 < (a.numbers == 79).sum() < 600
    s1, s2 = structures.atomic_square()
    assert np.allclose(s2.positions, s1.positions * .75)


def test_atoms_stand_in():
    a = ase_shim.Atoms('Au4', [[0, 0, 0], [3, 0, 0], [0, 3, 0], [3, 3, 0]])
    a.center()
    assert np.allclose(a.positions.mean(0), 0)
    assert list(a.numbers) == [79] * 4 and a.get_chemical_symbols()[0] == 'Au'
    a.set_momenta(np.ones((4, 3)))
    assert np.allclose(a.get_velocities(), 1 / 196.966569)
    b = copy.deepcopy(a)
    b.positions += 1
    assert not np.allclose(a.positions, b.positions)
    assert len(ase_shim.Atoms('AuPt2')) == 3
    ase_shim.MaxwellBoltzmannDistribution(a, temp=0.1, force_temp=True)
    assert abs(a.get_kinetic_energy() / 4 / 1.5 - 0.1) < 1e-12
    with pytest.raises(RuntimeError):
        a.get_forces()


class Harmonic(ase_shim.Calculator):
    """CPU toy calculator (0.5 k |x|^2) to exercise the sampler logic."""
    implemented_properties = ['energy', 'forces']
    ncalls = 0

    def calculate(self, atoms=None, properties=['energy'], system_changes=[]):
        ase_shim.Calculator.calculate(self, atoms, properties, system_changes)
        Harmonic.ncalls += 1
        self.results['energy'] = 0.5 * 2.0 * float((self.atoms.positions ** 2).sum())
        self.results['forces'] = -2.0 * self.atoms.positions


def test_leapfrog_properties():
    """tests/test_sim/test_leapfrog.py:25-73: no-force no-move, momentum
    move, time reversibility."""
    a = ase_shim.Atoms('Au4', [[0, 0, 0], [3, 0, 0], [0, 3, 0], [3, 3, 0]])
    a.center()
    a.set_calculator(Harmonic())
    a.set_momenta(np.ones((4, 3)))
    n0 = Harmonic.ncalls
    a.get_forces()
    b = sim.leapfrog(a, 0.1, False)
    assert Harmonic.ncalls == n0 + 2  # one evaluation per leapfrog (+ the first)
    c = sim.leapfrog(b, -0.1, False)
    assert np.allclose(c.positions, a.positions, atol=1e-12)
    assert np.allclose(c.get_momenta(), a.get_momenta(), atol=1e-12)
    assert abs(b.get_total_energy() - a.get_total_energy()) < 1e-2
    z = ase_shim.Atoms('Au2', [[0, 0, 0], [0, 0, 0]])
    z.set_calculator(Harmonic())
    assert np.allclose(sim.leapfrog(z, 1, False).positions, 0)


def test_classical_dynamics_conserves_energy():
    """pyiid/sim/dynamics.py: n leapfrog steps, n + 1 frames, start included;
    the Hamiltonian of the harmonic stand-in is conserved to O(step^2)."""
    from pyiid.sim.dynamics import classical_dynamics
    a = ase_shim.Atoms('Au3', np.random.RandomState(2).normal(size=(3, 3)))
    a.set_calculator(Harmonic())
    a.set_momenta(np.random.RandomState(3).normal(size=(3, 3)))
    traj = classical_dynamics(a, 0.05, 20)
    assert len(traj) == 21 and traj[0] is a
    # the first step re-centres the cluster in its cell (leapfrog's default),
    # which the toy well -- unlike Rw -- is not invariant to
    e = [t.get_total_energy() for t in traj[1:]]
    assert max(e) - min(e) < 1e-2 * abs(e[0])
    assert not np.allclose(traj[-1].positions, traj[0].positions)


def test_nuts_samples_a_harmonic_well():
    a = ase_shim.Atoms('Au3', np.random.RandomState(0).normal(size=(3, 3)))
    a.set_calculator(Harmonic())
    np.random.seed(0)
    ens = sim.NUTSCanonicalEnsemble(a, temperature=300, escape_level=6, seed=1)
    traj, meta = ens.run(25)
    assert meta['accepted_samples'] >= 5 and meta['samples_total'] > 0
    assert len(traj) == meta['accepted_samples'] + 1
    assert ens.step_size > 0 and np.isfinite(ens.step_size)
    pe = [t.get_potential_energy() for t in traj]
    assert np.mean(pe[len(pe) // 2:]) < pe[0]


def test_calc1d_argument_checks():
    s = ElasticScatter()
    with pytest.raises(NotImplementedError):
        Calc1D(target_data=np.zeros((2, 2)), exp_function=s.get_pdf,
               exp_grad_function=s.get_grad_pdf)
    with pytest.raises(NotImplementedError):
        Calc1D(target_data=np.zeros(4))
    with pytest.raises(NotImplementedError):
        Calc1D(target_data=np.zeros(4), exp_function=s.get_pdf,
               exp_grad_function=s.get_grad_pdf, potential='xyz')
    c = Calc1D(target_data=np.zeros(4000), exp_function=s.get_pdf,
               exp_grad_function=s.get_grad_pdf, conv=100)
    assert c._fused is s and c.rw_to_eV == 100
    c2 = Calc1D(target_data=np.zeros(250), exp_function=s.get_fq,
                exp_grad_function=s.get_grad_fq)
    assert c2._fused is None


def test_empty_and_single_atom_inputs_need_no_device():
    """k_max == 0 gives zeros (flat_multi_cpu_wrap.py:85-86); handled on the
    host, so it works without a GPU."""
    s = ElasticScatter()
    for n in (0, 1):
        atoms = ase_shim.Atoms(numbers=[79] * n, positions=np.zeros((n, 3)))
        fq = s.get_fq(atoms)
        assert fq.shape == (250,) and fq.dtype == np.float32 and not np.any(fq)
        g = s.get_grad_fq(atoms)
        assert g.shape == (n, 3, 250) and not np.any(g)
        pdf = s.get_pdf(atoms)
        assert pdf.shape == (4000,) and not np.any(pdf)


def test_sampler_state_slots_and_helpers():
    """Host side of the device-resident sampler: slots return to the pool with
    the last reference to their state, the U-turn test on cached velocities is
    the Atoms-level test, exp() guards keep the reference's conventions."""
    pool = sim._SlotPool(4)
    q = np.arange(6.).reshape(2, 3)
    a = sim._DevState(q, q + 1, 1.0, None, 2.0, pool.take(), pool)
    b = sim._DevState(q, q + 1, 1.0, None, 2.0, pool.take(), pool)
    assert (a.slot, b.slot) == (0, 1) and len(pool.free) == 2 and a.total == 3.0
    alias = a
    del a
    assert len(pool.free) == 2  # still referenced
    del alias, b
    assert sorted(pool.free) == [0, 1, 2, 3]
    for _ in range(4):
        pool.take()
    with pytest.raises(RuntimeError):
        pool.take()
    # U-turn criterion: array-level == Atoms-level (nuts_hmc.py:84-86)
    rs = np.random.RandomState(0)
    atoms = structures.random_atoms(7, 1)
    masses = atoms.get_masses().reshape(-1, 1)
    for _ in range(20):
        ends = []
        for _ in range(2):
            x = atoms.copy()
            x.positions += rs.normal(0, 1, (7, 3))
            x.set_momenta(rs.normal(0, 1, (7, 3)))
            ends.append(x)
        sts = [sim._State(x.get_positions(), x.get_momenta(), 0., None, 0.) for x in ends]
        assert sim._no_u_turn_states(sts[0], sts[1], masses) == sim._no_u_turn(*ends)
        assert sts[0].velocities(masses) is sts[0].velocities(masses)
        assert np.array_equal(sts[0].velocities(masses), ends[0].get_velocities().ravel())
    assert sim._safe_exp(0.) == 1.0 and sim._safe_exp(1e4) == np.inf
    assert sim._safe_exp(-1e4) == 0.0 and sim._safe_exp(float('nan')) == 0.0
    assert abs(sim._safe_exp(-3.5) - np.exp(-3.5)) < 1e-16
