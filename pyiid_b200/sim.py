"""Leapfrog and the No-U-Turn sampler over the ASE calculator protocol.

Interface mirror of ``pyiid/sim/__init__.py`` (``leapfrog :10-38``,
``Ensemble :41-82``) and ``pyiid/sim/nuts_hmc.py`` (``buildtree :15-88``,
``NUTSCanonicalEnsemble :91-244``) so scripts written against ``pyiid.sim``
run on the B200 calculator.  The reference's own sampler source also runs
unchanged on this package's ``Calc1D`` (it only speaks the ASE protocol);
``tests/test_reference_sim.py`` exercises that when the reference mount is
present.  The algorithm is the slice-sampling NUTS of Hoffman & Gelman
(2014, Alg. 6) with dual-averaging step-size adaptation, in the reference's
conventions: Hamiltonian = ``atoms.get_total_energy()`` in eV, the temperature
only enters through the Maxwell-Boltzmann momentum refresh, tree doubling
stops at ``escape_level``.
"""
from __future__ import print_function

import math
from collections import namedtuple
from copy import deepcopy as dc
from time import time

import collections

import numpy as np
from numpy.random import RandomState

from . import ase_shim
from . import _lib

if ase_shim.have_real_ase():  # pragma: no cover
    from ase.optimize.optimize import Optimizer
    from ase.md.velocitydistribution import MaxwellBoltzmannDistribution
    from ase.units import kB, fs
else:
    Optimizer = ase_shim.Optimizer
    MaxwellBoltzmannDistribution = ase_shim.MaxwellBoltzmannDistribution
    kB, fs = ase_shim.units.kB, ase_shim.units.fs

__all__ = ['leapfrog', 'Ensemble', 'NUTSCanonicalEnsemble', 'buildtree',
           'classical_dynamics']

Emax = 200  # energy error beyond which a trajectory counts as divergent


def leapfrog(atoms, step, center=True):
    """One kick-drift-kick step on a copy of ``atoms``
    (``pyiid/sim/__init__.py:10-38``).  ``get_forces`` is evaluated once at the
    new positions; the first half kick reuses the cached forces."""
    latoms = dc(atoms)
    latoms.set_momenta(latoms.get_momenta() + 0.5 * step * latoms.get_forces())
    latoms.positions += step * latoms.get_velocities()
    latoms.set_momenta(latoms.get_momenta() + 0.5 * step * latoms.get_forces())
    if center:
        latoms.center()
    return latoms


def classical_dynamics(atoms, stepsize, n_steps):
    """Hamiltonian dynamics by repeated leapfrog steps; returns the list of
    configurations, the start included (``pyiid/sim/dynamics.py:5-30``).  With
    the fused device calculator the states stay on the device between steps
    (:class:`_DeviceSystem`); an Atoms object is built per returned frame."""
    atoms.get_forces()
    traj = [atoms]
    if _DeviceSystem.usable(atoms):
        system = _DeviceSystem(atoms)
        st = system.state_of(atoms)
        done = 0
        while done < n_steps:
            # a batch of steps as chains on the device, then their frames (the
            # download of a frame's forces is another call on the handle)
            batch = min(64, n_steps - done)
            system.expect(batch)
            states = []
            for _ in range(batch):
                st = system.leapfrog(st, stepsize)
                states.append(st)
            traj.extend(system.to_atoms(s) for s in states)
            done += batch
        system.close()
        return traj
    for _ in range(n_steps):
        traj.append(leapfrog(traj[-1], stepsize))
    return traj


class Ensemble(Optimizer):
    """Base class of the samplers (``pyiid/sim/__init__.py:41-82``)."""

    def __init__(self, atoms, restart=None, logfile=None, trajectory=None,
                 seed=None, verbose=False):
        Optimizer.__init__(self, atoms, restart, logfile, trajectory)
        atoms.get_forces()
        atoms.get_potential_energy()
        if seed is None:
            seed = np.random.randint(0, 2 ** 31)
        self.verbose = verbose
        self.random_state = RandomState(seed)
        self.starting_atoms = dc(atoms)
        self.traj = [dc(atoms)]
        self.pe = []
        self.metadata = {'seed': seed}

    def check_eq(self, eq_steps, tol):
        ret = np.cumsum(self.pe, dtype=float)
        ret[eq_steps:] = ret[eq_steps:] - ret[:-eq_steps]
        ret = ret[eq_steps - 1:] / eq_steps
        return np.sum(np.gradient(ret[eq_steps:])) < tol

    def run(self, steps=100000000, eq_steps=None, eq_tol=None, **kwargs):
        self.metadata['planned iterations'] = steps
        try:
            for i in range(steps):
                if eq_steps is not None and self.check_eq(eq_steps, eq_tol):
                    break
                if self.verbose:
                    print('iteration number', i)
                self.step()
        except KeyboardInterrupt:
            print('Interupted, returning data')
        return self.traj, self.metadata

    def step(self):
        pass

    def estimate_simulation_duration(self, atoms, iterations):
        pass


_Tree = namedtuple('_Tree', 'minus plus proposal n_valid keep_going '
                            'accept_sum n_leaves')


def _no_u_turn(minus, plus):
    span = (plus.positions - minus.positions).ravel()
    return (span.dot(minus.get_velocities().ravel()) >= 0) and \
        (span.dot(plus.get_velocities().ravel()) >= 0)


def _safe_exp(x):
    """exp(x) with overflow -> inf and a non-finite argument -> 0 (no numpy
    error-state context per call: this runs three times per leapfrog)."""
    try:
        v = math.exp(x)
    except OverflowError:
        return np.inf
    return v if v == v else 0.0


def buildtree(input_atoms, u, v, j, e, e0, rs, beta=1):
    """Recursive doubling of the NUTS trajectory (``nuts_hmc.py:15-88``).
    Returns (neg_atoms, pos_atoms, proposal, n_valid, keep_going, accept_sum,
    n_leaves)."""
    if j == 0:
        leaf = leapfrog(input_atoms, v * e)
        neg_delta = e0 - leaf.get_total_energy()
        n_valid = int(u <= _safe_exp(neg_delta))
        keep_going = int(u < _safe_exp(Emax + neg_delta))
        accept = min(1., _safe_exp(input_atoms.get_total_energy() -
                                   leaf.get_total_energy()))
        return _Tree(leaf, leaf, leaf, n_valid, keep_going, accept, 1)
    first = buildtree(input_atoms, u, v, j - 1, e, e0, rs, beta)
    if first.keep_going != 1:
        return first
    if v == -1:
        second = buildtree(first.minus, u, v, j - 1, e, e0, rs, beta)
        minus, plus = second.minus, first.plus
    else:
        second = buildtree(first.plus, u, v, j - 1, e, e0, rs, beta)
        minus, plus = first.minus, second.plus
    proposal = first.proposal
    total = first.n_valid + second.n_valid
    if rs.uniform() < float(second.n_valid) / max(total, 1):
        proposal = second.proposal
    keep_going = int(second.keep_going and _no_u_turn(minus, plus))
    return _Tree(minus, plus, proposal, total, keep_going,
                 first.accept_sum + second.accept_sum,
                 first.n_leaves + second.n_leaves)


class _State(object):
    """Phase-space point of the array-level sampler path: positions, momenta,
    potential energy, forces, kinetic energy (no Atoms object per leapfrog)."""
    __slots__ = ('q', 'p', 'pe', 'f', 'ke', '_v')

    def __init__(self, q, p, pe, f, ke):
        self.q, self.p, self.pe, self.f, self.ke = q, p, pe, f, ke
        self._v = None

    def velocities(self, masses):
        """p / m, flattened; kept: a state is an end point of several sub-trees."""
        if self._v is None:
            self._v = (self.p / masses).ravel()
        return self._v

    @property
    def total(self):
        return self.pe + self.ke


class _FastSystem(object):
    """The same dynamics as :func:`leapfrog` + ``get_total_energy`` for atoms
    whose calculator is the fused device ``Calc1D``: one native
    energy+forces call per leapfrog on plain arrays.  The reference's sampler
    deep-copies the Atoms object, its arrays and its calculator on every
    leapfrog (``pyiid/sim/__init__.py:29``); at a few hundred atoms that host
    work, not the scattering arithmetic, bounds the sampling rate (SURVEY.md
    section 8f-2)."""

    def __init__(self, atoms):
        calc = atoms.get_calculator()
        # Calc1D alone, or a MultiCalc of one Calc1D + rep / att springs (the
        # springs ride in the same device sequence, multi_calc.py)
        plan = getattr(calc, '_plan', None)
        self.springs = plan[1] if plan is not None else []
        calc = plan[0] if plan is not None else calc
        self.calc = calc
        self.scat = calc._fused
        self.template = atoms
        self.masses = atoms.get_masses().reshape(-1, 1)
        self.cell_centre = 0.5 * np.asarray(atoms.cell).sum(0)
        self.scat._ensure_wrapped(atoms)
        self.evals = 0

    @staticmethod
    def usable(atoms):
        from .calc import Calc1D
        from .multi_calc import MultiCalc
        calc = atoms.get_calculator() if hasattr(atoms, 'get_calculator') else None
        if isinstance(calc, MultiCalc):
            return calc._plan is not None
        return isinstance(calc, Calc1D) and calc._fused is not None

    def evaluate(self, q):
        scat, calc = self.scat, self.calc
        be = scat._load(self.template, scat.pdf_qbin, 'PDF')
        be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), scat.exp['qmin'])
        be.set_restraints(self.springs)
        e, scale, f, _ = be.energy_forces(q, calc.target_data, calc.potential_name,
                                          calc.rw_to_eV, True)
        self.evals += 1
        return float(e) + be.restraint_energy, f

    def expect(self, n):
        """Hint: the next ``n`` leapfrog calls continue one trajectory (used
        by the device-resident system to compute several steps per call)."""

    def close(self):
        """End of the system's use (the device-resident system returns the
        slots of steps it enqueued ahead)."""

    def kinetic(self, p):
        return 0.5 * float(np.vdot(p, p / self.masses))

    def state_of(self, atoms):
        q = atoms.get_positions()
        p = atoms.get_momenta()
        return _State(q, p, float(atoms.get_potential_energy()), atoms.get_forces(),
                      self.kinetic(p))

    def leapfrog(self, st, step, center=True):
        p = st.p + 0.5 * step * st.f
        q = st.q + step * (p / self.masses)
        pe, f = self.evaluate(q)
        p = p + 0.5 * step * f
        if center:
            q = q + (self.cell_centre - 0.5 * (q.min(0) + q.max(0)))
        return _State(q, p, pe, f, self.kinetic(p))

    def to_atoms(self, st):
        """An Atoms object with this state and a calculator whose result cache
        already holds its energy and forces."""
        atoms = self.template.copy()
        atoms.set_positions(st.q)
        atoms.set_momenta(st.p)
        calc = dc(self.calc)
        calc.atoms = atoms.copy()
        calc.results = {'energy': st.pe, 'forces': st.f.copy()}
        atoms.set_calculator(calc)
        return atoms


class _SlotPool(object):
    """Free list of the device slots that hold phase-space points."""

    def __init__(self, n):
        self.free = list(range(n - 1, -1, -1))

    def take(self):
        if not self.free:
            raise RuntimeError('out of device state slots (tree deeper than planned)')
        return self.free.pop()

    def give(self, slot):
        self.free.append(slot)


class _DevState(_State):
    """A :class:`_State` whose (q, p, f) also live in a device slot; ``q`` and
    ``p`` are host mirrors (U-turn test, proposals), ``f`` stays on the device
    until an accepted sample needs it.  The slot returns to the pool when the
    last reference to the state goes away."""
    __slots__ = ('slot', 'pool')

    def __init__(self, q, p, pe, f, ke, slot, pool):
        _State.__init__(self, q, p, pe, f, ke)
        self.slot, self.pool = slot, pool

    def __del__(self):
        try:
            if self.slot is not None:
                self.pool.give(self.slot)
                self.slot = None
        except Exception:  # interpreter shutdown: the pool may be gone already
            pass


class _DeviceSystem(_FastSystem):
    """:class:`_FastSystem` with the integrator on the device: a leapfrog step
    is ONE native call (``iid_leapfrog_host``: kick, drift, fused energy +
    forces, kick, centring, kinetic energy, replayed from a CUDA graph) between
    device-resident states; only q, p and three scalars come back per step.
    Same operation order as the host arithmetic (positions and momenta are
    bit-identical to :meth:`_FastSystem.leapfrog`); the kinetic-energy sum is
    reduced in a different order (last-bit differences)."""
    N_SLOTS = 256  # >= 3 live states per tree level; escape_level <= 13 in practice

    def __init__(self, atoms):
        _FastSystem.__init__(self, atoms)
        scat = self.scat
        be = scat._load(self.template, scat.pdf_qbin, 'PDF')
        be.set_transform(scat.exp['rstep'], scat.pdf_qbin, scat.get_r(), scat.exp['qmin'])
        be.set_restraints(self.springs)
        masses = np.ascontiguousarray(self.masses.reshape(-1), dtype=np.float64)
        key = (be.n, be._skey, masses.tobytes(), self.cell_centre.tobytes())
        if getattr(be, '_sampler_key', None) != key:
            be.sampler_setup(self.N_SLOTS, masses, self.cell_centre)
            be._sampler_key = key
            be._slot_pool = _SlotPool(self.N_SLOTS)
        self.be = be
        self.pool = be._slot_pool
        # look-ahead: buildtree grows a subtree of depth j by 2**j leapfrogs in a
        # row from the tree's edge; `expect(n)` announces them, and `leapfrog`
        # then ENQUEUES up to CHAIN of them in one native call (one cooperative
        # launch walks the whole chain) and collects them one by one as the
        # device completes them, so the tree logic of step i runs while the
        # device computes step i + 1.  Steps enqueued but not asked for (the
        # subtree stopped early) are dropped unseen: results and the order of
        # the random numbers are those of the step-by-step walk.
        self._ahead = collections.deque()  # destination slots of the chain in flight
        self._ahead_prev = None            # the state the next step of the chain starts from
        self._ahead_step = 0.
        self._ahead_id = 0
        self._expected = 0

    CHAIN = 64

    def expect(self, n):
        """The next ``n`` leapfrog calls continue one trajectory."""
        self._drop_ahead()
        self._expected = int(n)

    def _drop_ahead(self):
        """Forget the chain in flight (the native library waits for it and drops
        its uncollected steps at the next call)."""
        while self._ahead:
            self.pool.give(self._ahead.popleft())
        self._ahead_prev = None

    def close(self):
        """Return the slots of steps enqueued ahead but never asked for (a
        subtree that stopped early at the end of an iteration)."""
        self._expected = 0
        self._drop_ahead()

    def __del__(self):
        try:
            self._drop_ahead()
        except Exception:  # interpreter shutdown: the pool may be gone already
            pass

    @staticmethod
    def usable(atoms):
        """The fused calculator on a single rank (the sharded evaluation goes
        through torch.distributed collectives on the host path)."""
        from .backend import _dist_state
        if not _FastSystem.usable(atoms) or _dist_state()[1] != 1:
            return False
        # the one-process multi-GPU handle keeps sampler states on device 0, which
        # needs a structure that one device evaluates (below 2 x 1500 atoms)
        calc = atoms.get_calculator()
        plan = getattr(calc, '_plan', None)
        scat = (plan[0] if plan is not None else calc)._fused
        return not (getattr(scat, '_device_sel', None) == 'multi' and len(atoms) >= 3000)

    def state_of(self, atoms):
        self._drop_ahead()  # (the upload below is another call on the handle)
        st = _FastSystem.state_of(self, atoms)
        slot = self.pool.take()
        self.be.state_upload(slot, st.q, st.p, st.f)
        return _DevState(st.q, st.p, st.pe, st.f, st.ke, slot, self.pool)

    def leapfrog(self, st, step, center=True):
        calc = self.calc
        if self._ahead and not (self._ahead_prev is st and self._ahead_step == step and center):
            self._drop_ahead()  # another trajectory: what was enqueued ahead is void
        res = None
        if self._ahead:
            try:
                res = self.be.leapfrog_chain_next(self._ahead_id)
            except _lib.ChainDropped:  # another call on the backend came in between
                self._drop_ahead()
        if res is None:
            n = max(1, min(self.CHAIN, self._expected)) if center else 1
            dsts = [self.pool.take() for _ in range(n)]
            try:
                self._ahead_id = self.be.leapfrog_chain_begin(
                    st.slot, dsts, step, center, calc.target_data, calc.potential_name,
                    calc.rw_to_eV)
                self._ahead.extend(dsts)
                self._ahead_step = step
                res = self.be.leapfrog_chain_next(self._ahead_id)
            except Exception:
                if not self._ahead:
                    for d in dsts:
                        self.pool.give(d)
                self._drop_ahead()
                raise
        e, _, es, ke, q, p = res
        new = _DevState(q, p, float(e) + float(es), None, float(ke), self._ahead.popleft(),
                        self.pool)
        self._ahead_prev = new if self._ahead else None
        self._expected -= 1
        self.evals += 1
        return new

    def to_atoms(self, st):
        if st.f is None:
            self._drop_ahead()  # (the download is another call on the handle)
            st.f = self.be.state_download(st.slot, want=('f',))['f']
        return _FastSystem.to_atoms(self, st)


def _no_u_turn_states(minus, plus, masses):
    span = (plus.q - minus.q).ravel()
    return (span.dot(minus.velocities(masses)) >= 0) and \
        (span.dot(plus.velocities(masses)) >= 0)


def _buildtree_states(system, st, u, v, j, e, e0, rs):
    """:func:`buildtree` on :class:`_State` objects (same random-number
    consumption, same arithmetic)."""
    if j == 0:
        leaf = system.leapfrog(st, v * e)
        neg_delta = e0 - leaf.total
        n_valid = int(u <= _safe_exp(neg_delta))
        keep_going = int(u < _safe_exp(Emax + neg_delta))
        accept = min(1., _safe_exp(st.total - leaf.total))
        return _Tree(leaf, leaf, leaf, n_valid, keep_going, accept, 1)
    first = _buildtree_states(system, st, u, v, j - 1, e, e0, rs)
    if first.keep_going != 1:
        return first
    if v == -1:
        second = _buildtree_states(system, first.minus, u, v, j - 1, e, e0, rs)
        minus, plus = second.minus, first.plus
    else:
        second = _buildtree_states(system, first.plus, u, v, j - 1, e, e0, rs)
        minus, plus = first.minus, second.plus
    proposal = first.proposal
    total = first.n_valid + second.n_valid
    if rs.uniform() < float(second.n_valid) / max(total, 1):
        proposal = second.proposal
    keep_going = int(second.keep_going and _no_u_turn_states(minus, plus, system.masses))
    return _Tree(minus, plus, proposal, total, keep_going,
                 first.accept_sum + second.accept_sum,
                 first.n_leaves + second.n_leaves)


class NUTSCanonicalEnsemble(Ensemble):
    """No-U-Turn sampler in the canonical ensemble (``nuts_hmc.py:91-244``).

    ``fast`` (default: automatic) selects the array-level path
    (:class:`_FastSystem`) when the atoms carry the fused device ``Calc1D``;
    it draws the same random numbers and performs the same arithmetic as the
    Atoms-level path, so both produce the same trajectory.  ``device_states``
    (default: automatic) additionally keeps every phase-space point of the
    tree on the device (:class:`_DeviceSystem`)."""

    def __init__(self, atoms, restart=None, logfile=None, trajectory=None,
                 temperature=100, escape_level=13, accept_target=.65,
                 momentum=None, seed=None, verbose=False, fast=None,
                 device_states=None):
        Ensemble.__init__(self, atoms, restart, logfile, trajectory, seed,
                          verbose)
        if fast is None:
            fast = _FastSystem.usable(atoms)
        self.fast = bool(fast) and _FastSystem.usable(atoms)
        # phase-space points resident on the device (one native call per
        # leapfrog) whenever the array-level path runs on a single rank
        if device_states is None:
            device_states = True
        self.device_states = self.fast and bool(device_states) and \
            _DeviceSystem.usable(atoms)
        self.accept_target = accept_target
        self.temp = temperature
        self.thermal_nrg = self.temp * kB
        self.momentum = momentum
        self.step_size = self._find_step_size(atoms, self.thermal_nrg)
        self.mu = np.log(10 * self.step_size)
        self.sim_hbar = 0
        self.gamma = 0.05
        self.t0 = 10
        self.metadata.update({'samples_total': 0, 'accepted_samples': 0})
        self.escape_level = escape_level
        self.m = 0
        self.leapfrogs = 0

    def _refresh_momenta(self, atoms):
        if self.momentum is None:
            MaxwellBoltzmannDistribution(atoms, self.thermal_nrg,
                                         force_temp=True)
        else:
            atoms.set_momenta(self.random_state.normal(0, 1, (len(atoms), 3)))

    def _find_step_size(self, input_atoms, thermal_nrg=None, momentum=None):
        """Hoffman & Gelman Alg. 4 (``nuts_hmc.py:113-158``): double or halve
        until the one-step acceptance crosses 1/2."""
        atoms = dc(input_atoms)
        step_size = .5
        self._refresh_momenta(atoms)
        e_start = atoms.get_total_energy()

        def ratio(eps):
            return _safe_exp(e_start - leapfrog(atoms, eps).get_total_energy())

        a = 1 if ratio(step_size) > 0.5 else -1
        while ratio(step_size) ** a > 2. ** -a:
            step_size *= 2. ** a
            if self.verbose:
                print('trying step size', step_size)
            if step_size < 1e-7 or step_size > 1e7:
                step_size = 1.
                break
        if self.verbose:
            print('optimal step size', step_size)
        return step_size

    def step(self):
        """One NUTS iteration (``nuts_hmc.py:160-232``); returns the list of
        accepted configurations or None."""
        if self.fast:
            return self._step_fast()
        return self._step_atoms()

    def _step_fast(self):
        current = self.traj[-1]
        system = (_DeviceSystem if self.device_states else _FastSystem)(current)
        accepted = []
        if self.verbose:
            print('\ttime step size', self.step_size / fs, 'fs')
        self._refresh_momenta(current)
        u = self.random_state.uniform(0, 1)
        start = system.state_of(current)
        e0 = start.total
        e = self.step_size
        n, keep_going, depth = 1, 1, 0
        minus, plus = start, start
        acc_sum, leaves = 0., 1
        while keep_going == 1:
            v = self.random_state.choice([-1, 1])
            system.expect(2 ** depth)  # that many leapfrogs in a row from the tree's edge
            tree = _buildtree_states(system, minus if v == -1 else plus, u, v, depth, e,
                                     e0, self.random_state)
            if v == -1:
                minus = tree.minus
            else:
                plus = tree.plus
            acc_sum, leaves = tree.accept_sum, tree.n_leaves
            self.leapfrogs += tree.n_leaves
            if tree.keep_going == 1 and self.random_state.uniform() < min(
                    1, tree.n_valid * 1. / n):
                sample = system.to_atoms(tree.proposal)
                self.traj += [sample]
                self.metadata['accepted_samples'] += 1
                accepted.append(sample)
                self.call_observers()
            n += tree.n_valid
            keep_going = int(tree.keep_going and
                             _no_u_turn_states(minus, plus, system.masses))
            depth += 1
            if self.verbose:
                print('\t \tdepth', depth, 'samples', 2 ** depth)
            self.metadata['samples_total'] += 2 ** depth
            if depth >= self.escape_level:
                if self.verbose:
                    print('\t \t \tjmax emergency escape at {}'.format(depth))
                keep_going = 0
        system.close()  # steps enqueued ahead of a subtree that stopped early
        w = 1. / (self.m + self.t0)
        self.sim_hbar = (1 - w) * self.sim_hbar + \
            w * (self.accept_target - acc_sum / leaves)
        self.step_size = np.exp(self.mu - (self.m ** .5 / self.gamma) *
                                self.sim_hbar)
        self.m += 1
        return accepted if accepted else None

    def _step_atoms(self):
        current = self.traj[-1]
        accepted = []
        if self.verbose:
            print('\ttime step size', self.step_size / fs, 'fs')
        self._refresh_momenta(current)
        u = self.random_state.uniform(0, 1)
        e0 = current.get_total_energy()
        e = self.step_size
        n, keep_going, depth = 1, 1, 0
        minus = dc(current)
        plus = dc(current)
        acc_sum, leaves = 0., 1
        while keep_going == 1:
            v = self.random_state.choice([-1, 1])
            tree = buildtree(minus if v == -1 else plus, u, v, depth, e, e0,
                             self.random_state, 1 / self.thermal_nrg)
            if v == -1:
                minus = tree.minus
            else:
                plus = tree.plus
            acc_sum, leaves = tree.accept_sum, tree.n_leaves
            self.leapfrogs += tree.n_leaves
            if tree.keep_going == 1 and self.random_state.uniform() < min(
                    1, tree.n_valid * 1. / n):
                self.traj += [tree.proposal]
                self.metadata['accepted_samples'] += 1
                accepted.append(tree.proposal)
                tree.proposal.get_forces()
                tree.proposal.get_potential_energy()
                self.call_observers()
            n += tree.n_valid
            keep_going = int(tree.keep_going and _no_u_turn(minus, plus))
            depth += 1
            if self.verbose:
                print('\t \tdepth', depth, 'samples', 2 ** depth)
            self.metadata['samples_total'] += 2 ** depth
            if depth >= self.escape_level:
                if self.verbose:
                    print('\t \t \tjmax emergency escape at {}'.format(depth))
                keep_going = 0
        # dual averaging of the step size (Hoffman & Gelman Alg. 6)
        w = 1. / (self.m + self.t0)
        self.sim_hbar = (1 - w) * self.sim_hbar + \
            w * (self.accept_target - acc_sum / leaves)
        self.step_size = np.exp(self.mu - (self.m ** .5 / self.gamma) *
                                self.sim_hbar)
        self.m += 1
        return accepted if accepted else None

    def estimate_simulation_duration(self, atoms, iterations):
        t0 = time()
        atoms.get_forces()
        tf = time() - t0
        t2 = time()
        atoms.get_potential_energy()
        te = time() - t2
        return iterations * (tf * 2 + te) * 2 ** self.escape_level
