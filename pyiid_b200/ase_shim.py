"""Minimal stand-ins for the parts of ASE the hot path's callers touch.

The reference builds on the Atomic Simulation Environment (``ase.Atoms``,
``ase.calculators.calculator.Calculator``, ``ase.optimize.optimize.Optimizer``,
``ase.md.velocitydistribution.MaxwellBoltzmannDistribution``, ``ase.units``);
it is a protocol, not arithmetic on the path.  ASE is
not installed in this image, so when ``import ase`` fails :func:`install`
registers these duck-types under the ``ase.*`` module names the reference's
``pyiid/sim`` and ``pyiid/calc`` import.  With a real ASE present nothing here
is used.
"""
import copy
import re
import sys
import types

import numpy as np

chemical_symbols = [
    'X', 'H', 'He', 'Li', 'Be', 'B', 'C', 'N', 'O', 'F', 'Ne', 'Na', 'Mg',
    'Al', 'Si', 'P', 'S', 'Cl', 'Ar', 'K', 'Ca', 'Sc', 'Ti', 'V', 'Cr', 'Mn',
    'Fe', 'Co', 'Ni', 'Cu', 'Zn', 'Ga', 'Ge', 'As', 'Se', 'Br', 'Kr', 'Rb',
    'Sr', 'Y', 'Zr', 'Nb', 'Mo', 'Tc', 'Ru', 'Rh', 'Pd', 'Ag', 'Cd', 'In',
    'Sn', 'Sb', 'Te', 'I', 'Xe', 'Cs', 'Ba', 'La', 'Ce', 'Pr', 'Nd', 'Pm',
    'Sm', 'Eu', 'Gd', 'Tb', 'Dy', 'Ho', 'Er', 'Tm', 'Yb', 'Lu', 'Hf', 'Ta',
    'W', 'Re', 'Os', 'Ir', 'Pt', 'Au', 'Hg', 'Tl', 'Pb', 'Bi', 'Po', 'At',
    'Rn', 'Fr', 'Ra', 'Ac', 'Th', 'Pa', 'U']
atomic_numbers = {s: z for z, s in enumerate(chemical_symbols)}

# standard atomic weights (u) for the elements the benchmarks and tests use;
# others default to 2.5 * Z (only the thermostat's time scale depends on it)
_MASS = {1: 1.008, 6: 12.011, 7: 14.007, 8: 15.999, 14: 28.085, 26: 55.845,
         28: 58.6934, 29: 63.546, 46: 106.42, 47: 107.8682, 78: 195.084,
         79: 196.966569}


def _mass(z):
    return _MASS.get(int(z), 2.5 * int(z))


_MASS_CACHE = {}


# ase.units (CODATA 2014 as in ASE 3.x)
class units(object):
    kB = 8.6173303e-05
    _e = 1.6021766208e-19
    _amu = 1.660539040e-27
    second = 1e10 * np.sqrt(_e / _amu)
    fs = 1e-15 * second


def _parse_symbols(symbols):
    if isinstance(symbols, str):
        out = []
        for sym, cnt in re.findall(r'([A-Z][a-z]?)(\d*)', symbols):
            out.extend([atomic_numbers[sym]] * (int(cnt) if cnt else 1))
        return out
    out = []
    for s in symbols:
        out.append(atomic_numbers[s] if isinstance(s, str) else int(s))
    return out


class Atom(object):
    def __init__(self, symbol='X', position=(0, 0, 0), index=None):
        self.number = (atomic_numbers[symbol] if isinstance(symbol, str)
                       else int(symbol))
        self.position = np.array(position, dtype=float)
        self.index = index

    @property
    def symbol(self):
        return chemical_symbols[self.number]


# Per-atom scatter-factor tables are written once by ElasticScatter._wrap_atoms
# (always replaced wholesale, never edited in place) and are by far the largest
# arrays on the object; copies of an Atoms share them instead of duplicating
# them on every leapfrog (pyiid/sim/__init__.py:29 deep-copies the atoms).
SHARED_ARRAYS = ('F(Q) scatter', 'PDF scatter')


class Atoms(object):
    def __init__(self, symbols=None, positions=None, numbers=None, cell=None,
                 pbc=False, momenta=None, calculator=None, info=None):
        if isinstance(symbols, Atoms):
            other = symbols
            symbols, numbers = None, other.numbers.copy()
            positions = other.positions.copy()
        if numbers is None:
            numbers = _parse_symbols(symbols if symbols is not None else [])
        numbers = np.array(numbers, dtype=int)
        n = len(numbers)
        if positions is None:
            positions = np.zeros((n, 3))
        self.arrays = {'numbers': numbers,
                       'positions': np.array(positions, dtype=float).reshape(n, 3)}
        self.cell = np.zeros((3, 3)) if cell is None else np.array(cell, float)
        if self.cell.shape == (3,):
            self.cell = np.diag(self.cell)
        self.pbc = np.array([bool(pbc)] * 3) if np.isscalar(pbc) else np.array(pbc)
        self.info = {} if info is None else dict(info)
        self._calc = None
        if momenta is not None:
            self.set_momenta(momenta)
        if calculator is not None:
            self.set_calculator(calculator)

    # --- containers ------------------------------------------------------
    def __len__(self):
        return len(self.arrays['positions'])

    @property
    def positions(self):
        return self.arrays['positions']

    @positions.setter
    def positions(self, value):
        self.arrays['positions'][:] = value

    @property
    def numbers(self):
        return self.arrays['numbers']

    @numbers.setter
    def numbers(self, value):
        self.arrays['numbers'][:] = value

    def get_positions(self):
        return self.arrays['positions'].copy()

    def set_positions(self, value):
        self.arrays['positions'][:] = value

    def get_atomic_numbers(self):
        return self.arrays['numbers'].copy()

    def get_chemical_symbols(self):
        return [chemical_symbols[z] for z in self.arrays['numbers']]

    def get_tags(self):
        return self.arrays.get('tags', np.zeros(len(self), int)).copy()

    def set_array(self, name, a, dtype=None, shape=None):
        if a is None:
            self.arrays.pop(name, None)
            return
        a = np.array(a, dtype)
        if len(a) != len(self):
            raise ValueError('Array has wrong length: %d != %d.' %
                             (len(a), len(self)))
        self.arrays[name] = a

    def get_array(self, name, copy=True):
        return self.arrays[name].copy() if copy else self.arrays[name]

    def has(self, name):
        return name in self.arrays

    def get_masses(self):
        if 'masses' in self.arrays:
            return self.arrays['masses'].copy()
        numbers = self.arrays['numbers']
        key = numbers.tobytes()
        m = _MASS_CACHE.get(key)
        if m is None:
            zs, inv = np.unique(numbers, return_inverse=True)
            m = np.array([_mass(z) for z in zs])[inv.reshape(-1)]
            if len(_MASS_CACHE) > 64:
                _MASS_CACHE.clear()
            _MASS_CACHE[key] = m
        return m.copy()

    def set_masses(self, masses):
        self.set_array('masses', masses, float)

    def get_momenta(self):
        if 'momenta' in self.arrays:
            return self.arrays['momenta'].copy()
        return np.zeros((len(self), 3))

    def set_momenta(self, momenta):
        self.arrays['momenta'] = np.array(momenta, dtype=float).reshape(len(self), 3)

    def get_velocities(self):
        return self.get_momenta() / self.get_masses().reshape(-1, 1)

    def set_velocities(self, v):
        self.set_momenta(self.get_masses()[:, None] * np.asarray(v))

    def get_kinetic_energy(self):
        p = self.get_momenta()
        return 0.5 * np.vdot(p, self.get_velocities())

    def get_center_of_mass(self):
        m = self.get_masses()
        return np.dot(m, self.arrays['positions']) / m.sum()

    def center(self, vacuum=None, axis=(0, 1, 2), about=None):
        """Centre the bounding box of the atoms in the cell (ASE semantics
        for ``vacuum=None``)."""
        p = self.arrays['positions']
        if len(p) == 0:
            return
        if vacuum is not None:
            # orthorhombic cell with `vacuum` on both sides of the atoms
            extent = p.max(0) - p.min(0)
            for a in np.atleast_1d(axis):
                self.cell[a] = 0.0
                self.cell[a, a] = extent[a] + 2.0 * vacuum
        box_centre = 0.5 * (p.min(0) + p.max(0))
        target = 0.5 * self.cell.sum(0) if about is None else np.array(about)
        shift = target - box_centre
        for a in np.atleast_1d(axis):
            p[:, a] += shift[a]

    def copy(self):
        new = object.__new__(self.__class__)
        new.arrays = {k: (v if k in SHARED_ARRAYS else v.copy())
                      for k, v in self.arrays.items()}
        new.cell = self.cell.copy()
        new.pbc = self.pbc.copy()
        # info holds small bookkeeping (the experiment dict, counters)
        new.info = {k: (dict(v) if isinstance(v, dict) else copy.deepcopy(v))
                    for k, v in self.info.items()}
        new._calc = None
        return new

    def __deepcopy__(self, memo):
        new = self.copy()
        memo[id(self)] = new
        new._calc = copy.deepcopy(self._calc, memo)
        return new

    def __add__(self, other):
        new = self.copy()
        new.extend(other)
        return new

    def __iadd__(self, other):
        self.extend(other)
        return self

    def get_cell(self):
        return self.cell.copy()

    def set_cell(self, cell):
        cell = np.array(cell, float)
        self.cell = np.diag(cell) if cell.shape == (3,) else cell

    def extend(self, other):
        if isinstance(other, Atom):
            other = Atoms(numbers=[other.number], positions=[other.position])
        n1, n2 = len(self), len(other)
        for name in list(self.arrays):
            a = self.arrays[name]
            if name in other.arrays:
                b = other.arrays[name]
            else:
                b = np.zeros((n2,) + a.shape[1:], a.dtype)
            self.arrays[name] = np.concatenate([a, b])
        assert len(self) == n1 + n2

    append = extend

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            n = len(self)
            if i < -n or i >= n:
                raise IndexError('Index out of range.')
            return Atom(int(self.arrays['numbers'][i]), self.arrays['positions'][i],
                        index=int(i) % n)
        new = self.__class__(numbers=self.arrays['numbers'][i],
                             positions=self.arrays['positions'][i],
                             cell=self.cell.copy(), pbc=self.pbc.copy(),
                             info=copy.deepcopy(self.info))
        for k, v in self.arrays.items():
            new.arrays[k] = v[i].copy()
        return new

    def __delitem__(self, i):
        mask = np.ones(len(self), bool)
        mask[i] = False
        for k in list(self.arrays):
            self.arrays[k] = self.arrays[k][mask]

    # --- calculator ------------------------------------------------------
    def set_calculator(self, calc=None):
        self._calc = calc

    def get_calculator(self):
        return self._calc

    calc = property(get_calculator, set_calculator)

    def _need_calc(self):
        if self._calc is None:
            raise RuntimeError('Atoms object has no calculator.')
        return self._calc

    def get_potential_energy(self):
        return self._need_calc().get_potential_energy(self)

    def get_forces(self):
        return self._need_calc().get_forces(self)

    def get_total_energy(self):
        return self.get_potential_energy() + self.get_kinetic_energy()


all_changes = ['positions', 'numbers', 'cell', 'pbc', 'charges', 'magmoms']


def compare_atoms(atoms1, atoms2, tol=1e-15):
    if atoms1 is None:
        return all_changes[:]
    changes = []
    if len(atoms1) != len(atoms2):
        return all_changes[:]
    if not np.array_equal(atoms1.positions, atoms2.positions):
        changes.append('positions')
    if not np.array_equal(atoms1.numbers, atoms2.numbers):
        changes.append('numbers')
    if not np.array_equal(atoms1.cell, atoms2.cell):
        changes.append('cell')
    if not np.array_equal(atoms1.pbc, atoms2.pbc):
        changes.append('pbc')
    return changes


class Calculator(object):
    """The slice of ase.calculators.calculator.Calculator that Calc1D uses:
    a results cache invalidated by ``check_state``."""
    implemented_properties = []

    def __init__(self, restart=None, ignore_bad_restart_file=False, label=None,
                 atoms=None, **kwargs):
        self.atoms = None
        self.results = {}
        self.parameters = dict(kwargs)
        self.label = label
        if atoms is not None:
            atoms.set_calculator(self)

    def reset(self):
        self.atoms = None
        self.results = {}

    def check_state(self, atoms, tol=1e-15):
        return compare_atoms(self.atoms, atoms)

    def calculate(self, atoms=None, properties=['energy'],
                  system_changes=all_changes):
        if atoms is not None:
            self.atoms = atoms.copy()

    def get_property(self, name, atoms=None, allow_calculation=True):
        if name not in self.implemented_properties:
            raise NotImplementedError('%s property not implemented' % name)
        if atoms is None:
            atoms = self.atoms
            system_changes = []
        else:
            system_changes = self.check_state(atoms)
            if system_changes:
                self.reset()
        if name not in self.results:
            if not allow_calculation:
                return None
            self.calculate(atoms, [name], system_changes)
        result = self.results[name]
        if isinstance(result, np.ndarray):
            result = result.copy()
        return result

    def get_potential_energy(self, atoms=None, force_consistent=False):
        return self.get_property('energy', atoms)

    def get_forces(self, atoms=None):
        return self.get_property('forces', atoms)


class Optimizer(object):
    """ase.optimize.optimize.Optimizer reduced to what pyiid.sim.Ensemble
    needs: positional (atoms, restart, logfile, trajectory), observers."""

    def __init__(self, atoms, restart=None, logfile=None, trajectory=None,
                 master=None, **kwargs):
        self.atoms = atoms
        self.restart = restart
        self.logfile = logfile
        self.trajectory = trajectory
        self.observers = []
        self.nsteps = 0

    def attach(self, function, interval=1, *args, **kwargs):
        self.observers.append((function, interval, args, kwargs))

    def call_observers(self):
        for function, interval, args, kwargs in self.observers:
            if interval > 0 and self.nsteps % interval == 0:
                function(*args, **kwargs)
        self.nsteps += 1


def MaxwellBoltzmannDistribution(atoms, temp=None, communicator=None,
                                 force_temp=False, rng=None, temperature_K=None):
    """Old-ASE signature: ``temp`` is k_B T in energy units (the reference
    passes ``temperature * kB``, pyiid/sim/nuts_hmc.py:135-136,166-167)."""
    if temperature_K is not None:
        temp = temperature_K * units.kB
    if rng is None:
        rng = np.random
    masses = atoms.get_masses()
    xi = rng.standard_normal((len(masses), 3))
    momenta = xi * np.sqrt(masses * temp)[:, np.newaxis]
    atoms.set_momenta(momenta)
    if force_temp:
        temp0 = atoms.get_kinetic_energy() / len(atoms) / 1.5
        if temp0 > 0:
            atoms.set_momenta(atoms.get_momenta() * np.sqrt(temp / temp0))


def have_real_ase():
    mod = sys.modules.get('ase')
    if mod is not None:
        return not getattr(mod, '__pyiid_b200_shim__', False)
    try:
        import ase  # noqa: F401
        return True
    except ImportError:
        return False


def install():
    """Register the stand-ins as ``ase.*`` unless a real ASE is importable."""
    if have_real_ase() or 'ase' in sys.modules:
        return False

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__pyiid_b200_shim__ = True
        sys.modules[name] = m
        return m

    ase = mod('ase', Atoms=Atoms, Atom=Atom)
    ase.__path__ = []
    ase.atoms = mod('ase.atoms', Atoms=Atoms)
    ase.atom = mod('ase.atom', Atom=Atom)
    ase.units = mod('ase.units', kB=units.kB, fs=units.fs, second=units.second)
    calcs = mod('ase.calculators')
    calcs.__path__ = []
    calcs.calculator = mod('ase.calculators.calculator', Calculator=Calculator,
                           all_changes=all_changes)
    ase.calculators = calcs
    opt = mod('ase.optimize')
    opt.__path__ = []
    opt.optimize = mod('ase.optimize.optimize', Optimizer=Optimizer)
    ase.optimize = opt
    md = mod('ase.md')
    md.__path__ = []
    md.velocitydistribution = mod(
        'ase.md.velocitydistribution',
        MaxwellBoltzmannDistribution=MaxwellBoltzmannDistribution)
    ase.md = md
    ase.data = mod('ase.data', chemical_symbols=chemical_symbols,
                   atomic_numbers=atomic_numbers)
    return True
