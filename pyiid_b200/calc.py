"""``Calc1D`` / ``PDFCalc``: the Rw / chi^2 ASE-calculator surface.

Mirror of ``pyiid/calc/calc_1d.py:9-101`` (class ``Calc1D``) and
``pyiid/calc/__init__.py:10-105`` (``wrap_rw``, ``wrap_chi_sq``,
``wrap_grad_rw``, ``wrap_grad_chi_sq``).  ``PDFCalc`` is the class's older
name, still used by ``benchmarks/time_comparison.py:6,45``
(``PDFCalc(obs_data=..., scatter=..., conv=..., potential=...)``).

When ``exp_function`` / ``exp_grad_function`` are ``get_pdf`` /
``get_grad_pdf`` of one :class:`~pyiid_b200.elasticscatter.ElasticScatter`,
energy and forces come from ONE fused device evaluation
(F(Q) pass -> G(r) -> Rw/chi^2 + chain-rule weights -> force pass) and both
properties are cached for that configuration; the reference runs
1 x grad-PDF + 2 x PDF for the same result (``calc_1d.py:78-95``).  Any other
pair of experiment functions takes the generic route below; its reductions
also run on the device.
"""
import copy
import ctypes

import numpy as np

from . import ase_shim
from ._lib import IID_POT_RW, IID_POT_CHI_SQ, check

if ase_shim.have_real_ase():  # pragma: no cover - ASE is absent in this image
    from ase.calculators.calculator import Calculator
else:
    Calculator = ase_shim.Calculator

__all__ = ['Calc1D', 'PDFCalc', 'wrap_rw', 'wrap_chi_sq', 'wrap_grad_rw',
           'wrap_grad_chi_sq']


def _potential(gcalc, gobs, potential, want_c=False):
    """(value, scale[, c]) on the device; see iid_rw_host."""
    from .backend import Backend
    be = Backend.get('fp64')
    gcalc = np.ascontiguousarray(gcalc, dtype=np.float64).ravel()
    gobs = np.ascontiguousarray(gobs, dtype=np.float64).ravel()
    if gcalc.shape != gobs.shape:
        raise ValueError('calculated and observed data differ in length')
    out = np.zeros(4)
    c = np.zeros(len(gcalc)) if want_c else None
    check(be.lib.iid_rw_host(be.h, gcalc.ctypes.data, gobs.ctypes.data,
                             len(gcalc), potential, 1.0, out.ctypes.data,
                             c.ctypes.data if want_c else None))
    return (out[2], out[1], c) if want_c else (out[2], out[1])


def _contract(grad_gcalc, c):
    from .backend import Backend
    be = Backend.get('fp64')
    g = np.asarray(grad_gcalc)
    if g.dtype != np.float32:
        g = g.astype(np.float64, copy=False)
    g = np.ascontiguousarray(g)
    rows = int(np.prod(g.shape[:-1]))
    out = np.zeros(rows)
    check(be.lib.iid_contract_host(be.h, g.ctypes.data,
                                   int(g.dtype == np.float32), rows,
                                   g.shape[-1], c.ctypes.data, out.ctypes.data))
    return out.reshape(g.shape[:-1])


def wrap_rw(gcalc, gobs):
    """(rw, scale) -- ``calc/__init__.py:10-31`` / ``master_kernel.get_rw``."""
    return _potential(gcalc, gobs, IID_POT_RW)


def wrap_chi_sq(gcalc, gobs):
    """(chi_sq, scale) -- ``calc/__init__.py:33-54``."""
    return _potential(gcalc, gobs, IID_POT_CHI_SQ)


def wrap_grad_rw(grad_gcalc, gcalc, gobs):
    """[N,3] gradient of Rw -- ``calc/__init__.py:56-79`` /
    ``master_kernel.get_grad_rw :293-347`` (linear in grad_gcalc)."""
    _, _, c = _potential(gcalc, gobs, IID_POT_RW, True)
    return _contract(grad_gcalc, c)


def wrap_grad_chi_sq(grad_gcalc, gcalc, gobs):
    """[N,3] gradient of chi^2 -- ``calc/__init__.py:82-105``."""
    _, _, c = _potential(gcalc, gobs, IID_POT_CHI_SQ, True)
    return _contract(grad_gcalc, c)


def _bound_to(func, name):
    """The ElasticScatter a bound method belongs to, if it is ``name``."""
    from .elasticscatter import ElasticScatter
    owner = getattr(func, '__self__', None)
    if isinstance(owner, ElasticScatter) and \
            getattr(func, '__func__', None) is getattr(ElasticScatter, name):
        return owner
    return None


def translation_aware_changes(calc, atoms, tol=1e-15):
    """Calculator.check_state that ignores a rigid translation of the atoms."""
    changes = Calculator.check_state(calc, atoms, tol)
    if changes == ['positions'] and calc.atoms is not None and \
            len(calc.atoms) == len(atoms) and len(atoms) > 0:
        old, new = calc.atoms.positions, atoms.positions
        if np.abs((new - new[0]) - (old - old[0])).max() <= 1e-10:
            return []
    return changes


class Calc1D(Calculator):
    """PDF / F(Q) based Rw or chi^2 calculator (``calc/calc_1d.py:9-101``)."""
    implemented_properties = ['energy', 'forces']

    def __init__(self, restart=None, ignore_bad_restart_file=False, label=None,
                 atoms=None, target_data=None, exp_function=None,
                 exp_grad_function=None, conv=1., potential='rw', **kwargs):
        Calculator.__init__(self, restart, ignore_bad_restart_file, label,
                            atoms, **kwargs)
        if target_data is None or len(np.shape(target_data)) != 1:
            raise NotImplementedError('Need a 1d array target data set')
        if exp_function is None or exp_grad_function is None:
            raise NotImplementedError('Need functions which return the '
                                      'simulated data associated with the '
                                      'experiment and its gradient')
        # a private read-only copy: the device-resident target is recognised by
        # object identity on the sampler's hot path, which is only safe if the
        # array cannot change in place (assign a new array to replace it)
        self.target_data = np.array(target_data)
        self.target_data.flags.writeable = False
        self.exp_function = exp_function
        self.exp_grad_function = exp_grad_function
        self.scale = 1
        self.rw_to_eV = conv
        if potential == 'chi_sq':
            self.potential = wrap_chi_sq
            self.grad = wrap_grad_chi_sq
        elif potential == 'rw':
            self.potential = wrap_rw
            self.grad = wrap_grad_rw
        else:
            raise NotImplementedError('Potential not implemented')
        self.potential_name = potential
        scat = _bound_to(exp_function, 'get_pdf')
        self._fused = scat if (scat is not None and scat is _bound_to(
            exp_grad_function, 'get_grad_pdf')) else None

    def __deepcopy__(self, memo):
        """Copies share the target data and the ElasticScatter (and through it
        the native handle); only the results cache and the reference atoms are
        private.  leapfrog deep-copies atoms + calculator on every step."""
        new = copy.copy(self)
        memo[id(self)] = new
        new.results = {k: (v.copy() if isinstance(v, np.ndarray) else v)
                       for k, v in self.results.items()}
        new.atoms = None if self.atoms is None else self.atoms.copy()
        if hasattr(self, 'parameters'):
            new.parameters = copy.copy(self.parameters)
        return new

    def check_state(self, atoms, tol=1e-15):
        """F(Q), G(r) and therefore Rw/chi^2 and the forces are invariant under
        rigid translation; leapfrog re-centres the atoms after its last force
        evaluation (pyiid/sim/__init__.py:36-37), which must not invalidate
        the cached results."""
        return translation_aware_changes(self, atoms, tol)

    def calculate(self, atoms=None, properties=['energy'],
                  system_changes=['positions', 'numbers', 'cell', 'pbc',
                                  'charges', 'magmoms']):
        if self._fused is not None and atoms is not None:
            # keep the scatter-factor arrays on the caller's atoms so the
            # per-evaluation copy below inherits them instead of re-wrapping
            self._fused._ensure_wrapped(atoms)
        Calculator.calculate(self, atoms, properties, system_changes)
        for prop in properties:
            if prop not in self.results:
                if prop == 'energy':
                    self.calculate_energy(self.atoms)
                if prop == 'forces':
                    self.calculate_forces(self.atoms)

    def _fused_eval(self, atoms, want_forces):
        e, scale, forces = self._fused.get_pdf_energy_forces(
            atoms, self.target_data, self.potential_name, self.rw_to_eV,
            want_forces)
        self.scale = scale
        self.results['energy'] = e
        if want_forces:
            self.results['forces'] = forces

    def calculate_energy(self, atoms):
        if self._fused is not None:
            return self._fused_eval(atoms, False)
        energy, scale = self.potential(self.exp_function(atoms),
                                       self.target_data)
        self.scale = scale
        self.results['energy'] = energy * self.rw_to_eV

    def calculate_forces(self, atoms):
        if self._fused is not None:
            # one evaluation serves both properties of this configuration
            return self._fused_eval(atoms, True)
        self.results['forces'] = self.grad(self.exp_grad_function(atoms),
                                           self.exp_function(atoms),
                                           self.target_data) * self.rw_to_eV


def PDFCalc(obs_data=None, scatter=None, conv=1., potential='rw', **kwargs):
    """Older constructor name (``benchmarks/time_comparison.py:45``)."""
    return Calc1D(target_data=obs_data, exp_function=scatter.get_pdf,
                  exp_grad_function=scatter.get_grad_pdf, conv=conv,
                  potential=potential, **kwargs)
