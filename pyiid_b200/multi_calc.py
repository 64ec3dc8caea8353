"""Sum of several calculators (reference ``pyiid/calc/multi_calc.py``).

Energies and forces of the sub-calculators are added, as in the reference.
The combination used in refinements -- one ``Calc1D`` on the PDF plus rep / att
``Spring`` restraints -- is evaluated as ONE device sequence: the spring pair
kernels run in the CUDA graph of the fused energy+forces call, the forces are
summed on the device and come back in a single copy
(``iid_set_restraints`` + ``iid_energy_forces_host``, ``include/iid_b200.h``).
Any other list falls back to evaluating each calculator in turn (each of them
on the GPU) exactly as ``multi_calc.py:58-83`` does.
"""
import copy

import numpy as np

from .ase_shim import Calculator
from .calc import Calc1D, translation_aware_changes
from .spring_calc import Spring
from ._lib import IID_MAX_RESTRAINTS

__all__ = ['MultiCalc']


class MultiCalc(Calculator):
    implemented_properties = ['energy', 'forces']

    def __init__(self, restart=None, ignore_bad_restart_file=False, label=None,
                 atoms=None, calc_list=None, **kwargs):
        Calculator.__init__(self, restart, ignore_bad_restart_file, label,
                            atoms, **kwargs)
        self.calc_list = calc_list
        self._plan = self._fusable(calc_list)

    @staticmethod
    def _fusable(calc_list):
        """(calc1d, [(sp_type, k, rt), ...]) when the list is one fused
        Calc1D plus rep / att springs of the same precision, else None."""
        if not calc_list:
            return None
        pdf = [c for c in calc_list if isinstance(c, Calc1D)]
        springs = [c for c in calc_list if isinstance(c, Spring)]
        if len(pdf) != 1 or len(pdf) + len(springs) != len(calc_list):
            return None
        scat = pdf[0]._fused
        if scat is None or not springs or len(springs) > IID_MAX_RESTRAINTS:
            return None
        if any(s._kind == 'com' or s.precision != scat.precision for s in springs):
            return None
        return pdf[0], [(s._kind, s.k, s.rt) for s in springs]

    def __deepcopy__(self, memo):
        new = copy.copy(self)
        memo[id(self)] = new
        new.results = {k: (v.copy() if isinstance(v, np.ndarray) else v)
                       for k, v in self.results.items()}
        new.atoms = None if self.atoms is None else self.atoms.copy()
        new.calc_list = [copy.deepcopy(c, memo) for c in self.calc_list or []]
        new._plan = self._fusable(new.calc_list)
        return new

    def check_state(self, atoms, tol=1e-15):
        # the fused combination (PDF + pair springs) is translation invariant;
        # leapfrog re-centres the atoms after its last force evaluation
        if self._plan is not None:
            return translation_aware_changes(self, atoms, tol)
        return Calculator.check_state(self, atoms, tol)

    def calculate(self, atoms=None, properties=['energy'],
                  system_changes=['positions', 'numbers', 'cell', 'pbc',
                                  'charges', 'magmoms']):
        if self._plan is not None and atoms is not None:
            self._plan[0]._fused._ensure_wrapped(atoms)
        Calculator.calculate(self, atoms, properties, system_changes)
        if len(system_changes) > 0:
            if 'energy' in properties:
                self.calculate_energy(self.atoms)
            if 'forces' in properties:
                self.calculate_forces(self.atoms)
        for prop in properties:
            if prop not in self.results:
                if prop == 'energy':
                    self.calculate_energy(self.atoms)
                if prop == 'forces':
                    self.calculate_forces(self.atoms)

    def _fused_eval(self, atoms, want_forces):
        calc, springs = self._plan
        e, scale, forces, e_spring = calc._fused.get_pdf_energy_forces(
            atoms, calc.target_data, calc.potential_name, calc.rw_to_eV,
            want_forces, restraints=springs)
        calc.scale = scale
        self.results['energy'] = e + e_spring
        if want_forces:
            self.results['forces'] = forces

    def calculate_energy(self, atoms):
        if self._plan is not None:
            return self._fused_eval(atoms, False)
        energy_list = []
        for calculator in self.calc_list:
            atoms.set_calculator(calculator)
            energy_list.append(atoms.get_potential_energy())
        self.results['energy'] = sum(energy_list)

    def calculate_forces(self, atoms):
        if self._plan is not None:
            return self._fused_eval(atoms, True)
        forces = np.zeros((len(atoms), 3))
        for calculator in self.calc_list:
            atoms.set_calculator(calculator)
            forces[:, :] += atoms.get_forces()
        self.results['forces'] = forces

    def calculate_voxel_energy(self, atoms, resolution):
        c = np.diagonal(atoms.get_cell())
        voxel_energy = np.zeros(tuple(int(v) for v in c / resolution))
        for calc in self.calc_list:
            try:
                voxel_energy += calc.calculate_voxel_energy(atoms, resolution)
            except AttributeError:
                pass  # as the reference: calculators without a voxel energy are skipped
        return voxel_energy

    def calculate_atomwise_energy(self, atoms):
        nrg = np.zeros(len(atoms))
        for calc in self.calc_list:
            try:
                nrg += calc.calculate_atomwise_energy(atoms)
            except AttributeError:
                pass
        return nrg
