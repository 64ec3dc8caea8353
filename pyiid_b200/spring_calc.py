"""Spring restraint calculators on the B200 (reference ``pyiid/calc/spring_calc.py``).

Same class, keyword arguments and module-level functions as the reference;
the N x N x 3 numpy arrays it builds per call are replaced by the pair kernels
of ``csrc/iid_spring.cuh`` behind ``iid_spring_host`` / ``iid_spring_voxel_host``
(``include/iid_b200.h``).  ``precision='fp32'`` reproduces the reference's
float32 pair arithmetic, ``'fp64'`` evaluates the same formulas in float64.
"""
import numpy as np

from .ase_shim import Calculator
from .backend import Backend

__all__ = ['Spring', 'spring_nrg', 'spring_force', 'voxel_spring_nrg',
           'atomwise_spring_nrg', 'com_spring_nrg', 'com_spring_force',
           'voxel_com_spring_nrg', 'atomwise_com_spring_nrg',
           'att_spring_nrg', 'att_spring_force', 'voxel_att_spring_nrg',
           'atomwise_att_spring_nrg']


def _backend(precision='fp32'):
    # the spring entry points need no structure: share the F(Q)-grid handle
    return Backend.get(precision, None, 'fq')


def _com(atoms, sp_type):
    return atoms.get_center_of_mass() if sp_type == 'com' else None


def _energy(atoms, k, rt, sp_type, precision='fp32'):
    if len(atoms) == 0:
        return 0.0
    e, _, _ = _backend(precision).spring(atoms.get_positions(), sp_type, k, rt,
                                         _com(atoms, sp_type))
    return e


def _force(atoms, k, rt, sp_type, precision='fp32'):
    if len(atoms) == 0:
        return np.zeros((0, 3))
    _, f, _ = _backend(precision).spring(atoms.get_positions(), sp_type, k, rt,
                                         _com(atoms, sp_type), want_energy=False,
                                         want_forces=True)
    return f


def _atomwise(atoms, k, rt, sp_type, precision='fp32'):
    if len(atoms) == 0:
        return np.zeros(0)
    _, _, a = _backend(precision).spring(atoms.get_positions(), sp_type, k, rt,
                                         _com(atoms, sp_type), want_energy=False,
                                         want_atomwise=True)
    return a


def _voxels(atoms, k, rt, resolution, sp_type, precision='fp32'):
    # spring_calc.py:151-152: np.zeros(c / resolution) (a float shape there)
    c = np.diagonal(atoms.get_cell())
    shape = tuple(int(v) for v in c / resolution)
    return _backend(precision).spring_voxels(atoms.get_positions(), sp_type, k, rt,
                                             resolution, shape, _com(atoms, sp_type))


# module-level functions, names and signatures of spring_calc.py:107-332
def spring_nrg(atoms, k, rt, precision='fp32'):
    return _energy(atoms, k, rt, 'rep', precision)


def spring_force(atoms, k, rt, precision='fp32'):
    return _force(atoms, k, rt, 'rep', precision)


def voxel_spring_nrg(atoms, k_const, rt, resolution, precision='fp32'):
    return _voxels(atoms, k_const, rt, resolution, 'rep', precision)


def atomwise_spring_nrg(atoms, k, rt, precision='fp32'):
    return _atomwise(atoms, k, rt, 'rep', precision)


def com_spring_nrg(atoms, k, rt, precision='fp32'):
    return _energy(atoms, k, rt, 'com', precision)


def com_spring_force(atoms, k, rt, precision='fp32'):
    return _force(atoms, k, rt, 'com', precision)


def voxel_com_spring_nrg(atoms, k_const, rt, resolution, precision='fp32'):
    return _voxels(atoms, k_const, rt, resolution, 'com', precision)


def atomwise_com_spring_nrg(atoms, k, rt, precision='fp32'):
    # the reference sums over the atoms here (spring_calc.py:259-266)
    return np.sum(_atomwise(atoms, k, rt, 'com', precision), axis=0)


def att_spring_nrg(atoms, k, rt, precision='fp32'):
    return _energy(atoms, k, rt, 'att', precision)


def att_spring_force(atoms, k, rt, precision='fp32'):
    return _force(atoms, k, rt, 'att', precision)


def voxel_att_spring_nrg(atoms, k_const, rt, resolution, precision='fp32'):
    return _voxels(atoms, k_const, rt, resolution, 'att', precision)


def atomwise_att_spring_nrg(atoms, k, rt, precision='fp32'):
    return _atomwise(atoms, k, rt, 'att', precision)


class Spring(Calculator):
    """Spring repulsion / attraction / centre-of-mass potential energy surface
    (``calc/spring_calc.py:10-104``)."""
    implemented_properties = ['energy', 'forces']

    def __init__(self, restart=None, ignore_bad_restart_file=False, label=None,
                 atoms=None, k=10, rt=1.5, sp_type='rep', precision='fp32',
                 **kwargs):
        Calculator.__init__(self, restart, ignore_bad_restart_file, label,
                            atoms, **kwargs)
        if sp_type not in ('rep', 'com', 'att'):
            # the reference silently keeps the 'rep' kernels for unknown types
            sp_type_eff = 'rep'
        else:
            sp_type_eff = sp_type
        self.sp_type = sp_type
        self._kind = sp_type_eff
        self.k = k
        self.rt = rt
        self.precision = precision

    def calculate(self, atoms=None, properties=['energy'],
                  system_changes=['positions', 'numbers', 'cell', 'pbc',
                                  'charges', 'magmoms']):
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

    def calculate_energy(self, atoms):
        self.results['energy'] = _energy(atoms, self.k, self.rt, self._kind,
                                         self.precision)

    def calculate_forces(self, atoms):
        self.results['forces'] = _force(atoms, self.k, self.rt, self._kind,
                                        self.precision)

    def calculate_voxel_energy(self, atoms, resolution):
        return _voxels(atoms, self.k, self.rt, resolution, self._kind,
                       self.precision)

    def calculate_atomwise_energy(self, atoms):
        a = _atomwise(atoms, self.k, self.rt, self._kind, self.precision)
        return np.sum(a, axis=0) if self._kind == 'com' else a
