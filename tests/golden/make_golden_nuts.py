"""Golden NUTS trajectory: the REFERENCE's own sampler source
(/root/reference/pyiid/sim/__init__.py + nuts_hmc.py, executed from the mount,
nothing copied) driving a CPU calculator built from the oracle (float64) on a
55-atom Au icosahedron.  Run in the build container (no GPU needed):

    python tests/golden/make_golden_nuts.py

tests/test_gpu_parity.py::test_nuts_trajectory_equals_the_reference_samplers
then runs pyiid_b200.sim's NUTS on the B200 Calc1D (FP64 mode) with the same
seeds and must land on the same samples.

The one override, as in tests/test_reference_sim.py: the reference's initial
step-size search evaluates ``2 ** -a`` with a numpy integer (nuts_hmc.py:139-
150), which numpy >= 1.12 rejects, so `_find_step_size` returns a fixed step.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import oracle  # noqa: E402
from pyiid_b200 import ase_shim, structures, formfactors  # noqa: E402
from test_reference_sim import load_reference_sim  # noqa: E402

STEP = 0.04
TEMP = 300.
ESCAPE = 3
SEED = 7
NP_SEED = 5
CONV = 100.
ITERS = 3


class OracleCalc1D(ase_shim.Calculator):
    """Calc1D (calc/calc_1d.py:78-95) with exp_function = get_pdf and
    exp_grad_function = get_grad_pdf evaluated by the float64 oracle."""
    implemented_properties = ['energy', 'forces']
    evaluations = 0

    def __init__(self, target, scatter_pdf, **kw):
        ase_shim.Calculator.__init__(self, **kw)
        self.target, self.sp = target, scatter_pdf

    def calculate(self, atoms=None, properties=['energy'], system_changes=[]):
        ase_shim.Calculator.calculate(self, atoms, properties, system_changes)
        e, f, _ = oracle.calc1d_energy_forces(
            self.atoms.get_positions(), self.sp, oracle.DEFAULT_EXP, self.target, 'rw', CONV,
            'fp64')
        OracleCalc1D.evaluations += 1
        self.results['energy'] = float(e)
        self.results['forces'] = np.array(f)


def main():
    sim, nuts = load_reference_sim()
    exp = oracle.DEFAULT_EXP
    ideal = structures.icosahedron('Au', 2)
    n = len(ideal)
    nq = int(np.floor(exp['qmax'] / oracle.pdf_qbin(exp)))
    sp = np.zeros((n, nq), np.float32)
    formfactors.get_scatter_array(sp, [79] * n, oracle.pdf_qbin(exp))
    target = oracle.experiment_pdf(ideal.get_positions(), sp, exp, 'fp64')
    start = ideal.copy()
    start.positions *= 1.02
    start.positions += np.random.RandomState(2).normal(0, 0.02, start.positions.shape)
    start.center()
    start_positions = start.get_positions().copy()
    start.set_calculator(OracleCalc1D(target, sp))

    class Ensemble(nuts.NUTSCanonicalEnsemble):
        def _find_step_size(self, input_atoms, thermal_nrg=None, momentum=None):
            return STEP

    np.random.seed(NP_SEED)
    ens = Ensemble(start, temperature=TEMP, escape_level=ESCAPE, seed=SEED)
    traj, meta = ens.run(ITERS)
    out = {
        'start_positions': start_positions, 'target': target,
        'masses': start.get_masses(),
        'step': STEP, 'temperature': TEMP, 'escape_level': ESCAPE, 'seed': SEED,
        'np_seed': NP_SEED, 'conv': CONV, 'iterations': ITERS,
        'traj_positions': np.array([a.get_positions() for a in traj]),
        'traj_momenta': np.array([a.get_momenta() for a in traj]),
        'traj_energy': np.array([a.get_potential_energy() for a in traj]),
        'samples_total': meta['samples_total'], 'accepted_samples': meta['accepted_samples'],
        'final_step_size': ens.step_size,
    }
    path = os.path.join(ROOT, 'tests', 'golden', 'nuts_au55.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, 'trajectory length', len(traj), 'accepted', meta['accepted_samples'],
          'samples_total', meta['samples_total'], 'oracle evaluations', OracleCalc1D.evaluations,
          'final step size', ens.step_size)


if __name__ == '__main__':
    main()
