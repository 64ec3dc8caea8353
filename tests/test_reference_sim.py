"""The reference's own sampler source (pyiid/sim/__init__.py, nuts_hmc.py) runs
UNCHANGED on this package's ASE stand-ins and calculator protocol.  Build
container only (needs the read-only reference mount); nothing is copied: the
files are executed from where they lie."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from pyiid_b200 import ase_shim

from pyiid.sim import reference

REF = reference.REF
pytestmark = pytest.mark.skipif(not reference.available(), reason='reference mount not present')


def load_reference_sim():
    """The reference's sim/__init__.py and nuts_hmc.py, executed from the mount
    by the package's own loader (pyiid/sim/reference.py)."""
    return reference.load()


class Harmonic(ase_shim.Calculator):
    implemented_properties = ['energy', 'forces']

    def calculate(self, atoms=None, properties=['energy'], system_changes=[]):
        ase_shim.Calculator.calculate(self, atoms, properties, system_changes)
        self.results['energy'] = float((self.atoms.positions ** 2).sum())
        self.results['forces'] = -2.0 * self.atoms.positions


def test_reference_leapfrog_and_nuts_run_on_the_stand_ins(capsys):
    sim, nuts = load_reference_sim()
    a = ase_shim.Atoms('Au3', np.random.RandomState(0).normal(size=(3, 3)))
    a.set_calculator(Harmonic())
    a.set_momenta(np.ones((3, 3)))
    b = sim.leapfrog(a, 0.1, False)
    c = sim.leapfrog(b, -0.1, False)
    assert np.allclose(c.positions, a.positions, atol=1e-12)
    # our leapfrog is the same map
    from pyiid_b200 import sim as mysim
    b2 = mysim.leapfrog(a, 0.1, False)
    assert np.allclose(b.positions, b2.positions) and np.allclose(b.get_momenta(), b2.get_momenta())
    np.random.seed(1)

    # The reference's initial step-size search evaluates `2 ** -a` / `2 ** a`
    # with a numpy integer a (nuts_hmc.py:139-150), which numpy >= 1.12 rejects
    # ("Integers to negative integer powers are not allowed"), so that one
    # method cannot run on this image under ANY calculator; everything else of
    # the class (step, buildtree, dual averaging) runs as shipped.
    class Ensemble(nuts.NUTSCanonicalEnsemble):
        def _find_step_size(self, input_atoms, thermal_nrg=None, momentum=None):
            return 0.05

    with pytest.raises(ValueError):
        nuts.NUTSCanonicalEnsemble(a, temperature=300, escape_level=4, seed=3)
    ens = Ensemble(a, temperature=300, escape_level=4, seed=3)
    traj, meta = ens.run(5)
    assert meta['samples_total'] > 0 and len(traj) >= 1
    # the packaged class: the reference's NUTS with only the step-size search
    # replaced (float exponent); pyiid.sim itself stays the B200 implementation
    np.random.seed(1)
    ens2 = reference.NUTSCanonicalEnsemble(a, temperature=300, escape_level=4, seed=3)
    assert isinstance(ens2, nuts.NUTSCanonicalEnsemble) and ens2.step_size > 0
    traj2, meta2 = ens2.run(3)
    assert meta2['samples_total'] > 0
    import pyiid.sim
    from pyiid_b200 import sim as mysim2
    assert pyiid.sim.leapfrog is mysim2.leapfrog
