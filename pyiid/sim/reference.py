"""The REFERENCE's own sampler source on the B200 calculator.

``pyiid/sim/__init__.py`` (leapfrog, Ensemble) and ``pyiid/sim/nuts_hmc.py``
(buildtree, NUTSCanonicalEnsemble) of ZhouHUB/pyIID depend only on ASE + numpy.
Nothing of them is copied into this repository: when a checkout of the
reference is available (``$PYIID_REFERENCE``, default ``/root/reference``)
:func:`load` executes the two files from where they lie, against this
package's ASE stand-ins (or a real ASE) and any calculator -- in particular
``pyiid_b200.Calc1D``.  SURVEY.md section 7 ("Running pyiid.sim unchanged").

    from pyiid.sim import reference
    sim, nuts = reference.load()
    ens = reference.NUTSCanonicalEnsemble(atoms, temperature=300, escape_level=8)

``reference.NUTSCanonicalEnsemble`` is the reference class with ONE method
replaced: its initial step-size search evaluates ``2 ** -a`` with a numpy
integer (``nuts_hmc.py:139-150``), which numpy >= 1.12 rejects under any
calculator; the replacement is the same search (Hoffman & Gelman Alg. 4) with a
float exponent.  Everything else -- ``step``, ``buildtree``, the dual averaging
-- runs as shipped.  ``tests/golden/make_golden_nuts.py`` uses this loader to
produce the trajectory that pins ``pyiid_b200.sim`` to the reference's sampler.
"""
import importlib.util
import os
import sys
import types

import numpy as np

from pyiid_b200 import ase_shim

REF = os.environ.get('PYIID_REFERENCE', '/root/reference')
_loaded = None


def available():
    return os.path.isfile(os.path.join(REF, 'pyiid/sim/nuts_hmc.py'))


def load():
    """(sim module, nuts_hmc module) executed from the reference checkout."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError('no reference checkout at %s (set PYIID_REFERENCE)' % REF)
    ase_shim.install()
    names = ('pyiid', 'pyiid.sim', 'pyiid.sim.nuts_hmc')
    saved = {k: sys.modules.get(k) for k in names}
    pkg = types.ModuleType('pyiid')
    pkg.__path__ = [os.path.join(REF, 'pyiid')]
    sys.modules['pyiid'] = pkg
    try:
        spec = importlib.util.spec_from_file_location(
            'pyiid.sim', os.path.join(REF, 'pyiid/sim/__init__.py'),
            submodule_search_locations=[os.path.join(REF, 'pyiid/sim')])
        sim = importlib.util.module_from_spec(spec)
        sys.modules['pyiid.sim'] = sim
        spec.loader.exec_module(sim)
        spec2 = importlib.util.spec_from_file_location(
            'pyiid.sim.nuts_hmc', os.path.join(REF, 'pyiid/sim/nuts_hmc.py'))
        nuts = importlib.util.module_from_spec(spec2)
        sys.modules['pyiid.sim.nuts_hmc'] = nuts
        spec2.loader.exec_module(nuts)
    finally:
        for k, v in saved.items():  # the alias package stays what `import pyiid` gives
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _loaded = (sim, nuts)
    return _loaded


def __getattr__(name):
    if name == 'NUTSCanonicalEnsemble':
        sim, nuts = load()

        class NUTSCanonicalEnsemble(nuts.NUTSCanonicalEnsemble):
            def _find_step_size(self, input_atoms, thermal_nrg=None, momentum=None):
                from copy import deepcopy as dc
                atoms = dc(input_atoms)
                step_size = .5
                ase_shim.MaxwellBoltzmannDistribution(atoms, temp=thermal_nrg, force_temp=True)
                e0 = atoms.get_total_energy()

                def ratio(eps):
                    with np.errstate(over='ignore'):
                        return np.exp(e0 - sim.leapfrog(atoms, eps).get_total_energy())

                a = 1. if ratio(step_size) > 0.5 else -1.
                while ratio(step_size) ** a > 2. ** -a:
                    step_size *= 2. ** a
                    if step_size < 1e-7 or step_size > 1e7:
                        step_size = 1.
                        break
                return step_size

        globals()['NUTSCanonicalEnsemble'] = NUTSCanonicalEnsemble
        return NUTSCanonicalEnsemble
    raise AttributeError(name)
