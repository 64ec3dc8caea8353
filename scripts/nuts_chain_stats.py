"""NUTS at Au561 (GPU box): how the look-ahead chains of _DeviceSystem are used
-- native calls, steps computed, steps consumed -- and the leapfrog rate with and
without the look-ahead."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyiid_b200 import ElasticScatter, Calc1D, structures, sim
from pyiid_b200.backend import Backend

ideal = structures.icosahedron('Au', 5)
scat = ElasticScatter(precision='fp32', device=0)
target = scat.get_pdf(ideal)
atoms = structures.icosahedron('Au', 5)
atoms.positions *= 1.05
calc = Calc1D(target_data=target, exp_function=scat.get_pdf,
              exp_grad_function=scat.get_grad_pdf, conv=100, potential='rw')
atoms.set_calculator(calc)
atoms.get_forces()

for chain in (1, 16, 64):
    sim._DeviceSystem.CHAIN = chain
    np.random.seed(0)
    a = atoms.copy()
    a.set_calculator(calc)
    ens = sim.NUTSCanonicalEnsemble(a, temperature=1000, escape_level=8, seed=0, fast=True,
                                    device_states=True)
    calls = []
    orig = Backend.leapfrog_chain

    def counted(self, src, dsts, *args, **kw):
        calls.append(len(dsts))
        return orig(self, src, dsts, *args, **kw)
    Backend.leapfrog_chain = counted
    ens.run(1)
    del calls[:]
    lf0, t = ens.leapfrogs, time.perf_counter()
    ens.run(12)
    dt = time.perf_counter() - t
    used = ens.leapfrogs - lf0
    print('CHAIN %2d: %5d leapfrogs used, %5d computed in %4d native calls, %.0f leapfrogs/s, '
          '%.1f us per used leapfrog' % (chain, used, sum(calls), len(calls), used / dt,
                                         dt / used * 1e6))
    Backend.leapfrog_chain = orig

# where the rest of an iteration goes
import cProfile
import pstats
sim._DeviceSystem.CHAIN = 16
np.random.seed(0)
a = atoms.copy()
a.set_calculator(calc)
ens = sim.NUTSCanonicalEnsemble(a, temperature=1000, escape_level=8, seed=0, fast=True,
                                device_states=True)
ens.run(1)
pr = cProfile.Profile()
pr.enable()
ens.run(12)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
