"""Block 0's phase times of the fused leapfrog launch at Au561 (GPU box):
IID_FUSED_STAMPS=1 python scripts/lf_phase_times.py [chain length]."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, Calc1D, structures, sim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ideal = structures.icosahedron('Au', 5)
scat = ElasticScatter(precision='fp32', device=0)
target = scat.get_pdf(ideal)
atoms = structures.icosahedron('Au', 5)
atoms.positions *= 1.05
calc = Calc1D(target_data=target, exp_function=scat.get_pdf, exp_grad_function=scat.get_grad_pdf,
              conv=100, potential='rw')
atoms.set_calculator(calc)
atoms.set_momenta(np.random.RandomState(0).normal(0, 1, (561, 3)))
atoms.get_forces()
dev = sim._DeviceSystem(atoms)
st = dev.state_of(atoms)
be = dev.be
be.set_option("graph", int(os.environ.get("LF_GRAPH", "0")))
slots = [dev.pool.take() for _ in range(65)]
for _ in range(int(os.environ.get("LF_REPS", "400"))):
    be.leapfrog_chain(st.slot, slots[1:1 + n], 1e-3, True, calc.target_data, 'rw', 100.)
