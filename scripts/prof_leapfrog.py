"""ncu target: a few device-resident leapfrog steps at Au561 (configs[1])."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, Calc1D, structures, sim
scat = ElasticScatter()
ideal = structures.icosahedron('Au', 5)
target = scat.get_pdf(ideal)
atoms = structures.icosahedron('Au', 5)
atoms.positions *= 1.05
atoms.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf,
                            exp_grad_function=scat.get_grad_pdf, conv=100, potential='rw'))
atoms.set_momenta(np.random.RandomState(0).normal(0, 1, (561, 3)))
atoms.get_forces()
dev = sim._DeviceSystem(atoms)
st = dev.state_of(atoms)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(4):  # eager, eager, capture, replay
    st = dev.leapfrog(st, 1e-3)
t = time.perf_counter()
for i in range(n):
    st = dev.leapfrog(st, 1e-3)
print('us per leapfrog', (time.perf_counter() - t) / n * 1e6)
