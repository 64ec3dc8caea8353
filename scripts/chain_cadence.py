"""When the steps of a 16-step chain become available to the host (GPU box):
time of iid_leapfrog_chain_begin and of each iid_leapfrog_chain_next."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, Calc1D, structures, sim

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
slots = [dev.pool.take() for _ in range(65)]
tgt = calc.target_data
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
acc = np.zeros(n + 1)
reps = 200
for rep in range(reps + 5):
    t0 = time.perf_counter()
    cid = be.leapfrog_chain_begin(st.slot, slots[1:1 + n], 1e-3, True, tgt, 'rw', 100.)
    ts = [time.perf_counter()]
    for i in range(n):
        be.leapfrog_chain_next(cid)
        ts.append(time.perf_counter())
    if rep >= 5:
        acc += np.array(ts) - t0
acc *= 1e6 / reps
print('begin returns after %.1f us; steps available after (us):' % acc[0])
print(' '.join('%.0f' % v for v in acc[1:]))
print('increments:', ' '.join('%.1f' % v for v in np.diff(acc)))
