"""Stress of the per-step completion flags (GPU box): every step of many chains,
collected while the launch is still running, must equal the state the device
slab holds once the chain has ended."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, Calc1D, structures, sim

shells = int(sys.argv[1]) if len(sys.argv) > 1 else 5
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 300
n = 32
ideal = structures.icosahedron('Au', shells)
scat = ElasticScatter(precision='fp32', device=0)
target = scat.get_pdf(ideal)
atoms = structures.icosahedron('Au', shells)
atoms.positions *= 1.05
calc = Calc1D(target_data=target, exp_function=scat.get_pdf, exp_grad_function=scat.get_grad_pdf,
              conv=100, potential='rw')
atoms.set_calculator(calc)
atoms.set_momenta(np.random.RandomState(0).normal(0, 1, (len(atoms), 3)))
atoms.get_forces()
dev = sim._DeviceSystem(atoms)
st = dev.state_of(atoms)
be = dev.be
slots = [dev.pool.take() for _ in range(n)]
tgt = calc.target_data
bad = 0
rs = np.random.RandomState(1)
t0 = time.perf_counter()
for c in range(chains):
    step = 1e-3 * (1 + c % 7)
    cid = be.leapfrog_chain_begin(st.slot, slots, step, True, tgt, 'rw', 100.)
    got = []
    for i in range(n):
        e, sc, es, ke, q, p = be.leapfrog_chain_next(cid)
        got.append((q.copy(), p.copy(), float(e)))
        if rs.rand() < 0.3:
            time.sleep(rs.rand() * 1e-4)  # the host falls behind now and then
    for i in range(n):
        d = be.state_download(slots[i], want=('q', 'p'))
        if not (np.array_equal(d['q'], got[i][0]) and np.array_equal(d['p'], got[i][1])):
            bad += 1
            print('chain %d step %d: mirror differs from the slab (max |dq| %.3e, |dp| %.3e)' % (
                c, i, np.abs(d['q'] - got[i][0]).max(), np.abs(d['p'] - got[i][1]).max()))
print('%d chains x %d steps, %d mismatches, %.1f s' % (chains, n, bad, time.perf_counter() - t0))
