import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from pyiid_b200 import ElasticScatter, Calc1D, structures, sim
ideal = structures.icosahedron('Au', 5)
scat = ElasticScatter(precision='fp32', device=0)
target = scat.get_pdf(ideal)
atoms = structures.icosahedron('Au', 5); atoms.positions *= 1.05
calc = Calc1D(target_data=target, exp_function=scat.get_pdf, exp_grad_function=scat.get_grad_pdf, conv=100, potential='rw')
atoms.set_calculator(calc)
atoms.set_momenta(np.random.RandomState(0).normal(0, 1, (561, 3)))
atoms.get_forces()
dev = sim._DeviceSystem(atoms)
st = dev.state_of(atoms)
be = dev.be
slots = [dev.pool.take() for _ in range(65)]
tgt = calc.target_data
for n in (1, 2, 4, 8, 16, 64):
    for _ in range(6):
        be.leapfrog_chain(st.slot, slots[1:1 + n], 1e-3, True, tgt, 'rw', 100.)
    t = time.perf_counter(); reps = 100
    for _ in range(reps):
        be.leapfrog_chain(st.slot, slots[1:1 + n], 1e-3, True, tgt, 'rw', 100.)
    dt = (time.perf_counter() - t) / reps
    print('chain of %2d: %.1f us per call, %.1f us per step' % (n, dt * 1e6, dt * 1e6 / n))
# through the system with look-ahead
for exp in (1, 8, 64):
    s0 = st
    for rep in range(4):  # graph capture of this chain length
        dev.expect(exp)
        s = s0
        for k in range(exp):
            s = dev.leapfrog(s, 1e-3)
    t = time.perf_counter(); cnt = 0
    for rep in range(40):
        dev.expect(exp)
        s = s0
        for k in range(exp):
            s = dev.leapfrog(s, 1e-3); cnt += 1
    print('system.leapfrog with expect(%d): %.1f us per step' % (exp, (time.perf_counter() - t) / cnt * 1e6))
