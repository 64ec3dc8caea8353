"""Where a leapfrog's time goes at Au561 (GPU box): native evaluation, the
device-resident leapfrog call, the two sampler systems, and a cProfile of a NUTS
run on each path."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyiid_b200 import ElasticScatter, Calc1D, structures, sim


def timeit(fn, n=300):
    for _ in range(5):
        fn()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t) / n * 1e6


ideal = structures.icosahedron('Au', 5)
scat = ElasticScatter(precision='fp32')
target = scat.get_pdf(ideal)
atoms = structures.icosahedron('Au', 5)
atoms.positions *= 1.05
calc = Calc1D(target_data=target, exp_function=scat.get_pdf,
              exp_grad_function=scat.get_grad_pdf, conv=100, potential='rw')
atoms.set_calculator(calc)
atoms.set_momenta(np.random.RandomState(0).normal(0, 1, (561, 3)))
atoms.get_forces()
be = scat.pdf_backend
pos = atoms.get_positions()
print('be.energy_forces           %7.1f us' % timeit(lambda: be.energy_forces(pos, target, 'rw', 100.)))
host, dev = sim._FastSystem(atoms), sim._DeviceSystem(atoms)
sh, sd = host.state_of(atoms), dev.state_of(atoms)
a, b = dev.pool.take(), dev.pool.take()
be.state_upload(a, sh.q, sh.p, sh.f)
print('be.leapfrog (native+wrapper) %5.1f us' % timeit(lambda: be.leapfrog(a, b, 1e-3, True, target, 'rw', 100.)))
import ctypes
out = np.empty(9); q = np.empty((561, 3)); p = np.empty((561, 3))
lib, h = be.lib, be.h
print('iid_leapfrog_host (ctypes)  %6.1f us' % timeit(lambda: lib.iid_leapfrog_host(
    h, a, b, 1e-3, 1, None, 0, 100., out.ctypes.data, q.ctypes.data, p.ctypes.data)))
print('iid_leapfrog_host no mirror %6.1f us' % timeit(lambda: lib.iid_leapfrog_host(
    h, a, b, 1e-3, 1, None, 0, 100., out.ctypes.data, None, None)))
f = np.empty((561, 3)); o4 = np.zeros(4)
print('iid_energy_forces_host      %6.1f us' % timeit(lambda: lib.iid_energy_forces_host(
    h, pos.ctypes.data, None, 0, 100., o4.ctypes.data, f.ctypes.data, None)))
print('host system.leapfrog        %6.1f us' % timeit(lambda: host.leapfrog(sh, 1e-3)))
print('device system.leapfrog      %6.1f us' % timeit(lambda: dev.leapfrog(sd, 1e-3)))
print('host evaluate()             %6.1f us' % timeit(lambda: host.evaluate(pos)))
print('u-turn test                 %6.1f us' % timeit(lambda: sim._no_u_turn_states(sh, sh, host.masses)))
print('safe_exp x3                 %6.1f us' % timeit(lambda: (sim._safe_exp(-1.), sim._safe_exp(2.), sim._safe_exp(0.1))))
for devs in (True, False):
    np.random.seed(0)
    a2 = atoms.copy()
    a2.set_calculator(calc)
    ens = sim.NUTSCanonicalEnsemble(a2, temperature=1000, escape_level=8, seed=0, fast=True,
                                    device_states=devs)
    ens.run(2)
    lf0 = ens.leapfrogs
    pr = cProfile.Profile()
    t = time.perf_counter()
    pr.enable()
    ens.run(4)
    pr.disable()
    dt = time.perf_counter() - t
    print('NUTS device_states=%s: %.1f us per leapfrog (under cProfile), %d leapfrogs' % (
        devs, dt / (ens.leapfrogs - lf0) * 1e6, ens.leapfrogs - lf0))
    pstats.Stats(pr).sort_stats('tottime').print_stats(12)
