"""A long NUTS run at Au561 (GPU box): rate, acceptance, and the accounting of the
device state slots per iteration."""
import gc
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, Calc1D, structures, sim

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 150
level = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ideal = structures.icosahedron("Au", 5)
scat = ElasticScatter(precision="fp32", device=0)
target = scat.get_pdf(ideal)
atoms = structures.icosahedron("Au", 5)
atoms.positions *= 1.05
calc = Calc1D(target_data=target, exp_function=scat.get_pdf, exp_grad_function=scat.get_grad_pdf,
              conv=100, potential="rw")
atoms.set_calculator(calc)
atoms.get_forces()
np.random.seed(0)
ens = sim.NUTSCanonicalEnsemble(atoms, temperature=1000, escape_level=level, seed=0, fast=True,
                                device_states=True)
pool = scat.pdf_backend._slot_pool if hasattr(scat.pdf_backend, '_slot_pool') else None
t = time.perf_counter()
low = 10 ** 9
for it in range(iters):
    lf0 = ens.leapfrogs
    try:
        ens.run(1)
    except RuntimeError as exc:
        live = sum(1 for o in gc.get_objects() if isinstance(o, sim._DevState))
        print('iteration %d failed after %d leapfrogs: %s; live device states %d, free slots %d' % (
            it, ens.leapfrogs - lf0, exc, live, len(scat.pdf_backend._slot_pool.free)))
        raise
    pool = scat.pdf_backend._slot_pool
    low = min(low, len(pool.free))
    if it % 25 == 0:
        live = sum(1 for o in gc.get_objects() if isinstance(o, sim._DevState))
        print('iteration %3d: %5d leapfrogs, free slots %3d, live device states %d, step %.4f' % (
            it, ens.leapfrogs - lf0, len(pool.free), live, ens.step_size))
dt = time.perf_counter() - t
print("%d leapfrogs in %.2f s = %.0f /s; accepted %d; free slots now %d, lowest seen between "
      "iterations %d of %d" % (ens.leapfrogs, dt, ens.leapfrogs / dt, ens.metadata["accepted_samples"],
                              len(pool.free), low, 256))
