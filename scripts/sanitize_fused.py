"""Small run of the fused evaluation kernel for compute-sanitizer (memcheck,
racecheck): plain evaluations with the table and the direct force pass, a
two-element structure, leapfrog chains inside one launch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, Calc1D, structures, sim

for name, atoms in (('Au55', structures.icosahedron('Au', 2)),
                    ('AuPt90', structures.alloy_sphere(90, seed=2))):
    scat = ElasticScatter(precision='fp32', device=0)
    target = scat.get_pdf(atoms)
    a = atoms.copy()
    a.positions *= 1.04
    calc = Calc1D(target_data=target, exp_function=scat.get_pdf,
                  exp_grad_function=scat.get_grad_pdf, conv=100, potential='rw')
    a.set_calculator(calc)
    a.set_momenta(np.random.RandomState(0).normal(0, 1, (len(a), 3)))
    a.get_forces()
    be = scat.pdf_backend
    pos = a.get_positions()
    for table in (1, 0):
        be.set_option('fused_table', table)
        for _ in range(3):
            e = be.energy_forces(pos, calc.target_data, 'rw', 100.)[0]
    be.set_option('fused_table', 1)
    dev = sim._DeviceSystem(a)
    st = dev.state_of(a)
    dev.expect(5)
    s = st
    for _ in range(5):
        s = dev.leapfrog(s, 0.01)
    print(name, e, s.pe, s.ke)
