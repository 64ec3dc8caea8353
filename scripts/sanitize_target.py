"""Small run of every kernel variant for compute-sanitizer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyiid_b200 import ElasticScatter, Calc1D, structures
for prec in ('fp32', 'fp64'):
    atoms = structures.alloy_sphere(150, seed=1)
    scat = ElasticScatter(precision=prec)
    fq = scat.get_fq(atoms); g = scat.get_grad_fq(atoms); pdf = scat.get_pdf(atoms)
    gp = scat.get_grad_pdf(atoms)
    target = scat.get_pdf(structures.alloy_sphere(150, seed=1, sigma=0.0))
    a = atoms.copy()
    a.set_calculator(Calc1D(target_data=target, exp_function=scat.get_pdf, exp_grad_function=scat.get_grad_pdf, conv=10.))
    for i in range(4):
        a.positions += 1e-4
        a.positions[0, 0] += 1e-3
        e = a.get_potential_energy(); f = a.get_forces()
    print(prec, float(np.abs(fq).max()), float(np.abs(g).max()), float(np.abs(pdf).max()), float(np.abs(gp).max()), e, float(np.abs(f).max()))
