"""Golden fixture for isotropic atomic displacement parameters, from the
REFERENCE's own (otherwise unused) ADP kernels.

Run in the build container only (needs /root/reference):

    PYTHONPATH=. python tests/golden/make_golden_adp.py

The reference carries two kernels that apply a Debye-Waller array tau to the
pair terms -- get_adp_fq (kernels/cpu_nxn.py:114-121: fq = norm * omega * tau)
and get_adp_grad_fq (kernels/cpu_flat.py:156-174: grad = norm * (tau *
grad_omega + omega * grad_tau)) -- but builds no tau itself (its wrappers pass
adps = None, cpu_wrappers/flat_multi_cpu_wrap.py:18-19).  Here tau[k, Q] =
exp(-(u_i^2 + u_j^2) Q^2 / 2), grad_tau = 0, is handed to THOSE kernels together
with the reference's own d, r, norm, omega, grad_omega, and the results are
summed and normalised as flat_serial_cpu_wrap.py:52-69,118-129 (float64
normaliser).  Inputs are stored next to the outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
import oracle  # noqa: E402
from pyiid_b200 import structures, formfactors  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
EXP = dict(oracle.DEFAULT_EXP)
QB = EXP['qbin']
NQ = int(np.floor(EXP['qmax'] / QB))


def ref_adp(pos, scat, adps, prec):
    base, flat, exp, mk = ref_shim.kernels(prec)
    dt = np.float32 if prec == 'fp32' else np.float64
    q, sc, d, r, norm, omega = ref_shim.ref_pair_arrays(pos, scat, QB, prec)
    n, nq = sc.shape
    k = n * (n - 1) // 2
    tau = oracle.adp_tau(adps, nq, QB, prec)
    # F(Q): get_adp_fq is an [n, n, Q] kernel (float32 signature only)
    if prec == 'fp32':
        nxn = ref_shim.nxn_kernels()
        i, j = oracle.pair_indices(n)

        def cube(a):  # flat pair list -> symmetric [n, n, Q], zero diagonal
            c = np.zeros((n, n, nq), dt)
            c[i, j] = a
            c[j, i] = a
            return c

        fq3 = np.zeros((n, n, nq), dt)
        nxn.get_adp_fq(fq3, cube(omega), cube(tau), cube(norm))
        s = fq3[i, j].sum(axis=0, dtype=np.float64).astype(dt)  # each pair once
    else:
        s = ((norm * omega) * tau).sum(axis=0, dtype=np.float64)
    na = np.mean(norm, axis=0, dtype=np.float64) * n
    with np.errstate(all='ignore'):
        fq = 2 * np.nan_to_num(s / na)
    # gradient: get_adp_grad_fq on the flat pair list
    go = np.zeros((k, 3, nq), dt)
    flat.get_grad_omega(go, omega, r, d, dt(QB))
    grad = np.zeros((k, 3, nq), dt)
    flat.get_adp_grad_fq(grad, omega, tau, go, np.zeros((k, 3, nq), dt), norm)
    rtn = np.zeros((n, 3, nq), dt)
    exp.experimental_sum_grad_cpu(rtn, grad, 0)
    with np.errstate(all='ignore'):
        rtn = np.nan_to_num(rtn / na)
    return fq, rtn


def main():
    assert ref_shim.available()
    rs = np.random.RandomState(7)
    base = structures.alloy_sphere(24, seed=4)
    pos = base.get_positions() + rs.normal(0, 0.05, (24, 3))
    numbers = np.array(base.get_atomic_numbers())
    scat = np.zeros((24, NQ), np.float32)
    formfactors.get_scatter_array(scat, numbers, QB)
    # one displacement per element, then two atoms of one element set apart
    adps = np.where(numbers == numbers.min(), 0.008, 0.012)
    adps[[3, 11]] = 0.02
    out = dict(positions=pos, numbers=numbers, scatter=scat, adps=adps, qbin=QB)
    for prec, tag in (('fp32', 'f32'), ('fp64', 'f64')):
        fq, grad = ref_adp(pos, scat, adps, prec)
        out['fq_' + tag] = fq
        out['grad_' + tag] = grad
    np.savez_compressed(os.path.join(HERE, 'adp_aupt24.npz'), **out)
    print({k: getattr(v, 'shape', v) for k, v in out.items()})


if __name__ == '__main__':
    main()
