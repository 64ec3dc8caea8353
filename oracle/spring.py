"""CPU restatement of the reference's spring restraint potentials.

TEST INFRASTRUCTURE ONLY (same rule as the rest of ``oracle/``): imported by
``tests/`` as the checker of the CUDA pair kernels in
``pyiid_b200/csrc/iid_spring.cuh``; never by the product path.

Follows ``pyiid/calc/spring_calc.py`` of the reference statement by statement,
including its mixed precision (float32 positions / differences / distances,
float64 energies and forces), on explicit position arrays instead of ASE
``Atoms``.  The N x N distance arrays come from the restatement of
``kernels/cpu_nxn.py:17-52`` below.  ``precision='fp64'`` is the same code
with float64 in place of float32 (what FP64-mode handles compute).

Pinned against the reference itself: ``tests/golden/make_golden.py`` runs the
reference's own ``spring_nrg`` / ``spring_force`` / ``atomwise_*`` functions
(loaded through ``oracle/ref_shim.py``) and stores their outputs in
``tests/golden/springs.npz``; ``tests/test_oracle.py`` checks this module
against them.  The reference's ``voxel_*`` functions cannot run under the
installed numpy (``np.zeros(c / resolution)`` with a float shape,
spring_calc.py:152): the voxel restatement is pinned by the reference's own
test property instead (voxel energy == energy change on adding an atom there,
tests/test_calc/test_spring.py:17-42).
"""
import numpy as np


def _dt(precision):
    return np.float32 if precision == 'fp32' else np.float64


def nxn_d_r(positions, precision='fp32'):
    """d[i,j] = q_j - q_i and r[i,j] (cpu_nxn.py:17-52): sequential
    accumulation of the three squares in the working precision."""
    dt = _dt(precision)
    q = np.asarray(positions, np.float64).astype(dt)
    d = (q[None, :, :] - q[:, None, :]).astype(dt)
    tmp = np.zeros(d.shape[:2], dt)
    for w in range(3):
        tmp = (tmp + d[:, :, w] * d[:, :, w]).astype(dt)
    return q, d, np.sqrt(tmp).astype(dt)


def _thresh(r, rt, sp_type):
    # spring_calc.py:115-117 (rep: np.less), :277-279 (att: np.greater)
    t = np.less(r, r.dtype.type(rt)) if sp_type == 'rep' else \
        np.greater(r, r.dtype.type(rt))
    t[np.diag_indices(len(r))] = False
    return t


def _mag(r, thresh, k, rt):
    # :119-120: the float32 product k * (r - rt) stored in a float64 array
    dt = r.dtype.type
    mag = np.zeros(r.shape)
    mag[thresh] = dt(k) * (r[thresh] - dt(rt))
    return mag


def pair_energy(positions, k, rt, sp_type='rep', precision='fp32'):
    """spring_nrg :107-123 / att_spring_nrg :269-285."""
    q, d, r = nxn_d_r(positions, precision)
    thresh = _thresh(r, rt, sp_type)
    mag = _mag(r, thresh, k, rt)
    return float(np.sum(mag[thresh] / 2. *
                        (r[thresh] - r.dtype.type(rt)).astype(np.float64)))


def pair_force(positions, k, rt, sp_type='rep', precision='fp32'):
    """spring_force :126-147 / att_spring_force :288-311."""
    q, d, r = nxn_d_r(positions, precision)
    n = len(q)
    thresh = _thresh(r, rt, sp_type)
    mag = _mag(r, thresh, k, rt)
    direction = np.zeros((n, n, 3))
    with np.errstate(all='ignore'):
        for tz in range(3):
            direction[thresh, tz] = (d[thresh, tz] / r[thresh]) * mag[thresh]
    direction[np.isnan(direction)] = 0.0
    return np.sum(direction, axis=1)


def pair_atomwise(positions, k, rt, sp_type='rep', precision='fp32'):
    """atomwise_spring_nrg :171-185 / atomwise_att_spring_nrg :320-332."""
    q, d, r = nxn_d_r(positions, precision)
    dt = r.dtype.type
    nrg = dt(.5 * k) * (r - dt(rt)) ** 2
    if sp_type == 'rep':
        nrg[np.where(r > dt(rt))] = 0.0
        nrg[np.diag_indices(len(r))] = 0.0
    else:
        nrg[np.where(r < dt(rt))] = 0.0
    return -np.sum(nrg, axis=0) * 2  # a float32 column sum in the reference


def com_energy(positions, com, k, rt, precision='fp32'):
    """com_spring_nrg :188-198."""
    q = np.asarray(positions, np.float64).astype(_dt(precision))
    disp = q - np.asarray(com, np.float64)
    dist = np.sqrt(np.sum(disp ** 2, axis=1))
    thresh = np.greater(dist, rt)
    mag = np.zeros(len(q))
    mag[thresh] = k * (dist[thresh] - rt)
    return float(np.sum(mag[thresh] / 2. * (dist[thresh] - rt)))


def com_force(positions, com, k, rt, precision='fp32'):
    """com_spring_force :201-217."""
    q = np.asarray(positions, np.float64).astype(_dt(precision))
    disp = q - np.asarray(com, np.float64)
    dist = np.sqrt(np.sum(disp ** 2, axis=1))
    thresh = np.greater(dist, rt)
    mag = np.zeros(len(q))
    mag[thresh] = k * (dist[thresh] - rt)
    direction = np.zeros(q.shape)
    for tz in range(3):
        direction[thresh, tz] = disp[thresh, tz] / dist[thresh] * mag[thresh]
    return direction * -1.


def com_atomwise(positions, com, k, rt, precision='fp32'):
    """atomwise_com_spring_nrg :259-266 (a scalar in the reference)."""
    q = np.asarray(positions, np.float64).astype(_dt(precision))
    disp = q - np.asarray(com, np.float64)
    dist = np.sqrt(np.sum(disp ** 2, axis=1))
    nrg = .5 * k * (dist - rt) ** 2
    nrg[np.where(dist < rt)] = 0.0
    return np.sum(nrg, axis=0) * 2


def voxel_energy(positions, k_const, rt, resolution, shape, sp_type='rep',
                 com=None, precision='fp32'):
    """voxel_spring_nrg :150-168, voxel_com_spring_nrg :240-256,
    voxel_att_spring_nrg :314-332, with an integer grid shape.  Voxel centres
    are Python floats, positions float32 scalars: float64 arithmetic."""
    q = np.asarray(positions, np.float64).astype(_dt(precision)).astype(np.float64)
    im, jm, km = shape
    x = (np.arange(im) + .5) * resolution
    y = (np.arange(jm) + .5) * resolution
    z = (np.arange(km) + .5) * resolution
    X, Y, Z = np.meshgrid(x, y, z, indexing='ij')
    vox = np.zeros(shape)
    if sp_type == 'com':
        c = np.asarray(com, np.float64)
        temp = np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2)
        hit = temp > rt
        vox[hit] = .5 * k_const * (temp[hit] - rt) ** 2
        return vox * 2
    for l in range(len(q)):
        temp = np.sqrt((X - q[l, 0]) ** 2 + (Y - q[l, 1]) ** 2 + (Z - q[l, 2]) ** 2)
        hit = temp < rt if sp_type == 'rep' else temp > rt
        vox[hit] += .5 * k_const * (temp[hit] - rt) ** 2
    return vox * 2
