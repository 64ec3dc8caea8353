"""X-ray form factors f_Z(Q) for the scatter-factor arrays.

The reference calls ``xraylib.FF_Rayl(Z, q)`` with ``q = kq*qbin/(4 pi)``
(``pyiid/experiments/elasticscatter/kernels/master_kernel.py:14-36``).  xraylib
is not vendored by the reference and is not installed here, so the table is
restated from its published parametrisation: the Waasmaier-Kirfel (1995)
five-Gaussian fit  f(s) = sum_i a_i exp(-b_i s^2) + c,  s = Q/(4 pi).

Pinning: the reference's only golden vector for this function is
``pyiid/tests/test_master/c60_scat.txt`` (carbon, 250 bins of 0.1 1/A, its
tolerance rtol 1e-2); the carbon row below reproduces it to 3e-9.  For every
other element parity with xraylib is UNPINNED (only sum(a)+c ~= Z is checked).
Form factors are inputs to both the oracle and the kernels in the parity
tests, and cancel exactly from F(Q) for single-element structures.
"""
import numpy as np

# Z: (a1..a5, b1..b5, c)
WK95 = {
    1: ((0.413048, 0.294953, 0.187491, 0.080701, 0.023736),
        (15.569946, 32.398468, 5.711404, 61.889874, 1.334118), 0.000049),
    6: ((2.657506, 1.078079, 1.490909, -4.241070, 0.713791),
        (14.780758, 0.776775, 42.086843, -0.000294, 0.239535), 4.297983),
    8: ((2.960427, 2.508818, 0.637853, 0.722838, 1.142756),
        (14.182259, 5.936858, 0.112726, 34.958481, 0.390240), 0.027014),
    14: ((5.275329, 3.191038, 1.511514, 1.356849, 2.519114),
         (2.631338, 33.730728, 0.081119, 86.288640, 1.170087), 0.145073),
    78: ((31.273891, 18.445441, 17.063745, 5.555933, 1.575270),
         (1.316992, 8.797154, 0.124741, 40.177994, 1.316997), 4.050394),
    79: ((16.777389, 19.317156, 32.979682, 5.595453, 10.576854),
         (0.122737, 8.621570, 1.256902, 38.008821, 0.000601), -6.279078),
}

_custom = {}


def register_form_factor(z, func):
    """Register ``func(q_array) -> f`` (Q in 1/A) for atomic number ``z``."""
    _custom[int(z)] = func


def form_factor(z, q):
    """f_Z(Q) at scattering-vector magnitudes ``q`` (1/A), float64."""
    z = int(z)
    q = np.asarray(q, dtype=np.float64)
    if z in _custom:
        return np.asarray(_custom[z](q), dtype=np.float64)
    if z not in WK95:
        raise KeyError(
            'no form-factor table for Z=%d; add one with '
            'pyiid_b200.formfactors.register_form_factor' % z)
    a, b, c = WK95[z]
    s2 = (q / (4 * np.pi)) ** 2
    f = np.full(q.shape, c, dtype=np.float64)
    for ai, bi in zip(a, b):
        f += ai * np.exp(-bi * s2)
    return f


def get_scatter_array(scatter_array, numbers, qbin):
    """Fill ``scatter_array[i, kq] = f_{numbers[i]}(kq*qbin)`` -- the corrected
    reading of master_kernel.get_scatter_array :14-36 (one row per entry of
    ``numbers``)."""
    nq = scatter_array.shape[1]
    q = np.arange(nq) * qbin
    for i, z in enumerate(numbers):
        scatter_array[i, :] = form_factor(z, q)
    return scatter_array
