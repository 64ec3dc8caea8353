"""X-ray form factors f_Z(Q) for the scatter-factor arrays.

The reference calls ``xraylib.FF_Rayl(Z, q)`` with ``q = kq*qbin/(4 pi)``
(``pyiid/experiments/elasticscatter/kernels/master_kernel.py:14-36``) and so
accepts any element.  xraylib is not vendored by the reference and is not
installed here; its ``FF_Rayl`` evaluates the Waasmaier-Kirfel (1995)
five-Gaussian fit  f(s) = sum_i a_i exp(-b_i s^2) + c,  s = Q/(4 pi).  Sources,
in this order:

1. a function registered with :func:`register_form_factor`;
2. ``xraylib.FF_Rayl`` itself when xraylib is importable (exact reference
   behaviour; set ``IID_NO_XRAYLIB=1`` to skip it);
3. a table loaded with :func:`load_table` / the ``IID_FORMFACTOR_TABLE``
   environment variable (same column layout as the embedded file);
4. the embedded table ``data/f0_waaskirf.txt``, Z = 1 .. 98.

Pinning: the reference's only golden vector for this function is
``pyiid/tests/test_master/c60_scat.txt`` (carbon, 250 bins of 0.1 1/A, its
tolerance rtol 1e-2); the carbon row reproduces it to 3e-9.  Every embedded
row satisfies sum(a) + c = Z to 0.04 and gives a positive, monotone f(Q) that
varies smoothly with Z (``tests/test_formfactors.py``); beyond that, parity of
the other elements with xraylib is UNPINNED (no second source in this image).
Form factors are inputs to both the oracle and the kernels in the parity
tests, and cancel exactly from F(Q) for single-element structures.
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
TABLE_PATH = os.path.join(_HERE, 'data', 'f0_waaskirf.txt')


def read_table(path):
    """{Z: (a[5], b[5], c)} from a text file with the columns
    ``Z symbol a1..a5 c b1..b5`` (``#`` comments)."""
    table = {}
    with open(path) as fh:
        for line in fh:
            line = line.split('#', 1)[0].split()
            if not line:
                continue
            if len(line) != 13:
                raise ValueError('form-factor table %s: expected 13 columns, got %r'
                                 % (path, line))
            v = [float(x) for x in line[2:]]
            table[int(line[0])] = (tuple(v[0:5]), tuple(v[6:11]), v[5])
    return table


WK95 = read_table(TABLE_PATH)
if os.environ.get('IID_FORMFACTOR_TABLE'):
    WK95.update(read_table(os.environ['IID_FORMFACTOR_TABLE']))

_custom = {}
_xraylib = None


def _get_xraylib():
    global _xraylib
    if _xraylib is None:
        _xraylib = False
        if not os.environ.get('IID_NO_XRAYLIB'):
            try:
                import xraylib
                _xraylib = xraylib
            except ImportError:
                pass
    return _xraylib


def register_form_factor(z, func):
    """Register ``func(q_array) -> f`` (Q in 1/A) for atomic number ``z``."""
    _custom[int(z)] = func


def load_table(path):
    """Add / replace rows of the five-Gaussian table from ``path``."""
    WK95.update(read_table(path))


def source(z):
    """Which of the sources above serves atomic number ``z``."""
    z = int(z)
    if z in _custom:
        return 'registered'
    if _get_xraylib():
        return 'xraylib'
    if z in WK95:
        return 'table'
    return None


def form_factor(z, q):
    """f_Z(Q) at scattering-vector magnitudes ``q`` (1/A), float64."""
    z = int(z)
    q = np.asarray(q, dtype=np.float64)
    if z in _custom:
        return np.asarray(_custom[z](q), dtype=np.float64)
    xrl = _get_xraylib()
    if xrl:
        # master_kernel.get_scatter_array :35-36: FF_Rayl(Z, Q / 4 pi)
        flat = [xrl.FF_Rayl(z, float(v) / (4 * np.pi)) for v in q.ravel()]
        return np.asarray(flat, dtype=np.float64).reshape(q.shape)
    if z not in WK95:
        raise KeyError(
            'no form-factor table for Z=%d (embedded: Z = 1..98); add one with '
            'pyiid_b200.formfactors.register_form_factor or load_table' % z)
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
