import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# On a box with several GPUs ElasticScatter() would take all of them (the
# one-process multi-GPU handle); the device-level checks of this suite address
# ONE handle on ONE device.  The multi-GPU tests ask for 'Multi-GPU' explicitly.
os.environ.setdefault('IID_PROCESSOR', 'B200')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu)')


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a B200: skip (not fail) the gpu tests."""
    if has_gpu():
        return
    skip = pytest.mark.skip(reason='no sm_100 GPU on this box')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def has_gpu():
    try:
        import ctypes
        from pyiid_b200 import _lib
        lib = _lib.load()
        n = ctypes.c_int(0)
        return lib.iid_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


@pytest.fixture(scope='session', autouse=True)
def _built():
    """Build the CUDA library and the oracle once per session (nvcc and gcc
    cross-compile without a GPU)."""
    import __graft_entry__ as entry
    entry.build()


def golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        return {k: z[k] for k in z.files}


def nerr(a, b):
    """max |a-b| / max |b|: the normalised max-norm used for every floating
    point parity bound (F(Q) and the gradients cross zero, so element-wise
    relative error is meaningless)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


# Tolerances.  north_star: FP32 mode matches the reference's CPU path to rel
# 1e-5 on F(Q) / G(r) / forces, FP64 mode to 1e-10.
TOL32 = 1e-5
TOL64 = 1e-10
# The full gradient array has no Q-averaging; the float32 reference itself is
# only this close to its own float64 variant (measured: 0.7e-5 at N=55,
# 1.8e-5 at N=1000; it grows like sqrt(N) Q^2 2^-24), so parity of grad F(Q)
# against the float32 oracle is bounded by the reference's own noise floor.
TOL32_GRAD_VS_F32 = 5e-5
