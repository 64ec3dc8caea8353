"""pyiid_b200 -- the pyIID elastic-scattering hot path on NVIDIA B200.

Public surface (mirrors the reference's import paths, see also the ``pyiid``
alias package at the repository root):

* :class:`pyiid_b200.elasticscatter.ElasticScatter`
  (reference ``pyiid.experiments.elasticscatter.ElasticScatter``)
* :class:`pyiid_b200.calc.Calc1D`, :func:`pyiid_b200.calc.PDFCalc`
  (reference ``pyiid.calc.calc_1d.Calc1D``)
* :class:`pyiid_b200.spring_calc.Spring`, :class:`pyiid_b200.multi_calc.MultiCalc`
  (reference ``pyiid.calc.spring_calc`` / ``pyiid.calc.multi_calc``)
* :mod:`pyiid_b200.sim` -- ``leapfrog``, ``NUTSCanonicalEnsemble``
  (reference ``pyiid.sim``)

All numerical work runs in hand-written sm_100a CUDA behind the C ABI in
``include/iid_b200.h`` (``pyiid_b200/libiid_b200.so``).  There is no CPU
fallback: without the built library or without an sm_100 GPU the compute
entry points raise.
"""
from . import ase_shim

ase_shim.install()

from .elasticscatter import ElasticScatter, wrap_atoms  # noqa: E402
from .calc import Calc1D, PDFCalc  # noqa: E402
from .spring_calc import Spring  # noqa: E402
from .multi_calc import MultiCalc  # noqa: E402

__all__ = ['ElasticScatter', 'wrap_atoms', 'Calc1D', 'PDFCalc', 'Spring', 'MultiCalc']
__version__ = '0.1.0'
