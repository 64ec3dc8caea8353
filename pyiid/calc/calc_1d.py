"""``pyiid.calc.calc_1d`` -> :mod:`pyiid_b200.calc`."""
from pyiid_b200.calc import Calc1D, PDFCalc  # noqa: F401
