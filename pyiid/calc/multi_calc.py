"""``pyiid.calc.multi_calc`` -> :mod:`pyiid_b200.multi_calc`."""
from pyiid_b200.multi_calc import MultiCalc  # noqa: F401
