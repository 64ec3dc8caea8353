"""``pyiid.calc.spring_calc`` -> :mod:`pyiid_b200.spring_calc`."""
from pyiid_b200.spring_calc import *  # noqa: F401,F403
from pyiid_b200.spring_calc import Spring  # noqa: F401
