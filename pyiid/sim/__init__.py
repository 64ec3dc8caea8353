"""``pyiid.sim`` -> :mod:`pyiid_b200.sim`."""
from pyiid_b200.sim import leapfrog, Ensemble  # noqa: F401
