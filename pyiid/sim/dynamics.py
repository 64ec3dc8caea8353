"""``pyiid.sim.dynamics`` -> :mod:`pyiid_b200.sim`."""
from pyiid_b200.sim import classical_dynamics  # noqa: F401
