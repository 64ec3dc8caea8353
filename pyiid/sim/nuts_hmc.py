"""``pyiid.sim.nuts_hmc`` -> :mod:`pyiid_b200.sim`."""
from pyiid_b200.sim import NUTSCanonicalEnsemble, buildtree, Emax  # noqa: F401
