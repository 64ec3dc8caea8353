"""``pyiid.calc`` -> :mod:`pyiid_b200.calc`."""
from pyiid_b200.calc import (wrap_rw, wrap_chi_sq, wrap_grad_rw,  # noqa: F401
                             wrap_grad_chi_sq)
