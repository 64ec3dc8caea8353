"""Import-path alias: ``pyiid.*`` names of the reference resolved to the B200
implementation in ``pyiid_b200`` (option A)."""
import pyiid_b200  # noqa: F401  (installs the ASE stand-ins when ASE is absent)
