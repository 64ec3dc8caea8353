"""``pyiid.experiments.elasticscatter`` -> :mod:`pyiid_b200.elasticscatter`."""
from pyiid_b200.elasticscatter import ElasticScatter, wrap_atoms  # noqa: F401
