"""smoke() on the GPU, as the driver runs it."""
import pytest


@pytest.mark.gpu
def test_smoke():
    import __graft_entry__ as entry
    entry.smoke()
