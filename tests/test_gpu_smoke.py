"""The driver's entry points: build() leaves a loadable library, smoke() runs the hot path once against the oracle."""
import pytest


@pytest.mark.gpu
def test_smoke_entry_point_runs():
    import __graft_entry__ as g
    g.build()
    g.smoke()
