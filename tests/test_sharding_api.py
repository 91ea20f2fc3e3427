"""The `sharding=` argument of time_integration (time_integration.py:61-68): what is accepted in place of the
reference's jax NamedSharding, and what is refused."""
import pytest

import astronomix_b200 as ax
from astronomix_b200 import stepping as ti


def test_slab_sharding_shorthands():
    assert ti._as_slab_sharding(ax.SlabSharding()) == ax.SlabSharding()
    assert ti._as_slab_sharding("x") == ax.SlabSharding()
    assert ti._as_slab_sharding((None, "x", None, None)) == ax.SlabSharding()
    assert ti._as_slab_sharding([None, "x", None]) == ax.SlabSharding()
    s = ax.SlabSharding(local_input=True, local_output=True)
    assert ti._as_slab_sharding(s) is s and s.local_input and s.local_output and s.group is None


@pytest.mark.parametrize("bad", ["y", (None, None, "y", None), ("x", None, None, None), (None, "x", "y", None), 3])
def test_other_decompositions_are_refused(bad):
    with pytest.raises(NotImplementedError):
        ti._as_slab_sharding(bad)


def test_weak_series_geometry():
    """bench.py --gpus N: 512^3 cells per GPU, 8 GPUs = the north star's 1024^3 as 128 x 1024^2 slabs"""
    from astronomix_b200.slab import WEAK_SERIES
    for world, g in WEAK_SERIES.items():
        assert g[0] * g[1] * g[2] == world * 512 ** 3 and g[0] % world == 0
    assert WEAK_SERIES[8] == (1024, 1024, 1024)
