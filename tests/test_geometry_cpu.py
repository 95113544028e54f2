"""CPU: the vectorised obstacle preprocessing reproduces the reference's link lists and IBB
distances exactly (fixtures from tests/golden/make_golden.py; counts 234 / 468 are the
reference's own test, lbm/tst/lattice/test_lattice.py:27,36)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from lbm_b200 import geometry


class _Lat:
    IBB = True


class _Obs:
    pass


@pytest.mark.parametrize("name,count", [("turek100", 234), ("turek200", 468), ("array", 928)])
def test_links_identical_to_reference(name, count):
    z = np.load(os.path.join(os.path.dirname(GOLDEN), os.pardir, "lbm_b200", "data", "links_%s.npz" % name))
    lat = _Lat()
    for k in ("nx", "ny"):
        setattr(lat, k, int(z[k]))
    for k in ("x_min", "x_max", "y_min", "y_max", "dx"):
        setattr(lat, k, float(z[k]))
    lat.lattice = np.zeros((lat.nx, lat.ny))
    off, poff = z["offsets"], z["polygon_offsets"]
    total = 0
    for o in range(len(off) - 1):
        obs = _Obs()
        obs.tag = o + 1
        obs.polygon = z["polygon"][poff[o]:poff[o + 1]]
        area, bnd, ibb = geometry.add_obstacle(lat, obs)
        assert np.array_equal(bnd, z["boundary"][off[o]:off[o + 1]])
        assert np.array_equal(ibb, z["ibb"][off[o]:off[o + 1]])
        total += len(bnd)
    assert total == count
    assert np.array_equal(np.argwhere(lat.lattice > 0), z["solid"])
