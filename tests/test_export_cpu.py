"""CPU: the plotting-free field writers (lbm_b200/export.py)."""
import numpy as np

from lbm_b200 import export


def test_raw_round_trip(tmp_path):
    rng = np.random.default_rng(5)
    for dt in (np.float64, np.float32):
        a = rng.standard_normal((7, 5)).astype(dt)
        p = tmp_path / ("f_%s.raw" % np.dtype(dt).name)
        export.write_raw(p, a)
        b = export.read_raw(p)
        assert b.dtype == a.dtype and np.array_equal(a, b)


def test_vtk_layout(tmp_path):
    nx, ny = 4, 3
    s = np.arange(nx * ny, dtype=np.float64).reshape(nx, ny)
    u = np.stack([s + 100.0, s + 200.0])
    p = tmp_path / "step.vtk"
    export.write_vtk(p, {"speed": s, "u": u}, dx=0.5, origin=(1.0, 2.0))
    raw = p.read_bytes()
    head, rest = raw.split(b"LOOKUP_TABLE default\n", 1)
    assert b"DIMENSIONS 4 3 1" in head and b"SPACING 0.5 0.5 1" in head and b"POINT_DATA 12" in head
    vals = np.frombuffer(rest[:8 * nx * ny], dtype=">f8").reshape(ny, nx)      # x fastest in the file
    assert np.array_equal(vals, s.T)
    vec = rest.split(b"VECTORS u double\n", 1)[1]
    v = np.frombuffer(vec[:8 * 3 * nx * ny], dtype=">f8").reshape(ny, nx, 3)
    assert np.array_equal(v[:, :, 0], u[0].T) and np.array_equal(v[:, :, 1], u[1].T) and not v[:, :, 2].any()
