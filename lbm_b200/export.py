"""Field export for the output step of the apps.

The reference draws |u| with matplotlib (lbm/src/plot/plot.py:9-32: plot_norm -> png); matplotlib is a
host-side dependency that stays where it is.  What the GPU path adds is (a) the speed field computed
on the device (lattice.speed(), one plane to the host instead of lattice.u's two + the mask logic),
and (b) writers that need no plotting package: a raw dump and a legacy-VTK structured-points file
that ParaView / VisIt open directly.  Arrays are [nx, ny] as in the reference (x first)."""
import numpy as np


def write_raw(path, field):
    """field[nx, ny] -> little-endian binary, C order, preceded by a one-line ASCII header."""
    a = np.ascontiguousarray(field)
    with open(path, "wb") as f:
        f.write(("# lbm_b200 raw %s %d %d\n" % (a.dtype.str, a.shape[0], a.shape[1])).encode())
        f.write(a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes())


def read_raw(path):
    with open(path, "rb") as f:
        _, _, _, dt, nx, ny = f.readline().decode().split()
        return np.frombuffer(f.read(), dtype=np.dtype(dt)).reshape(int(nx), int(ny)).copy()


def write_vtk(path, fields, dx=1.0, origin=(0.0, 0.0)):
    """Legacy VTK (binary STRUCTURED_POINTS).  fields: name -> [nx, ny] scalar or [2, nx, ny] vector."""
    first = next(iter(fields.values()))
    nx, ny = (first.shape[-2], first.shape[-1])
    with open(path, "wb") as f:
        f.write(b"# vtk DataFile Version 3.0\nlbm_b200\nBINARY\nDATASET STRUCTURED_POINTS\n")
        f.write(("DIMENSIONS %d %d 1\nORIGIN %.17g %.17g 0\nSPACING %.17g %.17g 1\nPOINT_DATA %d\n"
                 % (nx, ny, origin[0], origin[1], dx, dx, nx * ny)).encode())
        for name, a in fields.items():
            a = np.asarray(a, dtype=np.float64)
            if a.ndim == 2:                      # VTK runs x fastest: transpose the [nx, ny] array
                f.write(("SCALARS %s double 1\nLOOKUP_TABLE default\n" % name).encode())
                f.write(np.ascontiguousarray(a.T).astype(">f8").tobytes())
            else:
                v = np.zeros((ny, nx, 3))
                v[:, :, 0], v[:, :, 1] = a[0].T, a[1].T
                f.write(("VECTORS %s double\n" % name).encode())
                f.write(v.astype(">f8").tobytes())
            f.write(b"\n")


def write_step(lattice, path):
    """One output step of a run: speed (with the obstacle mask), density and velocity of the last macro()."""
    write_vtk(path, {"speed": lattice.speed(), "rho": lattice.rho, "u": lattice.u}, dx=lattice.dx,
              origin=(getattr(lattice, "x_min", 0.0), getattr(lattice, "y_min", 0.0)))
