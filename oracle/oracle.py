"""ctypes front-end of the CPU oracle (oracle/lbm_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke()
and the cpu_baseline / --impl reference legs of bench.py -- never by lbm_b200/.

``OracleLattice`` exposes the method/attribute surface of the reference's
``lattice`` class (/root/reference/lbm/src/core/lattice.py:15-286) over host
NumPy arrays, one C call per reference ``nb_*`` call, so that the reference's
driver loop (lbm/src/core/run.py:24-54) and app callbacks can run on it
unmodified and its results can be compared array by array.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblbm_oracle.so")
_lib = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_i64 = ctypes.c_int64
_dbl = ctypes.c_double


def build(force=False):
    """Compile liblbm_oracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "lbm_oracle.c")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liblbm_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = ctypes.CDLL(_LIB_PATH)
    L.orc_abi_version.restype = ctypes.c_int
    L.orc_get_threads.restype = ctypes.c_int
    L.orc_set_threads.argtypes = [ctypes.c_int]
    L.orc_macro.argtypes = [_i64, _i64, _f64p, _f64p, _f64p]
    L.orc_equilibrium.argtypes = [_i64, _i64, _f64p, _f64p, _f64p]
    L.orc_col_str.argtypes = [_i64, _i64, _f64p, _f64p, _f64p, _dbl, _dbl]
    L.orc_drag_lift.argtypes = [_i64, _i64, _i64, _i64p, _f64p, _f64p, _dbl, _dbl, _dbl, _f64p]
    L.orc_bounce_back_obstacle.argtypes = [_i64, _i64, ctypes.c_int, _i64, _i64p, _f64p, _f64p, _f64p]
    for name in ("left_wall_velocity", "right_wall_velocity", "top_wall_velocity",
                 "bottom_wall_velocity"):
        getattr(L, "orc_zou_he_" + name).argtypes = [_i64, _i64, _f64p, _f64p, _f64p, _f64p]
    L.orc_zou_he_right_wall_pressure.argtypes = [_i64, _i64, _f64p, _f64p, _f64p, _f64p, _f64p]
    for name in ("bottom_left", "top_left", "top_right", "bottom_right"):
        getattr(L, "orc_zou_he_%s_corner" % name).argtypes = [_i64, _i64, _f64p, _f64p, _f64p]
    _lib = L
    return L


def set_threads(n):
    lib().orc_set_threads(int(n))


def get_threads():
    return int(lib().orc_get_threads())


class OracleLattice:
    """Host-side lattice with the reference's names (lattice.py:117-174)."""

    def __init__(self, app):
        L = lib()
        self._L = L
        self.nx = int(app.nx)
        self.ny = int(app.ny)
        self.lx = self.nx - 1
        self.ly = self.ny - 1
        self.q = 9
        for k, default in (("name", "lattice"), ("x_min", 0.0), ("x_max", 1.0), ("y_min", 0.0),
                           ("y_max", 1.0), ("tau_lbm", 1.0), ("dx", 1.0), ("dt", 1.0),
                           ("u_lbm", 0.03), ("L_lbm", 100), ("rho_lbm", 1.0), ("IBB", False),
                           ("stop", "it"), ("it_max", 1000), ("obs_cv_ct", 1.0e-1),
                           ("obs_cv_nb", 500)):
            setattr(self, k, getattr(app, k, default))
        # TRT rates, lattice.py:127-131
        self.tau_p_lbm = self.tau_lbm
        self.lambda_trt = 1.0 / 4.0
        self.tau_m_lbm = self.lambda_trt / (self.tau_p_lbm - 0.5) + 0.5
        self.om_p_lbm = 1.0 / self.tau_p_lbm
        self.om_m_lbm = 1.0 / self.tau_m_lbm
        # tables, lattice.py:135-152
        self.c = np.array([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1],
                           [1, 1], [-1, -1], [-1, 1], [1, -1]], dtype=np.int64)
        self.w = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
        self.ns = np.array([0, 2, 1, 4, 3, 6, 5, 8, 7], dtype=np.int64)
        nx, ny = self.nx, self.ny
        self.g = np.zeros((9, nx, ny))
        self.g_eq = np.zeros((9, nx, ny))
        self.g_up = np.zeros((9, nx, ny))
        self.u_left = np.zeros((2, ny))
        self.u_right = np.zeros((2, ny))
        self.u_top = np.zeros((2, nx))
        self.u_bot = np.zeros((2, nx))
        self.rho_right = np.zeros(ny)
        self.lattice = np.zeros((nx, ny))
        self.rho = np.ones((nx, ny))
        self.u = np.zeros((2, nx, ny))
        self.output_dir = getattr(app, "output_dir", "./")
        self.png_dir = self.output_dir

    # --- phases -----------------------------------------------------------
    def macro(self):
        self._L.orc_macro(self.nx, self.ny, self.g, self.rho, self.u)

    def equilibrium(self):
        self._L.orc_equilibrium(self.nx, self.ny, self.u, self.rho, self.g_eq)

    def collision_stream(self):
        self._L.orc_col_str(self.nx, self.ny, self.g, self.g_eq, self.g_up,
                            self.om_p_lbm, self.om_m_lbm)

    def drag_lift(self, obs, R_ref, U_ref, L_ref):
        out = np.zeros(2)
        bnd = np.ascontiguousarray(obs.boundary, dtype=np.int64)
        self._L.orc_drag_lift(self.nx, self.ny, len(bnd), bnd.reshape(-1), self.g_up, self.g,
                              float(R_ref), float(U_ref), float(L_ref), out)
        return float(out[0]), float(out[1])

    def bounce_back_obstacle(self, obstacle):
        bnd = np.ascontiguousarray(obstacle.boundary, dtype=np.int64)
        ibb = np.ascontiguousarray(obstacle.ibb, dtype=np.float64)
        if not self.IBB:
            ibb = np.zeros(len(bnd))
        self._L.orc_bounce_back_obstacle(self.nx, self.ny, int(bool(self.IBB)), len(bnd),
                                         bnd.reshape(-1), ibb, self.g_up, self.g)

    def zou_he_left_wall_velocity(self):
        self._L.orc_zou_he_left_wall_velocity(self.nx, self.ny, self.u, self.u_left, self.rho, self.g)

    def zou_he_right_wall_velocity(self):
        self._L.orc_zou_he_right_wall_velocity(self.nx, self.ny, self.u, self.u_right, self.rho, self.g)

    def zou_he_right_wall_pressure(self):
        self._L.orc_zou_he_right_wall_pressure(self.nx, self.ny, self.u, self.rho_right,
                                               self.u_right, self.rho, self.g)

    def zou_he_top_wall_velocity(self):
        self._L.orc_zou_he_top_wall_velocity(self.nx, self.ny, self.u, self.u_top, self.rho, self.g)

    def zou_he_bottom_wall_velocity(self):
        self._L.orc_zou_he_bottom_wall_velocity(self.nx, self.ny, self.u, self.u_bot, self.rho, self.g)

    def zou_he_bottom_left_corner(self):
        self._L.orc_zou_he_bottom_left_corner(self.nx, self.ny, self.u, self.rho, self.g)

    def zou_he_top_left_corner(self):
        self._L.orc_zou_he_top_left_corner(self.nx, self.ny, self.u, self.rho, self.g)

    def zou_he_top_right_corner(self):
        self._L.orc_zou_he_top_right_corner(self.nx, self.ny, self.u, self.rho, self.g)

    def zou_he_bottom_right_corner(self):
        self._L.orc_zou_he_bottom_right_corner(self.nx, self.ny, self.u, self.rho, self.g)

    # --- host helpers used by app callbacks --------------------------------
    def get_coords(self, i, j):  # lattice.py:379-387
        dx = (self.x_max - self.x_min) / (self.nx - 1)
        dy = (self.y_max - self.y_min) / (self.ny - 1)
        return [self.x_min + i * dx, self.y_min + j * dy]

    def generate_image(self, obstacles):  # output writer: out of scope
        pass


def run_loop(lat, app, n_iters=None, on_step=None):
    """The reference's driver loop (run.py:12-61) without prints and outputs.

    Runs ``n_iters`` loop bodies (or until ``app.check_stop`` says stop when
    ``n_iters`` is None).  ``on_step(it)`` is called after ``observables``.
    """
    app.initialize(lat)
    it = 0
    while True:
        app.set_inlets(lat, it)
        lat.macro()
        lat.equilibrium()
        lat.collision_stream()
        app.set_bc(lat)
        app.observables(lat, it)
        if on_step is not None:
            on_step(it)
        it += 1
        if n_iters is not None:
            if it >= n_iters:
                break
        elif not app.check_stop(it - 1):
            break
    return it
