"""Host-side case descriptions (parameters, inlet ramp, boundary-condition sequence).

The reference keeps its configuration in per-app Python classes
(/root/reference/lbm/src/app/{cavity,turek,poiseuille,array}.py) that a driver
loop calls back into (lbm/src/core/run.py:24-54).  Those files stay the host
side of a real deployment: the ``lattice`` class of this
package is a drop-in for theirs.  This module restates the *inputs* those apps
produce -- sizes, relaxation time, ramp, wall profiles, the order of the
boundary calls -- behind the same callback protocol, so that tests and
bench.py can run the BASELINE configurations on machines that do not have the
reference checked out (the GPU box).  tests/test_oracle_golden.py pins every
case here against runs of the reference apps themselves.

Nothing in here computes lattice physics; it only fills wall-profile arrays
and calls ``lattice`` methods.
"""
import math
import os

import numpy as np

# package data: the obstacle link lists of the BASELINE configurations (Turek cylinder at ny = 100 / 200, the
# eight squares of the array case) as the reference's lattice.add_obstacle produces them (lattice.py:290-375)
# from its shape generator; written by tests/golden/make_golden.py, reproduced by lbm_b200/geometry.py
# (tests/test_geometry_cpu.py)
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class ConvergenceBuffer:
    """Triple moving average of an observable with the growth-rate convergence flag of the
    reference's buff class (/root/reference/lbm/src/utils/buff.py:8-77), same arithmetic on the
    same slices (np.sum is NumPy's pairwise sum in both), but on geometrically grown arrays instead
    of one np.append reallocation per value and level."""

    def __init__(self, name, dt, obs_cv_ct, obs_cv_nb, output_dir=None):
        self.name, self.dt = name, dt
        self.obs_cv_ct, self.obs_cv_nb = obs_cv_ct, obs_cv_nb
        self.it, self.obs, self.obs_cv_cnt, self.obs_cv = 0, 0.0, 0, False
        self._n = 2                                   # every level starts as [0, 0] (buff.py:18-21)
        self._lv = [np.zeros(1024) for _ in range(4)]  # raw values, avg1, avg2, avg3

    def _grow(self):
        if self._n + 1 > len(self._lv[0]):
            self._lv = [np.concatenate([a, np.zeros(len(a))]) for a in self._lv]

    def add(self, value):                             # buff.py:35-38
        self._grow()
        self._lv[0][self._n] = value
        self.it += 1

    def mv_avg(self):                                 # buff.py:51-77
        it_s, it_e = math.floor(self.it * 4 / 5), self.it
        n = self._n
        cnt = float(it_e - it_s + 1)
        self.obs = np.sum(self._lv[0][it_s:it_e]) / cnt
        for lv in (1, 2, 3):
            self._lv[lv][n] = self.obs
            self.obs = np.sum(self._lv[lv][it_s:it_e]) / cnt
        self._n = n + 1
        growth = 0.0
        if self.it > 10:
            a3 = self._lv[3]
            growth = (a3[it_e] - a3[it_s]) / ((it_e - it_s + 1) * self.dt)
            if abs(growth) < self.obs_cv_ct:
                self.obs_cv_cnt += 1
            else:
                self.obs_cv_cnt = 0
            if self.obs_cv_cnt > self.obs_cv_nb:
                self.obs_cv = True
        return self.obs, growth


class Obstacle:
    """Link list of one body: rows (i, j, q) with q pointing fluid -> solid and the
    IBB wall distance of each link (obstacle.py:3-25, produced by lattice.py:290-375)."""

    def __init__(self, boundary, ibb, tag=1):
        self.boundary = np.ascontiguousarray(boundary, dtype=np.int64).reshape(-1, 3)
        self.ibb = np.ascontiguousarray(ibb, dtype=np.float64).reshape(-1)
        self.tag = tag


def ramp(it, sigma):
    """Inlet ramp of every reference app (cavity.py:70-71, turek.py:99-100)."""
    return 1.0 - math.exp(-it ** 2 / (2.0 * sigma ** 2))


class Case:
    """Callback protocol of run.py: initialize / set_inlets / set_bc / observables /
    check_stop / finalize (+ printings / outputs as no-ops here)."""

    name = "case"
    IBB = False
    stop = "it"
    rho_lbm = 1.0
    right_wall = "pressure"

    def __init__(self):
        self.obstacles = []
        self.forces = []          # (Cx, Cy) of obstacle 0 per iteration
        self.output_freq = 1 << 62

    # -- sizes -----------------------------------------------------------
    def _finish(self):
        self.Cs = 1.0 / math.sqrt(3.0)
        self.dx = (self.y_max - self.y_min) / self.ny
        self.nx = math.floor(self.ny * (self.x_max - self.x_min) / (self.y_max - self.y_min))
        self.tau_lbm = 0.5 + self.nu_lbm / (self.Cs ** 2)
        self.it_max = math.floor(self.t_max / self.dt)
        self.sigma = math.floor(10 * self.nx)

    # -- protocol --------------------------------------------------------
    def initialize(self, lattice):
        self.forces = []
        self.set_inlets(lattice, 0)
        lattice.rho *= self.rho_lbm
        lattice.equilibrium()
        lattice.g = lattice.g_eq.copy()

    def printings(self, it):
        pass

    def outputs(self, lattice, it):
        pass

    def observables(self, lattice, it):
        pass

    def finalize(self, lattice):
        pass

    def check_stop(self, it):
        return it < self.it_max  # base_app.py:57-61 ('it' rule; the loop body runs it_max+1 times)

    def set_bc(self, lattice):
        for obs in self.obstacles:
            lattice.bounce_back_obstacle(obs)
        lattice.zou_he_bottom_wall_velocity()
        lattice.zou_he_left_wall_velocity()
        if self.right_wall == "pressure":      # channel order, turek.py:118-121
            lattice.zou_he_top_wall_velocity()
            lattice.zou_he_right_wall_pressure()
        else:                                  # cavity order, cavity.py:82-85
            lattice.zou_he_right_wall_velocity()
            lattice.zou_he_top_wall_velocity()
        lattice.zou_he_bottom_left_corner()
        lattice.zou_he_top_left_corner()
        lattice.zou_he_top_right_corner()
        lattice.zou_he_bottom_right_corner()


class Cavity(Case):
    """Lid-driven cavity (cavity.py:12-89): four velocity walls, ramped lid."""

    name = "cavity"
    right_wall = "velocity"

    def __init__(self, L_lbm=100, Re_lbm=100.0, u_lbm=0.2, t_max=20.0, sigma=None, tau_lbm=None):
        super().__init__()
        self.L_lbm, self.Re_lbm, self.u_lbm, self.t_max = L_lbm, Re_lbm, u_lbm, t_max
        self.x_min, self.x_max, self.y_min, self.y_max = 0.0, 1.0, 0.0, 1.0
        self.ny = L_lbm
        self.nu_lbm = u_lbm * L_lbm / Re_lbm
        self.dt = Re_lbm * self.nu_lbm / L_lbm ** 2
        self._finish()
        if sigma is not None:
            self.sigma = sigma
        if tau_lbm is not None:
            self.tau_lbm = tau_lbm

    def set_inlets(self, lattice, it):
        lattice.u_top[0, :] = self.u_lbm * ramp(it, self.sigma)
        lattice.u_bot[0, :] = 0.0
        lattice.u_left[1, :] = 0.0
        lattice.u_right[1, :] = 0.0

    def line_fields(self, lattice):
        """Centre-line profiles u_y(x, ny/2)/u_lbm and u_x(nx/2, y)/u_lbm (cavity.py:110-133)."""
        u = lattice.u
        return u[1, :, self.ny // 2] / self.u_lbm, u[0, self.nx // 2, :] / self.u_lbm


class Channel(Case):
    """Channel flow with a ramped parabolic inlet on the left, no-slip top/bottom and a
    pressure outlet on the right; base of poiseuille.py, turek.py and array.py."""

    right_wall = "pressure"

    def poiseuille(self, pt):
        """Inlet profile at a point (turek.py:165-173)."""
        y = pt[1]
        H = self.y_max - self.y_min
        u = np.zeros(2)
        u[0] = 4.0 * (self.y_max - y) * (y - self.y_min) / H ** 2
        return u

    def inlet_shape(self, lattice):
        ny = self.ny
        dy = (self.y_max - self.y_min) / (ny - 1)        # lattice.get_coords, lattice.py:379-387
        prof = np.zeros(ny)
        for j in range(ny):
            prof[j] = self.poiseuille([self.x_min, self.y_min + j * dy])[0]
        return prof

    def set_inlets(self, lattice, it):
        if getattr(self, "_shape", None) is None:
            self._shape = self.inlet_shape(lattice)
        a = ramp(it, self.sigma) * self.u_lbm           # evaluated left to right, turek.py:104
        lattice.u_left[0, :] = a * self._shape
        lattice.u_left[1, :] = a * 0.0
        lattice.u_top[0, :] = 0.0
        lattice.u_bot[0, :] = 0.0
        lattice.u_right[1, :] = 0.0
        lattice.rho_right[:] = self.rho_lbm

    def observables(self, lattice, it):
        if self.obstacles and getattr(self, "track_forces", False):
            cx, cy = lattice.drag_lift(self.obstacles[0], self.rho_lbm, self.u_avg, self.D_lbm)
            self.forces.append((cx, cy))
            if self.stop == "obs":                   # turek.add_buff, turek.py:147-155
                self.drag_buff.add(cx)
                self.lift_buff.add(cy)
                self.avg_drag, _ = self.drag_buff.mv_avg()
                self.avg_lift, _ = self.lift_buff.mv_avg()

    def check_stop(self, it):
        if self.stop == "obs":                       # base_app.py:63-67
            return not (self.drag_buff.obs_cv and self.lift_buff.obs_cv)
        return it < self.it_max

    def initialize(self, lattice):
        if self.stop == "obs":                       # turek.py:77-87
            self.drag_buff = ConvergenceBuffer("drag", lattice.dt, lattice.obs_cv_ct, lattice.obs_cv_nb)
            self.lift_buff = ConvergenceBuffer("lift", lattice.dt, lattice.obs_cv_ct, lattice.obs_cv_nb)
        super().initialize(lattice)


class Poiseuille(Channel):
    """poiseuille.py:12-65."""

    name = "poiseuille"

    def __init__(self, L_lbm=50, Re_lbm=100.0, u_lbm=0.1, t_max=15.0, sigma=None):
        super().__init__()
        self.L_lbm, self.Re_lbm, self.u_lbm, self.t_max = L_lbm, Re_lbm, u_lbm, t_max
        self.x_min, self.x_max, self.y_min, self.y_max = -0.2, 1.0, -0.2, 0.2
        self.ny = L_lbm
        self.u_avg = 2.0 * u_lbm / 3.0
        self.nu_lbm = self.u_avg * L_lbm / Re_lbm
        self.dt = Re_lbm * self.nu_lbm / L_lbm ** 2
        self._finish()
        if sigma is not None:
            self.sigma = sigma


class Turek(Channel):
    """Schaefer-Turek cylinder (turek.py:14-61): IBB cylinder, drag/lift of obstacle 0."""

    name = "turek"
    IBB = True
    track_forces = True

    obs_cv_ct = 1.0e-3      # turek.py:27-29
    obs_cv_nb = 1000

    def __init__(self, L_lbm=100, Re_lbm=20.0, u_lbm=0.05, sigma=None, links=None, stop="it"):
        super().__init__()
        self.stop = stop
        self.L_lbm, self.Re_lbm, self.u_lbm, self.t_max = L_lbm, Re_lbm, u_lbm, 0.02
        self.x_min, self.x_max, self.y_min, self.y_max = -0.2, 2.0, -0.2, 0.21
        self.ny = L_lbm
        self.u_avg = 2.0 * u_lbm / 3.0
        self.r_cyl = 0.1
        self.D_lbm = math.floor(self.ny * self.r_cyl / (self.y_max - self.y_min))
        self.nu_lbm = self.u_avg * self.D_lbm / Re_lbm
        self.dt = Re_lbm * self.nu_lbm / self.D_lbm ** 2
        self._finish()
        if sigma is not None:
            self.sigma = sigma
        if links is None:
            links = "links_turek%d.npz" % L_lbm
        self.obstacles = load_links(links)


class Array(Channel):
    """Ring of eight squares at Re=2000 (array.py:14-68)."""

    name = "array"
    IBB = True

    def __init__(self, L_lbm=200, Re_lbm=2000.0, u_lbm=0.025, sigma=None, links="links_array.npz"):
        super().__init__()
        self.L_lbm, self.Re_lbm, self.u_lbm, self.t_max = L_lbm, Re_lbm, u_lbm, 7.5
        self.x_min, self.x_max, self.y_min, self.y_max = -1.0, 8.0, -1.0, 1.0
        self.ny = L_lbm
        self.u_avg = 2.0 * u_lbm / 3.0
        self.D_lbm = math.floor(self.ny * 0.1 / (self.y_max - self.y_min))
        self.nu_lbm = self.u_avg * L_lbm / Re_lbm
        self.dt = Re_lbm * self.nu_lbm / L_lbm ** 2
        self._finish()
        if sigma is not None:
            self.sigma = sigma
        self.obstacles = load_links(links)


def load_links(path):
    """Obstacle link lists of lbm_b200/data/ (outputs of the reference's lattice.add_obstacle)."""
    if isinstance(path, (list, tuple)):
        return list(path)
    if not os.path.isabs(path):
        path = os.path.join(_DATA, path)
    z = np.load(path)
    off = z["offsets"]
    return [Obstacle(z["boundary"][off[k]:off[k + 1]], z["ibb"][off[k]:off[k + 1]], tag=k + 1)
            for k in range(len(off) - 1)]
