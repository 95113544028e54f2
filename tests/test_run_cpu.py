"""CPU: the batched driver loop (lbm_b200/run.py) -- batching, replay of the per-iteration callbacks,
the exact stop-rule rollback -- driven by a lattice that has the lazy / batched surface of
lbm_b200.lattice.lattice but executes the oracle's phases on the host.  The result must equal the
reference's plain loop (oracle.run_loop) bit for bit: same iteration count, same drag/lift series,
same final arrays.  (On the GPU the same driver is covered by tests/test_gpu_run.py.)"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from lbm_b200 import cases
from lbm_b200.run import run
from oracle import oracle as orc

_BC = ("zou_he_left_wall_velocity", "zou_he_right_wall_velocity", "zou_he_right_wall_pressure",
       "zou_he_top_wall_velocity", "zou_he_bottom_wall_velocity", "zou_he_bottom_left_corner",
       "zou_he_top_left_corner", "zou_he_top_right_corner", "zou_he_bottom_right_corner")


class LazyOracleLattice(orc.OracleLattice):
    """Oracle phases behind the deferred-execution interface of the GPU lattice: boundary calls are
    recorded (with the wall arrays of that moment), the work of an iteration is executed by the next
    macro() or by batch_updates()."""

    def __init__(self, app):
        super().__init__(app)
        self._state = "fresh"            # fresh | macro_done | streamed
        self._recorded = []              # [(method name, obstacle or None)] of the open BC window
        self._row = np.zeros(5 * self.ny + 4 * self.nx)
        self._replay = None
        self._advanced = False           # collide + stream + BCs of the current iteration already executed
        self._link_obstacles = []
        self.batches = []

    # -- wall rows ------------------------------------------------------------------------
    def snapshot_walls(self):
        nx, ny = self.nx, self.ny
        return np.concatenate([self.u_left.reshape(-1), self.u_right.reshape(-1), self.u_top.reshape(-1),
                               self.u_bot.reshape(-1), self.rho_right])

    def _load_row(self, row):
        nx, ny = self.nx, self.ny
        self.u_left[:] = row[0:2 * ny].reshape(2, ny)
        self.u_right[:] = row[2 * ny:4 * ny].reshape(2, ny)
        self.u_top[:] = row[4 * ny:4 * ny + 2 * nx].reshape(2, nx)
        self.u_bot[:] = row[4 * ny + 2 * nx:4 * ny + 4 * nx].reshape(2, nx)
        self.rho_right[:] = row[4 * ny + 4 * nx:]

    # -- execution ------------------------------------------------------------------------
    def _advance(self):
        """collide + stream of the current iteration, then its recorded boundary conditions with the
        recorded wall row; returns the momentum-exchange sums per obstacle."""
        now = self.snapshot_walls()
        self._load_row(self._row)
        orc.OracleLattice.equilibrium(self)
        orc.OracleLattice.collision_stream(self)
        obstacles = []
        for name, obs in self._recorded:
            if name == "bounce_back_obstacle":
                orc.OracleLattice.bounce_back_obstacle(self, obs)
                obstacles.append(obs)
            else:
                getattr(orc.OracleLattice, name)(self)
        self._link_obstacles = obstacles
        f = np.zeros((max(len(obstacles), 1), 2))
        for k, obs in enumerate(obstacles):
            cx, cy = orc.OracleLattice.drag_lift(self, obs, 1.0, 1.0, 1.0)      # C = -2 f
            f[k] = (-0.5 * cx, -0.5 * cy)
        self._load_row(now)
        return f

    def macro(self):
        if self._state == "macro_done":
            return
        if self._state == "streamed" and not self._advanced:
            self._advance()
        orc.OracleLattice.macro(self)
        self._state, self._advanced = "macro_done", False

    def equilibrium(self):
        if self._state == "fresh":
            orc.OracleLattice.equilibrium(self)

    def collision_stream(self):
        assert self._state == "macro_done"
        self._state, self._recorded = "streamed", []

    def _record(self, name, obs=None):
        assert self._state == "streamed" and not self._advanced
        self._recorded.append((name, obs))
        self._row = self.snapshot_walls()

    def bounce_back_obstacle(self, obstacle):
        self._record("bounce_back_obstacle", obstacle)

    def drag_lift(self, obs, R_ref, U_ref, L_ref):
        if self._replay is not None:
            k = [i for i, o in enumerate(self._link_obstacles) if o is obs][0]
            fx, fy = float(self._replay[k, 0]), float(self._replay[k, 1])
        else:
            assert self._state == "streamed"
            if not self._advanced:
                self._last_f = self._advance()
                self._advanced = True
            k = [i for i, o in enumerate(self._link_obstacles) if o is obs][0]
            fx, fy = self._last_f[k]
        return (-2.0 * fx / (R_ref * L_ref * U_ref ** 2), -2.0 * fy / (R_ref * L_ref * U_ref ** 2))

    def batch_updates(self, rows):
        assert self._state == "streamed"
        rows = np.asarray(rows).reshape(-1, self._row.size)
        self.batches.append(len(rows))
        forces = []
        for k, row in enumerate(rows):
            self._row = row.copy()
            if self._advanced:                       # (drag_lift already executed this iteration's tail)
                f, self._advanced = self._last_f, False
            else:
                f = self._advance()
            forces.append(f)
            orc.OracleLattice.macro(self)
            self.fields_stale = False
            self._state = "macro_done"
            if k + 1 < len(rows):
                self._state = "streamed"             # same BC set for the following update
        return np.array(forces)

    def batch_updates_ramp(self, base_row, scales):
        """What the library does with a ramp table: velocity entries of the base row x scale, one rounded product."""
        nvel = 4 * self.ny + 4 * self.nx
        rows = np.tile(np.asarray(base_row, dtype=np.float64), (len(scales), 1))
        rows[:, :nvel] = np.asarray(scales)[:, None] * rows[:, :nvel]
        self.ramp_batches = getattr(self, "ramp_batches", 0) + 1
        return self.batch_updates(rows)

    # the two halves of a batch (lbm_b200.lattice: enqueue without waiting, look at the sums later); here the work is
    # done at enqueue time, and a token may only be redeemed once and in order -- what the pinned buffers of the real
    # lattice allow
    def can_pipeline(self):
        return True

    def batch_enqueue_ramp(self, base_row, scales):
        self.in_flight = getattr(self, "in_flight", 0) + 1
        assert self.in_flight <= 2
        self.max_in_flight = max(getattr(self, "max_in_flight", 0), self.in_flight)
        return [self.batch_updates_ramp(base_row, scales)]

    def batch_result(self, token):
        self.in_flight -= 1
        return token.pop()

    def save_state(self, slot=0):
        self.__dict__.setdefault("_saved", {})[slot] = tuple(a.copy() for a in (self.g, self.g_up, self.rho, self.u)) + (self._advanced,)

    def restore_state(self, slot=0):
        """(The real lattice restores the populations only: rho / u on the device stay those of the last update that was
        EXECUTED.  Here they are inputs of the next update and have to come back, so staleness is tracked by a flag that
        the next executed update clears: a run must not end with it set.)"""
        self.rollbacks = getattr(self, "rollbacks", 0) + 1
        for a, b in zip((self.g, self.g_up, self.rho, self.u), self._saved[slot][:4]):
            a[:] = b
        self._advanced = self._saved[slot][4]
        self.fields_stale = True


for _name in _BC:
    setattr(LazyOracleLattice, _name, (lambda n: lambda self: self._record(n))(_name))


def _cases():
    z = np.load(os.path.join(GOLDEN, "run_turek30.npz"))

    def turek(stop):
        c = cases.Turek(L_lbm=30, Re_lbm=20.0, sigma=15, links=[cases.Obstacle(z["boundary"], z["ibb"])], stop=stop)
        if stop == "obs":
            c.obs_cv_ct, c.obs_cv_nb = 5.0e-2, 40          # converges after a few hundred iterations
        else:
            c.it_max = 230
        return c
    return {"cavity": lambda: _with(cases.Cavity(L_lbm=24, sigma=20), it_max=157),
            "turek_it": lambda: turek("it"), "turek_obs": lambda: turek("obs"),
            "turek_out": lambda: _with(turek("it"), output_freq=100)}


def _with(c, **kw):
    for k, v in kw.items():
        setattr(c, k, v)
    return c


@pytest.mark.parametrize("name", ["cavity", "turek_it", "turek_obs", "turek_out"])
@pytest.mark.parametrize("batch,pipeline", [(1, True), (7, False), (7, True), (64, True)])
def test_batched_driver_equals_the_plain_loop(name, batch, pipeline):
    """pipeline: the next batch is enqueued before the callbacks of the previous one are replayed (the drag/lift of a
    batch's last iteration arrives with the next batch); turek_out: with output iterations, where the chain is broken."""
    mk = _cases()[name]
    ca, cb = mk(), mk()
    la, lb = LazyOracleLattice(ca), orc.OracleLattice(cb)
    na = run(la, ca, batch=batch, quiet=True, pipeline=pipeline)
    nb = orc.run_loop(lb, cb)
    assert na == nb and na > 50
    assert getattr(la, "max_in_flight", 0) == (2 if pipeline else 0)
    if batch > 1:
        assert max(la.batches) > 1                                   # really batched
        assert getattr(la, "ramp_batches", 0) > 0                    # ... through the inlet model (one scalar per iteration)
        if name == "turek_obs":
            assert getattr(la, "rollbacks", 0) == 1                  # the stop rule fired inside a batch
    if getattr(ca, "forces", None):
        assert np.array_equal(np.array(ca.forces), np.array(cb.forces))
    # the plain loop ends after set_bc of the last iteration; bring the lazy lattice to the same point
    if la._state == "streamed" and not la._advanced:
        la._advance()
    for k in ("g", "g_up", "rho", "u"):
        assert np.array_equal(getattr(la, k), getattr(lb, k)), k


@pytest.mark.parametrize("j_stop", [5, 6, 7, 8, 13, 14, 15, 20, 21, 22])
def test_pipelined_stop_on_every_position_of_a_batch(j_stop):
    """A stop rule that fires on iteration j_stop, batches of 7 updates (iterations 1-7, 8-14, 15-21, ..): the stop on the
    last iteration of a batch is seen one batch later (its drag/lift is slot 0 of the NEXT batch: the roll-back then re-runs
    nothing), on the first iteration it discards a batch that was speculated in vain.  Same iteration count, force series
    and final arrays as the plain loop."""
    z = np.load(os.path.join(GOLDEN, "run_turek30.npz"))

    class StopAt(cases.Turek):
        def check_stop(self, it):
            return it != j_stop

    def mk():
        return StopAt(L_lbm=30, Re_lbm=20.0, sigma=15, links=[cases.Obstacle(z["boundary"], z["ibb"])], stop="obs")
    ca, cb = mk(), mk()
    la, lb = LazyOracleLattice(ca), orc.OracleLattice(cb)
    na = run(la, ca, batch=7, quiet=True)
    nb = orc.run_loop(lb, cb)
    assert na == nb == j_stop + 1
    assert getattr(la, "rollbacks", 0) == 1 and la.max_in_flight == 2
    assert not la.fields_stale                  # rho / u belong to the iteration the run stopped on
    assert np.array_equal(np.array(ca.forces), np.array(cb.forces))
    if la._state == "streamed" and not la._advanced:
        la._advance()
    for k in ("g", "g_up", "rho", "u"):
        assert np.array_equal(getattr(la, k), getattr(lb, k)), k


def test_inlet_model_reproduces_the_apps_bit_for_bit():
    """InletModel (run.py): closed form of app.set_inlets for the restated cases and -- where the reference is
    checked out -- for the reference's own cavity / turek / poiseuille / array apps: every wall row it predicts
    equals what the app writes, bitwise, on iterations that are not among the detection probes."""
    from lbm_b200.run import InletModel
    apps = [("cases.cavity", cases.Cavity(L_lbm=40)), ("cases.turek", _cases()["turek_it"]()), ("cases.poiseuille", cases.Poiseuille(L_lbm=20))]
    from oracle import refload
    if refload.available():
        import contextlib
        import io
        ns = refload.load()
        with refload.in_scratch(), contextlib.redirect_stdout(io.StringIO()):
            for name in ("cavity", "poiseuille", "turek", "array"):
                a = ns.app.app_factory.create(name)
                if name == "turek":
                    a.L_lbm = 60
                    a.compute_lbm_parameters()
                apps.append(("ref." + name, a))
    for name, app in apps:
        lat = LazyOracleLattice(app)
        m = InletModel.detect(lat, app)
        assert m is not None, name
        for it in (4, 5, 77, 500, 2222, 40000):
            app.set_inlets(lat, it)
            assert np.array_equal(m.row(it), lat.snapshot_walls()), (name, it)
        its = [4, 77, 2222, int(8.9 * m.sigma), int(9.0 * m.sigma), int(9.0 * m.sigma) + 1, int(40 * m.sigma)]
        assert np.array_equal(m.scales(its), np.array([m.scale(it) for it in its]))


def test_inlet_model_rejects_an_app_that_does_not_fit():
    from lbm_b200.run import InletModel

    class Odd(cases.Cavity):
        def set_inlets(self, lattice, it):
            super().set_inlets(lattice, it)
            lattice.u_top[0, :] += 1.0e-3 * np.sin(0.1 * it)          # not profile x ramp
    app = Odd(L_lbm=24, sigma=20)
    app.it_max = 60
    lat = LazyOracleLattice(app)
    assert InletModel.detect(lat, app) is None
    cb = Odd(L_lbm=24, sigma=20)
    cb.it_max = 60
    lb = orc.OracleLattice(cb)
    assert run(lat, app, batch=16, quiet=True) == orc.run_loop(lb, cb)
    assert getattr(lat, "ramp_batches", 0) == 0
    lat._advance()
    assert np.array_equal(lat.g, lb.g)
