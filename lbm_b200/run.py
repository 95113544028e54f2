"""Driver loop: same phase order and callbacks as the reference's run()
(/root/reference/lbm/src/core/run.py:12-61), executed in batches.

The reference calls every phase of every iteration from Python.  Once an update costs
microseconds that per-phase traffic dominates, so this driver executes up to
`batch` iterations with ONE library call and then replays the host-side callbacks of those
iterations (printings, observables, check_stop) in order with the per-iteration drag/lift sums
the updates stored.  What the app sees is identical to the per-phase loop:

  * wall profiles of iteration `it` are what app.set_inlets(lattice, it) leaves in the arrays;
  * app.outputs(lattice, it) runs with lattice.u / lattice.rho of iteration `it` whenever
    it % app.output_freq == 0 (batches end on those iterations);
  * app.observables(lattice, it) gets the drag/lift of iteration `it` from lattice.drag_lift;
  * with stop == 'obs' the run ends on exactly the iteration the reference would stop at: if the
    rule fires inside a batch, the populations are rolled back to the batch start and re-run up to
    that iteration (the update is deterministic).

The reference's own run() also works with lbm_b200.lattice.lattice (one update per iteration).
"""
import math
import time

import numpy as np


class InletModel:
    """Closed form of what app.set_inlets(lattice, it) leaves in the wall-profile arrays, for apps whose inlets are
    a fixed profile times the ramp of the reference apps,

        velocity entries(it) = fl( s(it) * B ),   s(it) = fl( ret(it) * u_lbm ),   ret(it) = 1 - exp(-it^2 / (2 sigma^2))

    (cavity.py:70-73: u_top = u_lbm*ret, B = 1 on the lid; turek.py:99-104, poiseuille.py, array.py, step.py:
    u_left = ret*u_lbm*poiseuille(pt), B = the app's own poiseuille() profile).  The model is only used after it
    has reproduced the app's own set_inlets BIT FOR BIT on a set of probe iterations, and run() re-checks it against
    the app at the end of every batch; an app that does not fit keeps the per-iteration path.  With a model a
    batch of updates needs one scalar per iteration from the host (lbm_set_ramp) instead of one call of the app's
    Python set_inlets and one wall row per iteration."""

    PROBES = (0, 1, 2, 3, 10, 97, 1000, 54321, 10 ** 9)

    def __init__(self, base, sigma, u_lbm, nvel):
        self.base, self.sigma, self.u_lbm, self.nvel = base, float(sigma), float(u_lbm), nvel

    def scale(self, it):
        return (1.0 - math.exp(-it ** 2 / (2.0 * self.sigma ** 2))) * self.u_lbm

    def scales(self, its):
        # (scalar libm calls, like the apps': a vectorised exp may differ in the last bit; ~0.3 us per iteration)
        return np.array([self.scale(int(it)) for it in its], dtype=np.float64)

    def row(self, it):
        r = self.base.copy()
        r[:self.nvel] = self.scale(it) * self.base[:self.nvel]
        return r

    @classmethod
    def detect(cls, lattice, app):
        sigma, u_lbm = getattr(app, "sigma", None), getattr(app, "u_lbm", None)
        if not sigma or u_lbm is None:
            return None
        nx, ny = lattice.nx, lattice.ny
        nvel = 4 * ny + 4 * nx

        def actual(it):
            app.set_inlets(lattice, it)
            return lattice.snapshot_walls()
        try:
            full = actual(cls.PROBES[-1])                 # ret == 1.0 exactly
            candidates = []
            b = full.copy()
            b[:nvel] = (full[:nvel] != 0.0).astype(np.float64)      # constant profile (cavity lid)
            candidates.append(b)
            if hasattr(app, "poiseuille"):                # channel apps: the app's own inlet profile on the left wall
                b = full.copy()
                b[:nvel] = 0.0
                for j in range(ny):
                    p = np.asarray(app.poiseuille(lattice.get_coords(0, j)), dtype=np.float64)
                    b[j], b[ny + j] = p[0], p[1]
                candidates.append(b)
            for base in candidates:
                m = cls(base, sigma, u_lbm, nvel)
                if all(np.array_equal(m.row(it), actual(it)) for it in cls.PROBES):
                    return m
        except Exception:
            return None
        return None


def _tail_of_iteration(lattice, app, it):
    """run.py:36-51 after macro(): outputs, (equilibrium, collision_stream), set_bc, observables, stop."""
    app.outputs(lattice, it)
    lattice.equilibrium()
    lattice.collision_stream()
    app.set_bc(lattice)
    app.observables(lattice, it)
    return app.check_stop(it)


def run(lattice, app, batch=512, quiet=False, inlet_model=True):
    app.initialize(lattice)
    model = InletModel.detect(lattice, app) if inlet_model and hasattr(lattice, "batch_updates_ramp") else None
    start_time = time.time()
    if not quiet:
        print('### Solving')
    freq = int(getattr(app, "output_freq", 0) or 0)
    stop_on_it = getattr(app, "stop", "it") == "it"
    exact_stop = not stop_on_it

    # iteration 0 phase by phase (collide-only update; records the BC sequence)
    it = 0
    if not quiet:
        app.printings(it)
    app.set_inlets(lattice, it)
    lattice.macro()
    compute = _tail_of_iteration(lattice, app, it)
    it = 1

    while compute:
        # batch = iterations it .. it+n-1; it ends on the next output iteration / it_max
        n = max(1, int(batch))
        if freq > 0:
            n = min(n, (-it) % freq + 1 if it % freq else 1)
        if stop_on_it and hasattr(app, "it_max"):
            n = max(1, min(n, int(app.it_max) - it + 1))
        # wall rows: update `it+k` applies the boundary conditions of iteration it+k-1
        if model is not None:
            if not np.array_equal(lattice._row, model.row(it - 1)):
                raise RuntimeError("inlet model no longer matches app.set_inlets at iteration %d" % (it - 1))
            scales = model.scales(np.arange(it - 1, it - 1 + n))
            rows = None
        else:
            rows = [lattice._row.copy()]
            for k in range(n - 1):
                app.set_inlets(lattice, it + k)
                rows.append(lattice.snapshot_walls())
        if exact_stop and n > 1:
            lattice.save_state()
        if model is not None:
            forces = lattice.batch_updates_ramp(model.base, scales)
        else:
            forces = lattice.batch_updates(np.stack(rows))   # slot k = iteration it+k-1
        last = it + n - 1
        stopped_at = None
        for k in range(n - 1):                                # replay iterations it .. last-1
            j = it + k
            if not quiet:
                app.printings(j)
            lattice._replay = forces[k + 1]
            try:
                app.outputs(lattice, j)                       # no-op off the output iterations
                app.observables(lattice, j)
            finally:
                lattice._replay = None
            if not app.check_stop(j):
                stopped_at = j
                break
        if stopped_at is not None:
            # the stop rule fired inside the batch: redo it .. stopped_at exactly
            m = stopped_at - it + 1
            lattice.restore_state()
            lattice._state = "streamed"
            if model is not None:
                lattice.batch_updates_ramp(model.base, scales[:m])
            else:
                lattice.batch_updates(np.stack(rows[:m]))
            lattice.collision_stream()
            app.set_inlets(lattice, stopped_at)
            app.set_bc(lattice)
            it = stopped_at + 1
            compute = False
            break
        if not quiet:
            app.printings(last)
        app.set_inlets(lattice, last)
        compute = _tail_of_iteration(lattice, app, last)
        it = last + 1

    if not quiet:
        print("# Loop time = {:f}".format(time.time() - start_time))
    app.finalize(lattice)
    return it
