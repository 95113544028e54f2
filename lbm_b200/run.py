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

Pipelining (default where the lattice offers batch_enqueue_ramp and the inlet model fits): a batch is enqueued and the
host does NOT wait for it; it records the boundary conditions of the batch's last iteration (set_inlets / set_bc: host
bookkeeping only), enqueues the NEXT batch, and only then fetches the first batch's drag/lift sums and replays its
callbacks -- while the device executes the next batch.  The drag/lift of a batch's last iteration is slot 0 of the
following batch (the sums of iteration i are produced by update i+1), so that iteration's observables / check_stop are
replayed with the next batch.  A whole run then costs max(device, host) per iteration instead of their sum.  The
speculation is given up where the app needs the fields (output iterations), at it_max, and when an 'obs' stop fires:
the populations are rolled back to the start of the batch the stop lies in (the starting states of three batches are
kept: the one replayed, the one before it -- a stop on ITS last iteration is only seen now -- and the one speculated
after it), exactly as without pipelining.  The app sees
the same calls with the same values; only set_inlets / set_bc of a batch's last iteration come before the replay of
that batch's earlier iterations.

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
        # (scalar libm calls, like the apps': a vectorised exp may differ in the last bit; ~0.3 us per iteration.  Beyond
        # 9 sigma exp(-40.5) = 2.6e-18 < 2^-53: ret is exactly 1.0 and the scale exactly u_lbm)
        its = np.asarray(its)
        out = np.full(its.shape, 1.0 * self.u_lbm, dtype=np.float64)
        for k in np.nonzero(its <= 9.0 * self.sigma)[0]:
            out[k] = self.scale(int(its[k]))
        return out

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


def _run_pipelined(lattice, app, model, batch, quiet, freq, stop_on_it, start_time):
    """The batched loop with one batch in flight while the previous one is replayed (module docstring)."""
    exact_stop = not stop_on_it
    it_max = int(app.it_max) if stop_on_it and hasattr(app, "it_max") else None

    def size(it):
        n = max(1, int(batch))
        if freq > 0:
            n = min(n, (-it) % freq + 1 if it % freq else 1)
        if it_max is not None:
            n = max(1, min(n, it_max - it + 1))
        return n

    serial = [0]

    def enqueue(it, carry):
        """Updates of iterations it .. it+n-1; needs the boundary conditions of iteration it-1 recorded."""
        n = size(it)
        if not np.array_equal(lattice._row, model.row(it - 1)):
            raise RuntimeError("inlet model no longer matches app.set_inlets at iteration %d" % (it - 1))
        scales = model.scales(np.arange(it - 1, it - 1 + n))
        slot = serial[0] % 3                                  # (this batch, the one before, the one speculated after)
        serial[0] += 1
        if exact_stop:
            lattice.save_state(slot)
        return dict(it=it, n=n, last=it + n - 1, scales=scales, slot=slot, carry=carry,
                    token=lattice.batch_enqueue_ramp(model.base, scales))

    def callbacks(j, f):
        if not quiet:
            app.printings(j)
        lattice._replay = f
        try:
            app.outputs(lattice, j)                           # no-op off the output iterations
            app.observables(lattice, j)
        finally:
            lattice._replay = None
        return app.check_stop(j)

    cur, prev = enqueue(1, False), None
    while True:
        it, n, last = cur["it"], cur["n"], cur["last"]
        chain = not (freq > 0 and last % freq == 0) and not (it_max is not None and last >= it_max)
        nxt = None
        if chain:
            # iteration `last`: its boundary conditions are host bookkeeping; its drag/lift arrives with the next batch
            app.set_inlets(lattice, last)
            lattice.collision_stream()
            app.set_bc(lattice)
            nxt = enqueue(last + 1, True)
        forces = lattice.batch_result(cur["token"])           # slot k = iteration it+k-1
        stopped_at = None
        for k in range(0 if cur["carry"] else 1, n):
            if not callbacks(it + k - 1, forces[k]):
                stopped_at = it + k - 1
                break
        if stopped_at is not None:
            # the stop rule fired: back to the start of this batch, redo it .. stopped_at exactly
            if nxt is not None:
                lattice.batch_result(nxt["token"])            # (speculated in vain)
            m = stopped_at - it + 1
            if m > 0:
                lattice.restore_state(cur["slot"])
                lattice._state = "streamed"
                lattice.batch_updates_ramp(model.base, cur["scales"][:m])
            else:
                # the last iteration of the batch BEFORE this one (its drag/lift came with this batch): that batch once
                # more from its own start, so that rho / u of its last update are the stored fields again
                lattice.restore_state(prev["slot"])
                lattice._state = "streamed"
                lattice.batch_updates_ramp(model.base, prev["scales"])
            lattice.collision_stream()
            app.set_inlets(lattice, stopped_at)
            app.set_bc(lattice)
            it = stopped_at + 1
            break
        if chain:
            cur, prev = nxt, cur
            continue
        if not quiet:
            app.printings(last)
        app.set_inlets(lattice, last)
        if not _tail_of_iteration(lattice, app, last):
            it = last + 1
            break
        cur, prev = enqueue(last + 1, False), None

    if not quiet:
        print("# Loop time = {:f}".format(time.time() - start_time))
    app.finalize(lattice)
    return it


def run(lattice, app, batch=512, quiet=False, inlet_model=True, pipeline=True):
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
    if (compute and pipeline and model is not None and hasattr(lattice, "batch_enqueue_ramp") and lattice.can_pipeline()
            and (exact_stop or hasattr(app, "it_max"))):
        return _run_pipelined(lattice, app, model, batch, quiet, freq, stop_on_it, start_time)

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
