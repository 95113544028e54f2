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
import time

import numpy as np


def _tail_of_iteration(lattice, app, it):
    """run.py:36-51 after macro(): outputs, (equilibrium, collision_stream), set_bc, observables, stop."""
    app.outputs(lattice, it)
    lattice.equilibrium()
    lattice.collision_stream()
    app.set_bc(lattice)
    app.observables(lattice, it)
    return app.check_stop(it)


def run(lattice, app, batch=512, quiet=False):
    app.initialize(lattice)
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
        rows = [lattice._row.copy()]
        for k in range(n - 1):
            app.set_inlets(lattice, it + k)
            rows.append(lattice.snapshot_walls())
        if exact_stop and n > 1:
            lattice.save_state()
        forces = lattice.batch_updates(np.stack(rows))       # slot k = iteration it+k-1
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
