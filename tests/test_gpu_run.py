"""GPU: the batched driver (lbm_b200.run.run) reproduces the per-phase loop: same final fields,
same per-iteration drag/lift, same stop iteration for a force-dependent stop rule."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from lbm_b200 import cases
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _turek30(cls=cases.Turek, **kw):
    z = np.load(os.path.join(GOLDEN, "run_turek30.npz"))
    return cls(L_lbm=30, Re_lbm=20.0, sigma=15, links=[cases.Obstacle(z["boundary"], z["ibb"])], **kw)


class ObsStop(cases.Turek):
    """Force-dependent stop rule (the reference's 'obs' mode, base_app.py:63-67): stop as soon as the
    drag changes by less than 2e-4 between two iterations, after the ramp."""
    stop = "obs"

    def check_stop(self, it):
        f = self.forces
        return not (len(f) > 40 and abs(f[-1][0] - f[-2][0]) < 2.0e-4)


@pytest.mark.parametrize("batch,pipeline", [(1, True), (7, False), (7, True), (64, True)])
def test_batched_run_equals_per_phase_loop(batch, pipeline):
    """pipeline: the next batch runs on the device while the callbacks of the previous one are replayed (run.py)."""
    from lbm_b200.lattice import lattice
    from lbm_b200.run import run
    cg, co = _turek30(), _turek30()
    cg.it_max = co.it_max = 130
    lg = lattice(cg, make_dirs=False, arith="strict")
    n = run(lg, cg, batch=batch, quiet=True, pipeline=pipeline)
    lo = orc.OracleLattice(co)
    n_ref = orc.run_loop(lo, co)
    assert n == n_ref == 131
    for k in ("g_up", "g", "rho", "u"):
        assert np.array_equal(getattr(lg, k), getattr(lo, k)), k
    f, fo = np.array(cg.forces), np.array(co.forces)
    assert f.shape == fo.shape == (131, 2)
    assert np.max(np.abs(f - fo)) <= 1e-13 * np.max(np.abs(fo))


@pytest.mark.parametrize("pipeline", [False, True])
def test_batched_run_stops_on_the_same_iteration(pipeline):
    from lbm_b200.lattice import lattice
    from lbm_b200.run import run
    cg, co = _turek30(ObsStop), _turek30(ObsStop)
    lg = lattice(cg, make_dirs=False, arith="strict")
    n = run(lg, cg, batch=50, quiet=True, pipeline=pipeline)
    lo = orc.OracleLattice(co)
    n_ref = orc.run_loop(lo, co)
    assert n == n_ref and 41 < n < 5000
    assert n % 50 not in (0, 1)          # the rule fired inside a batch, so the rollback path ran
    assert len(cg.forces) == len(co.forces) == n
    for k in ("g_up", "g", "rho", "u"):
        assert np.array_equal(getattr(lg, k), getattr(lo, k)), k


def test_outputs_see_the_fields_of_their_iteration():
    from lbm_b200.lattice import lattice
    from lbm_b200.run import run

    class Probe(cases.Cavity):
        def outputs(self, lat, it):
            if it % self.output_freq == 0:
                self.seen.append((it, lat.u.copy(), lat.rho.copy()))
    cg, co = Probe(L_lbm=32, sigma=20), Probe(L_lbm=32, sigma=20)
    for c in (cg, co):
        c.output_freq, c.seen, c.it_max = 25, [], 90
    lg = lattice(cg, make_dirs=False, arith="strict")
    run(lg, cg, batch=40, quiet=True)
    lo = orc.OracleLattice(co)
    co.initialize(lo)
    for it in range(91):                    # run.py order, with outputs after macro
        co.set_inlets(lo, it)
        lo.macro()
        co.outputs(lo, it)
        lo.equilibrium(); lo.collision_stream(); co.set_bc(lo)
    assert [s[0] for s in cg.seen] == [s[0] for s in co.seen] == [0, 25, 50, 75]
    for a, b in zip(cg.seen, co.seen):
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    vx, uy = cg.line_fields(lg)
    vxo, uyo = co.line_fields(lo)
    assert np.array_equal(vx, vxo) and np.array_equal(uy, uyo)


def test_checkpoint_restart_continues_bit_for_bit(tmp_path):
    """Stop after 60 iterations, write a checkpoint, restore it into a fresh lattice and continue:
    identical to the uninterrupted run (the reference has no restart; its state is one array)."""
    from lbm_b200.lattice import lattice
    ca, cb = _turek30(), _turek30()
    la = lattice(ca, make_dirs=False)
    orc.run_loop(la, ca, n_iters=100)
    lb = lattice(cb, make_dirs=False)
    orc.run_loop(lb, cb, n_iters=60)
    ck = str(tmp_path / "state.npz")
    lb.save_checkpoint(ck)
    lc = lattice(cb, make_dirs=False)
    cb.initialize(lc)                       # same app: allocates, uploads a start state (overwritten)
    lc.macro(); lc.equilibrium(); lc.collision_stream(); cb.set_bc(lc)   # records obstacles / BC set
    lc.load_checkpoint(ck)
    for it in range(60, 100):
        cb.set_inlets(lc, it)
        lc.macro(); lc.equilibrium(); lc.collision_stream(); cb.set_bc(lc)
        cb.observables(lc, it)
    for k in ("g_up", "g", "rho", "u"):
        assert np.array_equal(getattr(lc, k), getattr(la, k)), k
    assert np.array_equal(np.array(cb.forces[-40:]), np.array(ca.forces[-40:]))


def test_checkpoint_restart_in_a_fresh_process_state(tmp_path):
    """The real restart case: the checkpoint is loaded into a lattice that never ran the app (no obstacles
    recorded, no links uploaded).  The link lists travel in the checkpoint, so the first update after loading
    bounces back on the cylinder exactly as the uninterrupted run does."""
    from lbm_b200.lattice import lattice
    ca, cb, cc = _turek30(), _turek30(), _turek30()
    la = lattice(ca, make_dirs=False)
    orc.run_loop(la, ca, n_iters=100)
    lb = lattice(cb, make_dirs=False)
    orc.run_loop(lb, cb, n_iters=60)
    ck = str(tmp_path / "state.npz")
    lb.save_checkpoint(ck)
    lb.close()
    lc = lattice(cc, make_dirs=False)           # fresh: nothing initialised, nothing recorded
    lc.load_checkpoint(ck)
    for it in range(60, 100):
        cc.set_inlets(lc, it)
        lc.macro(); lc.equilibrium(); lc.collision_stream(); cc.set_bc(lc)
        cc.observables(lc, it)
    for k in ("g_up", "g", "rho", "u"):
        assert np.array_equal(getattr(lc, k), getattr(la, k)), k
    assert np.array_equal(np.array(cc.forces), np.array(ca.forces[-40:]))


def test_stop_rule_rollback_on_a_lattice_where_launches_take_long():
    """save_state / restore_state are ordered on the library's stream: the exact-stop rollback of the
    batched driver on a lattice large enough for a batch to be still running when the host gets to the copy."""
    from lbm_b200.lattice import lattice
    from lbm_b200.run import run

    class Stop(cases.Cavity):
        stop = "obs"

        def observables(self, lat, it):
            self.seen = it

        def check_stop(self, it):
            return it < 37
    cg, co = Stop(L_lbm=1536, sigma=10), Stop(L_lbm=1536, sigma=10)
    lg = lattice(cg, make_dirs=False, arith="strict")
    n = run(lg, cg, batch=24, quiet=True)
    lo = orc.OracleLattice(co)
    n_ref = orc.run_loop(lo, co)
    assert n == n_ref == 38
    assert np.array_equal(lg.g_up, lo.g_up)


def test_graph_replay_equals_plain_launches():
    """Batches of >= 16 updates on small lattices are captured into a CUDA graph and replayed
    (lbm_step); populations, stored drag/lift sums and the macro fields must not notice."""
    import os
    from lbm_b200.solver import Solver
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "run_turek30.npz"))
    nx, ny, n = 160, 30, 40
    res = []
    for graph in (1, 0):
        s = Solver(nx, ny, tau=0.62, right_wall="pressure")
        s.set_tuning("graph", graph)
        s.set_tuning("resident", 0)                # (the default on this size: one resident launch per batch, test_gpu_resident.py)
        s.set_links([cases.Obstacle(z["boundary"], z["ibb"])])
        s.init_equilibrium(1.0, 0.03, 0.0)
        yy = np.linspace(0.0, 1.0, ny)
        rows = np.zeros((n, s.row_len))
        for k in range(n):
            rows[k, 0:ny] = 0.03 * (1.0 + 0.01 * k) * 4.0 * yy * (1.0 - yy)
            rows[k, 4 * ny + 4 * nx:] = 1.0
        s.set_walls(rows)
        s.step(1)
        out = []
        for rep in range(3):                       # the second and third call replay the cached graph
            l0 = s.launches
            s.step(n, 0, 1, macro_last=True)
            assert s.launches - l0 == n
            out.append((s.populations("post_collision"), s.forces(0, n), s.macro()))
        res.append(out)
        s.close()
    for a, b in zip(*res):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.array_equal(a[2][0], b[2][0]) and np.array_equal(a[2][1], b[2][1])
    assert np.max(np.abs(res[0][-1][1])) > 1e-6


def test_speed_field_on_device_matches_plot_norm():
    """lbm_get_speed == what plot_norm (plot.py:12-15) computes from lattice.u and lattice.lattice."""
    import os
    from lbm_b200.solver import Solver
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "run_turek30.npz"))
    nx, ny = 160, 30
    for dtype in ("f64", "f32"):
        s = Solver(nx, ny, tau=0.62, right_wall="pressure", dtype=dtype)
        s.set_links([cases.Obstacle(z["boundary"], z["ibb"])])
        s.init_equilibrium(1.0, 0.03, 0.002)
        row = s.wall_row(rho_right=np.ones(ny))
        s.set_walls(row[None, :])
        s.step(1)
        s.step(25, 0, 0, macro_last=True)
        rho, u = s.macro()
        solid = np.zeros((nx, ny), dtype=np.uint8)
        solid[40:50, 10:20] = 1
        v = s.speed(solid)
        ref = np.sqrt(u[0] ** 2 + u[1] ** 2)
        ref[solid != 0] = -1.0
        assert v.dtype == ref.dtype and np.array_equal(v, ref)
        assert np.array_equal(s.speed(), np.sqrt(u[0] ** 2 + u[1] ** 2))
        s.close()
