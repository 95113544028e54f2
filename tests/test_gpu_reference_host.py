"""GPU: the reference's OWN host side -- its app classes (lbm/src/app/*.py), its driver loop
(lbm/src/core/run.py:12-61) and its known-answer tests -- on the B200 path, with lbm_b200.lattice.lattice
in place of lbm.src.core.lattice.lattice and nothing else changed (north_star: "the Python host side
stays as it is").

The reference sources come from /root/reference (build container) or from the git-ignored copy
baseline/_ref/ (tools/make_ref_copy.py), which travels to the GPU box with the gpurun snapshot;
skipped where neither exists."""
import contextlib
import io
import os

import numpy as np
import pytest

from lbm_b200 import cases
from oracle import oracle as orc
from oracle import refload

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refload.available(), reason="no reference sources (run tools/make_ref_copy.py)")]


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def ours(app, **kw):
    from lbm_b200.lattice import lattice
    return lattice(app, **kw)


def test_reference_test_cavity_on_the_gpu_path():
    """lbm/tst/cavity/test_cavity.py:12-28 verbatim, with the B200 lattice: 10 001 iterations of the
    100 x 100 cavity through the reference's run(), then the four known answers (+-1e-6)."""
    ns = refload.load()
    with refload.in_scratch(), quiet():
        app = ns.app.cavity()
        app.output_freq = 10000
        ltc = ours(app)
        ns.run.run(ltc, app)
        vx, uy = app.line_fields(ltc)
    assert ltc.updates == 10001
    assert abs(vx[10] - 0.12493089684236539) < 1.0e-6
    assert abs(vx[50] - 0.05295104908939561) < 1.0e-6
    assert abs(uy[10] + 0.05968571489630510) < 1.0e-6
    assert abs(uy[50] + 0.19792323493599165) < 1.0e-6
    # and the full profiles against the oracle driven by the restated case (same inputs)
    case = cases.Cavity()
    case.output_freq = 1 << 62
    lo = orc.OracleLattice(case)
    orc.run_loop(lo, case)
    vxo, uyo = case.line_fields(lo)
    assert np.max(np.abs(vx - vxo)) < 1e-11 and np.max(np.abs(uy - uyo)) < 1e-11


def test_reference_test_poiseuille_on_the_gpu_path():
    """lbm/tst/poiseuille/test_poiseuille.py:12-25 with the B200 lattice: L1 error of the mid-channel
    profile against the analytic parabola below 1e-3."""
    ns = refload.load()
    with refload.in_scratch(), quiet():
        app = ns.app.poiseuille()
        ltc = ours(app)
        ns.run.run(ltc, app)
        l1_error = app.compute_error(ltc)
    assert l1_error < 1.0e-3
    # the same number from the oracle on the same inputs
    case = cases.Poiseuille()
    lo = orc.OracleLattice(case)
    orc.run_loop(lo, case)
    u = ltc.u
    assert np.max(np.abs(u - lo.u)) < 1e-11 * max(np.max(np.abs(lo.u)), 1e-30) + 1e-13


def _ref_turek(ns, n_it):
    app = ns.app.turek()
    app.L_lbm, app.Re_lbm = 100, 20.0
    app.compute_lbm_parameters()
    app.stop, app.it_max, app.output_freq = "it", n_it - 1, 1 << 62
    app.sigma = 80
    return app


def test_reference_turek_app_per_phase_and_batched():
    """The reference's turek app (ny = 100, Re = 20: BASELINE config 2 inputs, fast ramp) for 1500 iterations:
    (a) through the reference's run() (one fused update per macro() call) and (b) through the batched driver
    lbm_b200.run.run -- same unmodified app object protocol.  Cd/Cl of every iteration within 1e-6 of the
    oracle, the two `drag_lift` logs (turek.py:157-162) byte-identical, link list as the reference's own."""
    from lbm_b200.run import run as batched_run
    ns = refload.load()
    n_it = 1500
    logs, lats = [], []
    for mode in ("per_phase", "batched"):
        with refload.in_scratch(), quiet():
            app = _ref_turek(ns, n_it)
            ltc = ours(app, arith="strict")
            if os.path.exists(ltc.output_dir + "drag_lift"):     # results/<timestamp to the second>/ may be shared
                os.remove(ltc.output_dir + "drag_lift")
            if mode == "per_phase":
                ns.run.run(ltc, app)
            else:
                batched_run(ltc, app, batch=256, quiet=True)
            with open(ltc.output_dir + "drag_lift") as f:
                logs.append(f.read())
            lats.append((ltc.g_up.copy(), ltc.u.copy()))
        assert len(app.obstacles[0].boundary) == 234          # lbm/tst/lattice/test_lattice.py:27
    assert logs[0].count("\n") == n_it
    assert logs[0] == logs[1]
    assert np.array_equal(lats[0][0], lats[1][0]) and np.array_equal(lats[0][1], lats[1][1])
    # against the oracle on the restated inputs
    case = cases.Turek(L_lbm=100, Re_lbm=20.0, sigma=80)
    lo = orc.OracleLattice(case)
    orc.run_loop(lo, case, n_iters=n_it)
    f = np.array([[float(x) for x in l.split()[1:3]] for l in logs[0].splitlines()])
    fo = np.array(case.forces)
    assert f.shape == fo.shape == (n_it, 2)
    assert np.max(np.abs(f - fo)) < 1e-6
    assert np.array_equal(lats[0][0], lo.g_up)                # STRICT arithmetic: bit-identical populations


def test_reference_array_app_links_and_short_run():
    """The reference's array app (BASELINE config 4: 8 squares, IBB, Re = 2000) initialised through the B200
    lattice: add_obstacle yields the reference's 8 x 116 links; 300 iterations within 1e-12 of the oracle."""
    ns = refload.load()
    with refload.in_scratch(), quiet():
        app = ns.app.array()
        app.sigma, app.stop, app.it_max, app.output_freq = 40, "it", 299, 1 << 62
        ltc = ours(app)
        ns.run.run(ltc, app)
        g_up, u = ltc.g_up.copy(), ltc.u.copy()
    assert [len(o.boundary) for o in app.obstacles] == [116] * 8
    case = cases.Array(sigma=40)
    lo = orc.OracleLattice(case)
    orc.run_loop(lo, case, n_iters=300)
    assert np.max(np.abs(g_up - lo.g_up)) / np.max(np.abs(lo.g_up)) < 1e-12
    assert np.max(np.abs(u - lo.u)) / np.max(np.abs(lo.u)) < 1e-11
