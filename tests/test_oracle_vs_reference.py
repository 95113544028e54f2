"""CPU, build container only: runs the UNMODIFIED reference (Numba path) beside the
oracle on identical inputs.  Skipped where /root/reference is absent (GPU box);
tests/test_oracle_golden.py carries the same pin through committed vectors."""
import contextlib
import io

import numpy as np
import pytest

import refload
from lbm_b200 import cases
from oracle import oracle as orc

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not refload.available(), reason="reference checkout not present")]


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _ref_run(ns, app, n):
    with refload.in_scratch(), contextlib.redirect_stdout(io.StringIO()):
        lat = ns.lattice.lattice(app)
        forces = []
        app.initialize(lat)
        for it in range(n):
            app.set_inlets(lat, it)
            lat.macro()
            lat.equilibrium()
            lat.collision_stream()
            app.set_bc(lat)
            if getattr(app, "obstacles", None):
                forces.append(lat.drag_lift(app.obstacles[0], app.rho_lbm, app.u_avg, app.D_lbm))
    return lat, np.array(forces)


def test_turek_re20_ny100_live():
    """BASELINE config 2 inputs, 300 iterations with a fast ramp (sigma=40)."""
    ns = refload.load()
    app = ns.app.app_factory.create("turek")
    app.L_lbm, app.Re_lbm = 100, 20.0
    app.compute_lbm_parameters()
    app.sigma = 40
    ref, f_ref = _ref_run(ns, app, 300)
    assert len(app.obstacles[0].boundary) == 234
    case = cases.Turek(L_lbm=100, Re_lbm=20.0, sigma=40)
    # the committed link fixture is what the reference generates now
    assert np.array_equal(case.obstacles[0].boundary, app.obstacles[0].boundary)
    assert np.array_equal(case.obstacles[0].ibb, app.obstacles[0].ibb)
    assert case.tau_lbm == app.tau_lbm and (case.nx, case.ny) == (app.nx, app.ny)
    lat = orc.OracleLattice(case)
    orc.run_loop(lat, case, n_iters=300)
    for k in ("g", "g_up", "rho", "u"):
        assert rel(getattr(lat, k), getattr(ref, k)) < 1e-13, k
    f = np.array(case.forces)
    assert np.max(np.abs(f - f_ref)) < 1e-10 * np.max(np.abs(f_ref))


def test_cavity_200_live():
    """BASELINE config 1 inputs (nx=200), 400 iterations."""
    ns = refload.load()
    app = ns.app.app_factory.create("cavity")
    app.L_lbm = 200
    app.compute_lbm_parameters()
    ref, _ = _ref_run(ns, app, 400)
    case = cases.Cavity(L_lbm=200)
    assert case.tau_lbm == app.tau_lbm and case.sigma == app.sigma and case.it_max == app.it_max
    lat = orc.OracleLattice(case)
    orc.run_loop(lat, case, n_iters=400)
    for k in ("g", "g_up", "rho", "u"):
        assert rel(getattr(lat, k), getattr(ref, k)) < 1e-13, k


def test_array_live():
    """BASELINE config 4 inputs (8 squares, 928 links), 150 iterations, fast ramp."""
    ns = refload.load()
    app = ns.app.app_factory.create("array")
    app.sigma = 30
    ref, _ = _ref_run(ns, app, 150)
    case = cases.Array(sigma=30)
    for o, r in zip(case.obstacles, app.obstacles):
        assert np.array_equal(o.boundary, r.boundary) and np.array_equal(o.ibb, r.ibb)
    assert case.tau_lbm == app.tau_lbm
    lat = orc.OracleLattice(case)
    orc.run_loop(lat, case, n_iters=150)
    for k in ("g", "g_up", "rho", "u"):
        assert rel(getattr(lat, k), getattr(ref, k)) < 1e-12, k  # tau=0.505 amplifies rounding (1.6e-13 seen)


def test_dropin_lattice_has_the_reference_surface():
    """Every public method of the reference's lattice class exists on lbm_b200.lattice.lattice."""
    import inspect
    ns = refload.load()
    from lbm_b200.lattice import lattice as ours
    ref_methods = [n for n, f in inspect.getmembers(ns.lattice.lattice, inspect.isfunction)
                   if not n.startswith("_") and n != "set_default_lbm"]
    assert len(ref_methods) >= 18
    for n in ref_methods:
        assert hasattr(ours, n), "missing method " + n
        a = list(inspect.signature(getattr(ns.lattice.lattice, n)).parameters)
        b = list(inspect.signature(getattr(ours, n)).parameters)
        assert a == b, (n, a, b)
