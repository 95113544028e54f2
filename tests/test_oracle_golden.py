"""CPU: pins the oracle (oracle/lbm_oracle.c) and lbm_b200.cases against the golden
vectors generated from the reference by tests/golden/make_golden.py, and against
the reference's own known-answer test (lbm/tst/cavity/test_cavity.py:25-28)."""
import os

import numpy as np
import pytest

from lbm_b200 import cases
from oracle import oracle as orc
from conftest import GOLDEN


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="module")
def ph():
    return np.load(os.path.join(GOLDEN, "phases.npz"))


class _P:
    pass


def _lat(ph):
    p = _P()
    p.nx, p.ny = int(ph["nx"]), int(ph["ny"])
    p.tau_lbm = 0.62
    return orc.OracleLattice(p)


def test_equilibrium_phase(ph):
    lat = _lat(ph)
    lat.rho[:], lat.u[:] = ph["rho0"], ph["u0"]
    lat.equilibrium()
    assert rel(lat.g_eq, ph["g_eq"]) < 5e-16


def test_macro_phase(ph):
    lat = _lat(ph)
    lat.g[:] = ph["g_in"]
    lat.macro()
    assert rel(lat.rho, ph["macro_rho"]) < 5e-16
    assert rel(lat.u, ph["macro_u"]) < 5e-15


def test_collide_stream_phase(ph):
    lat = _lat(ph)
    lat.g[:] = ph["g_in"]
    lat.macro()
    lat.equilibrium()
    assert abs(lat.om_p_lbm - float(ph["om_p"])) == 0.0 and abs(lat.om_m_lbm - float(ph["om_m"])) == 0.0
    lat.collision_stream()
    assert rel(lat.g_eq, ph["cs_g_eq"]) < 1e-15
    assert rel(lat.g_up, ph["cs_g_up"]) < 1e-15
    assert rel(lat.g, ph["cs_g"]) < 1e-15


def _after_cs(ph):
    lat = _lat(ph)
    lat.g[:] = ph["cs_g"]
    lat.g_up[:] = ph["cs_g_up"]
    lat.rho[:], lat.u[:] = ph["macro_rho"], ph["macro_u"]
    for k in ("u_left", "u_right", "u_top", "u_bot", "rho_right"):
        getattr(lat, k)[:] = ph[k]
    return lat


@pytest.mark.parametrize("name,method", [
    ("left", "zou_he_left_wall_velocity"), ("right", "zou_he_right_wall_velocity"),
    ("rightp", "zou_he_right_wall_pressure"), ("top", "zou_he_top_wall_velocity"),
    ("bottom", "zou_he_bottom_wall_velocity")])
def test_zou_he_walls(ph, name, method):
    lat = _after_cs(ph)
    getattr(lat, method)()
    assert rel(lat.g, ph["zh_%s_g" % name]) < 1e-15
    assert rel(lat.u, ph["zh_%s_u" % name]) < 1e-14
    assert rel(lat.rho, ph["zh_%s_rho" % name]) < 1e-15


def test_corners(ph):
    lat = _after_cs(ph)
    lat.zou_he_bottom_wall_velocity()
    lat.zou_he_top_wall_velocity()
    lat.zou_he_bottom_left_corner()
    lat.zou_he_top_left_corner()
    lat.zou_he_top_right_corner()
    lat.zou_he_bottom_right_corner()
    assert rel(lat.g, ph["corner_g"]) < 1e-15
    assert rel(lat.u, ph["corner_u"]) < 1e-15
    assert rel(lat.rho, ph["corner_rho"]) < 1e-15
    # the two tangent diagonals of every corner are exactly zero (nb.py:266-267 ...)
    for (i, j, qs) in ((0, 0, (7, 8)), (0, -1, (5, 6)), (-1, -1, (7, 8)), (-1, 0, (5, 6))):
        for q in qs:
            assert lat.g[q, i, j] == 0.0


@pytest.mark.parametrize("ibb,key", [(True, "bb_ibb_g"), (False, "bb_plain_g")])
def test_bounce_back(ph, ibb, key):
    lat = _after_cs(ph)
    lat.IBB = ibb
    obs = cases.Obstacle(ph["bb_boundary"], ph["bb_ibb"])
    lat.bounce_back_obstacle(obs)
    assert rel(lat.g, ph[key]) < 1e-15
    if ibb:
        cx, cy = lat.drag_lift(obs, 1.0, 0.03, 7.0)
        assert abs(cx - ph["drag_lift"][0]) < 1e-11 * abs(ph["drag_lift"][0])
        assert abs(cy - ph["drag_lift"][1]) < 1e-11 * abs(ph["drag_lift"][1])


def _golden_case(name):
    z = np.load(os.path.join(GOLDEN, "run_%s.npz" % name))
    if name == "cavity32":
        case = cases.Cavity(L_lbm=32, sigma=20)
    elif name == "turek30":
        case = cases.Turek(L_lbm=30, Re_lbm=20.0, sigma=15,
                           links=[cases.Obstacle(z["boundary"], z["ibb"])])
    else:
        case = cases.Poiseuille(L_lbm=20, sigma=10)
    return z, case


@pytest.mark.parametrize("name", ["cavity32", "turek30", "poiseuille20"])
def test_free_running_vs_reference_run(name):
    """The oracle, driven through lbm_b200.cases, reproduces a free-running reference run."""
    z, case = _golden_case(name)
    assert case.nx == int(z["nx"]) and case.ny == int(z["ny"])
    assert case.tau_lbm == float(z["tau"])
    lat = orc.OracleLattice(case)
    orc.run_loop(lat, case, n_iters=int(z["n_iters"]))
    for k in ("g", "g_up", "rho", "u"):
        assert rel(getattr(lat, k), z[k]) < 1e-13, k
    if name == "turek30":
        f = np.array(case.forces)
        assert f.shape == z["forces"].shape
        assert np.max(np.abs(f - z["forces"])) < 1e-10 * np.max(np.abs(z["forces"]))


def test_link_fixture_counts():
    """lbm/tst/lattice/test_lattice.py:27,36 -- 234 / 468 links for the Turek cylinder."""
    assert len(cases.Turek(L_lbm=100).obstacles[0].boundary) == 234
    assert len(cases.Turek(L_lbm=200).obstacles[0].boundary) == 468
    arr = cases.Array()
    assert len(arr.obstacles) == 8 and sum(len(o.boundary) for o in arr.obstacles) == 928
    assert (arr.nx, arr.ny) == (900, 200)
    t = cases.Turek(L_lbm=200, Re_lbm=100.0)
    assert (t.nx, t.ny) == (1073, 200)


def test_reference_known_answer_cavity():
    """lbm/tst/cavity/test_cavity.py:12-28: cavity 100x100, 10 001 iterations, four centre-line
    values to 1e-6 (HEAD itself reproduces them to 7e-7)."""
    case = cases.Cavity()
    assert (case.nx, case.ny, case.it_max, case.sigma) == (100, 100, 10000, 1000)
    assert abs(case.tau_lbm - 1.1) < 1e-12
    lat = orc.OracleLattice(case)
    n = orc.run_loop(lat, case)
    assert n == 10001
    vx, uy = case.line_fields(lat)
    assert abs(vx[10] - 0.12493089684236539) < 1.0e-6
    assert abs(vx[50] - 0.05295104908939561) < 1.0e-6
    assert abs(uy[10] + 0.05968571489630510) < 1.0e-6
    assert abs(uy[50] + 0.19792323493599165) < 1.0e-6


def test_convergence_buffer_matches_reference_buff():
    """cases.ConvergenceBuffer == lbm/src/utils/buff.py on a stored series (values, growth, flag)."""
    z = np.load(os.path.join(GOLDEN, "buff.npz"))
    b = cases.ConvergenceBuffer("drag", float(z["dt"]), float(z["ct"]), int(z["nb"]))
    for k, v in enumerate(z["x"]):
        b.add(v)
        obs, growth = b.mv_avg()
        assert obs == z["obs"][k] and growth == z["growth"][k] and b.obs_cv == bool(z["flag"][k]), k
    assert z["flag"].any()
