"""GPU: full-length runs of the BASELINE configurations through the batched driver, against the
oracle run on the host cores in the same test (north_star: Turek Cd/Cl within 1e-6, driven-cavity
centre-line profiles matching) and against the reference's own known answers."""
import numpy as np
import pytest

from lbm_b200 import cases
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _gpu_run(case, batch=1024, **kw):
    from lbm_b200.lattice import lattice
    from lbm_b200.run import run
    lat = lattice(case, make_dirs=False, **kw)
    n = run(lat, case, batch=batch, quiet=True)
    return lat, n


def test_cavity_100_known_answer_and_profiles():
    """lbm/tst/cavity/test_cavity.py:12-28 on the GPU path: 10 001 iterations at 100 x 100."""
    cg, co = cases.Cavity(), cases.Cavity()
    lat, n = _gpu_run(cg)
    assert n == 10001
    vx, uy = cg.line_fields(lat)
    assert abs(vx[10] - 0.12493089684236539) < 1.0e-6
    assert abs(vx[50] - 0.05295104908939561) < 1.0e-6
    assert abs(uy[10] + 0.05968571489630510) < 1.0e-6
    assert abs(uy[50] + 0.19792323493599165) < 1.0e-6
    lo = orc.OracleLattice(co)
    orc.run_loop(lo, co)
    vxo, uyo = co.line_fields(lo)
    assert np.max(np.abs(vx - vxo)) < 1e-11 and np.max(np.abs(uy - uyo)) < 1e-11


def test_cavity_200_config1_profiles():
    """BASELINE config 1: Re=100, nx=200, 20 001 iterations; centre-line profiles vs the oracle."""
    cg, co = cases.Cavity(L_lbm=200), cases.Cavity(L_lbm=200)
    assert cg.it_max == 20000
    lat, n = _gpu_run(cg)
    lo = orc.OracleLattice(co)
    assert orc.run_loop(lo, co) == n == 20001
    vx, uy = cg.line_fields(lat)
    vxo, uyo = co.line_fields(lo)
    assert np.max(np.abs(vx - vxo)) < 1e-11 and np.max(np.abs(uy - uyo)) < 1e-11
    for k in ("rho", "u", "g_up"):
        a, b = getattr(lat, k), getattr(lo, k)
        assert np.max(np.abs(a - b)) / np.max(np.abs(b)) < 1e-11, k
    # primary vortex of the Re=100 cavity: minimum of u_x on the vertical centre line ~ -0.21 u_lid
    assert -0.23 < uy.min() < -0.19


def test_turek_2d1_config2_converged_drag_lift():
    """BASELINE config 2: Re=20, ny=100, IBB cylinder, stop rule 'obs' (moving-average convergence).
    Same stop iteration as the oracle, Cd/Cl within 1e-6 (HEAD of the reference gives
    Cd = -5.687793, Cl = +0.039803 after 36 454 iterations)."""
    cg, co = cases.Turek(L_lbm=100, Re_lbm=20.0, stop="obs"), cases.Turek(L_lbm=100, Re_lbm=20.0, stop="obs")
    lat, n = _gpu_run(cg, batch=2048)
    lo = orc.OracleLattice(co)
    n_ref = orc.run_loop(lo, co)
    assert n == n_ref
    assert abs(n - 36454) <= 1
    assert abs(cg.avg_drag - co.avg_drag) < 1e-6 and abs(cg.avg_lift - co.avg_lift) < 1e-6
    assert abs(cg.avg_drag + 5.687793) < 2e-6 and abs(cg.avg_lift - 0.039803) < 2e-6
    f, fo = np.array(cg.forces), np.array(co.forces)
    assert np.max(np.abs(f - fo)) < 1e-6


def test_turek_2d2_config3_unsteady_drag_lift_series():
    """BASELINE config 3: Re=100, ny=200 (1073 x 200), first 6000 iterations with the standard ramp;
    the whole Cd/Cl series and its extrema within 1e-6 of the oracle."""
    cg, co = cases.Turek(L_lbm=200, Re_lbm=100.0), cases.Turek(L_lbm=200, Re_lbm=100.0)
    cg.it_max = co.it_max = 5999
    lat, n = _gpu_run(cg, batch=2048)
    lo = orc.OracleLattice(co)
    assert orc.run_loop(lo, co) == n == 6000
    f, fo = np.array(cg.forces), np.array(co.forces)
    assert np.max(np.abs(f - fo)) < 1e-6
    assert abs(np.abs(f[:, 0]).max() - np.abs(fo[:, 0]).max()) < 1e-6
    assert abs(np.abs(f[:, 1]).max() - np.abs(fo[:, 1]).max()) < 1e-6


def test_array_config4_many_obstacles():
    """BASELINE config 4: 8 squares, 928 IBB links, Re=2000 (tau=0.505): 2500 iterations."""
    cg, co = cases.Array(), cases.Array()
    cg.it_max = co.it_max = 2499
    lat, n = _gpu_run(cg, batch=1024)
    lo = orc.OracleLattice(co)
    assert orc.run_loop(lo, co) == n == 2500
    # the default ramp (sigma = 10 nx = 9000) leaves max|u| ~ 1e-3 after 2500 iterations, while the
    # pressure outlet computes u_x = sum/rho - 1 with an absolute rounding floor of ~1e-14: measure u
    # against the lattice velocity scale u_lbm, not against the still tiny max|u|
    for k in ("rho", "u", "g_up", "g"):
        a, b = getattr(lat, k), getattr(lo, k)
        scale = max(np.max(np.abs(b)), cg.u_lbm) if k == "u" else np.max(np.abs(b))
        assert np.max(np.abs(a - b)) / scale < 1e-10, k
