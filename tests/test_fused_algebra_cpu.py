"""CPU: the moment form of the fused collision (lbm_b200/csrc/d2q9.cuh: collide_fused) restated in
NumPy and compared with the oracle's three phases lattice.macro -> nb_equilibrium -> TRT collision
(oracle/lbm_oracle.c, pinned to the reference).  This guards the ALGEBRA of the production
arithmetic on the CPU; the GPU parity tests measure the kernels themselves."""
import numpy as np
import pytest

from lbm_b200 import cases
from oracle import oracle as orc

W = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)


def collide_fused_numpy(G, om_p, om_m):
    """d2q9.cuh: collide_fused, expression by expression (without FMA contraction)."""
    one_m_omp = 1.0 - om_p
    a_self, a_opp = 1.0 - 0.5 * (om_p + om_m), 0.5 * (om_p - om_m)
    wp = om_p * W                      # om_p w
    wq = 4.5 * om_p * W                # 4.5 om_p w
    wm = 3.0 * om_m * W                # 3 om_m w
    s = (((G[0] + G[1]) + (G[2] + G[3])) + ((G[4] + G[5]) + (G[6] + G[7]))) + G[8]
    d56, d78 = G[5] - G[6], G[7] - G[8]
    mx = ((G[1] - G[2]) + d56) - d78
    my = ((G[3] - G[4]) + d56) + d78
    y = 1.0 / s
    ms = [mx, my, mx + my, my - mx]
    h = 1.5 * (mx * mx + my * my)
    F = np.empty_like(G)
    F[0] = (one_m_omp * G[0] + s * wp[0]) - (h * wp[0]) * y
    for k in range(4):
        q, qb = 2 * k + 1, 2 * k + 2
        K = wq[q] * (ms[k] * ms[k]) - h * wp[q]
        M = wm[q] * ms[k]
        rp = s * wp[q]
        F[q] = K * y + (a_self * G[q] + ((rp + M) - a_opp * G[qb]))
        F[qb] = K * y + (a_self * G[qb] + ((rp - M) - a_opp * G[q]))
    return F, s, mx * y, my * y


@pytest.mark.parametrize("tau", [0.505, 0.56, 0.62, 1.7])
def test_moment_form_equals_macro_equilibrium_trt(tau):
    case = cases.Cavity(L_lbm=24, tau_lbm=tau)
    lo = orc.OracleLattice(case)
    rng = np.random.default_rng(17)
    rho = 1.0 + 0.05 * rng.standard_normal((lo.nx, lo.ny))
    u = 0.08 * rng.standard_normal((2, lo.nx, lo.ny))
    lo.rho[:], lo.u[:] = rho, u
    lo.equilibrium()
    lo.g[:] = lo.g_eq * (1.0 + 0.02 * rng.standard_normal(lo.g.shape))      # off-equilibrium populations
    G = lo.g.copy()
    lo.macro()
    lo.equilibrium()
    lo.collision_stream()                                                       # g_up = post-collision
    F, r, ux, uy = collide_fused_numpy(G, lo.om_p_lbm, lo.om_m_lbm)
    scale = np.max(np.abs(lo.g_up))
    assert np.max(np.abs(F - lo.g_up)) / scale < 5e-15
    assert np.max(np.abs(r - lo.rho)) < 5e-15
    assert np.max(np.abs(ux - lo.u[0])) < 5e-16 + 1e-15 * np.max(np.abs(lo.u))
    assert np.max(np.abs(uy - lo.u[1])) < 5e-16 + 1e-15 * np.max(np.abs(lo.u))
    # conservation: the collision keeps mass and momentum of every cell
    assert np.max(np.abs(F.sum(axis=0) - G.sum(axis=0))) < 1e-14
