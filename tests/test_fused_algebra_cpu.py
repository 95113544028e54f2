"""CPU: the moment form of the fused collision (lbm_b200/csrc/d2q9.cuh: collide_fused) restated in
NumPy and compared with the oracle's three phases lattice.macro -> nb_equilibrium -> TRT collision
(oracle/lbm_oracle.c, pinned to the reference).  This guards the ALGEBRA of the production
arithmetic on the CPU; the GPU parity tests measure the kernels themselves."""
import numpy as np
import pytest

from lbm_b200 import cases
from oracle import oracle as orc

W = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)


def collide_fused_numpy(G, om_p, om_m, dev=False):
    """d2q9.cuh: collide_fused, expression by expression (without FMA contraction); dev: deviation
    storage (G = f - w, rho = 1 + sum)."""
    one_m_omp = 1.0 - om_p
    cs, cd = 0.5 * (1.0 - om_p), 0.5 * (1.0 - om_m)
    wp = om_p * W                      # om_p w
    wq = 4.5 * om_p * W                # 4.5 om_p w
    wm = 3.0 * om_m * W                # 3 om_m w
    S = [G[1] + G[2], G[3] + G[4], G[5] + G[6], G[7] + G[8]]
    D = [G[1] - G[2], G[3] - G[4], G[5] - G[6], G[7] - G[8]]
    s = G[0] + ((S[0] + S[1]) + (S[2] + S[3]))
    r = 1.0 + s if dev else s
    mx = (D[0] + D[2]) - D[3]
    my = (D[1] + D[2]) + D[3]
    y = 1.0 / r
    ms = [mx, my, mx + my, my - mx]
    m2 = mx * mx + my * my
    E = s - (1.5 * m2) * y
    F = np.empty_like(G)
    F[0] = one_m_omp * G[0] + E * wp[0]
    for k in range(4):
        q, qb = 2 * k + 1, 2 * k + 2
        Fs = (wq[q] * y) * (ms[k] * ms[k]) + (cs * S[k] + E * wp[q])
        Fd = cd * D[k] + wm[q] * ms[k]
        F[q] = Fs + Fd
        F[qb] = Fs - Fd
    return F, r, mx * y, my * y


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


def test_deviation_storage_is_the_same_update():
    """f32 state is stored as h_q = f_q - w_q (d2q9.cuh: Stored<float>): the fused collision on h, with
    rho = 1 + sum(h) and om_p w dr in place of om_p w rho, yields F - w (checked in double precision)."""
    rng = np.random.default_rng(3)
    om_p, om_m = 1.0 / 0.56, 1.0 / (0.25 / 0.06 + 0.5)
    G = W[:, None] * (1.0 + 0.03 * rng.standard_normal((9, 500)))
    F, r, ux, uy = collide_fused_numpy(G, om_p, om_m)
    H = G - W[:, None]
    Fh, rho, _, _ = collide_fused_numpy(H, om_p, om_m, dev=True)
    assert np.max(np.abs(rho - r)) < 1e-15
    assert np.max(np.abs((Fh + W[:, None]) - F)) < 1e-15     # a few ulps of the populations (0.03 .. 0.44)
