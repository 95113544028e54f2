"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

The reference is imported through tests/refload.py (scratch copy + matplotlib
stub).  Every array written here is an output of the reference's own code
(lbm/src/core/nb.py, lattice.py, the app classes) on inputs that are stored
beside it, so the fixtures can be replayed anywhere (the GPU box has no
reference).  Files:

  links_*.npz    obstacle link lists + IBB distances + polygons (written to lbm_b200/data/: the restated
                 BASELINE cases of lbm_b200/cases.py load them)
                 (lattice.add_obstacle, lattice.py:290-375)
  phases.npz     one call of every nb_* kernel / lattice.macro on seeded inputs
  run_*.npz      free-running driver loops (run.py:24-54 order): final g, g_up,
                 rho, u and the per-iteration drag/lift series
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refload  # noqa: E402


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def make_lattice(ns, app):
    with refload.in_scratch(), quiet():
        return ns.lattice.lattice(app)


def ref_loop(ns, lat, app, n_iters, with_forces=False):
    """run.py:24-54 without printing / outputs; returns drag-lift series."""
    series = []
    with refload.in_scratch(), quiet():
        app.initialize(lat)
        for it in range(n_iters):
            app.set_inlets(lat, it)
            lat.macro()
            lat.equilibrium()
            lat.collision_stream()
            app.set_bc(lat)
            if with_forces:
                series.append(lat.drag_lift(app.obstacles[0], app.rho_lbm, app.u_avg, app.D_lbm))
    return np.array(series, dtype=np.float64).reshape(-1, 2)


def links_fixture(ns, name, app):
    lat = make_lattice(ns, app)
    with refload.in_scratch(), quiet():
        app.add_obstacles(lat, app.obstacles)
    bnd = [np.asarray(o.boundary, dtype=np.int32) for o in app.obstacles]
    ibb = [np.asarray(o.ibb, dtype=np.float64) for o in app.obstacles]
    poly = [np.asarray(o.polygon, dtype=np.float64) for o in app.obstacles]
    off = np.cumsum([0] + [len(b) for b in bnd]).astype(np.int64)
    poff = np.cumsum([0] + [len(p) for p in poly]).astype(np.int64)
    np.savez_compressed(
        os.path.join(os.path.dirname(os.path.dirname(HERE)), "lbm_b200", "data", "links_%s.npz" % name),
        boundary=np.concatenate(bnd), ibb=np.concatenate(ibb), offsets=off,
        polygon=np.concatenate(poly), polygon_offsets=poff,
        solid=np.argwhere(lat.lattice > 0).astype(np.int32),
        nx=lat.nx, ny=lat.ny, x_min=lat.x_min, x_max=lat.x_max, y_min=lat.y_min,
        y_max=lat.y_max, dx=lat.dx)
    print(name, "links", off[-1], "solid", int((lat.lattice > 0).sum()))
    return lat


def phases_fixture(ns):
    """One call of each reference kernel on the I-rand inputs of SURVEY.md 8(d)."""
    nb = ns.nb
    nx, ny = 20, 14
    lx, ly = nx - 1, ny - 1
    rng = np.random.default_rng(1234)
    rho = 1.0 + 0.01 * rng.standard_normal((nx, ny))
    u = 0.02 * rng.standard_normal((2, nx, ny))
    c = np.array([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1], [1, 1], [-1, -1], [-1, 1], [1, -1]])
    w = np.array([4. / 9.] + [1. / 9.] * 4 + [1. / 36.] * 4)
    ns_tab = np.array([0, 2, 1, 4, 3, 6, 5, 8, 7])
    out = dict(nx=nx, ny=ny, rho0=rho.copy(), u0=u.copy())
    g_eq = np.zeros((9, nx, ny))
    nb.nb_equilibrium(u, c, w, rho, g_eq)
    out["g_eq"] = g_eq.copy()
    # a non-equilibrium g: equilibrium plus seeded noise
    g = g_eq * (1.0 + 0.05 * rng.standard_normal((9, nx, ny)))
    out["g_in"] = g.copy()
    # macro through the reference lattice class (NumPy sum + tensordot)

    class _P:
        pass
    p = _P()
    p.nx, p.ny = nx, ny
    lat = make_lattice(ns, p)
    lat.g = g.copy()
    lat.macro()
    out["macro_rho"], out["macro_u"] = lat.rho.copy(), lat.u.copy()
    # collide + stream
    om_p, om_m = 1.0 / 0.62, 1.0 / (0.25 / (0.62 - 0.5) + 0.5)
    out["om_p"], out["om_m"] = om_p, om_m
    g2, g_up = g.copy(), np.zeros((9, nx, ny))
    geq2 = np.zeros((9, nx, ny))
    nb.nb_equilibrium(lat.u, c, w, lat.rho, geq2)
    nb.nb_col_str(g2, geq2, g_up, om_p, om_m, c, ns_tab, nx, ny, lx, ly)
    out["cs_g"], out["cs_g_up"], out["cs_g_eq"] = g2.copy(), g_up.copy(), geq2.copy()
    # wall profiles
    u_left = np.zeros((2, ny)); u_right = np.zeros((2, ny))
    u_top = np.zeros((2, nx)); u_bot = np.zeros((2, nx))
    yy = np.linspace(0.0, 1.0, ny)
    u_left[0] = 0.05 * 4.0 * yy * (1.0 - yy)
    u_left[1] = 0.003 * rng.standard_normal(ny)
    u_right[0] = 0.01 * rng.standard_normal(ny)
    u_right[1] = 0.002 * rng.standard_normal(ny)
    u_top[0] = 0.1
    u_top[1] = 0.004 * rng.standard_normal(nx)
    u_bot[0] = 0.02 * rng.standard_normal(nx)
    u_bot[1] = 0.003 * rng.standard_normal(nx)
    rho_right = 1.0 + 0.01 * rng.standard_normal(ny)
    out.update(u_left=u_left, u_right=u_right, u_top=u_top, u_bot=u_bot, rho_right=rho_right)
    for name, fn, args in (
            ("left", nb.nb_zou_he_left_wall_velocity, (u_left,)),
            ("right", nb.nb_zou_he_right_wall_velocity, (u_right,)),
            ("rightp", nb.nb_zou_he_right_wall_pressure, (rho_right, u_right)),
            ("top", nb.nb_zou_he_top_wall_velocity, (u_top,)),
            ("bottom", nb.nb_zou_he_bottom_wall_velocity, (u_bot,))):
        gg, uu, rr = g2.copy(), lat.u.copy(), lat.rho.copy()
        if name == "rightp":
            fn(lx, ly, uu, args[0], args[1], rr, gg)
        else:
            fn(lx, ly, uu, args[0], rr, gg)
        out["zh_%s_g" % name], out["zh_%s_u" % name], out["zh_%s_rho" % name] = gg, uu, rr
    # corners, applied after bottom+top walls (they read the wall's rho/u)
    gg, uu, rr = g2.copy(), lat.u.copy(), lat.rho.copy()
    nb.nb_zou_he_bottom_wall_velocity(lx, ly, uu, u_bot, rr, gg)
    nb.nb_zou_he_top_wall_velocity(lx, ly, uu, u_top, rr, gg)
    nb.nb_zou_he_bottom_left_corner_velocity(lx, ly, uu, rr, gg)
    nb.nb_zou_he_top_left_corner_velocity(lx, ly, uu, rr, gg)
    nb.nb_zou_he_top_right_corner_velocity(lx, ly, uu, rr, gg)
    nb.nb_zou_he_bottom_right_corner_velocity(lx, ly, uu, rr, gg)
    out["corner_g"], out["corner_u"], out["corner_rho"] = gg, uu, rr
    # bounce-back on a synthetic link list well inside the box (both IBB branches)
    K = 40
    bi = rng.integers(4, nx - 4, K); bj = rng.integers(4, ny - 4, K); bq = rng.integers(1, 9, K)
    bnd = np.unique(np.stack([bi, bj, bq], axis=1), axis=0).astype(np.int64)
    pib = rng.uniform(0.01, 1.03, len(bnd))
    out["bb_boundary"], out["bb_ibb"] = bnd, pib
    for flag, key in ((True, "bb_ibb_g"), (False, "bb_plain_g")):
        gg = g2.copy()
        nb.nb_bounce_back_obstacle(flag, bnd, ns_tab, c, pib, g_up, gg, lat.u, lat.lattice)
        out[key] = gg
    gg = out["bb_ibb_g"]
    cx, cy = nb.nb_drag_lift(bnd, ns_tab, c, g_up, gg, 1.0, 0.03, 7.0)
    out["drag_lift"] = np.array([cx, cy])
    np.savez_compressed(os.path.join(HERE, "phases.npz"), **out)
    print("phases ok")


def run_fixture(ns, name, app, n_iters, forces):
    lat = make_lattice(ns, app)
    series = ref_loop(ns, lat, app, n_iters, with_forces=forces)
    out = dict(n_iters=n_iters, g=lat.g, g_up=lat.g_up, rho=lat.rho, u=lat.u, forces=series,
               nx=lat.nx, ny=lat.ny, tau=lat.tau_lbm)
    if forces:
        out["boundary"] = np.asarray(app.obstacles[0].boundary, dtype=np.int32)
        out["ibb"] = np.asarray(app.obstacles[0].ibb, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "run_%s.npz" % name), **out)
    print(name, lat.nx, lat.ny, "iters", n_iters, "max|u|", float(np.abs(lat.u).max()))


def buff_fixture(ns):
    """buff.add / buff.mv_avg (lbm/src/utils/buff.py) on a seeded, converging series."""
    rng = np.random.default_rng(5)
    b = ns.buff.buff("drag", 0.004, 1.0e-2, 200, "./")
    n = 2500
    x = -5.7 + 2.0 * np.exp(-np.arange(n) / 120.0) + 1e-4 * rng.standard_normal(n) * np.exp(-np.arange(n) / 400.0)
    obs, growth, flag = np.zeros(n), np.zeros(n), np.zeros(n, dtype=bool)
    for k in range(n):
        b.add(x[k])
        obs[k], growth[k] = b.mv_avg()
        flag[k] = b.obs_cv
    np.savez_compressed(os.path.join(HERE, "buff.npz"), x=x, obs=obs, growth=growth, flag=flag,
                        dt=0.004, ct=1.0e-2, nb=200)
    print("buff: first converged at", int(np.argmax(flag)) if flag.any() else None)


def main():
    ns = refload.load()
    buff_fixture(ns)
    A = ns.app.app_factory.create
    # --- link lists of the BASELINE configs 2, 3, 4 -------------------------
    for name, L in (("turek100", 100), ("turek200", 200)):
        app = A("turek"); app.L_lbm = L; app.compute_lbm_parameters()
        links_fixture(ns, name, app)
    links_fixture(ns, "array", A("array"))
    # --- single calls ---------------------------------------------------------
    phases_fixture(ns)
    # --- free-running loops ---------------------------------------------------
    app = A("cavity"); app.L_lbm = 32; app.compute_lbm_parameters(); app.sigma = 20
    run_fixture(ns, "cavity32", app, 150, False)
    app = A("turek"); app.L_lbm = 30; app.Re_lbm = 20.0; app.compute_lbm_parameters(); app.sigma = 15
    run_fixture(ns, "turek30", app, 120, True)
    app = A("poiseuille"); app.L_lbm = 20; app.compute_lbm_parameters(); app.sigma = 10
    run_fixture(ns, "poiseuille20", app, 100, False)


if __name__ == "__main__":
    main()
