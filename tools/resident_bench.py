"""Device time per update of BASELINE configs 1-4 (with their obstacle links): resident batches (stepr_kernel, 1 / 2 / 3
blocks per SM) against CUDA-graph replay of one launch per update.
    python tools/resident_bench.py [n_updates_per_batch]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbm_b200 import cases
from lbm_b200.solver import Solver


def run(c, resident, blocks, n=1024, reps=6, dtype="f64"):
    s = Solver(c.nx, c.ny, tau=0.56, right_wall="pressure" if c.obstacles else "velocity", dtype=dtype)
    s.set_tuning("resident", resident)
    if blocks:
        s.set_tuning("resident_blocks", blocks)
    if c.obstacles:
        s.set_links(c.obstacles)
    if c.obstacles:
        yy = np.linspace(0, 1, c.ny)
        u_left = np.zeros((2, c.ny)); u_left[0] = 0.05 * 4 * yy * (1 - yy)
        s.set_wall_profiles(u_left=u_left, rho_right=np.ones(c.ny))
    else:
        u_top = np.zeros((2, c.nx)); u_top[0] = 0.1
        s.set_wall_profiles(u_top=u_top)
    s.set_ramp(1.0 - np.exp(-np.arange(n) ** 2 / 2e4), 0)
    s.init_equilibrium(1.0)
    s.step(1)
    s.step(n, 0, 1); s.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        s.step(n, 0, 1)
    s.sync()
    dt = (time.perf_counter() - t0) / (reps * n) * 1e6
    chk = s.checksum()
    s.close()
    return dt, chk


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    for name, mk in (("cavity 200x200", lambda: cases.Cavity(L_lbm=200)), ("turek 2D-1 536x100", lambda: cases.Turek(L_lbm=100)),
                     ("turek 2D-2 1073x200", lambda: cases.Turek(L_lbm=200, Re_lbm=100.0)), ("array 900x200", lambda: cases.Array())):
        c = mk()
        g, cg = run(c, 0, 0, n)
        out = ["%-20s graph %.2f us" % (name, g)]
        for b in (1, 2, 3):
            r, cr = run(c, 1, b, n)
            out.append("resident/%d %.2f us%s" % (b, r, "" if cr == cg else " CHECKSUM DIFFERS"))
        r32, _ = run(c, 1, 0, n, dtype="f32")
        out.append("f32 %.2f us" % r32)
        print("  ".join(out), flush=True)
