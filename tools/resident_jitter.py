"""Per-launch times of resident batches (stepr_kernel): looks for rare slow launches.
    python tools/resident_jitter.py [launches]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbm_b200 import cases
from lbm_b200.solver import Solver

if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    n = 1024
    for name, mk in (("turek 2D-1", lambda: cases.Turek(L_lbm=100)), ("turek 2D-2", lambda: cases.Turek(L_lbm=200, Re_lbm=100.0)),
                     ("array", lambda: cases.Array())):
        c = mk()
        for resident in (1, 0):
            s = Solver(c.nx, c.ny, tau=0.56, right_wall="pressure")
            s.set_tuning("resident", resident)
            s.set_links(c.obstacles)
            yy = np.linspace(0, 1, c.ny)
            u_left = np.zeros((2, c.ny)); u_left[0] = 0.05 * 4 * yy * (1 - yy)
            s.set_wall_profiles(u_left=u_left, rho_right=np.ones(c.ny))
            s.set_ramp(1.0 - np.exp(-np.arange(n) ** 2 / 2e4), 0)
            s.init_equilibrium(1.0)
            s.step(1)
            t = []
            for _ in range(reps):
                s.step(n, 0, 1)
                s.sync()
                t.append(s.last_step_ms() / n * 1e3)
            t = np.array(t)
            print("%-10s resident=%d  us/update: first %.2f  min %.2f  median %.2f  p90 %.2f  max %.2f  (> 1.5 x median: %d of %d)"
                  % (name, resident, t[0], t.min(), np.median(t), np.percentile(t, 90), t.max(), int((t > 1.5 * np.median(t)).sum()), reps), flush=True)
            s.close()
