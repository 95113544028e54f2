"""Small driver for ncu captures of stepr_kernel (resident batches): one launch of `n` updates on a BASELINE config.
    python tools/prof_resident.py [turek100 | turek200 | array | cavity] [n]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbm_b200 import cases
from lbm_b200.solver import Solver

which = sys.argv[1] if len(sys.argv) > 1 else "turek200"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
c = {"turek100": lambda: cases.Turek(L_lbm=100), "turek200": lambda: cases.Turek(L_lbm=200, Re_lbm=100.0),
     "array": lambda: cases.Array(), "cavity": lambda: cases.Cavity(L_lbm=200)}[which]()
s = Solver(c.nx, c.ny, tau=0.56, right_wall="pressure" if c.obstacles else "velocity")
s.set_tuning("resident", 1)
if c.obstacles:
    s.set_links(c.obstacles)
    yy = np.linspace(0, 1, c.ny)
    u_left = np.zeros((2, c.ny)); u_left[0] = 0.05 * 4 * yy * (1 - yy)
    s.set_wall_profiles(u_left=u_left, rho_right=np.ones(c.ny))
else:
    u_top = np.zeros((2, c.nx)); u_top[0] = 0.1
    s.set_wall_profiles(u_top=u_top)
s.set_ramp(1.0 - np.exp(-np.arange(n) ** 2 / 2e4), 0)
s.init_equilibrium(1.0)
s.step(1)
s.step(n, 0, 1)
s.sync()
print("%s %dx%d: %.2f us per update" % (which, c.nx, c.ny, s.last_step_ms() / n * 1e3))
