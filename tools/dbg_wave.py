import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lbm_b200.solver import Solver
nx, ny, depth, dtype, rows = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
s = Solver(nx, ny, tau=0.58, dtype=dtype)
s.set_temporal_blocking(-1)
s.set_temporal_depth(depth)
s.set_tuning("wave_rows", rows)
s.init_equilibrium(1.0)
s.set_walls(s.wall_row(u_top=np.stack([np.full(nx, 0.1), np.zeros(nx)])))
s.step(1)
try:
    s.step(depth)
    s.sync()
    print("OK", nx, ny, depth, dtype, rows, float(s.populations().sum()))
except Exception as e:
    print("FAIL", nx, ny, depth, dtype, rows, str(e)[-60:])
