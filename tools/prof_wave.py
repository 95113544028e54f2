import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lbm_b200.solver import Solver
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dtype = sys.argv[3] if len(sys.argv) > 3 else "f64"
s = Solver(n, n, tau=0.56, dtype=dtype)
s.set_temporal_blocking(1 if depth > 1 else 0)
s.set_temporal_depth(depth)
if os.environ.get('WAVE_ROWS'):
    s.set_tuning('wave_rows', int(os.environ['WAVE_ROWS']))
s.init_equilibrium(1.0)
s.set_walls(s.wall_row(u_top=np.stack([np.full(n, 0.1), np.zeros(n)])))
s.step(1)
s.step(2 * depth)
s.sync()
print(s.last_step_ms() / (2 * depth))
