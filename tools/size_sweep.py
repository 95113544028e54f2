"""MLUPS of 1 / 2 / 4 updates per launch over lattice sizes (which kernel should lbm_step pick?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lbm_b200.solver import Solver
dtype = sys.argv[1] if len(sys.argv) > 1 else "f64"
for (nx, ny) in ((512, 512), (1024, 1024), (2048, 1024), (2048, 2048), (4096, 2048), (4096, 4096), (8192, 8192)):
    out = []
    for depth in (1, 2, 4):
        s = Solver(nx, ny, tau=0.56, dtype=dtype)
        s.set_temporal_blocking(1 if depth > 1 else 0)
        s.set_temporal_depth(depth)
        s.init_equilibrium(1.0)
        s.set_walls(s.wall_row(u_top=np.stack([np.full(nx, 0.1), np.zeros(nx)])))
        s.step(1 + depth)
        s.sync()
        steps = max(8, min(400, int(4e9 / (nx * ny)) // 4 * 4))
        best = 1e30
        for _ in range(3):
            s.step(steps)
            best = min(best, s.last_step_ms())
        out.append(nx * ny * steps / (best * 1e-3) / 1e6)
        s.close()
    print("%5d x %5d %s: depth1 %8.0f  depth2 %8.0f  depth4 %8.0f MLUPS" % (nx, ny, dtype, *out), flush=True)
