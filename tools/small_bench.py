"""Device time per update of small (L2-resident) lattices for the launch variants: single updates, forced pairs
(step2_kernel), forced wavefront groups; CUDA-graph replay of 1024-update batches.
    python tools/small_bench.py [nx ny] ..."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbm_b200.solver import Solver


def run(nx, ny, temporal, depth, n=1024, reps=4, pdl=1):
    s = Solver(nx, ny, tau=0.6)
    s.set_tuning("pdl", pdl)
    s.set_temporal_blocking(temporal)
    s.set_temporal_depth(depth)
    u_top = np.zeros((2, nx)); u_top[0] = 0.1
    s.set_wall_profiles(u_top=u_top)
    s.set_ramp(np.linspace(0, 1, n), 0)
    s.init_equilibrium(1.0)
    s.step(1)
    s.step(n, 0, 1); s.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        s.step(n, 0, 1)
    s.sync()
    dt = (time.perf_counter() - t0) / (reps * n) * 1e6
    s.close()
    return dt


if __name__ == "__main__":
    sizes = [(200, 200), (536, 100), (1073, 200), (900, 200)]
    if len(sys.argv) > 2:
        sizes = [(int(sys.argv[1]), int(sys.argv[2]))]
    for nx, ny in sizes:
        print(nx, ny, "single %.2f us (no PDL %.2f)  pairs %.2f us" % (run(nx, ny, 0, 1), run(nx, ny, 0, 1, pdl=0), run(nx, ny, -1, 2)), flush=True)
