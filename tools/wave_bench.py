"""Device-side MLUPS of the multi-update launches (development aid; bench.py is the contract).
usage: wave_bench.py N dtype depth[,depth..] [chunk[,chunk..]] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lbm_b200.solver import Solver

def run(n, dtype, depth, chunk, steps):
    s = Solver(n, n, tau=0.56, dtype=dtype)
    s.set_temporal_blocking(1 if depth > 1 else 0)
    s.set_temporal_depth(depth)
    if chunk:
        s.set_tuning("wave_chunk", chunk)
    if os.environ.get("WAVE_ROWS"):
        s.set_tuning("wave_rows", int(os.environ["WAVE_ROWS"]))
    s.init_equilibrium(1.0)
    s.set_walls(s.wall_row(u_top=np.stack([np.full(n, 0.1), np.zeros(n)])))
    s.step(1 + depth)
    s.sync()
    best = 1e30
    for _ in range(3):
        s.step(steps)
        best = min(best, s.last_step_ms())
    mlups = n * n * steps / (best * 1e-3) / 1e6
    bpl = (144 if dtype == "f64" else 72) / depth
    print("%6d^2 %s depth=%d chunk=%s: %8.3f ms/update %9.1f MLUPS  HBM %7.1f GB/s" %
          (n, dtype, depth, chunk, best / steps, mlups, mlups * bpl / 1e3), flush=True)
    s.close()

if __name__ == "__main__":
    n = int(sys.argv[1]); dtype = sys.argv[2]
    depths = [int(v) for v in sys.argv[3].split(",")]
    chunks = [int(v) for v in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 12
    for d in depths:
        for c in (chunks if d > 2 else [0]):
            run(n, dtype, d, c, steps)
