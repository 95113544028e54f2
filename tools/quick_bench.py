"""Quick device-side MLUPS probe (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lbm_b200.solver import Solver

def run(nx, ny, steps, dtype="f64", arith="fused", temporal=True):
    s = Solver(nx, ny, tau=0.56, dtype=dtype, arith=arith)
    s.set_temporal_blocking(-1 if temporal else 0)
    s.init_equilibrium(1.0)
    s.set_walls(s.wall_row(u_top=np.stack([np.full(nx, 0.1), np.zeros(nx)])))
    s.step(3)
    s.sync()
    s.step(steps)
    ms = s.last_step_ms()
    mlups = nx * ny * steps / (ms * 1e-3) / 1e6
    bpl = 144 if dtype == "f64" else 72
    print("%6d x %6d %s %s tb=%d: %8.3f ms/step %9.1f MLUPS %7.1f GB/s" % (nx, ny, dtype, arith, temporal, ms / steps, mlups, mlups * bpl / 1e3), flush=True)
    s.close()

if __name__ == "__main__":
    import sys
    variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1]
    sizes = ((16384, 16384, 10),) if len(sys.argv) > 1 else ((200, 200, 2000), (1073, 200, 2000), (4096, 4096, 50), (16384, 16384, 10), (32768, 32768, 6))
    for (nx, ny, st) in sizes:
        for dt in ("f64", "f32"):
            for tb in variants:
                run(nx, ny, st, dt, "fused", tb)
