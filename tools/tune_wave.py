"""Launch-shape sweep of the wavefront kernel on one GPU: ms per launch and GLUPS of `depth`-update launches on an
nx x ny lattice (e.g. 4096 x 32768 = the slab of an 8-GPU run) for chunk widths and tail-chunk widths.

    python tools/tune_wave.py nx ny [depth] [chunks: 0=auto,...] [tails: -1=auto,0=off,...] [launches] [key=value ...]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from lbm_b200.solver import Solver
    nx, ny = int(sys.argv[1]), int(sys.argv[2])
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    chunks = [int(c) for c in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
    tails = [int(c) for c in sys.argv[5].split(",")] if len(sys.argv) > 5 else [-1]
    n = int(sys.argv[6]) if len(sys.argv) > 6 else 8
    extra = [a.split("=") for a in sys.argv[7:]]
    dtype = "f64"
    for chunk in chunks:
        for tail in tails:
            s = Solver(nx, ny, tau=0.56, dtype=dtype)
            if chunk:
                s.set_tuning("wave_chunk", chunk)
            s.set_tuning("wave_tail", tail)
            for k, v in extra:
                if k == "dtype":
                    continue
                s.set_tuning(k, int(v))
            u_top = np.zeros((2, nx)); u_top[0] = 0.1
            s.set_wall_profiles(u_top=u_top)
            s.set_ramp(np.linspace(0.0, 1.0, depth * (n + 3) + 1), 0)
            s.init_equilibrium(1.0)
            s.step_columns(0, nx, 0, 0); s.flip()
            ms = []
            for i in range(n + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s.stream)
                s.stepn_columns(0, nx, [1 + depth * i + k for k in range(depth)])
                e1.record(s.stream)
                s.flip()
                s.sync()
                ms.append(e0.elapsed_time(e1))
            ms = ms[2:]
            if len(ms) >= 24:                    # sustained run: the later half (clocks settle under the power cap)
                ms = ms[len(ms) // 2:]
            med = float(np.median(ms))
            print(json.dumps({"nx": nx, "ny": ny, "depth": depth, "chunk": chunk, "tail": tail, "extra": dict(extra),
                              "ms_per_launch": med, "min_ms": min(ms), "glups": nx * ny * depth / med / 1e6}), flush=True)
            s.close()
            del s
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
