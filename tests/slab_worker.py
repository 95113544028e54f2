"""torchrun worker for tests/test_gpu_slab.py: runs a slab-decomposed cavity and, on rank 0,
compares the gathered populations with a single-GPU run of the same binary, bit for bit."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from lbm_b200.slab import SlabSolver
    from lbm_b200.solver import Solver
    nx, ny, n_upd, overlap = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] == "1"
    # updates per launch: 0 -> 1, 1 -> 2 (step2_kernel), 3 / 4 -> wavefront launches
    tcode = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    depth = {0: 1, 1: 2}.get(tcode, tcode)
    temporal = depth > 1
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tau = 0.56
    rows = np.zeros((n_upd, 5 * ny + 4 * nx))
    for it in range(n_upd):
        rows[it, 4 * ny:4 * ny + nx] = 0.1 * (1.0 - math.exp(-it ** 2 / (2.0 * 6.0 ** 2)))
        rows[it, 0:ny] = 0.01 * np.sin(np.arange(ny))      # some inflow on the left too
    s = SlabSolver(nx, ny, tau, dist, rank, world, local, overlap=overlap)
    s.init_equilibrium(1.0)
    s.set_walls(rows)
    s.update(0, next_depth=depth)
    s.advance(0, n_upd - 1, depth)
    F = s.gather_populations()
    ok, err = True, 0.0
    if rank == 0:
        one = Solver(nx, ny, tau=tau, device=local)
        one.set_temporal_blocking(False)
        one.init_equilibrium(1.0)
        one.set_walls(rows)
        one.step(1)
        one.step(n_upd - 1, 0, 1)
        ref = one.populations("post_collision")
        ok = bool(np.array_equal(F, ref))
        err = float(np.max(np.abs(F - ref)))
        print(json.dumps({"ok": ok, "max_abs_diff": err, "world": world, "overlap": overlap, "temporal": temporal, "depth": depth,
                          "checksum": float(np.sum(ref))}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
