"""torchrun worker for tests/test_gpu_slab.py: a channel with an IBB cylinder that straddles a slab
interface (the Turek 2D-1 link list of the reference, shifted to the middle of the channel), run
slab-decomposed and, on rank 0, on a single GPU: populations bitwise equal, drag/lift sums of every
update equal to rounding (the per-rank partial sums are added in a different order)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from lbm_b200 import cases
    from lbm_b200.slab import SlabSolver
    from lbm_b200.solver import Solver
    n_upd = int(sys.argv[1])
    exchange = sys.argv[2] if len(sys.argv) > 2 else "peer"
    place = sys.argv[3] if len(sys.argv) > 3 else "interface"     # cylinder on a slab interface | inside one slab
    depth = int(sys.argv[4]) if len(sys.argv) > 4 else 1          # updates per launch group (4: obstacle band + wavefront launches)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, tau = 536, 100, 0.62
    z = np.load(os.path.join(ROOT, "lbm_b200", "data", "links_turek100.npz"))
    bnd = z["boundary"].copy()
    # cylinder centre -> column nx/2 (an interface for 2 and 4 slabs), or into the middle of the first slab
    shift = (nx // 2 if place == "interface" else nx // (2 * world)) - int(round(bnd[:, 0].mean()))
    bnd[:, 0] += shift
    obstacles = [cases.Obstacle(bnd, z["ibb"])]
    yy = np.linspace(0.0, 1.0, ny)
    rows = np.zeros((n_upd, 5 * ny + 4 * nx))
    for it in range(n_upd):
        a = 1.0 - np.exp(-(it + 1.0) ** 2 / 50.0)
        rows[it, 0:ny] = 0.05 * a * 4.0 * yy * (1.0 - yy)   # parabolic inlet, pressure outlet
        rows[it, 4 * ny + 4 * nx:] = 1.0
    s = SlabSolver(nx, ny, tau, dist, rank, world, local, right_wall="pressure", exchange=exchange)
    s.set_links(obstacles)
    straddles = bool((bnd[:, 0] < s.x0).any() and (bnd[:, 0] >= s.x0).any()) if rank == world // 2 else False
    multi_ok = s.multi_ok
    s.init_equilibrium(1.0, 0.03, 0.0)                      # uniform flow: the cylinder feels a force from the first update
    s.set_walls(rows)
    s.update(0, next_depth=depth)                           # iteration 0: collide only (the next launch reads `depth` halo columns)
    if depth > 1:
        forces = s.advance(0, n_upd - 1, depth, collect_forces=True)
    else:
        forces = []
        for it in range(1, n_upd):
            s.update(it - 1, slot=0)
            forces.append(s.forces(0, 1)[0])
        forces = np.array(forces)
    F = s.gather_populations()
    flags = [None] * world
    dist.all_gather_object(flags, straddles)
    s.close()
    ok, out = True, {}
    if rank == 0:
        one = Solver(nx, ny, tau=tau, device=local, right_wall="pressure")
        one.set_temporal_blocking(False)
        one.set_links(obstacles)
        one.init_equilibrium(1.0, 0.03, 0.0)
        one.set_walls(rows)
        one.step(1)
        one.step(n_upd - 1, 0, 1)
        ref = one.populations("post_collision")
        fref = one.forces(0, n_upd - 1)
        df = float(np.max(np.abs(forces - fref)))
        ok = bool(np.array_equal(F, ref)) and df < 1e-12 and (any(flags) or place != "interface") and float(np.max(np.abs(fref))) > 1e-6
        out = {"ok": ok, "place": place, "depth": depth, "multi_ok": multi_ok, "exchange": exchange, "pop_equal": bool(np.array_equal(F, ref)), "max_force_diff": df, "world": world,
               "straddles": any(flags), "max_force": float(np.max(np.abs(fref)))}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
