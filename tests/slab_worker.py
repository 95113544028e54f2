"""torchrun worker for tests/test_gpu_slab.py: runs a slab-decomposed cavity and, on rank 0,
compares it with a single-GPU run of the same binary, bit for bit: the gathered populations on small
lattices, the wrap-around sums of their bit patterns (lbm_state_checksum) on large ones.

    slab_worker.py nx ny n_updates overlap depth_code [exchange] [ramp]
"""
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
    exchange = sys.argv[6] if len(sys.argv) > 6 else "peer"
    use_ramp = len(sys.argv) > 7 and sys.argv[7] == "1"
    depth = {0: 1, 1: 2}.get(tcode, tcode)
    temporal = depth > 1
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tau = 0.56
    ret = np.array([1.0 - math.exp(-it ** 2 / (2.0 * 6.0 ** 2)) for it in range(n_upd)])
    u_top = np.zeros((2, nx)); u_top[0, :] = 0.1
    u_left = np.zeros((2, ny)); u_left[0, :] = 0.01 * np.sin(np.arange(ny))      # some inflow on the left too
    rows = np.zeros((n_upd, 5 * ny + 4 * nx))
    for it in range(n_upd):
        rows[it, 4 * ny:4 * ny + nx] = u_top[0] * ret[it]
        rows[it, 0:ny] = u_left[0] * ret[it]

    def feed(solver):
        if use_ramp:
            solver.set_wall_profiles(u_top=u_top, u_left=u_left)
            solver.set_ramp(ret, 0)
        else:
            solver.set_walls(rows)
    s = SlabSolver(nx, ny, tau, dist, rank, world, local, overlap=overlap, exchange=exchange)
    s.init_equilibrium(1.0)
    feed(s.s)
    s.update(0, next_depth=depth)
    s.advance(0, n_upd - 1, depth)
    s.finish()
    small = nx * ny <= (1 << 20)
    parts = [None] * world
    dist.all_gather_object(parts, s.s.checksum())
    cs = sum(parts) & 0xFFFFFFFFFFFFFFFF
    F = s.gather_populations() if small else None
    s.close()
    ok, err = True, 0.0
    if rank == 0:
        one = Solver(nx, ny, tau=tau, device=local)
        one.set_temporal_blocking(False)
        one.init_equilibrium(1.0)
        feed(one)
        one.step(1)
        one.step(n_upd - 1, 0, 1)
        ok = one.checksum() == cs
        if small:
            ref = one.populations("post_collision")
            ok = ok and bool(np.array_equal(F, ref))
            err = float(np.max(np.abs(F - ref)))
        print(json.dumps({"ok": ok, "max_abs_diff": err, "world": world, "overlap": overlap, "temporal": temporal, "depth": depth,
                          "exchange": exchange, "peer": bool(exchange == "peer"), "ramp": use_ramp, "checksum": "0x%016x" % cs,
                          "compared": "populations + checksum" if small else "checksum"}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
