"""Where the time of a slab-decomposed launch goes (per-phase CUDA events, one process per GPU):

    python -m torch.distributed.run --nproc-per-node N tools/slab_timeline.py [nx ny launches depth]

For the bench workload (32768^2 f64 cavity over N slabs, four-update wavefront launches) it measures, per rank,
  kernel    the launch alone: the same slab shape on a handle without neighbours (no halo traffic at all);
  nccl      launch, then pack -> NCCL send/recv -> unpack on the side stream: launch time, exchange time
            (launch end -> halos landed) and the period from launch start to launch start;
  peer      launch with fused peer stores + device-side flag hand-shake: period from launch start to launch start
            (the wait for the neighbours is inside it).
Rank 0 prints one JSON line with the max / mean over ranks: the difference between `period` and `kernel` is what
the exchange mechanism costs, the difference between `kernel` x N-slab and the one-GPU time is the launch tail."""
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
    from lbm_b200.slab import SlabSolver, slab_bounds
    from lbm_b200.solver import Solver
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    n_launch = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    depth = int(sys.argv[4]) if len(sys.argv) > 4 else 4
    tails = [int(t) for t in sys.argv[5].split(",")] if len(sys.argv) > 5 else [-1]
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_rows = 1 + depth * (n_launch + 2)
    sigma = math.floor(10 * nx)
    ret = np.array([1.0 - math.exp(-it ** 2 / (2.0 * sigma ** 2)) for it in range(n_rows)])
    u_top = np.zeros((2, nx)); u_top[0] = 0.1
    x0, nxl = slab_bounds(nx, world, rank)
    res = {"nx": nx, "ny": ny, "world": world, "depth": depth, "launches": n_launch, "nxl": nxl}

    def stats(ms):
        t = torch.tensor(ms, dtype=torch.float64, device=dev)
        mx, mean = t.clone(), t.clone()
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(mean, op=dist.ReduceOp.SUM)
            mean /= world
        return {"max_over_ranks_ms": float(mx.median()), "mean_over_ranks_ms": float(mean.median()),
                "min_launch_ms": float(mx.min()), "max_launch_ms": float(mx.max())}

    def feed(s):
        s.set_wall_profiles(u_top=u_top)
        s.set_ramp(ret, 0)
        s.init_equilibrium(1.0)

    # (a) the kernel alone on this slab shape
    for tail in tails:
        one = Solver(nx, ny, tau=0.56, device=local, x0=x0, nxl=nxl)
        one.set_tuning("wave_tail", tail)
        feed(one)
        one.step_columns(0, nxl, 0, 0); one.flip()
        ms = []
        for i in range(n_launch + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(one.stream)
            one.stepn_columns(0, nxl, [1 + depth * i + k for k in range(depth)])
            e1.record(one.stream)
            one.flip()
            one.sync()
            ms.append(e0.elapsed_time(e1))
        res["kernel_tail%d" % tail] = stats(ms[2:])
        one.close()
        del one
        torch.cuda.empty_cache()
    if world > 1:
        for exchange in ("nccl", "peer"):
            dist.barrier()
            s = SlabSolver(nx, ny, 0.56, dist, rank, world, local, exchange=exchange)
            feed(s.s)
            s.update(0, next_depth=depth)
            if exchange == "nccl":
                s.sync_halos()
            starts, ends, halos = [], [], []
            for i in range(n_launch + 2):
                if exchange == "nccl" and s._halo_ready is not None:
                    s.compute.wait_event(s._halo_ready)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s.compute)
                starts.append(e0)
                s.updaten([1 + depth * i + k for k in range(depth)])
                # launch end: in nccl mode `updaten` recorded its own event right after the launch; this one follows the
                # (asynchronous) exchange enqueue on the compute stream, i.e. it is the launch end as well
                e1.record(s.compute)
                ends.append(e1)
                if exchange == "nccl":
                    eh = torch.cuda.Event(enable_timing=True)
                    eh.record(s.comm)
                    halos.append(eh)
            s.finish()
            torch.cuda.synchronize(dev)
            launch = [starts[i].elapsed_time(ends[i]) for i in range(2, n_launch + 2)]
            period = [starts[i].elapsed_time(starts[i + 1]) for i in range(2, n_launch + 1)]
            r = {"launch": stats(launch), "period": stats(period)}
            if exchange == "nccl":
                r["exchange_after_launch"] = stats([ends[i].elapsed_time(halos[i]) for i in range(2, n_launch + 2)])
            res[exchange] = r
            s.close()
            del s
            torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
