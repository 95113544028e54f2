#!/usr/bin/env python
"""bench.py -- headline benchmark of the D2Q9 TRT fused collide-stream path.

    python bench.py --gpus N --steps K --warmup W            (N=1: one process)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W    (CPU arm: the reference's algorithm)

Workload (BASELINE.json configs[4]): synthetic lid-driven cavity 32768 x 32768, f64, TRT with
tau = 0.56, ramped lid (u_lbm = 0.1, sigma = 10 nx), initial state g = w_q rho; slab-decomposed
along x over N GPUs.  A "step" is one lattice update of the whole domain.  Prints ONE JSON line.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_LUP = {"f64": 144, "f32": 72}      # 9 population reads + 9 writes
TAU, U_LID = 0.56, 0.1


def lid_rows(solver_like, nx, ny, its, row_len):
    """Wall rows of the cavity for iterations `its` (cavity.py:65-76): only u_top[0,:] is non-zero."""
    sigma = math.floor(10 * nx)
    rows = np.zeros((len(its), row_len))
    for k, it in enumerate(its):
        ret = 1.0 - math.exp(-it ** 2 / (2.0 * sigma ** 2))
        rows[k, 4 * ny:4 * ny + nx] = U_LID * ret
    return rows


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.2] or [l for (_, l) in self.lines]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(nx, ny, dtype, n_gpus, kernel="step"):
    """DRAM bytes per launch of the step kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            for e in json.load(open(p)):
                if (e["nx"], e["ny"], e["dtype"], e["n_gpus"], e.get("kernel", "step")) == (nx, ny, dtype, n_gpus, kernel):
                    return e["dram_bytes_per_launch"]
        except Exception:
            pass
    return None


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's per-phase algorithm (oracle port, OpenMP over all host cores)
# ------------------------------------------------------------------------------------------
def cpu_run(nx, ny, steps, warmup, threads=None):
    from lbm_b200 import cases
    from oracle import oracle as orc
    if threads:
        orc.set_threads(threads)
    case = cases.Cavity(L_lbm=ny, u_lbm=U_LID, tau_lbm=TAU)
    case.nx, case.x_max = nx, float(nx) / ny
    case.sigma = math.floor(10 * nx)
    lat = orc.OracleLattice(case)
    case.initialize(lat)

    def one(it):
        case.set_inlets(lat, it)
        lat.macro()
        lat.equilibrium()
        lat.collision_stream()
        case.set_bc(lat)
    for it in range(warmup):
        one(it)
    t0 = time.perf_counter()
    for it in range(warmup, warmup + steps):
        one(it)
    dt = time.perf_counter() - t0
    return nx * ny * steps / dt / 1e6, dt, orc.get_threads()


def cpu_sample(budget_s, steps_hint=None):
    """Pick a bounded sample of the workload: a square sub-lattice and a step count that take
    about budget_s seconds on this host."""
    rate, _, cores = cpu_run(1024, 1024, 3, 1)              # calibration (also warms the library)
    n = 4096
    while n > 1024 and (steps_hint or 4) * n * n / (rate * 1e6) > budget_s:
        n //= 2
    steps = steps_hint or max(2, min(400, int(budget_s * rate * 1e6 / (n * n))))
    return n, steps, cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    n, _, cores = cpu_sample(90.0, steps_hint=total)
    mlups, dt, cores = cpu_run(n, n, args.steps, args.warmup)
    sample = ("%d x %d sub-lattice of the %d x %d cavity, %d timed + %d warm-up per-phase iterations "
              "(macro, equilibrium, collide-stream, Zou-He), OpenMP on %d threads" %
              (n, n, args.nx, args.ny, args.steps, args.warmup, cores))
    out = {"impl": "reference", "metric": "MLUPS (f64)", "value": mlups, "unit": "MLUPS",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "lid-driven cavity %dx%d f64 TRT (BASELINE configs[4])" % (args.nx, args.ny),
                      "sample": sample},
           "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from lbm_b200.slab import SlabSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d processes (one per GPU)" % (args.gpus, args.gpus))
        raise SystemExit("WORLD_SIZE=%d does not match --gpus %d" % (world, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, K, W = args.nx, args.ny, args.steps, args.warmup
    s = SlabSolver(nx, ny, TAU, dist, rank, world, local, dtype=args.dtype, overlap=not args.no_overlap)
    depth = 1 if args.no_temporal else max(1, min(4, args.depth))
    temporal = depth > 1

    def advance(first_row, n):
        """n lattice updates, up to `depth` consecutive ones per launch (temporal blocking)."""
        s.advance(first_row, n, depth)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    row_len = s.s.row_len
    n_rows = W + K + 1
    rows = lid_rows(s, nx, ny, list(range(n_rows)), row_len)
    s.init_equilibrium(1.0)
    s.set_walls(rows)                       # resident before the timed region
    s.update(0, next_depth=depth)           # iteration 0: collide-only
    if world > 1:                           # NCCL channel set-up outside the timed region
        s.update(0, next_depth=depth)
    advance(0, W)
    s.finish()
    barrier()

    # ---- device-resident throughput ("value") ---------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    l0 = s.s.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(s.compute)
    advance(W, K)
    s.finish()
    e1.record(s.compute)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    t_wall1 = time.time()
    barrier()
    launches = int(sum_over_ranks(s.s.launches - l0))
    ms = max_over_ranks(ms)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    value = nx * ny * K / (ms * 1e-3) / 1e6

    # ---- end to end through the public API with host buffers ("e2e") -----------------------
    # every step: H2D of that step's wall profiles from pinned host memory, one update, D2H of the
    # step's result (the two centre lines of rho,u that cavity.line_fields reads).
    pinned = torch.from_numpy(lid_rows(s, nx, ny, list(range(W + K, W + 2 * K)), row_len)).pin_memory()
    xmid, ymid = nx // 2, ny // 2
    owns_mid = s.x0 <= xmid < s.x0 + s.nxl
    Ke = K
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    per = depth                              # updates per launch = per e2e step
    it = 0
    while it < Ke:
        m = min(per, Ke - it)
        s.set_walls(pinned[it:it + m])
        s.advance(0, m, depth)
        s.finish()
        line_y = s.s.probe_line(1, ymid, m - 1)             # row y = ny/2, this slab's columns
        d2h = line_y.nbytes
        if owns_mid:
            line_x = s.s.probe_line(0, xmid - s.x0, m - 1)  # column x = nx/2
            d2h += line_x.nbytes
        it += m
    torch.cuda.synchronize(dev)
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = nx * ny * Ke / t_e2e / 1e6
    h2d = int(sum_over_ranks(row_len * 8 * per))
    d2h = int(sum_over_ranks(d2h))

    # algorithmic bytes: one read + one write of the nine populations per cell and LAUNCH; a
    # two-update launch (temporal blocking) serves two lattice updates with them.
    upl = min(depth, K)
    bpl = BYTES_PER_LUP[args.dtype] / upl
    peak, peak_src = measured_peak()
    kname = {1: "step", 2: "step2"}.get(upl, "stepw%d" % upl)
    kdesc = {1: "lbm::step_kernel<%s,fused>",
             2: "lbm::step2_kernel<%s,fused,8,64> (two updates per launch; the populations cross HBM once per launch)"}.get(
        upl, "lbm::stepw_kernel<%%s,fused,%d,64> (%d updates per launch, wavefront temporal blocking; the populations "
             "cross HBM once per launch; FP64-issue bound, not HBM bound)" % (upl, upl))
    achieved = bpl * s.nxl * ny / (ms / K * 1e-3) / 1e9      # this rank's kernel: bytes per launch / duration
    out = {"metric": "MLUPS (%s)" % args.dtype, "value": value, "unit": "MLUPS", "n_gpus": world,
           "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
           "config": {"workload": "lid-driven cavity %dx%d %s TRT tau=%.2f (BASELINE configs[4]), x-slabs over %d GPU(s)"
                                  % (nx, ny, args.dtype, TAU, world),
                      "parallelism": "slab%d" % world, "l2": "working set %.1f GB per GPU >> 126 MB L2, no flush needed"
                                  % (2 * 9 * s.nxl * ny * (8 if args.dtype == "f64" else 4) / 1e9),
                      "halo_overlap": bool(s.overlap and (depth <= 2 or s.overlap_wave)),
                      "halo_exchange": "none (one GPU)" if world == 1 else (
                          "NCCL send/recv of %d whole columns per side, one packed message per direction, after each launch" % depth
                          if depth >= 3 else "NCCL send/recv of the populations crossing the interface, overlapped with the interior launch"),
                      "temporal_blocking": ("%d updates per launch (%s)" % (depth, "step2_kernel" if depth == 2 else "stepw_kernel"))
                                           if temporal else "off"},
           "e2e": {"value": e2e_value, "unit": "MLUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "updates_per_step": per,
                   "note": "per launch (%d update(s)): pinned-host wall profiles -> device, the update(s), centre-line "
                           "rho/u -> host; byte counts are per launch; populations stay resident as in the "
                           "reference's in-place time stepping" % per},
           "gpu_launches": launches,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": profiled_traffic(nx, ny, args.dtype, world, kname), "peak_source": peak_src,
                        "bytes_per_lattice_update": bpl, "updates_per_launch": upl,
                        "bytes_per_launch": BYTES_PER_LUP[args.dtype] * s.nxl * ny, "kernel": kdesc % args.dtype,
                        "single_update_roofline_mlups": peak * 1e3 / BYTES_PER_LUP[args.dtype],
                        "value_over_single_update_roofline": value / world / (peak * 1e3 / BYTES_PER_LUP[args.dtype]),
                        "per": "rank 0 slab, bytes per launch / (timed region / launches)"},
           "clocks": clocks}
    if rank == 0:
        if world == 1 and not args.no_cpu:
            n, st, cores = cpu_sample(args.cpu_budget)
            mlups, dt, cores = cpu_run(n, n, st, 1)
            out["cpu_baseline"] = {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port",
                                   "sample": "%d x %d sub-lattice of the workload, %d timed per-phase iterations "
                                             "(reference algorithm restated in C, OpenMP on %d threads), %.1f s"
                                             % (n, n, st, cores, dt)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=32768)
    ap.add_argument("--ny", type=int, default=32768)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--no-temporal", action="store_true")
    ap.add_argument("--depth", type=int, default=4, help="lattice updates per launch (1, 2 = step2_kernel, 3/4 = stepw_kernel)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        reference_arm(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        gpu_arm(args)


if __name__ == "__main__":
    main()
