#!/usr/bin/env python
"""bench.py -- headline benchmark of the D2Q9 TRT fused collide-stream path.

    python bench.py --gpus N --steps K --warmup W            (N=1: one process)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W    (CPU arm: the reference's Numba path)

Workload (BASELINE.json configs[4]): synthetic lid-driven cavity 32768 x 32768, f64, TRT with
tau = 0.56, ramped lid (u_lbm = 0.1, sigma = 10 nx), initial state g = w_q rho; slab-decomposed
along x over N GPUs.  A "step" is one lattice update of the whole domain.  Prints ONE JSON line.

Timed region: the K-step block is repeated R = ceil(min_updates / K) times (default min_updates = 128,
so that the region lasts >= 1 s on one GPU), every block bracketed by barrier + synchronize, timed with
CUDA events on the launching stream, max over ranks; `value` / `ms_per_step` are the MEDIAN block.
R depends only on the arguments, so the number of updates -- and with it `parity.state_bits_sum`, the
wrap-around sum of the bit patterns of the final populations -- is the same at every N.
"""
import argparse
import contextlib
import io
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# The CPU legs use every host thread they can get: Numba's pool is sized to the cores this process may run
# on; OpenBLAS (np.tensordot in the reference's lattice.macro) is held at one thread -- with its default pool
# it fights Numba's and the reference loop gets 30-50x slower.  Both must be set before
# NumPy / Numba are imported.  torchrun exports OMP_NUM_THREADS=1 to its workers, which would throttle the
# Numba kernels (omp threading layer) and the OpenMP port of the reference arm: undone for that arm.
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("NUMBA_NUM_THREADS", str(host_threads()))
if "--impl" in sys.argv and "reference" in sys.argv:
    os.environ["OMP_NUM_THREADS"] = str(host_threads())

import numpy as np  # noqa: E402

BYTES_PER_LUP = {"f64": 144, "f32": 72}      # 9 population reads + 9 writes
TAU, U_LID = 0.56, 0.1
FP64_PER_UPDATE = 59                          # FP64 instructions of one fused cell update (d2q9.cuh: collide_fused)


def lid_ramp(nx, its):
    """ret(it) of the cavity's lid (cavity.py:70-71), sigma = floor(10 nx)."""
    sigma = math.floor(10 * nx)
    return np.array([1.0 - math.exp(-it ** 2 / (2.0 * sigma ** 2)) for it in its])


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
    """DRAM bytes per launch of the step kernel from the committed ncu capture (profiles/traffic.json):
    dram__bytes_read.sum + dram__bytes_write.sum cannot be measured outside a profiler."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            for e in json.load(open(p)):
                if (e["nx"], e["ny"], e["dtype"], e["n_gpus"], e.get("kernel", "step")) == (nx, ny, dtype, n_gpus, kernel):
                    return e["dram_bytes_per_launch"], e.get("source", "profiles/traffic.json")
        except Exception:
            pass
    return None, None


# ------------------------------------------------------------------------------------------
# CPU arm.  "reference": the UNMODIFIED reference (Numba kernels + NumPy macro, run.py phase order) from
# /root/reference or its git-ignored copy baseline/_ref/.  "port": the same algorithm restated in C with
# OpenMP (oracle/), used where the reference copy is absent.
# ------------------------------------------------------------------------------------------
def port_run(nx, ny, steps, warmup, threads=None):
    from lbm_b200 import cases
    from oracle import oracle as orc
    orc.set_threads(threads or host_threads())
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


def numba_available():
    try:
        from oracle import refload
        if not refload.available():
            return False
        import numba  # noqa: F401
        return True
    except Exception:
        return False


def numba_run(n, steps, warmup):
    """The reference's own loop body (run.py:27-48) on its own cavity app and lattice class, n x n."""
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    os.environ.setdefault("NUMBA_NUM_THREADS", str(host_threads()))
    from oracle import refload
    ns = refload.load()
    import numba
    with refload.in_scratch(), contextlib.redirect_stdout(io.StringIO()):
        app = ns.app.cavity()
        app.L_lbm, app.u_lbm, app.output_freq = n, U_LID, 1 << 62
        app.compute_lbm_parameters()
        app.tau_lbm = TAU
        lat = ns.lattice.lattice(app)
        app.initialize(lat)

        def one(it):
            app.set_inlets(lat, it)
            lat.macro()
            lat.equilibrium()
            lat.collision_stream()
            app.set_bc(lat)
        for it in range(max(warmup, 3)):          # JIT compilation happens here
            one(it)
        t0 = time.perf_counter()
        for it in range(max(warmup, 3), max(warmup, 3) + steps):
            one(it)
        dt = time.perf_counter() - t0
    return n * n * steps / dt / 1e6, dt, int(numba.get_num_threads())


def cpu_baseline(budget_s, steps=None, warmup=1):
    """Bounded sample of the workload on the host cores: a square sub-lattice of the cavity."""
    if numba_available():
        n = 2048                                   # 13-15 MLUPS: ~0.3 s per iteration
        st = steps or max(4, min(40, int(budget_s / 0.3)))
        mlups, dt, cores = numba_run(n, st, warmup)
        kind = "reference"
        what = ("the unmodified reference (Numba nb_* kernels + NumPy macro, run.py phase order, "
                "OPENBLAS_NUM_THREADS=%s, NUMBA_NUM_THREADS=%d)" % (os.environ.get("OPENBLAS_NUM_THREADS", "default"), cores))
    else:
        rate, _, cores = port_run(1024, 1024, 3, 1)
        n = 4096
        while n > 1024 and (steps or 4) * n * n / (rate * 1e6) > budget_s:
            n //= 2
        st = steps or max(2, min(400, int(budget_s * rate * 1e6 / (n * n))))
        mlups, dt, cores = port_run(n, n, st, warmup)
        kind = "port"
        what = "the reference algorithm restated in C (oracle/), OpenMP on %d threads" % cores
    sample = "%d x %d sub-lattice of the cavity, %d timed per-phase iterations after warm-up, %s, %.1f s" % (n, n, st, what, dt)
    return {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": kind, "sample": sample}, dt / st


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, s_per_step = cpu_baseline(60.0, steps=args.steps, warmup=args.warmup)
    out = {"impl": "reference", "metric": "MLUPS (f64)", "value": base["value"], "unit": "MLUPS",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "lid-driven cavity %dx%d f64 TRT (BASELINE configs[4])" % (args.nx, args.ny),
                      "sample": base["sample"]},
           "cpu_baseline": base,
           "e2e": {"value": base["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------
# BASELINE configs 1-4 (the reference's own small cases) on one GPU: device stepping and the batched driver
# ------------------------------------------------------------------------------------------
def small_configs(n_dev=4096, n_drv=8192):
    """Per config: device us per update (lbm_step batches, drag/lift of every update stored on the device)
    and us per iteration of whole runs through lbm_b200.run.run with the app's per-iteration observers: a run of n_drv
    and one of 5 n_drv iterations -- their difference is what an iteration costs once a run is under way, the rest the
    one-off cost of a run (library handle, inlet-model detection, iteration 0 phase by phase, graph capture)."""
    import torch
    from lbm_b200 import _capi as C
    from lbm_b200 import cases
    from lbm_b200.lattice import lattice
    from lbm_b200.run import run
    makers = (("1 cavity Re=100 nx=200", lambda: cases.Cavity(L_lbm=200)),
              ("2 turek 2D-1 Re=20 ny=100 IBB", lambda: cases.Turek(L_lbm=100, Re_lbm=20.0)),
              ("3 turek 2D-2 Re=100 ny=200 IBB", lambda: cases.Turek(L_lbm=200, Re_lbm=100.0)),
              ("4 array Re=2000 ny=200 IBB", lambda: cases.Array()))
    out = []
    for name, mk in makers:
        res = None
        t_runs = {}
        for n_it in (64, n_drv, 5 * n_drv):               # first pass: warm-up (allocations, module state)
            c = mk()
            c.it_max = n_it - 1
            lat = lattice(c, make_dirs=False)
            t0 = time.perf_counter()
            n = run(lat, c, batch=1024, quiet=True)
            torch.cuda.synchronize()
            t_runs[n] = time.perf_counter() - t0
            if n_it != n_drv:
                lat.close()
                continue
            L, h = lat._L, lat._h
            row = np.ascontiguousarray(lat._row[None, :])
            C.check(L.lbm_set_walls(h, 1, row.ctypes.data))
            C.check(L.lbm_sync(h))
            m = 1024
            C.check(L.lbm_step(h, m, 0, 0, 0))            # (graph capture / table build)
            C.check(L.lbm_sync(h))
            t0 = time.perf_counter()
            done = 0
            while done < n_dev:
                C.check(L.lbm_step(h, m, 0, 0, 0))
                done += m
            C.check(L.lbm_sync(h))
            t_raw = time.perf_counter() - t0
            res = {"config": name, "nx": c.nx, "ny": c.ny, "links": int(sum(len(o.boundary) for o in c.obstacles)),
                   "device_us_per_update": t_raw / done * 1e6, "device_mlups": c.nx * c.ny * done / t_raw / 1e6}
            lat.close()
        (n1, t1), (n2, t2) = sorted(t_runs.items())[1:]
        per_it = (t2 - t1) / (n2 - n1)
        res.update({"driver_us_per_iteration": per_it * 1e6, "driver_mlups": c.nx * c.ny / per_it / 1e6,
                    "driver_one_off_ms": (t1 - n1 * per_it) * 1e3, "driver_whole_run_us_per_iteration": t2 / n2 * 1e6})
        # the drop-in mode: the reference's run() phase order (run.py:27-48), one library call per phase from Python
        c = mk()
        lat = lattice(c, make_dirs=False)
        c.initialize(lat)
        n_pp = 1500
        for it in range(n_pp + 100):
            if it == 100:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            c.set_inlets(lat, it)
            lat.macro()
            lat.equilibrium()
            lat.collision_stream()
            c.set_bc(lat)
            c.observables(lat, it)
        torch.cuda.synchronize()
        res["per_phase_us_per_iteration"] = (time.perf_counter() - t0) / n_pp * 1e6
        lat.close()
        out.append(res)
    return {"configs": out, "note": "L2-resident lattices (2.9-15.5 MB): launch/latency bound, a percentage of the HBM roofline is nominal there; "
                                    "device = lbm_step batches of 1024 updates, drag/lift of every update summed on the device: CUDA-graph replay "
                                    "(obstacle-free lattices) or one resident launch per batch (stepr_kernel, lattices with obstacle links); "
                                    "driver = whole runs of %d and %d iterations through lbm_b200.run.run (batches of 1024 updates, one ramp scalar "
                                    "per iteration from the host, per-iteration callbacks of the app replayed on the host while the device "
                                    "executes the next batch): driver_us_per_iteration = difference of the two runs per iteration, "
                                    "driver_one_off_ms = the rest (handle, inlet-model detection, iteration 0, graph capture), "
                                    "driver_whole_run = the longer run all told; per_phase = the reference's own run() loop order on the drop-in lattice class (one fused "
                                    "update per macro() call, drag/lift fetched every iteration)" % (n_drv, 5 * n_drv)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from lbm_b200.slab import SlabSolver
    from lbm_b200.solver import Solver

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
    R = max(1, -(-args.min_updates // K))                  # timed blocks of K steps
    s = SlabSolver(nx, ny, TAU, dist, rank, world, local, dtype=args.dtype, overlap=not args.no_overlap,
                   exchange=args.exchange)
    depth = 1 if args.no_temporal else max(1, min(4, args.depth))
    temporal = depth > 1
    for key, val in (("wave_tail", args.wave_tail), ("wave_chunk", args.wave_chunk)):
        if val is not None:
            s.s.set_tuning(key, val)
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce(x, dist.ReduceOp.MAX) if world > 1 else x

    def sum_over_ranks(x):
        return reduce(x, dist.ReduceOp.SUM) if world > 1 else x

    # inputs: base wall profiles once (u_top[0,:] = u_lbm, cavity.py:73) + the ramp scalar of every update
    n_its = 1 + W + R * K + K + 1
    u_top = np.zeros((2, nx))
    u_top[0, :] = U_LID
    s.s.set_wall_profiles(u_top=u_top)
    ramp_host = torch.from_numpy(lid_ramp(nx, range(n_its))).pin_memory()
    s.s.set_ramp(ramp_host, 0)                 # resident before the timed region
    s.init_equilibrium(1.0)
    s.update(0, next_depth=depth)              # iteration 0: collide-only
    it = 0                                     # ramp index of the next update: update k+1 applies the walls of iteration k (run.py:45)
    if world > 1 and not s.peer:               # NCCL channel set-up outside the timed region
        s.sync_halos()
    s.advance(it, W, depth)
    it += W
    s.finish()
    barrier()

    # ---- device-resident throughput ("value") ---------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    l0 = s.s.launches
    t_wall0 = time.time()
    block_ms = []
    for r in range(R):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(s.compute)
        s.advance(it, K, depth)
        e1.record(s.compute)
        s.finish()
        torch.cuda.synchronize(dev)
        block_ms.append(max_over_ranks(e0.elapsed_time(e1)))
        it += K
    t_wall1 = time.time()
    barrier()
    launches = int(sum_over_ranks(s.s.launches - l0))
    ms = float(np.median(block_ms))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    value = nx * ny * K / (ms * 1e-3) / 1e6

    # ---- end to end through the public API with host buffers ("e2e") -----------------------
    # every launch: H2D of that launch's ramp scalars from pinned host memory (8 bytes per update: the
    # reference's per-step host input is the scalar ret(it), cavity.py:70-73), the updates, D2H of the
    # result (the two centre lines of rho,u that cavity.line_fields reads).
    xmid, ymid = nx // 2, ny // 2
    owns_mid = s.x0 <= xmid < s.x0 + s.nxl
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    buf_y = torch.empty((3, s.nxl), dtype=tdt).pin_memory()       # results land in pinned host memory
    buf_x = torch.empty((3, ny), dtype=tdt).pin_memory()
    barrier()
    t0 = time.perf_counter()
    d2h, done = 0, 0
    while done < K:
        m = min(depth, K - done)
        s.s.set_ramp(ramp_host[it:it + m], it)
        s.advance(it, m, depth)
        line_y = s.probe_line(1, ymid, it + m - 1, out=buf_y)  # row y = ny/2, this slab's columns (waits for the launch)
        d2h = line_y.numel() * line_y.element_size()
        if owns_mid:
            line_x = s.probe_line(0, xmid - s.x0, it + m - 1, out=buf_x)  # column x = nx/2
            d2h += line_x.numel() * line_x.element_size()
        it += m
        done += m
    torch.cuda.synchronize(dev)
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = nx * ny * K / t_e2e / 1e6
    h2d = int(sum_over_ranks(8 * depth))
    d2h = int(sum_over_ranks(d2h))

    # ---- parity, visible to the driver ------------------------------------------------------
    # (a) fingerprint of the final populations: identical at every N for the same --steps/--warmup
    cs = s.s.checksum()
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, cs)
        cs = sum(parts)
    parity = {"state_bits_sum": "0x%016x" % (cs & 0xFFFFFFFFFFFFFFFF), "updates": it + 1,
              "what": "wrap-around 64-bit sum of the bit patterns of the post-collision populations after all "
                      "%d updates of this run (1 + warm-up + %d x %d timed + %d e2e); must be identical at N = 1/2/4/8" % (it + 1, R, K, K)}
    upl = min(depth, K)
    s.close()
    del s
    torch.cuda.empty_cache()
    # (b) N > 1: a 4096 x 2048 slab run (1 + 16 updates, `depth` per launch, the benchmarked exchange) against a
    # single-GPU run of the same binary on rank 0, populations compared bit for bit
    if world > 1:
        cnx, cny, cn = 4096, 2048, 16
        t = SlabSolver(cnx, cny, TAU, dist, rank, world, local, dtype=args.dtype, exchange=args.exchange)
        prof = np.zeros((2, cnx)); prof[0, :] = U_LID
        left = np.zeros((2, cny)); left[0, :] = 0.01 * np.sin(np.arange(cny))
        ret = np.array([1.0 - math.exp(-k ** 2 / 72.0) for k in range(cn + 1)])
        t.s.set_wall_profiles(u_top=prof, u_left=left)
        t.s.set_ramp(ret, 0)
        t.init_equilibrium(1.0)
        t.update(0, next_depth=depth)
        t.advance(1, cn, depth)
        t.finish()
        F = torch.from_numpy(t.s.populations("post_collision")).to(dev)
        parts = [torch.empty_like(F) for _ in range(world)] if rank == 0 else None
        dist.gather(F, parts, dst=0)
        t.close()
        if rank == 0:
            one = Solver(cnx, cny, tau=TAU, dtype=args.dtype, device=local)
            one.set_temporal_blocking(False)
            one.set_wall_profiles(u_top=prof, u_left=left)
            one.set_ramp(ret, 0)
            one.init_equilibrium(1.0)
            one.step(1, 0, 1)
            one.step(cn, 1, 1)
            ref = torch.from_numpy(one.populations("post_collision")).to(dev)
            got = torch.cat(parts, dim=1)
            parity["slab_vs_single_bitwise"] = bool(torch.equal(got.view(torch.int64 if args.dtype == "f64" else torch.int32),
                                                                ref.view(torch.int64 if args.dtype == "f64" else torch.int32)))
            parity["slab_vs_single"] = ("%d x %d cavity + left inflow, 1 + %d updates, %d per launch over %d slabs (%s exchange) vs "
                                        "single-update launches on one GPU" % (cnx, cny, cn, depth, world, args.exchange))
            one.close()

    # ---- roofline ------------------------------------------------------------------------------
    # algorithmic bytes = 144 B (f64) per lattice update; a launch of `upl` updates processes
    # upl x cells units.  `achieved` / `frac` follow that definition; a multi-update launch moves the populations
    # through HBM only once, so it can exceed 1 -- `hbm_used_*` state what the launch really moves.
    nxl = slab_nxl(nx, world, rank)
    peak, peak_src = measured_peak()
    kname = {1: "step", 2: "step2"}.get(upl, "stepw%d" % upl)
    kdesc = {1: "lbm::step_kernel<%s,fused>",
             2: "lbm::step2_kernel<%s,fused,8,64> (two updates per launch)"}.get(
        upl, "lbm::stepw_kernel<%%s,fused,%d,64> (%d updates per launch, wavefront temporal blocking, TMA-fed)" % (upl, upl))
    t_launch = ms * 1e-3 / K * upl                           # average duration of one launch, from the timed region
    achieved = BYTES_PER_LUP[args.dtype] * nxl * ny * upl / t_launch / 1e9
    hbm_used = BYTES_PER_LUP[args.dtype] * nxl * ny / t_launch / 1e9
    traffic, traffic_src = profiled_traffic(nx, ny, args.dtype, world, kname)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp64_peak = 148 * 64 * sm_mhz * 1e6                      # FP64 lanes x clock: thread-instructions per second
    out = {"metric": "MLUPS (%s)" % args.dtype, "value": value, "unit": "MLUPS", "n_gpus": world,
           "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
           "config": {"workload": "lid-driven cavity %dx%d %s TRT tau=%.2f (BASELINE configs[4]), x-slabs over %d GPU(s)"
                                  % (nx, ny, args.dtype, TAU, world),
                      "parallelism": "slab%d" % world, "l2": "working set %.1f GB per GPU >> 126 MB L2, no flush needed"
                                  % (2 * 9 * nxl * ny * (8 if args.dtype == "f64" else 4) / 1e9),
                      "timing": "%d blocks of %d steps, each bracketed by barrier + synchronize; median block reported "
                                "(min %.3f / max %.3f ms per step)" % (R, K, min(block_ms) / K, max(block_ms) / K),
                      "halo_exchange": "none (one GPU)" if world == 1 else (
                          "peer stores over NVLink from the kernel's last stage into the neighbours' halo columns (CUDA IPC), "
                          "device-side flag hand-shake, no host or NCCL in the loop" if args.exchange == "peer" else
                          "NCCL send/recv of %d whole columns per side after each launch" % depth),
                      "temporal_blocking": ("%d updates per launch (%s)" % (depth, "step2_kernel" if depth == 2 else "stepw_kernel"))
                                           if temporal else "off"},
           "e2e": {"value": e2e_value, "unit": "MLUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "updates_per_step": depth,
                   "note": "per launch (%d update(s)): that launch's inlet-ramp scalars (8 B per update, the reference's "
                           "per-step host input ret(it)) from pinned host memory -> device, the update(s), centre-line rho/u -> "
                           "host; byte counts are per launch summed over ranks; base wall profiles and populations stay resident "
                           "as in the reference's in-place time stepping" % depth},
           "gpu_launches": launches,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                        "definition": "SURVEY 8(d): 144 B (f64) per lattice update x cells x updates per launch / launch duration",
                        "hbm_used_gbs": hbm_used, "frac_hbm_used": hbm_used / peak,
                        "bytes_per_launch_moved": BYTES_PER_LUP[args.dtype] * nxl * ny, "updates_per_launch": upl,
                        "kernel": kdesc % args.dtype, "launch_ms": t_launch * 1e3,
                        "fp64_pipe_frac": value / world * 1e6 * FP64_PER_UPDATE / fp64_peak,
                        "fp64_note": "share of the FP64 lanes the update arithmetic occupies: %d FP64 instructions per cell update "
                                     "(zero-redundancy count) x LUPS / (148 SMs x 64 lanes x %.0f MHz); the multi-update kernel is "
                                     "bound by its shared-memory data pipe (72 %%, ncu) and, sustained, by the board's power cap -- "
                                     "not by HBM and not by FP64 issue (DESIGN.md section 4)" % (FP64_PER_UPDATE, sm_mhz),
                        "per": "rank 0 slab, from the median timed block"},
           "parity": parity,
           "clocks": clocks}
    if rank == 0:
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"], _ = cpu_baseline(args.cpu_budget)
            if out["cpu_baseline"]["kind"] == "reference":    # the C/OpenMP port next to it, for reference
                mlups, dt, cores = port_run(2048, 2048, 8, 1)
                out["cpu_baseline_port"] = {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port",
                                            "sample": "2048 x 2048, 8 iterations, oracle/ C restatement with OpenMP"}
        if world == 1 and not args.no_small:
            try:
                out["small_configs"] = small_configs()
            except Exception as e:                            # never lose the headline line
                out["small_configs"] = {"error": repr(e)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def slab_nxl(nx, world, rank):
    from lbm_b200.slab import slab_bounds
    return slab_bounds(nx, world, rank)[1]


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
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="halo exchange of slab runs")
    ap.add_argument("--min-updates", type=int, default=128, help="timed updates: the K-step block is repeated ceil(this / K) times")
    ap.add_argument("--wave-tail", type=int, default=None)
    ap.add_argument("--wave-chunk", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-small", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=8.0)
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
