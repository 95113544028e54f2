"""BASELINE configs 1-4 (the reference's own cases; 40 k - 215 k cells, L2 resident, latency bound):
MLUPS of (a) the raw device stepping (lbm_step batches, drag/lift stored per update), (b) the whole
run through the batched driver with the app's per-iteration observers, (c) the CPU port of the
reference algorithm on the host cores.  One JSON line per config; development aid / evidence for
profiles/, bench.py is the contract."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from lbm_b200 import cases


def make(name):
    if name == "cavity200":
        return cases.Cavity(L_lbm=200)
    if name == "turek100":
        return cases.Turek(L_lbm=100, Re_lbm=20.0)
    if name == "turek200":
        return cases.Turek(L_lbm=200, Re_lbm=100.0)
    return cases.Array()


def gpu(name, n_it):
    from lbm_b200.lattice import lattice
    from lbm_b200.run import run
    import torch
    c = make(name)
    c.it_max = n_it - 1
    lat = lattice(c, make_dirs=False)
    t0 = time.perf_counter()
    n = run(lat, c, batch=2048, quiet=True)
    torch.cuda.synchronize()
    t_run = time.perf_counter() - t0
    # raw stepping: the same lattice object, n_it more updates in batches of 2048 with constant walls
    h = lat._handle() if hasattr(lat, "_handle") else None
    from lbm_b200 import _capi as C
    L = lat._L
    rows = np.repeat(lat._row[None, :], 1, axis=0)
    C.check(L.lbm_set_walls(lat._h, 1, rows.ctypes.data))
    C.check(L.lbm_sync(lat._h))
    t0 = time.perf_counter()
    done = 0
    while done < n_it:
        m = min(2048, n_it - done)
        C.check(L.lbm_step(lat._h, m, 0, 0, 0))
        done += m
    C.check(L.lbm_sync(lat._h))
    t_raw = time.perf_counter() - t0
    return c, n, t_run, t_raw


def cpu(name, n_it):
    from oracle import oracle as orc
    c = make(name)
    c.it_max = n_it - 1
    lo = orc.OracleLattice(c)
    t0 = time.perf_counter()
    n = orc.run_loop(lo, c)
    return n, time.perf_counter() - t0, orc.get_threads()


if __name__ == "__main__":
    n_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    n_cpu = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    for name in ("cavity200", "turek100", "turek200", "array"):
        gpu(name, 200)                                      # warm-up (module load, allocations)
        c, n, t_run, t_raw = gpu(name, n_gpu)
        cells = c.nx * c.ny
        nc, t_cpu, threads = cpu(name, n_cpu)
        print(json.dumps({"config": name, "nx": c.nx, "ny": c.ny, "cells": cells, "iterations": n,
                          "gpu_raw_us_per_update": t_raw / n_gpu * 1e6, "gpu_raw_mlups": cells * n_gpu / t_raw / 1e6,
                          "gpu_driver_us_per_iteration": t_run / n * 1e6, "gpu_driver_mlups": cells * n / t_run / 1e6,
                          "cpu_port_mlups": cells * nc / t_cpu / 1e6, "cpu_threads": threads,
                          "nominal_hbm_roofline_frac_raw": cells * n_gpu / t_raw * 144 / 6450.3e9,
                          "note": "L2-resident, launch-latency bound; drag/lift of every update stored on the device"}), flush=True)
