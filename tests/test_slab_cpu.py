"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path -- slab partition and
the halo exchange plan -- checked by doing a pull-stream on each slab with NumPy and comparing it
with the pull-stream of the undivided lattice."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lbm_b200 import slab  # noqa: E402

CX = [0, 1, -1, 0, 0, 1, -1, -1, 1]
CY = [0, 0, 0, 1, -1, 1, -1, 1, -1]


def test_slab_bounds_cover_the_domain():
    for nx, world in ((32768, 8), (1073, 4), (17, 8), (200, 3)):
        xs = [slab.slab_bounds(nx, world, r) for r in range(world)]
        assert xs[0][0] == 0 and xs[-1][0] + xs[-1][1] == nx
        for (a, n), (b, _) in zip(xs[:-1], xs[1:]):
            assert a + n == b and n >= 2
        assert max(n for _, n in xs) - min(n for _, n in xs) <= 1
    with pytest.raises(ValueError):
        slab.slab_bounds(7, 4, 0)


def _pull_interior(F):
    """G_q(x, y) = F_q(x - cx, y - cy) for 1 <= y < ny-1 and every x whose source column exists."""
    q, nxp, ny = F.shape
    G = np.full_like(F, np.nan)
    for k in range(9):
        for x in range(nxp):
            xs = x - CX[k]
            if 0 <= xs < nxp:
                G[k, x, 1:ny - 1] = F[k, xs, 1 - CY[k]:ny - 1 - CY[k]]
    return G


def _worker2(rank, world, port, nx, ny, out, depth=2, packed=False):
    """Depth-d exchange: after it, d consecutive pulls of the owned columns (the earlier ones also on
    the rim, as step2_kernel / stepw_kernel do) equal d pulls of the undivided lattice."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    F = rng.standard_normal((9, nx, ny))
    x0, nxl = slab.slab_bounds(nx, world, rank)
    H, pitch = max(2, depth), ny + 5
    view = torch.full((9, nxl + 2 * H, pitch), float("nan"), dtype=torch.float64)
    view[:, H:nxl + H, :ny] = torch.from_numpy(F[:, x0:x0 + nxl])
    if packed:
        slab.exchange_packed(view, nxl, rank, world, dist, H, depth, {})
    else:
        slab.exchange_halos(view, nxl, rank, world, dist, halo=H, depth=depth)
    G2, ref = view.numpy()[:, :, :ny], F
    for _ in range(depth):
        G2, ref = _pull_interior(G2), _pull_interior(ref)
    G2, ref = G2[:, H:nxl + H], ref[:, x0:x0 + nxl]
    ok = True
    for x in range(nxl):
        gx = x0 + x
        for k in range(9):
            if 0 <= gx - depth * CX[k] < nx:
                ok &= np.array_equal(G2[k, x, depth:ny - depth], ref[k, x, depth:ny - depth])
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,depth,packed", [(2, 12, 2, False), (3, 13, 2, False), (2, 17, 4, False), (3, 19, 3, False),
                                                   (2, 17, 4, True), (3, 19, 3, True), (3, 13, 2, True)])
def test_deep_exchange_feeds_several_updates(world, nx, depth, packed):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker2, args=(world, port, nx, 15, out, depth, packed), nprocs=world, join=True)
    assert all(out[r] for r in range(world)) and len(out) == world


def _worker(rank, world, port, nx, ny, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    F = rng.standard_normal((9, nx, ny))                     # same on every rank
    x0, nxl = slab.slab_bounds(nx, world, rank)
    pitch = ny + 3                                           # padded lines, like the device layout
    view = torch.full((9, nxl + 2, pitch), float("nan"), dtype=torch.float64)
    view[:, 1:nxl + 1, :ny] = torch.from_numpy(F[:, x0:x0 + nxl])
    slab.exchange_halos(view, nxl, rank, world, dist, halo=1, depth=1)
    local = view.numpy()[:, :, :ny]
    G = _pull_interior(local)[:, 1:nxl + 1]                  # owned columns
    ref = _pull_interior(F)[:, x0:x0 + nxl]
    ok = True
    for x in range(nxl):
        gx = x0 + x
        for k in range(9):
            if 0 <= gx - CX[k] < nx:                         # source exists in the global lattice
                ok &= np.array_equal(G[k, x, 1:ny - 1], ref[k, x, 1:ny - 1])
    # halo entries that no owned cell pulls from must not have been touched
    untouched = [k for k in range(9) if k not in slab.Q_RIGHT]
    if rank > 0:
        ok &= bool(np.isnan(view.numpy()[untouched, 0, :ny]).all())
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nx", [(2, 11), (3, 10)])
def test_halo_exchange_feeds_the_pull_stream(world, nx):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, nx, 9, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)) and len(out) == world


def test_launch_plan_groups_updates_and_announces_the_next_halo_depth():
    assert slab.launch_plan(10, 4) == [(4, 4), (4, 2), (2, 4)]
    assert slab.launch_plan(9, 4) == [(4, 4), (4, 1), (1, 4)]
    assert slab.launch_plan(7, 3) == [(3, 3), (3, 1), (1, 3)]
    assert slab.launch_plan(5, 2) == [(2, 2), (2, 1), (1, 2)]
    assert slab.launch_plan(3, 1) == [(1, 1), (1, 1), (1, 1)]
    assert slab.launch_plan(0, 4) == []
    for n in range(1, 40):
        for depth in (1, 2, 3, 4):
            plan = slab.launch_plan(n, depth)
            assert sum(d for d, _ in plan) == n and all(1 <= d <= depth for d, _ in plan)
            for (d, nxt), (d2, _) in zip(plan[:-1], plan[1:]):
                assert nxt == d2                     # every exchange prepares exactly the launch that follows
