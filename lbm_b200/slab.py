"""x-slab decomposition of the lattice over the GPUs of one box, one process per GPU.

The reference has no distributed path; the analogue of its single address
space is a domain split along x: rank r owns global columns [x0, x0+nxl) and keeps one halo
column on each side.  Because y is the contiguous axis (lattice.py:155), a halo is one contiguous
line per population.  After every update each interface exchanges the three populations that
cross it:

    to the right neighbour : q in {1, 5, 8} (c_x = +1) of the last owned column  -> its halo x = -1
    to the left  neighbour : q in {2, 6, 7} (c_x = -1) of the first owned column -> its halo x = nxl

The update itself is the same kernel with the same per-cell arithmetic, so a slab run is bitwise
identical to a single-GPU run.  Edge columns are updated first, their exchange (NCCL send/recv on a
side stream) overlaps the interior update.
"""
import numpy as np

Q_RIGHT = (1, 5, 8)   # populations moving towards +x
Q_LEFT = (2, 6, 7)    # populations moving towards -x


def slab_bounds(nx, world, rank):
    """Contiguous, balanced partition of nx columns; every slab at least 2 wide (corners need the
    x-neighbour on the same rank, nb.py:254-257)."""
    base, rem = divmod(nx, world)
    if base < 2:
        raise ValueError("need at least 2 columns per slab (nx=%d, world=%d)" % (nx, world))
    x0 = rank * base + min(rank, rem)
    return x0, base + (1 if rank < rem else 0)


def exchange_ops(view, nxl, rank, world, dist):
    """P2P ops that fill the halo columns of `view` ([9, nxl+2, pitch], column index = x + 1)."""
    ops = []
    if rank + 1 < world:
        for q in Q_RIGHT:
            ops.append(dist.P2POp(dist.isend, view[q, nxl], rank + 1))
        for q in Q_LEFT:
            ops.append(dist.P2POp(dist.irecv, view[q, nxl + 1], rank + 1))
    if rank > 0:
        for q in Q_LEFT:
            ops.append(dist.P2POp(dist.isend, view[q, 1], rank - 1))
        for q in Q_RIGHT:
            ops.append(dist.P2POp(dist.irecv, view[q, 0], rank - 1))
    return ops


def exchange_halos(view, nxl, rank, world, dist):
    ops = exchange_ops(view, nxl, rank, world, dist)
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class SlabSolver:
    """One rank's share of a slab-decomposed run (CUDA + NCCL)."""

    def __init__(self, nx, ny, tau, dist, rank, world, device, dtype="f64", arith="fused",
                 right_wall="velocity", overlap=True):
        import torch
        from .solver import Solver
        self.torch, self.dist = torch, dist
        self.rank, self.world = rank, world
        self.nx, self.ny = nx, ny
        self.x0, self.nxl = slab_bounds(nx, world, rank)
        self.compute = torch.cuda.Stream(device=device)
        self.comm = torch.cuda.Stream(device=device)
        self.s = Solver(nx, ny, tau=tau, dtype=dtype, arith=arith, right_wall=right_wall,
                        device=device, x0=self.x0, nxl=self.nxl, stream=self.compute)
        self.overlap = overlap and world > 1 and self.nxl >= 4
        self._halo_ready = None
        self.updates = 0

    def init_equilibrium(self, rho=1.0):
        self.s.init_equilibrium(rho)

    def set_walls(self, rows):
        self.s.set_walls(rows)

    def update(self, row=0):
        """One lattice update of the whole (distributed) domain."""
        torch = self.torch
        s, nxl = self.s, self.nxl
        if self.world == 1:
            s.step(1, row, 0)
            self.updates += 1
            return
        if self._halo_ready is not None:
            self.compute.wait_event(self._halo_ready)       # halos of the current array have landed
        _, oth = self.s.views()
        if self.overlap:
            s.step_columns(0, 1, row)
            s.step_columns(nxl - 1, nxl, row)
            edges_done = torch.cuda.Event()
            edges_done.record(self.compute)
            s.step_columns(1, nxl - 1, row)
            self.comm.wait_event(edges_done)
        else:
            s.step_columns(0, nxl, row)
            done = torch.cuda.Event()
            done.record(self.compute)
            self.comm.wait_event(done)
        with torch.cuda.stream(self.comm):
            exchange_halos(oth, nxl, self.rank, self.world, self.dist)
            self._halo_ready = torch.cuda.Event()
            self._halo_ready.record(self.comm)
        s.flip()
        self.updates += 1

    def finish(self):
        if self._halo_ready is not None:
            self.compute.wait_event(self._halo_ready)
        self.compute.synchronize()
        self.comm.synchronize()

    def gather_populations(self):
        """Post-collision populations of the whole domain on rank 0 (tests)."""
        self.finish()
        local = self.s.populations("post_collision")
        parts = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(local, parts, dst=0)
        if self.rank == 0:
            return np.concatenate(parts, axis=1)
        return None
