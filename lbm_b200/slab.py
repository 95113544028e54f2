"""x-slab decomposition of the lattice over the GPUs of one box, one process per GPU.

The reference has no distributed path; the analogue of its single address
space is a domain split along x: rank r owns global columns [x0, x0+nxl) and keeps halo columns
on each side (layout.halo = 4, the deepest multi-update launch).  Because y is the contiguous axis
(lattice.py:155), a halo is one contiguous line per population.  After a single update each
interface exchanges the three populations that cross it:

    to the right neighbour : q in {1, 5, 8} (c_x = +1) of the last owned column  -> its halo x = -1
    to the left  neighbour : q in {2, 6, 7} (c_x = -1) of the first owned column -> its halo x = nxl

Multi-update launches need deeper halos (see _plan and exchange_packed), obstacles two whole
columns.  The update itself is the same kernel with the same per-cell arithmetic, so a slab run
is bitwise identical to a single-GPU run.

Two exchange mechanisms (SlabSolver(exchange=...)):

  "peer"  (default on CUDA)  the neighbours' population buffers are mapped into this process (CUDA IPC,
          lbm_peer_* of include/lbm_b200.h).  A wavefront launch stores its four edge columns straight
          into the neighbour's halo columns from its last pipeline stage (NVLink peer stores, fused with
          the compute); single- and two-update launches are followed by one small copy kernel that does
          the same.  Ordering between the ranks is a pair of flag words per interface, written and
          polled on the device: no host, no NCCL and no staging copies in the loop.
  "nccl"  the halo columns travel as NCCL send/recv pairs on a high-priority side stream; for single-
          and two-update launches the edge columns are updated first and their exchange overlaps the
          interior update; a wavefront launch covers the slab in one go and the exchange follows it.
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


# what the neighbour needs after ONE update (its edge cells pull across the interface) and after a
# TWO-update launch (it recomputes the first update on my edge column, section "temporal blocking"
# of DESIGN.md).  Entries: (my column counted from the interface, populations, its halo depth).
_PLAN1 = {"right": [(0, Q_RIGHT)], "left": [(0, Q_LEFT)]}
_PLAN2 = {"right": [(0, (0, 3, 4) + Q_RIGHT), (1, Q_RIGHT)], "left": [(0, (0, 3, 4) + Q_LEFT), (1, Q_LEFT)]}
_ALL = tuple(range(9))


def _plan(depth):
    """depth >= 3 (wavefront launches): the neighbour recomputes depth-1 updates on my edge columns and
    streams whole columns in (TMA bulk copies of all nine planes), so it gets my last `depth` columns
    complete.  (Column d only needs the populations that can still reach the interface in depth-d
    steps; sending all nine keeps the copies contiguous and costs 9.4 MB per side at ny = 32768.)"""
    if depth == 1:
        return _PLAN1
    if depth == 2:
        return _PLAN2
    return None     # handled as blocks of `depth` consecutive columns per plane (exchange_ops)


def exchange_ops(view, nxl, rank, world, dist, halo=1, depth=1, full=False):
    """P2P ops that fill the halo columns of `view` ([9, nxl+2*halo, pitch], column index = x + halo)
    for the next launch: depth 1 = single update, depth 2 = two-update launch, 3/4 = wavefront launch.
    full: whole columns (all nine populations) at any depth -- interpolated bounce-back reads both
    directions of a link up to two columns away (nb.py:98-104)."""
    plan = None if full else _plan(depth)
    ops = []
    if plan is None:
        # [q][x][y] keeps the `depth` edge columns of one plane contiguous: one message per plane
        d = depth
        if rank + 1 < world:
            for q in range(9):
                ops.append(dist.P2POp(dist.isend, view[q, halo + nxl - d:halo + nxl], rank + 1))
                ops.append(dist.P2POp(dist.irecv, view[q, halo + nxl:halo + nxl + d], rank + 1))
        if rank > 0:
            for q in range(9):
                ops.append(dist.P2POp(dist.isend, view[q, halo:halo + d], rank - 1))
                ops.append(dist.P2POp(dist.irecv, view[q, halo - d:halo], rank - 1))
        return ops
    if rank + 1 < world:
        for d, qs in plan["right"]:
            for q in qs:
                ops.append(dist.P2POp(dist.isend, view[q, halo + nxl - 1 - d], rank + 1))
        for d, qs in plan["left"]:
            for q in qs:
                ops.append(dist.P2POp(dist.irecv, view[q, halo + nxl + d], rank + 1))
    if rank > 0:
        for d, qs in plan["left"]:
            for q in qs:
                ops.append(dist.P2POp(dist.isend, view[q, halo + d], rank - 1))
        for d, qs in plan["right"]:
            for q in qs:
                ops.append(dist.P2POp(dist.irecv, view[q, halo - 1 - d], rank - 1))
    return ops


def exchange_halos(view, nxl, rank, world, dist, halo=1, depth=1, full=False):
    ops = exchange_ops(view, nxl, rank, world, dist, halo, depth, full)
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def exchange_packed(view, nxl, rank, world, dist, halo, d, stage):
    """Whole halo columns (wavefront launches, obstacles): the d edge columns of the nine planes are
    packed into ONE message per direction (a strided copy into a staging buffer, 9.4 MB at d = 4,
    ny = 32768), sent with a single send/recv pair per neighbour and unpacked on arrival (18 P2P
    operations per neighbour and exchange cost ~0.5 ms per launch at 4 GPUs, measured).  `stage` is a
    dict that keeps the staging buffers between calls."""
    import torch
    h = halo
    key = (d, view.dtype, view.shape[2])
    if stage.get("key") != key:
        stage.clear()
        stage["key"] = key
        for name in ("send_r", "recv_r", "send_l", "recv_l"):
            stage[name] = torch.empty((9, d, view.shape[2]), dtype=view.dtype, device=view.device)
    ops = []
    if rank + 1 < world:
        stage["send_r"].copy_(view[:, h + nxl - d:h + nxl])
        ops += [dist.P2POp(dist.isend, stage["send_r"], rank + 1), dist.P2POp(dist.irecv, stage["recv_r"], rank + 1)]
    if rank > 0:
        stage["send_l"].copy_(view[:, h:h + d])
        ops += [dist.P2POp(dist.isend, stage["send_l"], rank - 1), dist.P2POp(dist.irecv, stage["recv_l"], rank - 1)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if rank + 1 < world:
        view[:, h + nxl:h + nxl + d].copy_(stage["recv_r"])
    if rank > 0:
        view[:, h - d:h].copy_(stage["recv_l"])


def launch_plan(n, depth):
    """How n updates are grouped into launches of at most `depth` updates: [(updates of the launch,
    updates of the launch after it)], the second entry being what the halo exchange that follows the
    launch has to prepare (the last launch prepares for a full-depth one: what comes next is unknown)."""
    plan, k = [], 0
    while k < n:
        d = min(depth, n - k)
        plan.append((d, min(depth, n - k - d) or depth))
        k += d
    return plan


class SlabSolver:
    """One rank's share of a slab-decomposed run (CUDA + NCCL)."""

    def __init__(self, nx, ny, tau, dist, rank, world, device, dtype="f64", arith="fused",
                 right_wall="velocity", overlap=True, exchange="peer"):
        import torch
        from .solver import Solver
        self.torch, self.dist = torch, dist
        self.rank, self.world = rank, world
        self.nx, self.ny = nx, ny
        self.x0, self.nxl = slab_bounds(nx, world, rank)
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.peer = exchange == "peer" and world > 1
        self.compute = torch.cuda.Stream(device=device)
        # the halo exchange must not queue behind the thousands of pending blocks of the interior
        # launch: high-priority stream, its kernels take the next free SM slots
        self.comm = torch.cuda.Stream(device=device, priority=-1)
        self.s = Solver(nx, ny, tau=tau, dtype=dtype, arith=arith, right_wall=right_wall,
                        device=device, x0=self.x0, nxl=self.nxl, stream=self.compute, own_buffers=self.peer)
        if world > 1 and self.nxl < self.s.layout.halo:
            raise ValueError("slab of %d columns is narrower than the halo (%d)" % (self.nxl, self.s.layout.halo))
        if self.peer:
            # every rank publishes the IPC handles of its buffers and flag words; each maps its two neighbours
            infos = [None] * world
            dist.all_gather_object(infos, self.s.peer_export())
            if rank > 0:
                self.s.peer_attach(0, infos[rank - 1])
            if rank + 1 < world:
                self.s.peer_attach(1, infos[rank + 1])
            dist.barrier()
        self.overlap = overlap and world > 1 and self.nxl >= 4 and not self.peer
        self._halo_ready = None
        self.edge = 16                   # columns of the edge launches of update2 (one tile)
        self.updates = 0
        self.n_obs = 0
        self._stage = {}
        self.multi_ok = True             # multi-update launches possible on every rank (see set_links)
        self.overlap_wave = False        # wavefront launches: one launch per slab, then the exchange (measured faster)

    def set_links(self, obstacles, use_ibb=True):
        """Obstacle link lists (GLOBAL column indices, the reference's obstacle.boundary / .ibb): every
        rank passes the whole list, the library keeps the links whose fluid node lies in its slab.
        With obstacles an update is one launch over the whole slab (the link blocks ride along), the
        halo exchange carries two whole columns (IBB stencil), and forces() sums the per-rank
        momentum-exchange sums over the ranks (one small all-reduce)."""
        self.s.set_links(obstacles, use_ibb)
        self.n_obs = len(obstacles) if obstacles else 0
        # multi-update launches with bodies: every rank must be able to (bodies in other slabs, or an obstacle band clear
        # of this slab's interfaces), because all ranks issue the same sequence of update groups
        ok = 1 if self.s.can_stepn() else 0
        if self.world > 1:
            t = self.torch.tensor([ok], dtype=self.torch.int32, device=self.s.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            ok = int(t.item())
        self.multi_ok = bool(ok)

    def forces(self, first, n):
        """[n, n_obs, 2] momentum-exchange sums of update slots first..first+n-1, whole domain."""
        torch = self.torch
        self.finish()
        f = self.s.forces(first, n)
        if self.world > 1:
            t = torch.from_numpy(f).to(self.s.device)
            self.dist.all_reduce(t)
            f = t.cpu().numpy()
        return f

    def init_equilibrium(self, rho=1.0, ux=0.0, uy=0.0):
        self.s.init_equilibrium(rho, ux, uy)

    def set_walls(self, rows):
        self.s.set_walls(rows)

    def close(self):
        """Unmap the neighbours' buffers before anybody frees them (collective)."""
        if self.s is None:
            return
        self.finish()
        if self.peer:
            self.dist.barrier()
            self.s.peer_detach()
            self.dist.barrier()
        self.s.close()
        self.s = None

    def _group_done(self):
        """Peer mode: end of one update group (launches + pushes): signal the neighbours, flip."""
        self.s.peer_signal()
        self.s.flip()

    def sync_halos(self):
        """Fill the neighbours' halos of the CURRENT array (after set_post_collision / a restart)."""
        if self.world == 1:
            return
        if self.peer:
            self.s.peer_push(0)
            self.s.peer_signal()
        else:
            cur, _ = self.s.views()
            ev = self.torch.cuda.Event()
            ev.record(self.compute)
            self._exchange(cur, ev, self.s.layout.halo)

    def _exchange(self, oth, after, depth):
        torch = self.torch
        self.comm.wait_event(after)
        full = self.n_obs > 0
        if full:
            depth = max(depth, 2)
        with torch.cuda.stream(self.comm):
            if full or depth >= 3:
                exchange_packed(oth, self.nxl, self.rank, self.world, self.dist, self.s.layout.halo, depth, self._stage)
            else:
                exchange_halos(oth, self.nxl, self.rank, self.world, self.dist, self.s.layout.halo, depth)
            self._halo_ready = torch.cuda.Event()
            self._halo_ready.record(self.comm)

    def update(self, row=0, next_depth=1, slot=0):
        """One lattice update of the whole (distributed) domain; next_depth = 2 if the next launch is a
        two-update one (it needs a deeper halo); slot = force slot of this update (obstacles)."""
        torch = self.torch
        s, nxl = self.s, self.nxl
        if self.world == 1:
            s.step_columns(0, nxl, row, slot)
            s.flip()
            self.updates += 1
            return
        if self.peer:
            s.step_columns(0, nxl, row, slot)
            s.peer_push(1)                                  # my four edge columns -> the neighbours' halos
            self._group_done()
            self.updates += 1
            return
        if self._halo_ready is not None:
            self.compute.wait_event(self._halo_ready)       # halos of the current array have landed
        _, oth = s.views()
        ev = torch.cuda.Event()
        # the exchange that follows copies my last `next_depth` columns (2 with obstacles): the edge launches
        # must cover all of them before `ev`, or the copy would race with the interior launch
        w = max(2, next_depth)
        if self.n_obs:
            s.step_columns(0, nxl, row, slot)               # link blocks ride along with the whole slab
            ev.record(self.compute)
        elif self.overlap and nxl >= 4 * w:
            s.step_columns(0, w, row)
            s.step_columns(nxl - w, nxl, row)
            ev.record(self.compute)
            s.step_columns(w, nxl - w, row)
        else:
            s.step_columns(0, nxl, row)
            ev.record(self.compute)
        self._exchange(oth, ev, next_depth)
        s.flip()
        self.updates += 1

    def update2(self, row1=0, row2=0, next_depth=2):
        """Two lattice updates in one launch per column range (temporal blocking)."""
        torch = self.torch
        s, nxl = self.s, self.nxl
        if self.world == 1:
            s.step2_columns(0, nxl, row1, row2)
            s.flip()
            self.updates += 2
            return
        if self.peer:
            s.step2_columns(0, nxl, row1, row2)
            s.peer_push(1)
            self._group_done()
            self.updates += 2
            return
        if self._halo_ready is not None:
            self.compute.wait_event(self._halo_ready)
        _, oth = s.views()
        ev = torch.cuda.Event()
        w = max(self.edge, next_depth)
        if self.overlap and nxl >= 4 * w:
            s.step2_columns(0, w, row1, row2)
            s.step2_columns(nxl - w, nxl, row1, row2)
            ev.record(self.compute)
            s.step2_columns(w, nxl - w, row1, row2)
        else:
            s.step2_columns(0, nxl, row1, row2)
            ev.record(self.compute)
        self._exchange(oth, ev, next_depth)
        s.flip()
        self.updates += 2

    def updaten(self, rows, next_depth=None):
        """len(rows) = 2..4 lattice updates in one wavefront launch per column range."""
        torch = self.torch
        s, nxl = self.s, self.nxl
        d = len(rows)
        if next_depth is None:
            next_depth = d
        if self.world == 1:
            s.stepn_columns(0, nxl, rows)
            s.flip()
            self.updates += d
            return
        if self.peer:
            s.stepn_columns(0, nxl, rows)                   # the last stage stores the edge columns into the neighbours' halos
            self._group_done()
            self.updates += d
            return
        if self._halo_ready is not None:
            self.compute.wait_event(self._halo_ready)
        _, oth = s.views()
        ev = torch.cuda.Event()
        # Measured at 4 GPUs (32768^2, d = 4): one launch over the slab followed by the exchange 479 GLUPS;
        # edge launches of 8 columns + overlapped exchange 472; edge launches of 256 columns 454 -- the
        # exchange (~0.1 ms) costs less than what splitting the launch costs, so overlap_wave is off.
        # With it on: edge launches first (the neighbour needs my last d columns), the exchange overlaps the interior.
        # A launch pays 3(d-1) pipeline fill/drain steps per chunk whatever its width, so the edges are
        # whole chunks of 256 columns where the slab is wide enough (8 columns cost 20 sweep steps, 256
        # cost 268), and the exchange still has the long interior launch to hide behind.
        w = 256 if nxl >= 2048 else max(8, d, next_depth)
        if self.overlap and self.overlap_wave and nxl >= 4 * w:
            s.stepn_columns(0, w, rows)
            s.stepn_columns(nxl - w, nxl, rows)
            ev.record(self.compute)
            s.stepn_columns(w, nxl - w, rows)
        else:
            s.stepn_columns(0, nxl, rows)
            ev.record(self.compute)
        self._exchange(oth, ev, next_depth)
        s.flip()
        self.updates += d

    def advance(self, first_row, n, depth, row_stride=1, collect_forces=False):
        """n lattice updates starting with wall row first_row, at most `depth` per launch.  collect_forces: return
        the momentum-exchange sums [n, n_obs, 2] of the updates (whole domain), fetched group by group."""
        k = 0
        out = []
        if self.n_obs and not self.multi_ok:
            depth = 1                    # bodies on a slab interface: single updates (the link blocks ride along)
        elif self.n_obs and depth == 2:
            depth = 1                    # (two-update launches have no obstacle path)
        for d, nxt in launch_plan(n, depth):
            rows = [first_row + (k + j) * row_stride for j in range(d)]
            if d >= 3 or (d == 2 and depth > 2):
                self.updaten(rows, next_depth=nxt)
            elif d == 2:
                self.update2(rows[0], rows[1], next_depth=nxt)
            else:
                self.update(rows[0], next_depth=nxt)
            if collect_forces:
                out.append(self.forces(0, d))
            k += d
        if collect_forces:
            return np.concatenate(out) if out else np.zeros((0, max(self.n_obs, 1), 2))

    def probe_line(self, axis, index, row=0, out=None):
        """(rho, ux, uy) along this slab's part of a lattice line (Solver.probe_line), after the halos of the
        current array have landed (the cells on the slab edges pull from them)."""
        if self._halo_ready is not None:
            self.compute.wait_event(self._halo_ready)
        return self.s.probe_line(axis, index, row, out)

    def finish(self):
        if self._halo_ready is not None:
            self.compute.wait_event(self._halo_ready)
        self.s.sync()                 # the compute stream; in peer mode also reports a timed-out flag wait
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
