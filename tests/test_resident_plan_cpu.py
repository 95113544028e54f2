"""CPU: the hand-shake of resident batches (lbm_b200/csrc/resident.cuh) replayed with random block timing.

The library plans which block owns which columns and which blocks a block waits for (lbm_resident_plan: the same host
function build_resident uses).  Here every block is a little state machine -- wait until all its dependencies have
published k, read, write, publish k+1 -- and a random scheduler decides who moves next, with reads and writes as
separate events so that a neighbour can be caught in the middle of either.  The two population buffers are modelled per
column as (version, readers): update k reads version k-1 from buffer k % 2 and writes version k into buffer (k+1) % 2.

  * a read must find the version it expects (read before the neighbour wrote: too old; neighbour already two updates
    further: too new);
  * a write must not hit a column that some block has started to read and not finished (write-after-read).

What a block reads is stated here independently of the plan: a column block its own columns +- 1 (pull streaming) and,
where it holds a lattice wall column, +- 2 on that side (a corner cell takes rho, u from its x-neighbour on the wall, whose
pulled populations reach one column further, nb.py:254-257); a link group its columns +- 2 (interpolated bounce-back,
nb.py:98-104).  The test also shows that it can fail: without the two-column reach at the walls in the PLAN (the bug
the first version of the kernel had) the replay finds the hazard."""
import ctypes
import random

import numpy as np
import pytest

from lbm_b200 import _capi as C


def _plan(nx, max_blocks, groups):
    L = C.lib()
    ng = len(groups)
    g0 = np.array([g[0] for g in groups] or [0], dtype=np.int32)
    g1 = np.array([g[1] for g in groups] or [0], dtype=np.int32)
    ncb = ctypes.c_int32()
    col_a = np.zeros(max_blocks + 1, dtype=np.int32)
    dep_off = np.zeros(max_blocks + 1, dtype=np.int32)
    dep = np.zeros(max_blocks * max_blocks, dtype=np.int32)
    p = lambda a: C.c_vp(a.ctypes.data)
    C.check(L.lbm_resident_plan(nx, max_blocks, ng, p(g0), p(g1), ctypes.byref(ncb), p(col_a), p(dep_off), p(dep), dep.size))
    n = ncb.value
    nb = n + ng
    deps = [list(dep[dep_off[b]:dep_off[b + 1]]) for b in range(nb)]
    return n, list(col_a[:n + 1]), deps


def _sets(nx, n, col_a, groups):
    """(reads, writes) per block as sets of columns -- from the stencils, not from the plan's own reach."""
    out = []
    for i in range(n):
        w = set(range(col_a[i], col_a[i + 1]))
        r = {c for x in w for c in (x - 1, x, x + 1)}
        if 0 in w:
            r |= {1, 2}
        if nx - 1 in w:
            r |= {nx - 2, nx - 3}
        out.append(({c for c in r if 0 <= c < nx}, w))
    for a, b in groups:
        w = set(range(a, b + 1))
        out.append(({c for c in range(a - 2, b + 3) if 0 <= c < nx}, w))
    return out


def _replay(nx, deps, sets, n_updates, seed):
    rng = random.Random(seed)
    nb = len(deps)
    version = [[-1] * nx, [None] * nx]          # buffer 0 holds the initial state (version -1), buffer 1 nothing yet
    reading = [[0] * nx, [0] * nx]              # blocks in the middle of reading that column of that buffer
    prog = [0] * nb                             # published progress
    k = [0] * nb                                # update the block works on
    phase = ["wait"] * nb                       # wait -> read -> write -> (publish) wait
    while any(kk < n_updates for kk in k):
        b = rng.choice([i for i in range(nb) if k[i] < n_updates])
        src, dst = k[b] % 2, (k[b] + 1) % 2
        r, w = sets[b]
        if phase[b] == "wait":
            if all(prog[d] >= k[b] for d in deps[b]):
                phase[b] = "read"
                for c in r:
                    if version[src][c] != k[b] - 1:
                        return "block %d, update %d: column %d holds version %s, expected %d" % (b, k[b], c, version[src][c], k[b] - 1)
                    reading[src][c] += 1
        elif phase[b] == "read":                # (the loads have returned: the results of a cell depend on all of them)
            for c in r:
                reading[src][c] -= 1
            phase[b] = "write"
        else:
            for c in w:
                if reading[dst][c]:
                    return "block %d, update %d: writes column %d of the buffer another block is reading" % (b, k[b], c)
                version[dst][c] = k[b]
            prog[b] = k[b] + 1
            k[b] += 1
            phase[b] = "wait"
    return None


CASES = [(200, 296, []), (200, 148, []), (17, 444, []), (3, 10, []), (536, 443, [(67, 91)]), (40, 12, [(10, 14), (15, 15), (30, 36)]),
         (1073, 442, [(140, 168), (169, 180)]), (64, 70, [(0, 3), (60, 63)])]


@pytest.mark.parametrize("nx,max_blocks,groups", CASES)
def test_hand_shake_is_hazard_free_under_random_timing(nx, max_blocks, groups):
    n, col_a, deps = _plan(nx, max_blocks, groups)
    assert col_a[0] == 0 and col_a[n] == nx and all(col_a[i] < col_a[i + 1] for i in range(n))
    assert all(b in deps[d] for b in range(len(deps)) for d in deps[b])          # the relation is symmetric
    sets = _sets(nx, n, col_a, groups)
    for seed in range(6):
        assert _replay(nx, deps, sets, 12, seed) is None


def test_replay_finds_the_missing_corner_dependency():
    """One column per block and the wall blocks waiting for their direct neighbours only -- what the first version of the
    plan did: the left corner cell reads column 2 while its owner may already be overwriting it."""
    nx = 24
    n, col_a, deps = _plan(nx, 100, [])
    assert n == nx and 2 in deps[0] and nx - 3 in deps[nx - 1]
    broken = [[d for d in ds if abs(d - b) <= 1] for b, ds in enumerate(deps)]
    sets = _sets(nx, n, col_a, [])
    assert any(_replay(nx, broken, sets, 12, seed) is not None for seed in range(40))
