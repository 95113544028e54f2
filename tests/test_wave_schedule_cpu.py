"""CPU: the ring schedule of stepw_kernel (lbm_b200/csrc/kernels.cuh), replayed symbolically.

The wavefront kernel keeps every time level in a ring of a few columns and lets all stages run in
the same sweep step, ordered only by one block barrier per step.  This test restates the schedule
(which stage touches which column of which ring in which step, what the producer warp loads when)
and checks, for chunks with and without walls, that
  * every column a stage reads is in its slot, was completed in an earlier step and is not being
    overwritten in the same step (the barrier is the only ordering there is),
  * the deferred left corner and the right corner find their x-neighbour's columns,
  * the last stage produces every output column of the chunk exactly once from valid inputs.
The constants are read from the source so that the model cannot drift from the kernel silently."""
import os
import re

import pytest

SRC = open(os.path.join(os.path.dirname(__file__), "..", "lbm_b200", "csrc", "kernels.cuh")).read()
HOST = open(os.path.join(os.path.dirname(__file__), "..", "lbm_b200", "csrc", "lbm_b200.cu")).read()
R = int(re.search(r"static constexpr int R = (\d+);", SRC).group(1))
LAG = int(re.search(r"static constexpr int LAG = (\d+);", SRC).group(1))
R0 = int(re.search(r"constexpr int kWaveR0 = (\d+);", HOST).group(1))
BIG = 1 << 20


def replay(D, ca, cb, nxl, has_left, has_right):
    x_lo, x_hi = (0 if has_left else -BIG), (nxl if has_right else nxl + BIG)
    x_wl, x_wr = (0 if has_left else -(1 << 30)), (nxl - 1 if has_right else -(1 << 30))
    xs0 = ca - (D - 1)
    c_first, c_last = xs0 - 2, cb + D - 1
    nsteps = (cb - ca) + 2 * (D - 1) + LAG * (D - 1)
    exists = lambda c: x_lo <= c < x_hi
    # ring contents: holds[level][slot] = (column, step at which it was completed) ; valid[level] = columns with valid data
    holds = [dict() for _ in range(D)]
    valid = [set() for _ in range(D + 1)]
    for c in range(c_first, min(c_first + R0, c_last + 1)):          # prologue of the producer warp
        holds[0][c % R0] = (c, -1)
        valid[0].add(c)
    produced = []
    for s in range(nsteps):
        writes = {}                                                  # (level, slot) written during this step
        if s >= 1 and xs0 + s + R0 - 3 <= c_last:                    # producer: one column ahead of the ring window
            c = xs0 + s + R0 - 3
            writes[(0, c % R0)] = c
        plan = []
        for k in range(D):
            x = xs0 + s - LAG * k
            lo, hi = max(ca - (D - 1 - k), x_lo), min(cb + (D - 1 - k), x_hi)
            if not (lo <= x < hi):
                continue
            cells = [x]                                              # columns of the cells computed in this step
            corner_cells = []
            if x == x_wl:
                corner_cells = []                                    # corner rows skip, other rows compute x
            if x == x_wl + 1 and x - 1 >= lo:
                corner_cells = [x - 1]                               # deferred left corner (edge rows only)
            plan.append((k, x, cells, corner_cells))
            if k < D - 1:
                for c in cells + corner_cells:
                    assert (k + 1, c % R) not in writes or writes[(k + 1, c % R)] == c
                    writes[(k + 1, c % R)] = c
        for k, x, cells, corner_cells in plan:
            mod = R0 if k == 0 else R
            need = set()
            for c in cells:
                need |= {c - 1, c, c + 1}
                if c == x_wr:
                    need |= {c - 2}                                  # right corner: pulled populations of x-1
            for c in corner_cells:                                   # left corner: its own pull and that of x+1
                need |= {c - 1, c, c + 1, c + 2}
            for c in need:
                if not exists(c):
                    continue                                         # overwritten by Zou-He, content irrelevant
                got = holds[k].get(c % mod)
                assert got is not None and got[0] == c, (D, s, k, x, c, got)
                assert got[1] < s, "read of a column completed in the same step"
                assert (k, c % mod) not in writes, (D, s, k, x, c, "slot overwritten while it is read")
                assert c in valid[k], (D, s, k, x, c, "input not valid")
            for c in cells + corner_cells:
                valid[k + 1].add(c)
                if k == D - 1 and c in cells:
                    produced.append(c)
        for (lvl, slot), c in writes.items():                        # the barrier: writes become visible
            holds[lvl][slot] = (c, s)
            if lvl == 0:
                valid[0].add(c)
    assert sorted(produced) == list(range(ca, cb)), (D, ca, cb)


@pytest.mark.parametrize("D", [2, 3, 4])
def test_ring_schedule_is_hazard_free(D):
    assert (R, LAG) == (4, 2) and R0 >= 8 and R0 & (R0 - 1) == 0
    nxl = 300
    for (ca, cb) in ((0, 16), (0, 2), (16, 48), (100, 117), (284, 300), (298, 300), (0, 300)):
        for has_left in (True, False):
            for has_right in (True, False):
                replay(D, ca, cb, nxl, has_left, has_right)
