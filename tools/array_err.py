import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lbm_b200 import cases
from lbm_b200.lattice import lattice
from lbm_b200.run import run
from oracle import oracle as orc
def rel(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
for arith in ("fused", "strict"):
    for n in (500, 1000, 1500, 2000, 2500):
        cg, co = cases.Array(), cases.Array()
        cg.it_max = co.it_max = n - 1
        lg = lattice(cg, make_dirs=False, arith=arith); run(lg, cg, batch=1024, quiet=True)
        lo = orc.OracleLattice(co); orc.run_loop(lo, co)
        print(arith, n, {k: "%.1e" % rel(getattr(lg, k), getattr(lo, k)) for k in ("g_up", "rho", "u")}, flush=True)
