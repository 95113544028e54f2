import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_gpu_parity import make_case, gpu_lattice, rel
from oracle import oracle as orc
for name, n in (("cavity32", 150), ("turek30", 120), ("poiseuille20", 100), ("cavity200", 2000), ("turek100", 2000)):
    cg, co = make_case(name), make_case(name)
    lg = gpu_lattice(cg, dtype="f32"); lo = orc.OracleLattice(co)
    orc.run_loop(lg, cg, n_iters=n); orc.run_loop(lo, co, n_iters=n)
    print(name, n, {k: "%.2e" % rel(getattr(lg, k), getattr(lo, k)) for k in ("g", "g_up", "rho", "u")},
          ("dF %.2e" % np.max(np.abs(np.array(cg.forces) - np.array(co.forces)))) if cg.forces else "", flush=True)
