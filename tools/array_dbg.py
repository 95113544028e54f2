import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lbm_b200 import cases
from lbm_b200.lattice import lattice
from lbm_b200.run import run
from oracle import oracle as orc
n = 300
cg, co = cases.Array(), cases.Array()
cg.it_max = co.it_max = n - 1
lg = lattice(cg, make_dirs=False); run(lg, cg, batch=1024, quiet=True)
lo = orc.OracleLattice(co); orc.run_loop(lo, co)
d = np.abs(lg.u - lo.u)
k = np.unravel_index(np.argmax(d), d.shape)
print("max u err", d.max(), "at", k, "gpu", lg.u[k], "ref", lo.u[k], "rho", lg.rho[k[1], k[2]], lo.rho[k[1], k[2]])
print("cells with err > 1e-13:", np.argwhere(d > 1e-13)[:20].tolist(), (d > 1e-13).sum())
print("is solid?", [int(any((o.boundary[:, 0] == k[1]) & (o.boundary[:, 1] == k[2])) ) for o in cg.obstacles])
