import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lbm_b200 import cases
from tools.resident_bench import run
for name, mk in (("cavity", lambda: cases.Cavity(L_lbm=200)), ("turek100", lambda: cases.Turek(L_lbm=100)), ("array", lambda: cases.Array())):
    for fl in (8, 8 + 6):
        print(name, "flags", fl, file=sys.stderr, flush=True)
        run(mk(), 1, 2, 512, reps=1, flags=fl)
