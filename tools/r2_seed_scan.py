"""Which cfg3 seeds does solve! not converge on (GPU scan)?  python tools/r2_seed_scan.py FIRST COUNT"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT
first, count = int(sys.argv[1]), int(sys.argv[2])
Ps = [lqc.cfg3(first + i) for i in range(count)]
k = BatchKKT(Ps[0], batch=count)
k.load_lq(Ps); k.initialize(np.stack([P.x0 for P in Ps])); k.lq_begin()
r = k.lq_solve(max_steps=150, check_every=150)
st = k.stats()
bad = np.nonzero(st["converged"] != 1)[0]
print(r, "not converged:", [(first + int(i), int(st["converged"][i]), int(st["status"][i]), int(st["total_iterations"][i]), int(st["outer"][i])) for i in bad])
its = st["total_iterations"] - 1
print("iterations: mean", its.mean(), "max", its.max(), "hist", np.bincount(np.minimum(its, 30)).tolist())
