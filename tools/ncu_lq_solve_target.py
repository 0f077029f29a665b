"""ncu target: ONE k_lq_step launch that carries a complete solve! of every instance (444 cfg3 instances, 16 distinct seeds):
ncu --set full --clock-control none -k regex:k_lq_step -c 1 python tools/ncu_lq_solve_target.py [BATCH]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT
B = int(sys.argv[1]) if len(sys.argv) > 1 else 444
Ps = [lqc.cfg3(i) for i in range(16)]
pl = [Ps[i % 16] for i in range(B)]
k = BatchKKT(Ps[0], batch=B)
k.load_lq(pl); k.initialize(np.stack([P.x0 for P in pl])); k.lq_begin()
r = k.lq_solve(max_steps=400, check_every=400)      # <- the captured launch
st = k.stats()
print("result", r, "newton iterations", int((st["total_iterations"] - 1).sum()), "factorizations", int(st["factorizations"].sum()),
      "reduced solves", int(st["solves"].sum()))
