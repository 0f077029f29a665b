"""The heavy kernels on a batch of BASELINE cfg2 instances (T=50, n_x=12, n_u=6, box + 20 SOC(3) cones) for an ncu capture:
python tools/ncu_cfg2_kernels.py [BATCH]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 444
Ps = [lqc.cfg2(i) for i in range(8)]
plist = [Ps[i % 8] for i in range(B)]
k = BatchKKT(Ps[0], batch=B)
print("cfg2:", k.info(), k.paths())
k.load_lq(plist)
k.initialize(np.stack([P.x0 for P in plist]))
k.lq_begin()
k.lq_step(3)                                   # k_lq_step (three Newton iterations per instance)
k.lq_evaluate(2 | 16 | 32)
k.cone(barrier=True, barrier_gradient=True, product=True)
k.residual()
k.search_direction()                           # k_search_direction
k.set_scalars(eps_p=1e-7, eps_d=1e-7)
k.kkt_factor_solve(1)                          # the KKT-solve unit
k.kkt_factor_solve(0)                          # factorisation only
k.synchronize()
r = k.lq_solve(max_steps=200, check_every=200)
print(r, "iterations", int((k.stats()["total_iterations"] - 1).sum()))
