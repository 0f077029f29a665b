"""Time k_kkt_factor_solve only (no convergence logic): python tools/kkt_time.py BATCH NSOLVES"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT
B = int(sys.argv[1]) if len(sys.argv) > 1 else 444
Ps = [lqc.cfg3(i) for i in range(4)]
k = BatchKKT(Ps[0], batch=B)
k.load_lq([Ps[i % 4] for i in range(B)])
k.initialize(np.stack([Ps[i % 4].x0 for i in range(B)]))
k.lq_begin()
k.set_scalars(eps_p=1e-7, eps_d=1e-7)
k.lq_evaluate(2 | 16 | 32); k.cone(barrier=True, barrier_gradient=True, product=True); k.residual()
for ns in [int(a) for a in sys.argv[2:]] or [0, 1, 5]:
    for _ in range(2): k.kkt_factor_solve(ns)
    k.synchronize(); t = time.time()
    for _ in range(5): k.kkt_factor_solve(ns)
    k.synchronize(); print(f"nsolves={ns}: {(time.time() - t) / 5 * 1e3:.3f} ms per launch")
