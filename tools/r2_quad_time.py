"""Throughput of complete solves on the quadruped-shaped instances: python tools/r2_quad_time.py [BATCH]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT
B = int(sys.argv[1]) if len(sys.argv) > 1 else 444
Ps = [lqc.quadruped_shape(i) for i in range(8)]
pl = [Ps[i % 8] for i in range(B)]
k = BatchKKT(Ps[0], batch=B)
print(k.paths(), k.info()["nnzL"])
k.load_lq(pl); X0 = np.stack([P.x0 for P in pl])
for rep in range(3):
    k.initialize(X0); k.lq_begin(); k.synchronize(); t = time.perf_counter()
    r = k.lq_solve(max_steps=400, check_every=400); k.synchronize(); dt = time.perf_counter() - t
its = int((k.stats()["total_iterations"] - 1).sum())
print(f"CB200_PLAN={os.environ.get('CB200_PLAN')} batch {B}: {dt*1e3:.1f} ms, {its/dt:.0f} it/s, {r}")
