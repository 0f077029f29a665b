"""Small driver for ncu captures: a few launches of the KKT factor+solve kernel and of the on-device Newton step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
nsolves = int(sys.argv[2]) if len(sys.argv) > 2 else 1
Ps = [lqc.cfg3(i) for i in range(4)]
k = BatchKKT(Ps[0], batch=B)
k.load_lq([Ps[i % 4] for i in range(B)])
k.initialize(np.stack([Ps[i % 4].x0 for i in range(B)]))
k.lq_begin()
k.lq_step(5)
k.set_scalars(eps_p=1e-7, eps_d=1e-7)
for _ in range(3):
    k.kkt_factor_solve(nsolves)
k.synchronize()
