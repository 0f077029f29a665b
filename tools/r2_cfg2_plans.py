"""cfg2 (small instances): complete batched solves with three 256-thread CTAs per SM against six 128-thread CTAs per SM
(build with -DCB_THREADS_NARROW=128 -DCB_NARROW_CTAS=6, CB200_PLAN=7).  python tools/r2_cfg2_plans.py LIB     Developer tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT
from tools.r2_ab_common import LooseBinding

lib = sys.argv[1]
B = 2664
Ps = [lqc.cfg2(i) for i in range(64)]
pl = [Ps[i % 64] for i in range(B)]
X0 = np.stack([P.x0 for P in pl])
for env in ({}, {"CB200_PLAN": "7", "CB200_PLAN_ONLY": "1"}):
    for kk in ("CB200_PLAN", "CB200_PLAN_ONLY"):
        os.environ.pop(kk, None)
    os.environ.update(env)
    k = BatchKKT(Ps[0], batch=B, binding=LooseBinding(lib))
    k.load_lq(pl)
    ts = []
    for rep in range(3):
        k.initialize(X0); k.lq_begin(); k.synchronize()
        t = time.perf_counter()
        r = k.lq_solve(max_steps=200, check_every=200)
        k.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    its = int((k.stats()["total_iterations"] - 1).sum())
    print(env, k.paths(), f"lq_solve batch {B}: {min(ts):.1f} ms, {its / min(ts) * 1e3:.0f} it/s, {r}", flush=True)
    k.close()
