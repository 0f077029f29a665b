"""Probe (GPU box): whole-solve differences between the product and the oracle on cfg3 / cfg2 / tiny instances."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT
from oracle import oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
make = getattr(lqc, sys.argv[2]) if len(sys.argv) > 2 else lqc.cfg3
Ps = [make(i) for i in range(n)]
k = BatchKKT(Ps[0], batch=n)
perm, _, _ = k.symbolic()
k.load_lq(Ps)
k.initialize(np.stack([P.x0 for P in Ps]))
k.lq_begin()
r = k.lq_solve(max_steps=400, check_every=400)
W, lam, st, sc = k.get("POINT"), k.get("DUAL"), k.stats(), k.scalars()
print(r)
worst = 0
for i, P in enumerate(Ps):
    t = time.time()
    o = orc.from_problem(P, perm=perm)
    o.use_superlu_fallback()
    o.initialize(P.x0)
    rc = o.solve()
    e = np.abs(W[i] - o.solution).max() / np.abs(o.solution).max()
    ex = np.abs(W[i][:P.n] - o.solution[:P.n]).max() / np.abs(o.solution[:P.n]).max()
    worst = max(worst, e)
    print(f"seed {i}: rc {rc} its gpu/oracle {st['total_iterations'][i]}/{o.stats['total_iterations']} outer {st['outer'][i]}/{o.stats['outer']} "
          f"fallbacks {st['fallbacks'][i]}/{o.stats['lu_fallbacks']} unrefined {st['unrefined_steps'][i]} rel err all {e:.2e} primal {ex:.2e} ({time.time()-t:.2f}s)")
print("worst rel err", worst)
