"""Phase counters (thread-0 clock64 accumulators) over a complete batched solve!: microseconds of CTA time per Newton iteration
by phase, in the real mix of phases (not the lock-step microbenchmark).  python tools/r2_solve_phases.py [BATCH]"""
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
k.profile()
r = k.lq_solve(max_steps=400, check_every=400)
prof = k.profile()
st = k.stats()
its = int((st["total_iterations"] - 1).sum())
print(r, "iterations", its, "factorizations", int(st["factorizations"].sum()), "solves", int(st["solves"].sum()))
print({kk: round(v / its / 1.965e3, 1) for kk, v in prof.items() if v})
