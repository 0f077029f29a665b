"""Experiment: effect of the GMRES restart length of the refinement-failure fallback on the number of reduced solves."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT, Options
B = 148
Ps = [lqc.cfg3(i) for i in range(16)]
for restart, cycles in ((30, 10), (60, 5), (100, 3), (150, 2)):
    k = BatchKKT(Ps[0], batch=B, options=Options(gmres_restart=restart, gmres_max_cycles=cycles))
    k.load_lq([Ps[i % 16] for i in range(B)]); X0 = np.stack([Ps[i % 16].x0 for i in range(B)])
    def solve():
        k.initialize(X0); k.lq_begin(); return k.lq_solve(max_steps=400, check_every=4)
    solve(); k.synchronize()
    t = time.time(); r = solve(); k.synchronize(); dt = time.time() - t
    st = k.stats()
    its = int((st['total_iterations'] - 1).sum())
    print(f"restart {restart:4d} x {cycles}: {r} it {its} time {dt*1e3:.1f} ms it/s {its/dt:.0f} solves/inst {st['solves'].sum()/B:.1f} fallbacks {int(st['fallbacks'].sum())} gmres_iters(last) {int(st['gmres_iters'].sum())}")
    k.close()
