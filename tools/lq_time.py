"""Time complete batched solves (cb200_lq_solve) for several convergence-check intervals, with the Newton iterations of
a check interval inside one launch (default) or one launch per iteration (CB200_LQ_LOCKSTEP=1):
python tools/lq_time.py [BATCH] [DISTINCT]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1332
distinct = int(sys.argv[2]) if len(sys.argv) > 2 else 64
Ps = [lqc.cfg3(i) for i in range(min(B, distinct))]
plist = [Ps[i % len(Ps)] for i in range(B)]
k = BatchKKT(Ps[0], batch=B)
k.load_lq(plist)
X0 = np.stack([P.x0 for P in plist])
ref = None
configs = [tuple(int(v) for v in c.split(":")) for c in os.environ.get("LQ_CONFIGS", "1:4,0:4,0:8,0:400").split(",")]
for lockstep, ce in configs:
    os.environ["CB200_LQ_LOCKSTEP"] = str(lockstep)
    ts = []
    for rep in range(3):
        k.initialize(X0); k.lq_begin(); k.synchronize()
        t = time.time()
        r = k.lq_solve(max_steps=400, check_every=ce)
        k.synchronize(); ts.append(time.time() - t)
    st = k.stats()
    its = int((st["total_iterations"] - 1).sum())
    w = k.get("POINT")
    if ref is None:
        ref = (its, w.copy())
    same = its == ref[0] and np.array_equal(w, ref[1])
    print(f"lockstep={lockstep} check_every={ce}: {min(ts[1:]) * 1e3:.1f} ms per solve, {its / min(ts[1:]):.0f} it/s, "
          f"iterations {its}, converged {r['converged']}, steps {r['steps']}, identical to first config: {same}", flush=True)
