"""Per-iteration time of the instances that are still running late in a batched solve! (lock-step launches of one Newton
iteration each; the last launches hold only the hardest instances), narrow (256-thread) and wide (512-thread) CTAs.

  python tools/r2_tail_probe.py [--rank 2]

Developer tool (not part of the product or of the tests)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT

ap = argparse.ArgumentParser()
ap.add_argument("--rank", type=int, default=2)
ap.add_argument("--batch", type=int, default=1332)
ap.add_argument("--distinct", type=int, default=166)
args = ap.parse_args()
B, D = args.batch, args.distinct
Ps = [lqc.cfg3(args.rank * D + i) for i in range(D)]
pl = [Ps[i % D] for i in range(B)]
X0 = np.stack([P.x0 for P in pl])
for threads in ("256", "512"):
    os.environ["CB200_THREADS"] = threads
    k = BatchKKT(Ps[0], batch=B)
    k.load_lq(pl)
    k.initialize(X0)
    k.lq_begin()
    k.synchronize()
    prev = k.stats()
    print(f"--- CB200_THREADS={threads}", flush=True)
    for it in range(1, 26):
        t = time.perf_counter()
        k.lq_step(1)
        k.synchronize()
        ms = (time.perf_counter() - t) * 1e3
        st = k.stats()
        running = int((st["converged"] == 0).sum())
        act = np.nonzero(st["total_iterations"] != prev["total_iterations"])[0]
        d = {kk: (st[kk][act] - prev[kk][act]) for kk in ("solves", "factorizations", "fallbacks")}
        print(f"iteration {it:2d}: {ms:7.2f} ms  instances that iterated {len(act):5d}  still running {running:5d}  per instance: "
              f"solves mean {d['solves'].mean() if len(act) else 0:.1f} max {d['solves'].max() if len(act) else 0}  "
              f"factorizations max {d['factorizations'].max() if len(act) else 0}  fallbacks max {d['fallbacks'].max() if len(act) else 0}",
              flush=True)
        prev = st
        if running == 0:
            break
    k.close()
