"""Drain tail of a batched solve!: the bench's per-rank instance block of a given rank (rank 2 of the 8-GPU run holds an
18-iteration instance) solved with different check intervals (after a check the survivors run on 512-thread CTAs).

  python tools/r2_tail_ab.py [--rank 2] [--checks 400,16,12,10,8]

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
ap.add_argument("--checks", default="400,16,12,10,8")
args = ap.parse_args()

B, D = args.batch, args.distinct
Ps = [lqc.cfg3(args.rank * D + i) for i in range(D)]
pl = [Ps[i % D] for i in range(B)]
k = BatchKKT(Ps[0], batch=B)
k.load_lq(pl)
X0 = np.stack([P.x0 for P in pl])
ref = None
for ce in [int(c) for c in args.checks.split(",")]:
    def solve():
        k.initialize(X0)
        k.lq_begin()
        return k.lq_solve(max_steps=400, check_every=ce)
    solve()
    k.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        r = solve()
    k.synchronize()
    ms = (time.perf_counter() - t) / 3 * 1e3
    st = k.stats()
    it = st["total_iterations"] - 1
    w = k.get("POINT")
    if ref is None:
        ref = (it.copy(), w.copy())
    dw = float(np.max(np.abs(w - ref[1]) / (1.0 + np.abs(ref[1]))))
    print(f"check_every {ce:4d}: {ms:8.2f} ms per solve  steps {r['steps']}  iterations sum {int(it.sum())} max {int(it.max())}  "
          f"iteration counts equal to first: {bool((it == ref[0]).all())}  max rel point diff {dw:.2e}", flush=True)
