"""Is the slowdown of a CTA under load SM-local or chip-wide?  k_kkt_factor_solve (factor / factor + 5 solves) and a J v heavy
search direction at three resident 256-thread CTAs per SM on (a) every SM, (b) half of the SMs while the other half is held by
spinning blocker CTAs (tools/microbench/sm_blocker.cu), (c) one CTA per SM on every SM.

  python tools/r2_load_probe.py          Developer tool."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

os.environ["CB200_THREADS"] = "256"            # the 256-thread instantiations at every batch size
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT

blk = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "microbench", "sm_blocker.so"))
Ps = [lqc.cfg3(i) for i in range(16)]


def measure(B, blocked):
    pl = [Ps[i % 16] for i in range(B)]
    k = BatchKKT(Ps[0], batch=B)
    k.load_lq(pl); k.initialize(np.stack([P.x0 for P in pl])); k.lq_begin(); k.lq_step(4)
    k.set_scalars(eps_p=1e-7, eps_d=1e-7)
    k.synchronize()
    seen = blk.blocker_start(blocked) if blocked else 0

    def timeit(fn, reps=4):
        fn(); k.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        k.synchronize()
        return (time.perf_counter() - t) / reps * 1e3
    t0 = timeit(lambda: k.kkt_factor_solve(0))
    t5 = timeit(lambda: k.kkt_factor_solve(5))
    tsd = timeit(lambda: k.search_direction())
    if blocked:
        assert blk.blocker_release() == 0
    k.close()
    print(f"batch {B:4d}, SMs held by blockers {seen:3d}: factor {t0:.3f} ms, one reduced solve {(t5 - t0) / 5:.3f} ms, "
          f"search direction {tsd:.3f} ms", flush=True)


measure(444, 0)        # three CTAs on each of 148 SMs
measure(222, 74)       # three CTAs on each of 74 SMs, the other 74 idle (spinning blockers)
measure(148, 0)        # one CTA per SM (the block scheduler spreads them)
measure(74, 74)        # one CTA per SM on 74 SMs
