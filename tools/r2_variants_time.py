"""Factor / solve / J v timings and phase counters of several builds (ablation variants under tools/variants/).

  python tools/r2_variants_time.py a.so b.so ...        Developer tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C

import numpy as np
import torch

from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT
from tools.r2_ab_common import LooseBinding

B = int(os.environ.get("BATCH", "444"))
Ps = [lqc.cfg3(i) for i in range(16)]
pl = [Ps[i % 16] for i in range(B)]
X0 = np.stack([P.x0 for P in pl])


def timeit(fn, k, reps=5):
    fn(); fn()
    k.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    k.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


for spec in sys.argv[1:]:
    path, _, envs = spec.partition(":")            # path[:NAME=VALUE,NAME=VALUE] -- environment for this build's handles
    for kv in [e for e in os.environ if e.startswith("CB200_PLAN")]:
        del os.environ[kv]
    for kv in filter(None, envs.split(",")):
        os.environ[kv.split("=")[0]] = kv.split("=")[1]
    k = BatchKKT(Ps[0], batch=B, binding=LooseBinding(path))
    k.load_lq(pl); k.initialize(X0); k.lq_begin()
    k.lq_step(4)
    k.set_scalars(eps_p=1e-7, eps_d=1e-7)
    t0, t1, t5 = (timeit(lambda ns=ns: k.kkt_factor_solve(ns), k) for ns in (0, 1, 5))
    C.cast(k.lib.cb200_jacobian_times, C.c_void_p)       # (exists in every build)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    stream = torch.cuda.ExternalStream(k.lib.cb200_stream(k.h))
    V = np.ones((B, k.total)); O = np.zeros((B, k.total))
    def ev_time(fn, reps=5):
        fn(); k.synchronize()
        ts = []
        for _ in range(reps):
            ev[0].record(stream); fn(); ev[1].record(stream); ev[1].synchronize(); ts.append(ev[0].elapsed_time(ev[1]))
        return min(ts)
    t_sd = ev_time(lambda: k.search_direction())
    t_ev = ev_time(lambda: k.lq_evaluate(2 | 16 | 32))
    k.profile()
    for _ in range(3):
        k.kkt_factor_solve(1)
    prof = k.profile()
    ph = {kk: round(v / B / 3 / 1.9e3, 1) for kk, v in prof.items() if v}
    print(f"{spec}: paths {k.paths()}")
    print(f"{path}: factor {t0:.3f} ms  one solve {(t5 - t0) / 5:.3f} ms  search_direction {t_sd:.3f} ms  lq_evaluate {t_ev:.3f} ms\n   phases us: {ph}", flush=True)
    k.close()
    if "abl" in os.path.basename(path):
        continue            # ablated builds give wrong results: no complete solves
    SB = 1332
    spl = [Ps[i % 16] for i in range(SB)]
    SX0 = np.stack([P.x0 for P in spl])
    k = BatchKKT(Ps[0], batch=SB, binding=LooseBinding(path))
    k.load_lq(spl)
    ts = []
    for rep in range(3):
        k.initialize(SX0); k.lq_begin(); k.synchronize()
        t = time.perf_counter()
        r = k.lq_solve(max_steps=60, check_every=60)
        k.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    its = int((k.stats()["total_iterations"] - 1).sum())
    print(f"   lq_solve(batch {SB}) {min(ts):.1f} ms  iterations {its}  {its / min(ts) * 1e3:.0f} it/s  {r}", flush=True)
    k.close()
