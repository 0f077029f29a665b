"""A/B of two builds of libcalipso_b200 on a B200: results (bitwise or to a tolerance) and timings of the KKT unit and of a
complete batched solve.

  python tools/r2_ab.py [--a tools/variants/libcalipso_b200_r1.so] [--b calipso_b200/libcalipso_b200.so] [--batch 444]

Developer tool (not part of the product or of the tests)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from calipso_b200 import _lib, lqc
from calipso_b200.solver import BatchKKT

ap = argparse.ArgumentParser()
ap.add_argument("--a", default="tools/variants/libcalipso_b200_r1.so")
ap.add_argument("--b", default="calipso_b200/libcalipso_b200.so")
ap.add_argument("--batch", type=int, default=444)
ap.add_argument("--distinct", type=int, default=16)
ap.add_argument("--solve-batch", type=int, default=1332)
ap.add_argument("--no-solve", action="store_true")
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()

Ps = [lqc.cfg3(i) for i in range(args.distinct)]


class LooseBinding(_lib.Binding):
    """binding of an older build: symbols it does not export yet are skipped"""

    def __init__(self, path):
        import ctypes as C
        self.path = path
        self.lib = C.CDLL(path)
        for name, (res, argtypes) in _lib.SYMBOLS.items():
            try:
                fn = getattr(self.lib, name)
            except AttributeError:
                continue
            fn.restype = res
            fn.argtypes = argtypes


def make(path, B):
    k = BatchKKT(Ps[0], batch=B, binding=LooseBinding(path))
    plist = [Ps[i % len(Ps)] for i in range(B)]
    k.load_lq(plist)
    k.X0 = np.stack([P.x0 for P in plist])
    k.initialize(k.X0)
    return k


def timeit(fn, k, reps=5):
    fn(); fn()
    k.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    k.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


res = {}
for tag, path in (("A", args.a), ("B", args.b)):
    if not os.path.exists(path):
        print(f"{tag}: {path} missing, skipped")
        continue
    k = make(path, args.batch)
    k.lq_begin()
    k.lq_evaluate(2 | 16 | 32)
    k.cone(barrier=True, barrier_gradient=True, product=True)
    k.residual()
    k.set_scalars(eps_p=1e-7, eps_d=1e-7)
    k.kkt_factor_solve(1)
    step = k.get("STEP")
    D = k.get("PIVOTS")
    st = k.stats()
    t0, t1, t5 = (timeit(lambda ns=ns: k.kkt_factor_solve(ns), k) for ns in (0, 1, 5))
    print(f"{tag} {path}: batch {args.batch}: factor {t0:.3f} ms, factor+1 solve {t1:.3f} ms, one solve {(t5 - t0) / 5:.3f} ms "
          f"(per instance: {1e3 * t0 / args.batch:.2f} / {1e3 * t1 / args.batch:.2f} / {1e3 * (t5 - t0) / 5 / args.batch:.2f} us)", flush=True)
    if args.profile:
        k.profile()
        for _ in range(3):
            k.kkt_factor_solve(1)
        prof = k.profile()
        print("   phase us/instance:", {kk: round(v / args.batch / 3 / 1.9e3, 1) for kk, v in prof.items() if v}, flush=True)
    # a full search direction from the same state (refinement, J v)
    k.search_direction()
    sd = k.get("STEP")
    st2 = k.stats()
    tsd = timeit(lambda: k.search_direction(), k, 3)
    print(f"   search_direction {tsd:.3f} ms per launch, refine passes mean {st2['n_refine'].mean():.2f}", flush=True)
    res[tag] = dict(step=step, D=D, sd=sd, refine=st2["n_refine"].copy(), inertia=st["inertia_pos"].copy())
    k.close()

if "A" in res and "B" in res:
    a, b = res["A"], res["B"]
    for name in ("D", "step", "sd"):
        same = np.array_equal(a[name], b[name])
        rel = np.abs(a[name] - b[name]).max() / np.abs(a[name]).max()
        print(f"compare {name}: bitwise identical {same}, max rel diff {rel:.3e}")
    print("refine counts equal:", np.array_equal(a["refine"], b["refine"]))

if not args.no_solve:
    out = {}
    for tag, path in (("A", args.a), ("B", args.b)):
        if not os.path.exists(path):
            continue
        B = args.solve_batch
        k = make(path, B)
        ts = []
        for rep in range(3):
            k.initialize(k.X0); k.lq_begin(); k.synchronize()
            t = time.perf_counter()
            r = k.lq_solve(max_steps=400, check_every=400)
            k.synchronize()
            ts.append(time.perf_counter() - t)
        st = k.stats()
        its = int((st["total_iterations"] - 1).sum())
        out[tag] = (k.get("POINT"), st["total_iterations"].copy())
        print(f"{tag}: lq_solve batch {B}: {min(ts) * 1e3:.1f} ms, {its / min(ts):.0f} it/s, iterations {its}, {r}", flush=True)
        k.close()
    if len(out) == 2:
        print("solutions bitwise identical:", np.array_equal(out["A"][0], out["B"][0]), "iterations equal:",
              np.array_equal(out["A"][1], out["B"][1]),
              "max rel diff", np.abs(out["A"][0] - out["B"][0]).max() / np.abs(out["A"][0]).max())
