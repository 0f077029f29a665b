"""Time cb200_scatter (k_scatter) on the cfg3 patterns with device-resident caches: python tools/scatter_time.py [BATCH]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT, A

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1332
P = lqc.cfg3(0)
k = BatchKKT(P, batch=B)


def keys(colptr, rowval):
    cols = np.repeat(np.arange(len(colptr) - 1), np.diff(colptr))
    return np.stack([np.asarray(rowval), cols], axis=1)


wk, gk, ck = keys(P.W_colptr, P.W_rowval), keys(P.G_colptr, P.G_rowval), keys(P.C_colptr, P.C_rowval)
# W: objective cache = every key, equality-dual cache = every key again in reverse order, cone-dual cache = the diagonal
diag = wk[wk[:, 0] == wk[:, 1]]
plans = {"W_VALUES": [wk, wk[::-1], diag], "G_VALUES": [gk], "C_VALUES": [ck]}
rng = np.random.default_rng(0)
for name, sp in plans.items():
    k.scatter_plan(name, sp)
    L = sum(len(s) for s in sp)
    nnz = k.length(name)
    k.scatter(name, rng.standard_normal((B, L)))              # H2D once; afterwards the caches stay on the device
    for _ in range(3):
        k.b.check(k.lib.cb200_scatter(k.h, A[name], None, 0, B))
    k.synchronize(); t = time.time()
    reps = 20
    for _ in range(reps):
        k.b.check(k.lib.cb200_scatter(k.h, A[name], None, 0, B))
    k.synchronize(); dt = (time.time() - t) / reps
    alg = B * (8 * (L + nnz) + 4 * len(sp) * nnz)
    print(f"{name}: nnz {nnz}, cache {L}, batch {B}: {dt * 1e3:.3f} ms per launch, {alg / dt / 1e9:.0f} GB/s algorithmic "
          f"(8(cache+nnz)+4 ncaches nnz per instance)")
