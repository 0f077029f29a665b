"""cb200_differentiate timing probe (8 instances x 64 parameters, pinned host buffers) for one or more builds.
python tools/r2_diff_probe.py a.so b.so ...     Developer tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from calipso_b200 import _lib, lqc
from calipso_b200.solver import BatchKKT
from tools.r2_ab_common import LooseBinding

nb, npar = 8, 64
Pd = [lqc.cfg3(i) for i in range(nb)]
for path in sys.argv[1:]:
    kd = BatchKKT(Pd[0], batch=nb, binding=LooseBinding(path))
    kd.load_lq(Pd)
    kd.initialize(np.stack([P.x0 for P in Pd]))
    kd.lq_begin()
    r = kd.lq_solve(max_steps=400, check_every=400)
    Hd = torch.empty((nb * npar, kd.total), dtype=torch.float64).pin_memory()
    Hd.numpy()[...] = np.random.default_rng(0).standard_normal(Hd.shape)
    Sd = torch.empty_like(Hd).pin_memory()
    ts = []
    for rep in range(6):
        t = time.perf_counter()
        kd.b.check(kd.lib.cb200_differentiate(kd.h, npar, _lib.C.cast(Hd.data_ptr(), _lib.c_dp), _lib.C.cast(Sd.data_ptr(), _lib.c_dp)))
        ts.append((time.perf_counter() - t) * 1e3)
    print(path, r, "ms per call:", [round(x, 2) for x in ts], "finite:", bool(np.isfinite(Sd.numpy()).all()), flush=True)
    kd.close()
