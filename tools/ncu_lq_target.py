"""ncu target: one k_lq_step launch in which every instance runs ONE Newton iteration at an interior point (the 6th):
python tools/ncu_lq_target.py [BATCH]"""
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
k.lq_step(5)
k.synchronize()
k.lq_step(1)      # <- the captured launch (second k_lq_step)
k.synchronize()
st = k.stats()
print("solves per instance in the run so far", st["solves"].mean(), "factorizations", st["factorizations"].mean())
