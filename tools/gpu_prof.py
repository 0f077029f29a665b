import sys, time, json
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
Ps = [lqc.cfg3(i) for i in range(8)]
k = BatchKKT(Ps[0], batch=B)
print(k.info())
k.load_lq([Ps[i % 8] for i in range(B)]); X0 = np.stack([Ps[i % 8].x0 for i in range(B)])
def solve():
    k.initialize(X0); k.lq_begin(); return k.lq_solve(max_steps=400, check_every=4)
solve(); k.profile()
t=time.time(); r = solve(); k.synchronize(); dt=time.time()-t
st = k.stats(); its = int((st['total_iterations']-1).sum())
prof = k.profile()
print('solve', r, 'time', dt, 'iters', its, 'it/s', its/dt)
tot = sum(v for kk,v in prof.items() if kk!='total')
for kk,v in prof.items(): print(f'  {kk:20s} {v/B/1.9e3:10.1f} us/instance  {100*v/max(tot,1):5.1f}%')
print('factorizations', int(st['factorizations'].sum())/B, 'solves', int(st['solves'].sum())/B)
# kkt factor solve timing
k.lq_begin(); k.lq_step(4); k.set_scalars(eps_p=1e-7, eps_d=1e-7)
for ns in (0, 1, 5):
    k.kkt_factor_solve(ns); k.synchronize(); k.profile()
    t=time.time()
    for _ in range(5): k.kkt_factor_solve(ns)
    k.synchronize(); dt=(time.time()-t)/5
    prof = k.profile()
    print(f'kkt_factor_solve nsolves={ns}: {dt*1e3:.2f} ms per launch', {kk: round(v/B/5/1.9e3,1) for kk,v in prof.items() if v})
