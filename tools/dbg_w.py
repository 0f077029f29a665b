import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import numpy as np
import problems, backends
from calipso_b200.solver import BatchKKT, Solver, initialize, solve
import calipso_b200.solver as S
P = problems.wachter()
s = Solver(P, P.callback, binding=backends.binding(sys.argv[1]))
k = s.kkt
print(k.info(), flush=True)
# wrap calls with prints
for name in ['cone','residual','search_direction','cone_search','apply_step']:
    f = getattr(k, name)
    def mk(f, name):
        def g(*a, **kw):
            print('  call', name, flush=True); r = f(*a, **kw); k.synchronize(); print('  done', name, {kk:int(v[0]) for kk,v in k.stats().items() if kk in ('n_trials','n_refine','status','inertia_pos','used_fallback','gmres_iters')}, flush=True); return r
        return g
    setattr(k, name, mk(f, name))
initialize(s, P.x0)
print('solve ->', solve(s), s.iterations, flush=True)
