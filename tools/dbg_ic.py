import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import numpy as np
import problems, backends
from calipso_b200.solver import BatchKKT
P = problems.maratos()
k = BatchKKT(P, binding=backends.binding(sys.argv[1]))
print(k.info(), flush=True)
from test_parity_kkt import push_state
from oracle import oracle as orc
o = orc.Oracle(P.n, P.m, P.p, P.num_nonnegative, P.soc_dims, P.W_colptr, P.W_rowval, P.G_colptr, P.G_rowval, P.C_colptr, P.C_rowval)
o.set_callback(P.callback)
o.initialize(np.array([0.5, 0.5])); o.solution[o.iy] = -10.0
o.set_scalars(kappa=1.0, rho=1.0); o.evaluate(511); o.residual_eval()
push_state(k, o)
print('pushed', flush=True)
k.kkt_factor_solve(0); k.synchronize(); print('factor ok', k.stats()['inertia_pos'], flush=True)
k.kkt_factor_solve(1); k.synchronize(); print('factor+solve ok', flush=True)
k.search_direction(); k.synchronize(); print('sd ok', {kk:int(v[0]) for kk,v in k.stats().items()}, flush=True)
