"""One launch of every kernel of libcalipso_b200 on a cfg3 batch (for `ncu --set full -k regex:^k_ ...` captures):
python tools/ncu_all_kernels.py [BATCH]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp

from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT, LDLSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 444
Ps = [lqc.cfg3(i) for i in range(8)]
plist = [Ps[i % 8] for i in range(B)]
k = BatchKKT(Ps[0], batch=B)
k.load_lq(plist)
k.initialize(np.stack([P.x0 for P in plist]))
k.lq_begin()                                   # k_expand, k_lq_begin
k.lq_step(3)                                   # k_lq_step (three Newton iterations per instance)
k.lq_evaluate(2 | 16 | 32)                     # k_lq_evaluate
k.cone(barrier=True, barrier_gradient=True, product=True)      # k_cone
k.residual()                                   # k_residual
k.search_direction()                           # k_search_direction
k.cone_search()                                # k_cone_search
k.lq_evaluate(1 | 4 | 8, at_candidate=True)
k.apply_step()                                 # k_apply_step
k.set_scalars(eps_p=1e-7, eps_d=1e-7)
k.kkt_factor_solve(1)                          # k_kkt_factor_solve (the KKT-solve unit)
k.kkt_factor_solve(0)                          # ... factorisation only
k.jacobian_times(np.ones((B, k.total)))        # k_jtimes
H = np.random.default_rng(0).standard_normal((B, k.total, 4))
k.differentiate(H)                             # k_differentiate (4 parameters)
P0 = Ps[0]
keysW = [(int(r), int(c)) for c in range(P0.n) for r in P0.W_rowval[P0.W_colptr[c]:P0.W_colptr[c + 1]]]
k.scatter_plan("W_VALUES", [keysW])
k.scatter("W_VALUES", np.stack([P.W_val for P in plist]))      # k_scatter
# the front end's stage loops: (g'y)_x through overlapping (x_t, u_t, x_{t+1}) index lists      k_stage_gather
T, nx, nu = 40, 36, 12
xuy = [list(range(t * (nx + nu), t * (nx + nu) + nx + nu + nx)) for t in range(T - 1)]
k.stage_plan("EQ_DUAL_GRAD", xuy, accumulate=True)
k.stage_scatter("EQ_DUAL_GRAD", np.random.default_rng(1).standard_normal((B, sum(len(i) for i in xuy))))
# filter line search over a block of 4 candidates (callback outputs uploaded)                   k_filter_reset, k_filter_search
k.filter_reset()
rng = np.random.default_rng(2)
k.filter_search(0, rng.standard_normal((B, 4)), rng.standard_normal((B, 4, k.m)), np.abs(rng.standard_normal((B, 4, k.p))) + 1.0)
k.synchronize()
# LinearSolver seam on a cfg3-shaped quasi-definite matrix: k_ldl_factor, k_ldl_solve
n, m, p = P0.n, P0.m, P0.p
K = sp.bmat([[P0.W_full() + 1e-7 * sp.eye(n), P0.G().T, P0.C().T], [P0.G(), -(1.0 + 1e-7) * sp.eye(m), None],
             [P0.C(), None, -0.7 * sp.eye(p)]]).tocsc()
s = LDLSolver(K, batch=B)
x = np.zeros((B, n + m + p))
s.linear_solve(x, K, np.ones((B, n + m + p)))
