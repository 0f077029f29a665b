"""options.warmstart (solve.jl:10-13): a second solve! that keeps slacks and duals of the
previous solution instead of re-initialising them -- product (on-device LQ loop) vs the oracle, MPC-style: the linear cost
is perturbed between the two solves."""
import numpy as np
import pytest

import backends
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT
from oracle import oracle as orc


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_warmstarted_resolve_matches_oracle(backend):
    P = lqc.tiny(2)
    k = BatchKKT(P, binding=backends.binding(backend))
    perm, _, _ = k.symbolic()
    k.load_lq(P)
    k.initialize(P.x0)
    k.lq_begin()
    assert k.lq_solve(max_steps=300, check_every=2)["converged"] == 1
    o = orc.from_problem(P, perm=perm)
    o.use_superlu_fallback()
    o.initialize(P.x0)
    assert o.solve() == 1
    cold_iterations = o.stats["total_iterations"]
    # perturb the linear cost, keep the point (primal, slacks, duals) and re-solve warm
    rng = np.random.default_rng(0)
    q2 = P.q + 0.05 * rng.standard_normal(P.n)
    k.set("LQ_Q", q2)
    k.lq_begin(warmstart=True)
    assert k.lq_solve(max_steps=300, check_every=2)["converged"] == 1
    o2 = orc.from_problem(P, perm=perm, options=dict(warmstart=1))
    o2.use_superlu_fallback()
    o2.set_lq(P.W_val, P.G_val, P.C_val, q2, P.g0, P.h0)
    o2.solution[:] = o.solution            # initialize! is not called: the previous solution is the warm start
    assert o2.solve() == 1
    st = {kk: int(v[0]) for kk, v in k.stats().items()}
    assert st["total_iterations"] == o2.stats["total_iterations"]
    assert np.abs(k.get("POINT")[0] - o2.solution).max() <= 1e-6 * max(1.0, np.abs(o2.solution).max())
    assert o2.stats["total_iterations"] <= cold_iterations + 2


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_schedule_hint_does_not_change_results(backend):
    """cb200_lq_set_order: the order in which the instances of a batch are started (longest first in an MPC-style re-solve)
    is a scheduling hint only -- every instance's iterates are bitwise those of the default order."""
    Ps = [lqc.tiny(i) for i in range(7)]
    k = BatchKKT(Ps[0], batch=7, binding=backends.binding(backend))
    k.load_lq(Ps)
    X0 = np.stack([P.x0 for P in Ps])
    out = []
    for order in (None, [3, 6, 0, 5, 1, 4, 2]):
        k.lq_set_order(order)
        k.initialize(X0)
        k.lq_begin()
        assert k.lq_solve(max_steps=300, check_every=300)["converged"] == 7
        out.append((k.get("POINT"), k.stats()["total_iterations"].copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    with pytest.raises(Exception, match="permutation"):
        k.lq_set_order([0, 0, 1, 2, 3, 4, 5])
