"""differentiate! (src/solver/differentiate.jl:1-61): solution sensitivities through the
C ABI vs the oracle's restatement, plus the property the reference's own test pins -- for an equality/nonnegative
problem the sensitivities solve J S = -dR/dtheta (test/solver/qp_equality.jl:116-122 compares with a dense solve)."""
import numpy as np
import pytest

import backends
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT
from test_parity_kkt import oracle_at_iteration, push_state, rel


def prepared(P, backend, iters):
    k = BatchKKT(P, binding=backends.binding(backend))
    perm, _, _ = k.symbolic()
    o = oracle_at_iteration(P, iters, perm=perm)
    o.evaluate(2 | 16 | 32)
    o.cone_eval(barrier=True, barrier_gradient=True)
    o.residual_eval()
    o.evaluate(64 | 128 | 256)
    o.cone_eval(jacobian=True)
    o.set_scalars(eps_p=1e-7, eps_d=1e-7)
    push_state(k, o)
    k.set_scalars(eps_p=1e-7, eps_d=1e-7)
    return k, o


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("make,iters", [(lqc.tiny, 6), (lqc.cfg2, 4)])
def test_sensitivities_match_oracle(backend, make, iters):
    P = make()
    k, o = prepared(P, backend, iters)
    rng = np.random.default_rng(5)
    H = rng.standard_normal((o.total, 3))
    S = k.differentiate(H)
    So = o.differentiate(H)
    assert S.shape == So.shape
    for i in range(H.shape[1]):
        assert rel(S[:, i], So[:, i]) < 1e-8


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_sensitivities_solve_the_newton_system_without_soc(backend):
    """No second-order cones => the reduced solve is exact: J S = -H to rounding (the reference's QP test idea)."""
    P = lqc.tiny(0, n_soc=0)
    k, o = prepared(P, backend, 5)
    rng = np.random.default_rng(7)
    H = rng.standard_normal((o.total, 4))
    S = k.differentiate(H)
    o.residual_jacobian_variables()
    J = o.dense_jacobian()
    assert np.abs(J @ S + H).max() <= 1e-7 * max(1.0, np.abs(H).max()) * np.abs(S).max()


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_batched_sensitivities(backend):
    Ps = [lqc.tiny(i) for i in range(3)]
    k = BatchKKT(Ps[0], batch=3, binding=backends.binding(backend))
    k.load_lq(Ps)
    k.initialize(np.stack([P.x0 for P in Ps]))
    k.lq_begin()
    k.lq_solve(max_steps=300, check_every=3)
    rng = np.random.default_rng(9)
    H = rng.standard_normal((3, k.total, 2))
    S = k.differentiate(H)
    for b, P in enumerate(Ps):            # each instance equals its own single-instance call
        k1 = BatchKKT(P, binding=backends.binding(backend))
        for name in ("POINT", "DUAL", "SCALARS", "W_VALUES", "G_VALUES", "C_VALUES"):
            k1.set(name, k.get(name, first=b, count=1))
        assert np.array_equal(k1.differentiate(H[b]), S[b])
