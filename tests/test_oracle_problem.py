"""Restatement of the reference's hot-path unit test, test/solver/problem.jl:1-212, against the oracle.

The reference test uses unseeded randn inputs, so what it pins are identities; the same identities are asserted
here (same tolerances) on seeded inputs, plus a second-order-cone variant the reference exercises only through
its convergence tests.
"""
import numpy as np
import pytest

import problems
from calipso_b200 import lqc
from oracle import oracle as orc

ALL = 1 | 2 | 4 | 8 | 16 | 32 | 64 | 128 | 256


def setup_qp(seed):
    P = problems.random_qp(seed)
    o = orc.Oracle(P.n, P.m, P.p, P.num_nonnegative, P.soc_dims, P.W_colptr, P.W_rowval, P.G_colptr, P.G_rowval,
                   P.C_colptr, P.C_rowval)
    o.set_callback(P.callback)
    rng = np.random.default_rng(100 + seed)
    w = o.solution
    w[o.ix] = rng.standard_normal(P.n)
    w[o.ir] = rng.random(P.m)
    w[o.is_] = rng.random(P.p)
    w[o.iy] = rng.standard_normal(P.m)
    w[o.iz] = rng.standard_normal(P.p)
    w[o.it] = rng.random(P.p)
    o.dual[:] = rng.standard_normal(P.m)
    o.set_scalars(kappa=0.17, rho=52.0, eps_p=0.12, eps_d=0.21)      # problem.jl:50-59
    o.evaluate(ALL)
    o.cone_eval(product=True, jacobian=True, target=True)
    o.residual_jacobian_variables()
    o.residual_jacobian_variables_symmetric()
    o.residual_eval()
    o.residual_symmetric_eval()
    return P, o


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_problem_jl_identities(seed):
    P, o = setup_qp(seed)
    n, m, p = P.n, P.m, P.p
    kappa, rho, ep, ed = 0.17, 52.0, 0.12, 0.21
    x, r, s, y, z, t = (o.solution[i] for i in (o.ix, o.ir, o.is_, o.iy, o.iz, o.it))
    lam = o.dual
    J, K = o.dense_jacobian(), o.dense_symmetric()
    ix, ir, is_, iy, iz, it = (np.arange(total)[sl] for total in [o.total] for sl in (o.ix, o.ir, o.is_, o.iy, o.iz, o.it))
    W = P.hess(x) + P.hess_gy(x, y) + P.hess_hz(x, z)
    G, C = P.jac_g(x), P.jac_h(x)
    tol = 1e-6
    nrm = np.linalg.norm
    # KKT matrix, problem.jl:112-142
    assert np.linalg.matrix_rank(J) == o.total
    assert nrm(J[np.ix_(ix, ix)] - (W + ep * np.eye(n))) < tol
    assert nrm(J[np.ix_(iy, ix)] - G) < tol and nrm(J[np.ix_(ix, iy)] - G.T) < tol
    assert nrm(J[np.ix_(iy, iy)] + ed * np.eye(m)) < tol
    assert nrm(J[np.ix_(iz, ix)] - C) < tol and nrm(J[np.ix_(ix, iz)] - C.T) < tol
    assert nrm(J[np.ix_(is_, iz)] + np.eye(p)) < tol and nrm(J[np.ix_(iz, is_)] + np.eye(p)) < tol
    assert nrm(J[np.ix_(is_, it)] + np.eye(p)) < tol
    assert nrm(J[np.ix_(it, is_)] - np.diag(t)) < tol
    assert nrm(J[np.ix_(it, it)] - (np.diag(s) - ed * np.eye(p))) < tol
    assert nrm(J[np.ix_(ir, iy)] + np.eye(m)) < tol and nrm(J[np.ix_(iy, ir)] + np.eye(m)) < tol
    assert nrm(J[np.ix_(ir, ir)] - (rho + ep) * np.eye(m)) < tol
    assert nrm(J[np.ix_(is_, is_)] - ep * np.eye(p)) < tol
    # symmetric KKT matrix, problem.jl:144-159
    assert np.linalg.matrix_rank(K) == o.N
    kx, ky, kz = np.arange(n), n + np.arange(m), n + m + np.arange(p)
    assert nrm(K[np.ix_(kx, kx)] - (W + ep * np.eye(n))) < tol
    assert nrm(K[np.ix_(ky, kx)] - G) < tol and nrm(K[np.ix_(kx, ky)] - G.T) < tol
    assert nrm(K[np.ix_(ky, ky)] - (-1.0 / (rho + ep) - ed) * np.eye(m)) < tol
    assert nrm(K[np.ix_(kz, kx)] - C) < tol and nrm(K[np.ix_(kx, kz)] - C.T) < tol
    assert nrm(K[np.ix_(kz, kz)] - np.diag(-1.0 * (s - ed) / (t + (s - ed) * ep) - ed)) < tol
    # residual, problem.jl:161-178
    R = o.residual
    assert nrm(R[ix] - (P.grad(x) + G.T @ y + C.T @ z)) < tol
    assert nrm(R[ir] - (lam + rho * r - y)) < tol
    assert nrm(R[is_] - (-z - t)) < tol
    assert nrm(R[iy] - (P.g(x) - r)) < tol
    assert nrm(R[iz] - (P.h(x) - s)) < tol
    assert nrm(R[it] - (s * t - kappa)) < tol
    # residual symmetric, problem.jl:180-189
    rs, rt = R[is_], R[it]
    Rs = o.residual_symmetric
    assert nrm(Rs[kx] - R[ix]) < tol
    assert nrm(Rs[ky] - (P.g(x) - r + R[ir] / (rho + ep))) < tol
    assert nrm(Rs[kz] - (P.h(x) - s + (rt + (s - ed) * rs) / (t + (s - ed) * ep))) < tol
    # step: reduced LDL + recovery == unsymmetric solve of J (UMFPACK in the reference), problem.jl:191-204
    delta = np.linalg.solve(J, R)
    o.search_direction_symmetric(factorize=True)
    assert nrm(delta - o.step) < tol
    # iterative refinement from a noise-corrupted step, problem.jl:206-211
    noisy = o.step + np.random.default_rng(5).standard_normal(o.total)
    o.iterative_refinement(noisy)
    assert nrm(R - J @ noisy) < 1.0e-10


def test_second_order_blocks_and_refinement():
    """SOC(3): cone Jacobians are arrow matrices (second_order.jl:19-22); the reduced block keeps only the upper
    triangle of -(T+S P)^-1 S + D (linear_solver.jl:23) so the LDL direction is inexact and
    refinement against the full J is what makes it right."""
    P = lqc.tiny()
    o = orc.from_problem(P)
    o.initialize(P.x0)
    o.solve_begin()
    for _ in range(3):
        if o.newton_iteration() == 2:
            o.outer_update()
    o.evaluate(2 | 16 | 32)
    o.cone_eval(barrier=True, barrier_gradient=True)
    o.residual_eval()
    o.evaluate(64 | 128 | 256)
    o.cone_eval(jacobian=True)
    ep = ed = 1e-7
    o.set_scalars(eps_p=ep, eps_d=ed)
    o.residual_jacobian_variables()
    o.residual_jacobian_variables_symmetric()
    J, K = o.dense_jacobian(), o.dense_symmetric()
    n, m, p, q = P.n, P.m, P.p, P.num_nonnegative
    s, t = o.solution[o.is_], o.solution[o.it]

    def arrow(v):
        A = np.eye(len(v)) * v[0]
        A[0, 1:] = v[1:]
        A[1:, 0] = v[1:]
        return A
    off = q
    it0, is0 = o.it.start, o.is_.start
    asym = 0.0
    for d in P.soc_dims:
        sl = slice(off, off + d)
        T, S = arrow(t[sl]), arrow(s[sl]) - ed * np.eye(d)
        assert np.allclose(J[it0 + off:it0 + off + d, is0 + off:is0 + off + d], T)
        assert np.allclose(J[it0 + off:it0 + off + d, it0 + off:it0 + off + d], S)
        B = -np.linalg.solve(T + S * ep, S) - ed * np.eye(d)
        Kb = K[n + m + off:n + m + off + d, n + m + off:n + m + off + d]
        assert np.allclose(np.triu(Kb), np.triu(B), atol=1e-12)       # upper triangle only
        assert np.allclose(Kb, Kb.T)
        asym = max(asym, np.abs(B - B.T).max())
        off += d
    assert asym > 1e-6                                                 # genuinely non-symmetric here
    exact = np.linalg.solve(J, o.residual)
    o.search_direction_symmetric(factorize=True)
    assert np.abs(o.step - exact).max() > 1e-8                         # inexact before refinement
    assert o.iterative_refinement()
    assert np.abs(o.residual - J @ o.step).max() <= 1e-10
    assert np.allclose(o.step, exact, rtol=1e-8, atol=1e-9)


def test_jacobian_coo_matches_dense():
    P, o = setup_qp(3)
    r, c, v = o.jacobian_coo()
    J = np.zeros((o.total, o.total))
    np.add.at(J, (r, c), v)
    assert np.allclose(J, o.dense_jacobian())


def test_inertia_correction_schedule():
    """inertia.jl:30-79 on an indefinite Hessian: IC-3 always takes max(1e-20, eps_last/3) (the Vector==0.0 test at
    :48 is always false), then x100 per failed trial while eps_last == 0, x8 afterwards."""
    P = problems.maratos()
    o = orc.Oracle(P.n, P.m, P.p, P.num_nonnegative, P.soc_dims, P.W_colptr, P.W_rowval, P.G_colptr, P.G_rowval,
                   P.C_colptr, P.C_rowval)
    o.set_callback(P.callback)
    o.initialize(np.array([0.5, 0.5]))
    o.solution[o.iy] = -10.0           # W = 4I + 2y I = -16 I: wrong inertia until eps_p > 16
    o.set_scalars(kappa=1.0, rho=1.0)
    o.evaluate(ALL)
    assert o.inertia_correction() == 0
    sc = o.scalars()
    # trials: 1e-7 (IC-1), then 1e-20 * 100^k until > 16  -> k = 11 gives 1e2
    assert sc["eps_p"] == pytest.approx(1e-20 * 100.0 ** 11, rel=1e-12)
    assert o.stats["n_trials"] == 1 + 12
    assert o.inertia == (P.n, P.m + P.p, 0)
    assert sc["eps_p_last"] == sc["eps_p"]
    # second call: eps_last != 0 -> start at eps_last/3, x8 per failure
    assert o.inertia_correction() == 0
    e = 1e2 / 3.0
    assert o.scalars()["eps_p"] == pytest.approx(e, rel=1e-12)
    assert o.stats["n_trials"] == 2


def test_cone_violation_is_nonstrict():
    P = lqc.tiny()
    o = orc.from_problem(P)
    x = np.ones(P.p)
    off = P.num_nonnegative
    for d in P.soc_dims:
        x[off + 1:off + d] = 0.1
        off += d
    tau = 0.99
    xh = x.copy()
    assert not o.cone_violation(xh, x, tau)
    xh[0] = (1.0 - tau) * x[0]                # equality counts as violation, nonnegative.jl:31
    assert o.cone_violation(xh, x, tau)
    xh = x.copy()
    off = P.num_nonnegative
    xh[off] = (1 - tau) * x[off]              # head - (1-tau) head = 0 <= ||tail difference||, second_order.jl:46
    assert o.cone_violation(xh, x, tau)
