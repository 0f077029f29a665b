"""Known answers / stopping criteria of the reference's solver tests, reached by the oracle's solve! restatement.

Mirrors test/solver/{wachter,maratos,knitro,friction_cone,portfolio,test1,test2,test3,test4}.jl: four stopping criteria at exit
(e.g. wachter.jl:36-45) plus the known optima.
"""
import itertools

import numpy as np
import pytest

import problems
from calipso_b200 import lqc
from oracle import oracle as orc


def solve(P, x0=None, lu=True):
    o = orc.Oracle(P.n, P.m, P.p, P.num_nonnegative, P.soc_dims, P.W_colptr, P.W_rowval, P.G_colptr, P.G_rowval,
                   P.C_colptr, P.C_rowval)
    o.set_callback(P.callback)
    if lu:
        o.use_superlu_fallback()
    o.initialize(P.x0 if x0 is None else x0)
    rc = o.solve()
    return o, rc


def check_criteria(o, tol=1e-4):
    R = o.residual
    assert np.abs(R).sum() / o.total < tol
    slack = max(np.abs(R[o.iy]).max(initial=0.0), np.abs(R[o.iz]).max(initial=0.0))
    assert slack < tol
    assert np.abs(o.equality).max(initial=0.0) <= tol
    assert np.abs(o.cone_product).max(initial=0.0) <= tol


def test_wachter():
    P = problems.wachter()
    o, rc = solve(P)
    assert rc == 1
    check_criteria(o)
    assert np.abs(o.solution[:3] - P.x_star).max() < 1e-3          # wachter.jl:47


def test_maratos():
    o, rc = solve(problems.maratos())
    assert rc == 1
    check_criteria(o)


def test_knitro():
    o, rc = solve(problems.knitro())
    assert rc == 1
    check_criteria(o)


def test_test1_to_test4():
    for P in (problems.test1(), problems.test2(), problems.test3(), problems.test4()):
        o, rc = solve(P)
        assert rc == 1
        check_criteria(o)


@pytest.mark.parametrize("v,mu,gamma", list(itertools.product(
    [[0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.0, 1.0, 1.0], [0.0, 10.0, 1.0]], [0.0, 0.5, 1.0], [0.0, 1.0])))
def test_friction_cone(v, mu, gamma):
    """friction_cone.jl:19-63 over the 30 (v, mu, gamma) combinations."""
    rng = np.random.default_rng(int(10 * mu + gamma + sum(v)))
    P = problems.friction(v, mu, gamma, rng.standard_normal(3))
    o, rc = solve(P)
    assert rc == 1
    check_criteria(o)
    assert not o.cone_violation(o.solution[o.is_], np.zeros(3), 0.0)
    x = o.solution[:3]
    v = np.asarray(v)
    if np.linalg.norm(v[1:]) > 0 and gamma > 0 and mu > 0:
        vdir = v[1:] / np.linalg.norm(v[1:])
        bdir = x[1:] / np.linalg.norm(x[1:])
        assert np.abs(vdir + bdir).max() < 1e-3
        assert np.linalg.norm(x[1:]) <= mu * gamma + 1e-6


def test_portfolio():
    P = problems.portfolio(seed=1)
    o, rc = solve(P)
    assert rc == 1
    check_criteria(o)
    s = o.solution[o.is_]
    assert np.all(s[:2] > -1e-5)
    assert np.linalg.norm(s[3:]) < s[2] + 1e-5
    assert np.abs(P.b - P.A @ o.solution[:P.n] - s).max() < 1e-4


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_qp_nonnegative(seed):
    """test/solver/qp_nonnegative.jl:49-62: stopping criteria, x >= -1e-4, ||A x - b||_inf < equality tolerance."""
    P = problems.qp_nonnegative(seed)
    o, rc = solve(P)
    assert rc == 1
    check_criteria(o)
    x = o.solution[:P.n]
    assert np.all(x > -1.0e-4)
    assert np.abs(P.A @ x - P.b).max() < 1e-4


@pytest.mark.parametrize("overwrite", [False, True])
def test_pendulum_readme_quickstart(overwrite):
    """README.md:129-176 / test/examples/pendulum.jl:64-73 (BASELINE cfg1): converges, dynamics feasible, goal reached --
    with the summed second derivatives and with the reference's actual last-write-wins scatter (Appendix A.17)."""
    P = problems.pendulum(overwrite=overwrite)
    o, rc = solve(P)
    assert rc == 1
    check_criteria(o)
    x = o.solution[:P.n]
    assert np.abs(P.g(x)).max() < 1e-4
    assert np.abs(x[P.ix[-1]:P.ix[-1] + 2] - P.x_goal).max() < 1e-4


def test_rocket_landing_example():
    """test/examples/rocket_landing.jl:72-91: converges, every thrust strictly inside its cone, stopping criteria."""
    P = problems.rocket_landing()
    o = orc.from_problem(P)
    o.use_superlu_fallback()
    o.initialize(P.x0)
    assert o.solve() == 1
    check_criteria(o)
    x = o.solution[:P.n]
    u = np.array([x[t * 9 + 6:t * 9 + 9] for t in range(P.meta["T"] - 1)])
    assert np.all(np.linalg.norm(u[:, :2], axis=1) < u[:, 2])               # rocket_landing.jl:80
    assert np.abs(x[-6:]).max() < 1e-4 and np.abs(x[:6] - [3.0, 2.0, 1.0, 0.0, 0.0, 0.0]).max() < 1e-4


def test_lqc_family_converges_and_is_feasible():
    for P in (lqc.tiny(), lqc.cfg2()):
        o = orc.from_problem(P)
        o.use_superlu_fallback()
        o.initialize(P.x0)
        assert o.solve() == 1
        check_criteria(o)
        x = o.solution[:P.n]
        assert np.abs(P.G() @ x + P.g0).max() < 1e-4
        h = P.C() @ x + P.h0
        assert h[:P.num_nonnegative].min() > -1e-4
        off = P.num_nonnegative
        for d in P.soc_dims:
            assert h[off] - np.linalg.norm(h[off + 1:off + d]) > -1e-4
            off += d


def test_reference_schedule_counts_factorisations():
    """Reference schedule (linear_solver.jl:56, iterative_refinement.jl:21-25): IC trials + 1 (linear_solve!) +
    one per refinement pass; dedup schedule: IC trials only.  Same numbers either way."""
    P = lqc.tiny()
    runs = []
    for ref in (0, 1):
        o = orc.from_problem(P, options=dict(reference_schedule=ref))
        o.initialize(P.x0)
        o.solve_begin()
        o.newton_iteration()            # inner-converged at kappa=1 (break) -> outer update
        o.outer_update()
        before = o.L.orc_qdldl_factor_count(o.ldl().h)
        assert o.newton_iteration() == 0
        st = o.stats
        runs.append((o.L.orc_qdldl_factor_count(o.ldl().h) - before, st["n_trials"], st["n_refine"], o.step.copy()))
    (f0, t0, r0, s0), (f1, t1, r1, s1) = runs
    assert (t0, r0) == (t1, r1)
    assert f0 == t0 and f1 == t1 + 1 + r1
    assert np.array_equal(s0, s1)
