"""The stage loops of the trajectory-optimisation front end (src/trajectory_optimization/
evaluate.jl:15-28,77-136,206-241,297-327, dynamics.jl:143-148,172-179, constraints.jl:169-176,203-212, costs.jl:115-120)
through the C ABI (cb200_stage_plan / cb200_stage_scatter) against the oracle's literal restatement.  Bit-exact: the
scatter moves values and adds them in the reference's program order.

Index tables follow trajectory_optimization/indices.jl: variables [x_1, u_1, ..., x_T] (dynamics.jl:333-340), stage t of the
dynamics owns (x_t, u_t, x_{t+1}) (state_action_next_state_indices), stage constraints and costs own (x_t, u_t);
equalities ordered dynamics, then stage constraints (data.jl:51-55)."""
import numpy as np
import pytest

import backends
import problems
from calipso_b200.solver import BatchKKT
from oracle import oracle as orc


def trajopt_indices(T, nx, nu):
    """0-based index tables of a T-stage problem with constant dimensions (indices.jl:136-139, 330-372)."""
    nz = nx + nu
    xu = [list(range(t * nz, t * nz + (nz if t < T - 1 else nx))) for t in range(T)]                  # state_action_indices
    xuy = [list(range(t * nz, t * nz + nz + nx)) for t in range(T - 1)]                               # state_action_next_state
    return xu, xuy


def pendulum_stage_caches(P, T, v, y):
    """What the generated stage functions of BASELINE cfg1 (README.md:129-176) leave in their caches at (v, y): the
    dynamics' constraint and dual-Jacobian caches for t = 1..T-1, then the two stage equalities (initial and goal state),
    and the costs' gradient caches."""
    nx, nu, h = 2, 1, 0.05
    ml2, grav_l, damp = 0.25, 9.81 / 0.5, 0.1 / 0.25
    nz = nx + nu
    dyn_g, dyn_dual = [], []
    for t in range(T - 1):
        x, u, yn = v[t * nz:t * nz + 2], v[t * nz + 2], v[(t + 1) * nz:(t + 1) * nz + 2]
        lam = y[2 * t:2 * t + 2]
        xm = 0.5 * (x + yn)
        fc = np.array([xm[1], u / ml2 - grav_l * np.sin(xm[0]) - damp * xm[1]])
        dyn_g.append(yn - (x + h * fc))
        # d = y - x - h f((x + y) / 2, u); rows of its Jacobian with respect to (x, u, y)
        A = np.array([[0.0, 1.0], [-grav_l * np.cos(xm[0]), -damp]])          # df/dxm
        Jx = -np.eye(2) - 0.5 * h * A
        Ju = -h * np.array([[0.0], [1.0 / ml2]])
        Jy = np.eye(2) - 0.5 * h * A
        J = np.hstack([Jx, Ju, Jy])
        dyn_dual.append(J.T @ lam)
    x_init, x_goal = np.array([0.0, 0.0]), np.array([np.pi, 0.0])
    eq_g = [v[0:2] - x_init, v[(T - 1) * nz:(T - 1) * nz + 2] - x_goal]
    eq_dual = [np.concatenate([y[2 * (T - 1):2 * (T - 1) + 2], np.zeros(nu)]), y[2 * (T - 1) + 2:2 * (T - 1) + 4]]
    cost_grad = [0.2 * v[t * nz:t * nz + (nz if t < T - 1 else nx)] for t in range(T)]
    return dyn_g, dyn_dual, eq_g, eq_dual, cost_grad


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_pendulum_stage_loops(backend):
    T = 11
    P = problems.pendulum(0, T)
    B = 4
    k = BatchKKT(P, batch=B, binding=backends.binding(backend))
    xu, xuy = trajopt_indices(T, 2, 1)
    # equality!: dynamics rows, then the two stage equalities (t = 1 and t = T)
    g_idx = [[2 * t, 2 * t + 1] for t in range(T - 1)] + [[2 * (T - 1), 2 * (T - 1) + 1], [2 * (T - 1) + 2, 2 * (T - 1) + 3]]
    # equality_dual_jacobian_variables!: dynamics over (x_t, u_t, x_{t+1}), then stage equalities over (x_t, u_t)
    d_idx = xuy + [xu[0], xu[T - 1]]
    k.stage_plan("EQUALITY", g_idx, accumulate=False)
    k.stage_plan("EQ_DUAL_GRAD", d_idx, accumulate=True)
    k.stage_plan("GRADIENT", xu, accumulate=True)
    rng = np.random.default_rng(5)
    pts = [(rng.standard_normal(P.n), rng.standard_normal(P.m)) for _ in range(B)]
    cg, cd, cf = [], [], []
    for v, y in pts:
        dyn_g, dyn_dual, eq_g, eq_dual, cost_grad = pendulum_stage_caches(P, T, v, y)
        cg.append(np.concatenate(dyn_g + eq_g))
        cd.append(np.concatenate(dyn_dual + eq_dual))
        cf.append(np.concatenate(cost_grad))
    k.stage_scatter("EQUALITY", np.stack(cg))
    k.stage_scatter("EQ_DUAL_GRAD", np.stack(cd))
    k.stage_scatter("GRADIENT", np.stack(cf))
    G, D, F = k.get("EQUALITY"), k.get("EQ_DUAL_GRAD"), k.get("GRADIENT")
    for b, (v, y) in enumerate(pts):
        dyn_g, dyn_dual, eq_g, eq_dual, cost_grad = pendulum_stage_caches(P, T, v, y)
        assert np.array_equal(G[b], orc.stage_scatter(P.m, g_idx, dyn_g + eq_g, False))
        assert np.array_equal(D[b], orc.stage_scatter(P.n, d_idx, dyn_dual + eq_dual, True))
        assert np.array_equal(F[b], orc.stage_scatter(P.n, xu, cost_grad, True))
        # ... and they are the flat callbacks' values: g(x), J(x)' y, grad f
        assert np.allclose(G[b], P.g(v), rtol=0, atol=1e-13)
        assert np.allclose(D[b], P.jac_g(v).T @ y, rtol=0, atol=1e-12)
        assert np.allclose(F[b], P.grad(v), rtol=0, atol=1e-15)


class _Pattern:
    """diagonal W, one entry per row of G: the stage loops do not depend on the matrix patterns"""

    def __init__(self, n, m):
        self.n, self.m, self.p, self.num_nonnegative, self.soc_dims = n, m, 0, 0, np.zeros(0, np.int32)
        self.W_colptr, self.W_rowval = np.arange(n + 1), np.arange(n)
        cols = np.arange(m) % n
        order = np.argsort(cols, kind="stable")
        self.G_rowval = np.arange(m)[order]
        self.G_colptr = np.concatenate([[0], np.cumsum(np.bincount(cols, minlength=n))])
        self.C_colptr, self.C_rowval = np.zeros(n + 1, np.int32), np.zeros(0, np.int32)


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("seed,T,nx,nu", [(0, 3, 2, 1), (1, 12, 5, 3), (2, 40, 12, 4)])
def test_random_overlaps_exact(backend, seed, T, nx, nu):
    """Random caches through overlapping index lists (x_{t+1} is written by stage t and by stage t+1): sums in program order,
    last write wins for `.=`, unwritten entries are 0.0, instances outside the range keep their values; a second plan replaces
    the first."""
    rng = np.random.default_rng(seed)
    xu, xuy = trajopt_indices(T, nx, nu)
    n = T * nx + (T - 1) * nu
    m = (T - 1) * nx + 2
    P = _Pattern(n, m)
    B = 5
    k = BatchKKT(P, batch=B, binding=backends.binding(backend))
    # (g'y)_x: dynamics blocks, then a stage constraint on every other stage, then a "general" list with repeats inside one list
    general = [list(rng.integers(0, n, size=2 * n))]
    d_idx = xuy + [xu[t] for t in range(0, T, 2)] + general
    k.stage_plan("EQ_DUAL_GRAD", d_idx, accumulate=True)
    # g(x) with `.=`: rows of the dynamics, then lists that overwrite some of them; the last row is never written
    g_idx = [list(range(t * nx, (t + 1) * nx)) for t in range(T - 1)] + [list(rng.integers(0, m - 1, size=m // 2))]
    k.stage_plan("EQUALITY", g_idx, accumulate=False)
    Ld, Lg = sum(len(i) for i in d_idx), sum(len(i) for i in g_idx)
    cd = rng.standard_normal((B, Ld)) * np.exp(rng.uniform(-8, 8, (B, Ld)))      # wide dynamic range: order matters
    cg = rng.standard_normal((B, Lg))
    sentinel = np.full((B, n), 7.25)
    k.set("EQ_DUAL_GRAD", sentinel)
    k.stage_scatter("EQ_DUAL_GRAD", cd[1:4], first=1)
    k.stage_scatter("EQUALITY", cg)
    D, G = k.get("EQ_DUAL_GRAD"), k.get("EQUALITY")
    assert np.array_equal(D[0], sentinel[0]) and np.array_equal(D[4], sentinel[4])

    def split(vals, idx):
        offs = np.cumsum([0] + [len(i) for i in idx])
        return [vals[offs[q]:offs[q + 1]] for q in range(len(idx))]

    order_matters = 0
    for b in range(B):
        ref_g = orc.stage_scatter(m, g_idx, split(cg[b], g_idx), False)
        assert np.array_equal(G[b], ref_g) and G[b][m - 1] == 0.0
        if 1 <= b <= 3:
            ref_d = orc.stage_scatter(n, d_idx, split(cd[b], d_idx), True)
            assert np.array_equal(D[b], ref_d)
            rev = orc.stage_scatter(n, d_idx[::-1], split(cd[b], d_idx)[::-1], True)
            order_matters += int(not np.array_equal(ref_d, rev))
    assert order_matters > 0          # the test would not notice a different summation order otherwise
    # re-plan: shorter list, other mode
    k.stage_plan("EQ_DUAL_GRAD", [xu[0]], accumulate=False)
    k.stage_scatter("EQ_DUAL_GRAD", np.arange(B * len(xu[0]), dtype=float).reshape(B, -1))
    D = k.get("EQ_DUAL_GRAD")
    for b in range(B):
        expect = np.zeros(n)
        expect[xu[0]] = np.arange(b * len(xu[0]), (b + 1) * len(xu[0]))
        assert np.array_equal(D[b], expect)


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_errors(backend):
    P = problems.pendulum(0, 5)
    k = BatchKKT(P, batch=2, binding=backends.binding(backend))
    from calipso_b200 import _lib
    with pytest.raises(_lib.CalipsoB200Error):
        k.stage_plan("EQUALITY", [[0, P.m]], accumulate=False)          # index out of range
    with pytest.raises(_lib.CalipsoB200Error):
        k.stage_plan("POINT", [[0]], accumulate=True)                   # not a stage-loop target
    k._stage_len = {"CONE": 0}
    with pytest.raises(_lib.CalipsoB200Error):
        k.stage_scatter("CONE", np.zeros((2, 0)))                       # no plan for this array


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_pendulum_solve_with_device_stage_loops(backend):
    """BASELINE cfg1 solved with the front end's stage loops on the device: the callbacks only fill the stage caches
    (what the generated functions do), cb200_stage_scatter forms grad f, g(x) and (g'y)_x.  Same Newton iterations as the
    flat callbacks, same solution (the sums differ from J' y in the last bits only)."""
    from calipso_b200.solver import Solver, initialize, solve
    T = 11
    P = problems.pendulum(0, T)
    b = backends.binding(backend)
    kd = BatchKKT(P, batch=1, binding=b)                  # the handle whose vectors the stage loops fill
    xu, xuy = trajopt_indices(T, 2, 1)
    g_idx = [[2 * t, 2 * t + 1] for t in range(T - 1)] + [[2 * (T - 1), 2 * (T - 1) + 1], [2 * (T - 1) + 2, 2 * (T - 1) + 3]]
    d_idx = xuy + [xu[0], xu[T - 1]]
    kd.stage_plan("EQUALITY", g_idx, accumulate=False)
    kd.stage_plan("EQ_DUAL_GRAD", d_idx, accumulate=True)
    kd.stage_plan("GRADIENT", xu, accumulate=True)
    calls = [0]

    def staged(flags, x, y, z, out):
        P.callback(flags & ~(2 | 4 | 16), x, y, z, out)            # objective value, Hessian and Jacobian caches as before
        if flags & (2 | 4 | 16):
            dyn_g, dyn_dual, eq_g, eq_dual, cost_grad = pendulum_stage_caches(P, T, x, y)
            if flags & 2:
                kd.stage_scatter("GRADIENT", np.concatenate(cost_grad)[None])
                out.gradient[:] = kd.get("GRADIENT")[0]
            if flags & 4:
                kd.stage_scatter("EQUALITY", np.concatenate(dyn_g + eq_g)[None])
                out.equality[:] = kd.get("EQUALITY")[0]
            if flags & 16:
                kd.stage_scatter("EQ_DUAL_GRAD", np.concatenate(dyn_dual + eq_dual)[None])
                out.eq_dual_grad[:] = kd.get("EQ_DUAL_GRAD")[0]
            calls[0] += 1

    s = Solver(P, staged, binding=b)
    initialize(s, P.x0)
    assert solve(s) is True
    ref = Solver(P, P.callback, binding=b)
    initialize(ref, P.x0)
    assert solve(ref) is True
    assert calls[0] > 0 and s.iterations == ref.iterations
    assert np.abs(s.solution - ref.solution).max() / (1.0 + np.abs(ref.solution).max()) < 1e-9
