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


def equality_qp(seed=0, n=10, m=5):
    """test/solver/qp_equality.jl:1-40: 1/2 x'P x + p'x with P diagonal, A x = b; theta = [diag P; p; vec A; b]."""
    import problems
    rng = np.random.default_rng(seed)
    Q = rng.random((n, n))
    Pd = np.diag(Q.T @ Q).copy()
    pv = rng.standard_normal(n)
    A = rng.random((m, n))
    b = A @ np.maximum(0.0, rng.standard_normal(n))
    z = problems._z
    P = problems.DenseNLP(
        "qp_equality", n, m, 0, 0, np.zeros(0, np.int32),
        f=lambda x: float(0.5 * x @ (Pd * x) + pv @ x), grad=lambda x: Pd * x + pv, hess=lambda x: np.diag(Pd),
        g=lambda x: A @ x - b, jac_g=lambda x: A, hess_gy=lambda x, y: z(n, n),
        h=lambda x: z(0), jac_h=lambda x: z(0, n), hess_hz=lambda x, zz: z(n, n), x0=rng.standard_normal(n))
    return P, Pd, pv, A, b


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_qp_equality_sensitivities(backend):
    """test/solver/qp_equality.jl:84-122: at the solution the sensitivities of x agree (1e-2, the reference's tolerance)
    with the closed form -[P A'; A 0]^-1 [d(Px+p)/dtheta + d(A'y)/dtheta; d(Ax-b)/dtheta] and with -J^-1 dR/dtheta."""
    from calipso_b200.solver import Solver, differentiate, initialize, residual_jacobian_parameters, solve
    from oracle import oracle as orc
    P, Pd, pv, A, b = equality_qp()
    n, m = P.n, P.m
    s = Solver(P, P.callback, binding=backends.binding(backend))
    initialize(s, P.x0)
    assert solve(s) is True
    x, y = s.solution[:n], s.solution[s.kkt.iy]
    assert np.abs(A @ x - b).max() < 1e-4                                   # qp_equality.jl:59-66
    nth = n + n + m * n + m
    Pxp = np.zeros((n, nth)); ATy = np.zeros((n, nth)); Axb = np.zeros((m, nth))
    Pxp[:, :n] = np.diag(x)                                                 # d(P x + p)/d diag(P)
    Pxp[:, n:2 * n] = np.eye(n)                                             # ... / dp
    for j in range(n):
        for i in range(m):
            col = 2 * n + i + j * m                                         # vec(A), column-major
            ATy[j, col] = y[i]                                              # d(A'y)_j / dA_ij
            Axb[i, col] = x[j]                                              # d(Ax - b)_i / dA_ij
    Axb[:, 2 * n + m * n:] = -np.eye(m)
    H = residual_jacobian_parameters(n, m, 0, Pxp, ATy, None, Axb, None)
    S = differentiate(s, H)
    assert S.shape == (s.total, nth)
    rz = np.block([[np.diag(Pd), A.T], [A, np.zeros((m, m))]])
    closed = -np.linalg.solve(rz, np.vstack([Pxp + ATy, Axb]))
    assert np.abs(closed[:n] - S[:n]).max() < 1e-2                          # qp_equality.jl:121
    o = orc.Oracle(P.n, P.m, P.p, P.num_nonnegative, P.soc_dims, P.W_colptr, P.W_rowval, P.G_colptr, P.G_rowval,
                   P.C_colptr, P.C_rowval, perm=s.kkt.symbolic()[0])
    o.set_callback(P.callback)
    o.use_superlu_fallback()
    o.initialize(P.x0)
    assert o.solve() == 1 and o.stats["total_iterations"] == s.iterations
    So = o.differentiate(H)
    assert np.abs(S - So).max() <= 1e-6 * max(1.0, np.abs(So).max())
    o.residual_jacobian_variables()
    full = -np.linalg.solve(o.dense_jacobian(), H)
    assert np.abs(full[:n] - S[:n]).max() < 1e-2                            # qp_equality.jl:120,122


def double_integrator(seed=0, T=5):
    """test/examples/double_integrator.jl:1-75 as a flat NLP: x_{t+1} = A x_t + B u_t, x_1 and x_T pinned, quadratic
    stage costs; theta = per-stage [vec A; B; diag Q; R (; x_init)] and terminal [diag QT; x_goal] (:24-32)."""
    import problems
    A = np.array([[1.0, 1.0], [0.0, 1.0]]); Bv = np.array([0.0, 1.0])
    Qd, R, QTd = np.array([1.0, 1.0]), 0.1, np.array([10.0, 10.0])
    x_init, x_goal = np.array([0.0, 0.0]), np.array([1.0, 0.0])
    nx, nu = 2, 1
    n, m = T * nx + (T - 1) * nu, (T - 1) * nx + 2 * nx
    ix = [t * 3 for t in range(T)]
    iu = [t * 3 + 2 for t in range(T - 1)]
    Gm = np.zeros((m, n)); g0 = np.zeros(m)
    for t in range(T - 1):
        r = 2 * t
        Gm[r:r + 2, ix[t + 1]:ix[t + 1] + 2] = np.eye(2)
        Gm[r:r + 2, ix[t]:ix[t] + 2] = -A
        Gm[r:r + 2, iu[t]] = -Bv
    r = 2 * (T - 1)
    Gm[r:r + 2, ix[0]:ix[0] + 2] = np.eye(2); g0[r:r + 2] = -x_init
    Gm[r + 2:r + 4, ix[T - 1]:ix[T - 1] + 2] = np.eye(2); g0[r + 2:r + 4] = -x_goal
    wd = np.zeros(n)
    for t in range(T - 1):
        wd[ix[t]:ix[t] + 2] = Qd
        wd[iu[t]] = R
    wd[ix[T - 1]:ix[T - 1] + 2] = QTd
    rng = np.random.default_rng(seed)
    x0 = np.zeros(n)
    for t in range(T):
        x0[ix[t]:ix[t] + 2] = x_init + (x_goal - x_init) * t / (T - 1)
    for t in range(T - 1):
        x0[iu[t]] = rng.standard_normal()
    z = problems._z
    P = problems.DenseNLP(
        "double_integrator", n, m, 0, 0, np.zeros(0, np.int32),
        f=lambda v: float(0.5 * v @ (wd * v)), grad=lambda v: wd * v, hess=lambda v: np.diag(wd),
        g=lambda v: Gm @ v + g0, jac_g=lambda v: Gm, hess_gy=lambda v, y: z(n, n),
        h=lambda v: z(0), jac_h=lambda v: z(0, n), hess_hz=lambda v, zz: z(n, n), x0=x0)
    return P, Gm, wd, ix, iu


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_double_integrator_sensitivities(backend):
    """test/examples/double_integrator.jl:64-165: solve with the example's tight tolerances, then the sensitivities of the
    primal variables agree with -L_zz^-1 L_z,theta (z = [x; y]) to 1e-3 (:164)."""
    from calipso_b200.solver import Options, Solver, differentiate, initialize, residual_jacobian_parameters, solve
    T = 5
    P, Gm, wd, ix, iu = double_integrator(0, T)
    n, m = P.n, P.m
    opt = Options(residual_tolerance=1.0e-12, equality_tolerance=1.0e-8, complementarity_tolerance=1.0e-8)
    s = Solver(P, P.callback, options=opt, binding=backends.binding(backend))
    initialize(s, P.x0)
    assert solve(s) is True
    v, y = s.solution[:n], s.solution[s.kkt.iy]
    R_ = s.residual
    assert max(np.abs(R_[:n]).max(), 0.0) < opt.optimality_tolerance               # :89-94
    assert np.abs(R_[s.kkt.iy]).max() < opt.slack_tolerance                          # :96-100
    assert np.abs(Gm @ v + P.g(np.zeros(n))).max() <= opt.equality_tolerance         # :102
    sizes = [11] + [9] * (T - 2) + [4]
    off = np.concatenate([[0], np.cumsum(sizes)])
    nth = int(off[-1])
    Oxp = np.zeros((n, nth)); Exp = np.zeros((n, nth)); Ep = np.zeros((m, nth))
    for t in range(T - 1):
        o_, lam = off[t], y[2 * t:2 * t + 2]
        Oxp[ix[t], o_ + 6] = v[ix[t]]; Oxp[ix[t] + 1, o_ + 7] = v[ix[t] + 1]       # d(Q x)/d diag Q
        Oxp[iu[t], o_ + 8] = v[iu[t]]                                               # d(R u)/dR
        for j in range(2):
            for i in range(2):
                Exp[ix[t] + j, o_ + i + 2 * j] = -lam[i]                            # d(-A'lambda)_j / dA_ij
                Ep[2 * t + i, o_ + i + 2 * j] = -v[ix[t] + j]                       # d(y - A x - B u)_i / dA_ij
        for i in range(2):
            Exp[iu[t], o_ + 4 + i] = -lam[i]                                        # d(-B'lambda) / dB_i
            Ep[2 * t + i, o_ + 4 + i] = -v[iu[t]]
    Ep[2 * (T - 1):2 * (T - 1) + 2, off[0] + 9:off[0] + 11] = -np.eye(2)            # x_1 - x_init
    oT = off[T - 1]
    Oxp[ix[T - 1], oT] = v[ix[T - 1]]; Oxp[ix[T - 1] + 1, oT + 1] = v[ix[T - 1] + 1]
    Ep[2 * (T - 1) + 2:2 * (T - 1) + 4, oT + 2:oT + 4] = -np.eye(2)                 # x_T - x_goal
    H = residual_jacobian_parameters(n, m, 0, Oxp, Exp, None, Ep, None)
    S = differentiate(s, H)
    Lzz = np.block([[np.diag(wd), Gm.T], [Gm, np.zeros((m, m))]])
    closed = -np.linalg.solve(Lzz, np.vstack([Oxp + Exp, Ep]))
    assert np.abs(closed[:n] - S[:n]).max() < 1.0e-3                                 # :164
    # finite-difference spot check of the closed form itself: perturb x_goal[0]
    col = oT + 2
    d = 1e-6
    g_shift = np.zeros(m); g_shift[2 * (T - 1) + 2] = -d
    KKT = Lzz
    rhs0 = -np.concatenate([np.zeros(n), P.g(np.zeros(n))])
    z0 = np.linalg.solve(KKT, rhs0)
    z1 = np.linalg.solve(KKT, rhs0 - np.concatenate([np.zeros(n), g_shift]))
    assert np.abs((z1 - z0)[:n] / d - closed[:n, col]).max() < 1e-5
