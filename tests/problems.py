"""Small NLPs from the reference's own test-suite, with hand-written analytic derivatives.

Test infrastructure shared by the oracle tests and the GPU parity tests.  Each problem mirrors one file under
/root/reference/test/solver/ (cited per function); the reference generates derivatives with Symbolics.jl, here they
are written out.  Patterns are dense (explicit zeros are harmless).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np


@dataclass
class DenseNLP:
    name: str
    n: int
    m: int
    p: int
    num_nonnegative: int
    soc_dims: np.ndarray
    f: Callable          # x -> float
    grad: Callable       # x -> (n,)
    hess: Callable       # x -> (n,n)
    g: Callable          # x -> (m,)
    jac_g: Callable      # x -> (m,n)
    hess_gy: Callable    # x,y -> (n,n)   sum_i y_i hess g_i
    h: Callable          # x -> (p,)
    jac_h: Callable      # x -> (p,n)
    hess_hz: Callable    # x,z -> (n,n)
    x0: np.ndarray
    x_star: np.ndarray | None = None

    # dense patterns (0-based CSC)
    def __post_init__(self):
        n, m, p = self.n, self.m, self.p
        iu = [(i, j) for j in range(n) for i in range(j + 1)]
        self.W_rowval = np.array([i for i, j in iu], dtype=np.int32)
        self.W_colptr = np.array([j * (j + 1) // 2 for j in range(n + 1)], dtype=np.int32)
        self._wi = (np.array([i for i, j in iu]), np.array([j for i, j in iu]))
        self.G_colptr = np.arange(0, m * n + 1, m if m else 1, dtype=np.int32) if m else np.zeros(n + 1, np.int32)
        self.G_rowval = np.tile(np.arange(m, dtype=np.int32), n)
        self.C_colptr = np.arange(0, p * n + 1, p if p else 1, dtype=np.int32) if p else np.zeros(n + 1, np.int32)
        self.C_rowval = np.tile(np.arange(p, dtype=np.int32), n)
        self.W_val = None

    def callback(self, flags, x, y, z, out):
        """evaluate! (src/solver/evaluate.jl) for the flag bits of oracle.h / include/calipso_b200.h."""
        if flags & 1:
            out.objective[0] = self.f(x)
        if flags & 2:
            out.gradient[:] = self.grad(x)
        if flags & 4 and self.m:
            out.equality[:] = self.g(x)
        if flags & 8 and self.p:
            out.cone[:] = self.h(x)
        if flags & 16:
            out.eq_dual_grad[:] = self.jac_g(x).T @ y if self.m else 0.0
        if flags & 32:
            out.cone_dual_grad[:] = self.jac_h(x).T @ z if self.p else 0.0
        if flags & 64:
            H = self.hess(x)
            if self.m:
                H = H + self.hess_gy(x, y)
            if self.p:
                H = H + self.hess_hz(x, z)
            out.W_val[:] = H[self._wi]
        if flags & 128 and self.m:
            out.G_val[:] = self.jac_g(x).flatten(order="F")
        if flags & 256 and self.p:
            out.C_val[:] = self.jac_h(x).flatten(order="F")


def _z(*shape):
    return np.zeros(shape)


def wachter() -> DenseNLP:
    """test/solver/wachter.jl:1-48 (also README.md:101-118, BASELINE cfg0): x* = [1, 0, 0.5]."""
    def hess_gy(x, y):
        H = _z(3, 3)
        H[0, 0] = 2.0 * y[0]
        return H
    return DenseNLP(
        "wachter", 3, 2, 2, 2, np.zeros(0, np.int32),
        f=lambda x: x[0], grad=lambda x: np.array([1.0, 0.0, 0.0]), hess=lambda x: _z(3, 3),
        g=lambda x: np.array([x[0] ** 2 - x[1] - 1.0, x[0] - x[2] - 0.5]),
        jac_g=lambda x: np.array([[2 * x[0], -1.0, 0.0], [1.0, 0.0, -1.0]]), hess_gy=hess_gy,
        h=lambda x: x[1:3].copy(), jac_h=lambda x: np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]),
        hess_hz=lambda x, z: _z(3, 3), x0=np.array([-2.0, 3.0, 1.0]), x_star=np.array([1.0, 0.0, 0.5]))


def maratos() -> DenseNLP:
    """test/solver/maratos.jl:1-31."""
    return DenseNLP(
        "maratos", 2, 1, 0, 0, np.zeros(0, np.int32),
        f=lambda x: 2.0 * (x[0] ** 2 + x[1] ** 2 - 1.0) - x[0],
        grad=lambda x: np.array([4 * x[0] - 1.0, 4 * x[1]]), hess=lambda x: 4.0 * np.eye(2),
        g=lambda x: np.array([x[0] ** 2 + x[1] ** 2 - 1.0]), jac_g=lambda x: np.array([[2 * x[0], 2 * x[1]]]),
        hess_gy=lambda x, y: 2.0 * y[0] * np.eye(2),
        h=lambda x: _z(0), jac_h=lambda x: _z(0, 2), hess_hz=lambda x, z: _z(2, 2), x0=np.array([2.0, 1.0]))


def knitro() -> DenseNLP:
    """test/solver/knitro.jl:1-45 (complementarity constraints as bilinear equalities)."""
    def g(x):
        return np.array([2 * (x[1] - 1) - 1.5 * x[1] + x[2] - 0.5 * x[3] + x[4],
                         3 * x[0] - x[1] - 3.0 - x[5], -x[0] + 0.5 * x[1] + 4.0 - x[6], -x[0] - x[1] + 7.0 - x[7],
                         x[2] * x[5], x[3] * x[6], x[4] * x[7]])

    def jac_g(x):
        J = _z(7, 8)
        J[0, [1, 2, 3, 4]] = [0.5, 1.0, -0.5, 1.0]
        J[1, [0, 1, 5]] = [3.0, -1.0, -1.0]
        J[2, [0, 1, 6]] = [-1.0, 0.5, -1.0]
        J[3, [0, 1, 7]] = [-1.0, -1.0, -1.0]
        J[4, 2], J[4, 5] = x[5], x[2]
        J[5, 3], J[5, 6] = x[6], x[3]
        J[6, 4], J[6, 7] = x[7], x[4]
        return J

    def hess_gy(x, y):
        H = _z(8, 8)
        for k, (a, b) in enumerate([(2, 5), (3, 6), (4, 7)]):
            H[a, b] += y[4 + k]
            H[b, a] += y[4 + k]
        return H
    Hf = _z(8, 8)
    Hf[0, 0], Hf[1, 1] = 2.0, 8.0
    return DenseNLP(
        "knitro", 8, 7, 8, 8, np.zeros(0, np.int32),
        f=lambda x: (x[0] - 5) ** 2 + (2 * x[1] + 1) ** 2,
        grad=lambda x: np.concatenate([[2 * (x[0] - 5), 4 * (2 * x[1] + 1)], np.zeros(6)]), hess=lambda x: Hf,
        g=g, jac_g=jac_g, hess_gy=hess_gy, h=lambda x: x.copy(), jac_h=lambda x: np.eye(8),
        hess_hz=lambda x, z: _z(8, 8), x0=np.zeros(8))


def friction(v, mu, gamma, x0) -> DenseNLP:
    """test/solver/friction_cone.jl:1-66: min v'x s.t. x[1] = mu*gamma, x in SOC(3)."""
    v = np.asarray(v, float)
    return DenseNLP(
        f"friction_{v}_{mu}_{gamma}", 3, 1, 3, 0, np.array([3], np.int32),
        f=lambda x: float(v @ x), grad=lambda x: v.copy(), hess=lambda x: _z(3, 3),
        g=lambda x: np.array([x[0] - mu * gamma]), jac_g=lambda x: np.array([[1.0, 0.0, 0.0]]),
        hess_gy=lambda x, y: _z(3, 3), h=lambda x: x.copy(), jac_h=lambda x: np.eye(3),
        hess_hz=lambda x, z: _z(3, 3), x0=np.asarray(x0, float))


def portfolio(seed=0, pdim=10) -> DenseNLP:
    """test/solver/portfolio.jl:1-63: 2 nonnegative rows + one SOC(12)."""
    rng = np.random.default_rng(seed)
    E = rng.standard_normal((pdim, pdim))
    Sigma = E.T @ E
    w, V = np.linalg.eigh(Sigma)
    Shalf = (V * np.sqrt(w)) @ V.T
    c = np.concatenate([np.zeros(pdim), [1.0]])
    G1 = np.block([[2.0 * Shalf, np.zeros((pdim, 1))], [np.zeros((1, pdim)), -np.ones((1, 1))]])
    hh = np.concatenate([np.zeros(pdim), [1.0]])
    q = np.concatenate([np.zeros(pdim), [1.0]])
    G2 = np.concatenate([np.ones(pdim), [0.0]])[None]
    G3 = np.concatenate([-np.ones(pdim), [0.0]])[None]
    A = np.vstack([G2, G3, -q[None], -G1])
    b = np.concatenate([[1.0, -1.0, 1.0], hh])
    n = pdim + 1
    P = DenseNLP(
        "portfolio", n, 0, 2 + pdim + 2, 2, np.array([pdim + 2], np.int32),
        f=lambda x: float(c @ x), grad=lambda x: c.copy(), hess=lambda x: _z(n, n),
        g=lambda x: _z(0), jac_g=lambda x: _z(0, n), hess_gy=lambda x, y: _z(n, n),
        h=lambda x: b - A @ x, jac_h=lambda x: -A, hess_hz=lambda x, z: _z(n, n), x0=rng.standard_normal(n))
    P.A, P.b = A, b
    return P


def random_qp(seed=0, n=10, m=5, p=5) -> DenseNLP:
    """generate_random_qp of test/solver/problem.jl:2-22 (objective z'Pz + q'z, A z = b, h - G z >= 0)."""
    rng = np.random.default_rng(seed)
    Pm = rng.standard_normal((n, n))
    Pm = Pm.T @ Pm
    q = rng.standard_normal(n)
    Gm = rng.standard_normal((p, n))
    x = rng.standard_normal(n)
    hv = Gm @ x + rng.random(p)
    A = rng.standard_normal((m, n))
    b = A @ x
    return DenseNLP(
        "random_qp", n, m, p, p, np.zeros(0, np.int32),
        f=lambda z: float(z @ Pm @ z + q @ z), grad=lambda z: 2 * Pm @ z + q, hess=lambda z: 2 * Pm,
        g=lambda z: A @ z - b, jac_g=lambda z: A, hess_gy=lambda z, y: _z(n, n),
        h=lambda z: hv - Gm @ z, jac_h=lambda z: -Gm, hess_hz=lambda z, y: _z(n, n), x0=rng.standard_normal(n))


def qp_nonnegative(seed=0, n=10, m=5) -> DenseNLP:
    """test/solver/qp_nonnegative.jl:1-62: 1/2 x'P x + p'x with P diagonal, A x = b, x >= 0 (seeded draws)."""
    rng = np.random.default_rng(seed)
    xh = np.maximum(0.0, rng.standard_normal(n))
    Q = rng.random((n, n))
    Pd = np.diag(Q.T @ Q).copy()
    pv = rng.standard_normal(n)
    A = rng.random((m, n))
    b = A @ xh
    P = DenseNLP(
        "qp_nonnegative", n, m, n, n, np.zeros(0, np.int32),
        f=lambda x: float(0.5 * x @ (Pd * x) + pv @ x), grad=lambda x: Pd * x + pv, hess=lambda x: np.diag(Pd),
        g=lambda x: A @ x - b, jac_g=lambda x: A, hess_gy=lambda x, y: _z(n, n),
        h=lambda x: x.copy(), jac_h=lambda x: np.eye(n), hess_hz=lambda x, z: _z(n, n), x0=rng.standard_normal(n))
    P.A, P.b = A, b
    return P


def test1() -> DenseNLP:
    """test/solver/test1.jl:1-35: 50 variables, 30 quadratic equalities, 3 inequalities."""
    n = 50

    def jac_g(x):
        J = _z(30, n)
        J[np.arange(30), np.arange(30)] = 2 * x[:30]
        return J

    def hess_gy(x, y):
        H = _z(n, n)
        H[np.arange(30), np.arange(30)] = 2 * y
        return H
    Jh = _z(3, n)
    Jh[0, 0], Jh[1, 1], Jh[2, 4] = 1.0, 1.0, -1.0
    return DenseNLP(
        "test1", n, 30, 3, 3, np.zeros(0, np.int32),
        f=lambda x: float(x @ x), grad=lambda x: 2 * x, hess=lambda x: 2 * np.eye(n),
        g=lambda x: x[:30] ** 2 - 1.2, jac_g=jac_g, hess_gy=hess_gy,
        h=lambda x: np.array([x[0] + 10.0, x[1] + 5.0, 20.0 - x[4]]), jac_h=lambda x: Jh,
        hess_hz=lambda x, z: _z(n, n), x0=np.ones(n))


def test2(seed=0) -> DenseNLP:
    """test/solver/test2.jl:1-32: bilinear objective, one quadratic and one linear inequality (x0 = rand(2), seeded)."""
    rng = np.random.default_rng(seed)
    return DenseNLP(
        "test2", 2, 0, 2, 2, np.zeros(0, np.int32),
        f=lambda x: -x[0] * x[1] + 2.0 / (3.0 * np.sqrt(3.0)), grad=lambda x: np.array([-x[1], -x[0]]),
        hess=lambda x: np.array([[0.0, -1.0], [-1.0, 0.0]]), g=lambda x: _z(0), jac_g=lambda x: _z(0, 2),
        hess_gy=lambda x, y: _z(2, 2),
        h=lambda x: np.array([-x[0] - x[1] ** 2 + 1.0, x[0] + x[1]]),
        jac_h=lambda x: np.array([[-1.0, -2.0 * x[1]], [1.0, 1.0]]),
        hess_hz=lambda x, z: z[0] * np.array([[0.0, 0.0], [0.0, -2.0]]), x0=rng.random(2))


def test3(seed=0) -> DenseNLP:
    """test/solver/test3.jl:1-32: Rosenbrock objective with a cubic and a linear inequality (x0 = rand(2), seeded)."""
    rng = np.random.default_rng(seed)
    return DenseNLP(
        "test3", 2, 0, 2, 2, np.zeros(0, np.int32),
        f=lambda x: 100.0 * (x[1] - x[0] ** 2) ** 2 + (1.0 - x[0]) ** 2,
        grad=lambda x: np.array([-400.0 * x[0] * (x[1] - x[0] ** 2) - 2.0 * (1.0 - x[0]), 200.0 * (x[1] - x[0] ** 2)]),
        hess=lambda x: np.array([[1200.0 * x[0] ** 2 - 400.0 * x[1] + 2.0, -400.0 * x[0]], [-400.0 * x[0], 200.0]]),
        g=lambda x: _z(0), jac_g=lambda x: _z(0, 2), hess_gy=lambda x, y: _z(2, 2),
        h=lambda x: np.array([-(x[0] - 1.0) ** 3 + x[1] - 1.0, -x[0] - x[1] + 2.0]),
        jac_h=lambda x: np.array([[-3.0 * (x[0] - 1.0) ** 2, 1.0], [-1.0, -1.0]]),
        hess_hz=lambda x, z: z[0] * np.array([[-6.0 * (x[0] - 1.0), 0.0], [0.0, 0.0]]), x0=rng.random(2))


def test4(seed=0) -> DenseNLP:
    """test/solver/test4.jl:1-33: linear objective on the unit ball."""
    rng = np.random.default_rng(seed)
    return DenseNLP(
        "test4", 3, 0, 1, 1, np.zeros(0, np.int32),
        f=lambda x: x[0] - 2.0 * x[1] + x[2] + np.sqrt(6.0), grad=lambda x: np.array([1.0, -2.0, 1.0]),
        hess=lambda x: _z(3, 3), g=lambda x: _z(0), jac_g=lambda x: _z(0, 3), hess_gy=lambda x, y: _z(3, 3),
        h=lambda x: np.array([1 - x @ x]), jac_h=lambda x: (-2 * x)[None], hess_hz=lambda x, z: -2 * z[0] * np.eye(3),
        x0=rng.random(3))


def pendulum(seed: int = 0, horizon: int = 11, overwrite: bool = False) -> DenseNLP:
    """README.md:129-176 / test/examples/pendulum.jl (BASELINE cfg1): pendulum swing-up, implicit-midpoint dynamics,
    variables [x_1, u_1, ..., x_{T-1}, u_{T-1}, x_T] (trajectory_optimization/dynamics.jl:333-340), equalities ordered
    dynamics, then stage constraints (data.jl:51-55): n = 32, m = 24, p = 0 for T = 11.  The action guess is seeded
    (the README draws it with randn).  Derivatives written out; second derivatives are summed where stage sparsities
    overlap (what the reference's own Hessian test does, hessian_lagrangian.jl:296-302; SURVEY.md Appendix A.17 notes
    that the reference's evaluate! overwrites instead -- that happens before the hot-path boundary; overwrite=True builds
    the equality-dual Hessian that way, i.e. the matrices the reference's solve! actually iterates with)."""
    T, h = horizon, 0.05
    ml2, grav_l, damp = 0.25, 9.81 / 0.5, 0.1 / 0.25
    nx, nu = 2, 1
    n = T * nx + (T - 1) * nu
    m = (T - 1) * nx + 2 * nx
    ix = [t * (nx + nu) for t in range(T)]                 # offset of x_t
    iu = [t * (nx + nu) + nx for t in range(T - 1)]        # offset of u_t
    x_init, x_goal = np.array([0.0, 0.0]), np.array([np.pi, 0.0])

    def f(v):
        c = 0.0
        for t in range(T):
            c += 0.1 * v[ix[t]:ix[t] + 2] @ v[ix[t]:ix[t] + 2]
        for t in range(T - 1):
            c += 0.1 * v[iu[t]] ** 2
        return c

    def grad(v):
        return 0.2 * v

    def hess(v):
        return 0.2 * np.eye(n)

    def g(v):
        out = np.zeros(m)
        for t in range(T - 1):
            x, y, u = v[ix[t]:ix[t] + 2], v[ix[t + 1]:ix[t + 1] + 2], v[iu[t]]
            xm = 0.5 * (x + y)
            fc = np.array([xm[1], u / ml2 - grav_l * np.sin(xm[0]) - damp * xm[1]])
            out[2 * t:2 * t + 2] = y - (x + h * fc)
        out[2 * (T - 1):2 * (T - 1) + 2] = v[ix[0]:ix[0] + 2] - x_init
        out[2 * (T - 1) + 2:] = v[ix[T - 1]:ix[T - 1] + 2] - x_goal
        return out

    def jac_g(v):
        J = np.zeros((m, n))
        for t in range(T - 1):
            x, y = v[ix[t]:ix[t] + 2], v[ix[t + 1]:ix[t + 1] + 2]
            c = np.cos(0.5 * (x[0] + y[0]))
            r = 2 * t
            # d1 = y1 - x1 - h * 0.5 (x2 + y2)
            J[r, ix[t]] = -1.0; J[r, ix[t] + 1] = -0.5 * h
            J[r, ix[t + 1]] = 1.0; J[r, ix[t + 1] + 1] = -0.5 * h
            # d2 = y2 - x2 - h (u / ml2 - grav_l sin(xm1) - damp xm2)
            J[r + 1, ix[t]] = 0.5 * h * grav_l * c; J[r + 1, ix[t + 1]] = 0.5 * h * grav_l * c
            J[r + 1, ix[t] + 1] = -1.0 + 0.5 * h * damp; J[r + 1, ix[t + 1] + 1] = 1.0 + 0.5 * h * damp
            J[r + 1, iu[t]] = -h / ml2
        r = 2 * (T - 1)
        J[r, ix[0]] = 1.0; J[r + 1, ix[0] + 1] = 1.0
        J[r + 2, ix[T - 1]] = 1.0; J[r + 3, ix[T - 1] + 1] = 1.0
        return J

    def hess_gy(v, y_):
        H = np.zeros((n, n))
        for t in range(T - 1):
            x, y = v[ix[t]:ix[t] + 2], v[ix[t + 1]:ix[t + 1] + 2]
            d2 = -0.25 * h * grav_l * np.sin(0.5 * (x[0] + y[0])) * y_[2 * t + 1]     # d^2 d2 / d(x1|y1)^2
            for a in (ix[t], ix[t + 1]):
                for b in (ix[t], ix[t + 1]):
                    if overwrite:
                        H[a, b] = d2      # evaluate!'s `=` scatter (src/solver/evaluate.jl:75-77): the later stage wins
                    else:
                        H[a, b] += d2
        return H

    rng = np.random.default_rng(seed)
    x0 = np.zeros(n)
    for t in range(T):
        x0[ix[t]:ix[t] + 2] = x_init + (x_goal - x_init) * t / (T - 1)       # linear_interpolation, utilities.jl:10
    for t in range(T - 1):
        x0[iu[t]] = rng.standard_normal()
    P = DenseNLP("pendulum", n, m, 0, 0, np.zeros(0, dtype=np.int32), f, grad, hess, g, jac_g, hess_gy,
                 lambda v: np.zeros(0), lambda v: np.zeros((0, n)), lambda v, z: np.zeros((n, n)), x0)
    P.x_goal, P.ix = x_goal, ix
    return P


def pendulum_overwrite() -> DenseNLP:
    """cfg1 with the reference's actual (last-write-wins) second-derivative scatter, SURVEY.md Appendix A.17."""
    P = pendulum(overwrite=True)
    P.name = "pendulum_overwrite"
    return P


def rocket_landing(seed: int = 0, horizon: int = 101):
    """test/examples/rocket_landing.jl:1-92 as a flat LQ-conic problem (every callback of this example is affine or
    quadratic): states (position, velocity) in R^6, thrust in R^3, implicit-midpoint dynamics of
    [v; (0, 0, -9.81) + f / mass] with h = 0.05 (:12-30), x_1 = (3, 2, 1, 0, 0, 0), x_T = 0 (:35-36, 45-49), cost
    |p|^2 + 0.1 |v|^2 + 0.1 |u|^2 per stage (:39-42), thrust cone (u_3; u_1; u_2) in SOC(3) at every stage (:51-62).
    Variables [x_1, u_1, ..., x_T] (trajectory_optimization/dynamics.jl:333-340), equalities ordered dynamics then
    stage constraints (data.jl:51-55).  The action guess 1e-3 randn (:69) is seeded."""
    import scipy.sparse as sp
    from calipso_b200.lqc import ConicProblem, _csc
    T, nx, nu, h, grav = horizon, 6, 3, 0.05, -9.81
    nz = nx + nu
    n = T * nx + (T - 1) * nu
    m = (T - 1) * nx + 2 * nx
    p = 3 * (T - 1)
    x_init, x_goal = np.array([3.0, 2.0, 1.0, 0.0, 0.0, 0.0]), np.zeros(6)
    G = sp.lil_matrix((m, n))
    g0 = np.zeros(m)
    I3 = np.eye(3)
    for t in range(T - 1):
        r, cx, cu, cy = t * nx, t * nz, t * nz + nx, (t + 1) * nz
        # y_p - x_p - h/2 (x_v + y_v)
        G[r:r + 3, cy:cy + 3] = I3; G[r:r + 3, cx:cx + 3] = -I3
        G[r:r + 3, cx + 3:cx + 6] = -0.5 * h * I3; G[r:r + 3, cy + 3:cy + 6] = -0.5 * h * I3
        # y_v - x_v - h ((0, 0, g) + u)
        G[r + 3:r + 6, cy + 3:cy + 6] = I3; G[r + 3:r + 6, cx + 3:cx + 6] = -I3
        G[r + 3:r + 6, cu:cu + 3] = -h * I3
        g0[r + 5] = -h * grav
    r = (T - 1) * nx
    G[r:r + 6, 0:6] = np.eye(6); g0[r:r + 6] = -x_init
    G[r + 6:r + 12, (T - 1) * nz:(T - 1) * nz + 6] = np.eye(6); g0[r + 6:r + 12] = -x_goal
    qd = np.zeros(n)
    for t in range(T):
        qd[t * nz:t * nz + 6] = 2.0 * np.array([1.0, 1.0, 1.0, 0.1, 0.1, 0.1])
        if t < T - 1:
            qd[t * nz + 6:t * nz + 9] = 0.2
    C = sp.lil_matrix((p, n))
    for t in range(T - 1):
        cu = t * nz + nx
        C[3 * t, cu + 2] = 1.0; C[3 * t + 1, cu] = 1.0; C[3 * t + 2, cu + 1] = 1.0
    rng = np.random.default_rng(seed)
    x0 = np.zeros(n)
    for t in range(T):
        x0[t * nz:t * nz + 6] = x_init + (x_goal - x_init) * t / (T - 1)      # linear_interpolation, utilities.jl:10
    for t in range(T - 1):
        x0[t * nz + 6:t * nz + 9] = 1.0e-3 * rng.standard_normal(3)
    Wp, Wi, Wv = _csc(sp.diags(qd).tocsc())
    Gp, Gi, Gv = _csc(G.tocsc())
    Cp, Ci, Cv = _csc(C.tocsc())
    return ConicProblem(n=n, m=m, p=p, num_nonnegative=0, soc_dims=np.full(T - 1, 3, dtype=np.int32),
                        W_colptr=Wp, W_rowval=Wi, G_colptr=Gp, G_rowval=Gi, C_colptr=Cp, C_rowval=Ci,
                        W_val=Wv, G_val=Gv, C_val=Cv, q=np.zeros(n), g0=g0, h0=np.zeros(p), x0=x0,
                        meta=dict(family="rocket_landing", T=T, n_x=nx, n_u=nu, n_soc=T - 1, seed=seed))
