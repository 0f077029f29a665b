"""Seeded synthetic LQ-conic trajectory-optimization instances ``LQC(T, n_x, n_u, n_soc, seed)``.

This is the input family the benchmark and the parity tests run on.  It is *not*
part of the reference: CALIPSO's own trajectory-optimization front end (``src/trajectory_optimization``)
produces the same kind of flat NLP -- variables ``[x_1,u_1,...,x_{T-1},u_{T-1},x_T]``
(``dynamics.jl:333-340``), equalities ordered dynamics-then-stage (``data.jl:51-55``), cones ordered
all-nonnegative-then-all-second-order (``methods.jl:46-50``) -- from user Julia callbacks through Symbolics
code generation.  Here the callbacks are affine/quadratic so that every derivative is a constant sparse
matrix and ``evaluate!`` (``src/solver/evaluate.jl``) reduces to sparse mat-vecs:

    f(v)  = 1/2 v' Q v + q' v          grad f = Q v + q       hess f = Q          (Q given as upper triangle)
    g(v)  = G v + g0   (m rows)         dynamics x_{t+1} - (A x_t + B u_t), then x_1 - xhat
    h(v)  = C v + h0   (p rows)         box rows [u - u_min ; u_max - u] per stage, then SOC(3) friction-shaped rows

All index arrays are 0-based int32 CSC with sorted rows (the reference is 1-based Int64; the C ABI takes 0-based).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


@dataclass
class ConicProblem:
    """Flat NLP pattern + (for LQ problems) constant values.  Shared by the oracle and the B200 path."""

    n: int
    m: int
    p: int
    num_nonnegative: int                 # cone rows [0, num_nonnegative) are the nonnegative orthant
    soc_dims: np.ndarray                 # int32[n_soc]; SOC blocks follow contiguously after the nonnegative rows
    # CSC patterns (0-based, sorted rows).  W is the UPPER triangle (incl. diagonal) of the Lagrangian Hessian.
    W_colptr: np.ndarray
    W_rowval: np.ndarray
    G_colptr: np.ndarray
    G_rowval: np.ndarray
    C_colptr: np.ndarray
    C_rowval: np.ndarray
    # LQ values (None for general nonlinear problems, where callbacks supply values each iteration)
    W_val: np.ndarray | None = None
    G_val: np.ndarray | None = None
    C_val: np.ndarray | None = None
    q: np.ndarray | None = None
    g0: np.ndarray | None = None
    h0: np.ndarray | None = None
    x0: np.ndarray | None = None
    meta: dict = field(default_factory=dict)

    @property
    def N(self) -> int:
        return self.n + self.m + self.p

    @property
    def total(self) -> int:
        return self.n + 2 * self.m + 3 * self.p

    def soc_offsets(self) -> np.ndarray:
        """Start row (within the cone block) of each SOC."""
        off = self.num_nonnegative + np.concatenate([[0], np.cumsum(self.soc_dims)[:-1]]) if len(self.soc_dims) else np.zeros(0)
        return off.astype(np.int32)

    def W_full(self) -> sp.csc_matrix:
        U = sp.csc_matrix((self.W_val, self.W_rowval, self.W_colptr), shape=(self.n, self.n))
        return (U + sp.triu(U, 1).T).tocsc()

    def G(self) -> sp.csc_matrix:
        return sp.csc_matrix((self.G_val, self.G_rowval, self.G_colptr), shape=(self.m, self.n))

    def C(self) -> sp.csc_matrix:
        return sp.csc_matrix((self.C_val, self.C_rowval, self.C_colptr), shape=(self.p, self.n))


def _csc(a: sp.spmatrix):
    a = sp.csc_matrix(a)
    a.sort_indices()
    return a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(np.float64)


def lqc(T: int, n_x: int, n_u: int, n_soc: int, seed: int, *, h: float = 0.05, u_max: float = 10.0,
        mu: float = 0.5, c0: float = 10.0, q_scale: float = 0.3) -> ConicProblem:
    """Build one ``LQC(T, n_x, n_u, n_soc, seed)`` instance; PCG64(seed).

    Two constants differ from SURVEY.md section 8(d) (c0 = 1, q ~ N(0,1)): with those the *reference algorithm
    itself* (as restated by the oracle, including its UMFPACK-style fallback) aborts with "cone search failure" on
    cfg3 after 53 iterations, because the SOC(3) reduced blocks are far from symmetric off the central path
    (SURVEY.md section 3.3 quirk); with c0 = 5 one cfg3 seed in 16 still fails to converge in 600 iterations.  With
    c0 = 10 and q ~ 0.3 N(0,1) the reference converges on every seed tried (cfg3 seeds 0..15: 7-10 Newton iterations,
    at most one fallback solve) while still needing several refinement passes per Newton step.  See DESIGN.md.
    """
    assert n_u % 3 == 0 or n_soc == 0
    rng = np.random.Generator(np.random.PCG64(seed))
    nz = n_x + n_u
    n = T * n_x + (T - 1) * n_u
    xs = lambda t: t * nz                # start of x_t (0-based stage t)
    us = lambda t: t * nz + n_x          # start of u_t

    # --- dynamics (dense A, B shared across stages like a time-invariant model) + initial condition
    Abar = rng.standard_normal((n_x, n_x)) / np.sqrt(n_x)
    A = np.eye(n_x) + h * Abar
    B = h * rng.standard_normal((n_x, n_u))
    xhat = rng.standard_normal(n_x)
    m = T * n_x
    ts = np.arange(T - 1)
    rows, cols, vals = [], [], []

    def dense_blocks(block, row0, col0):
        """COO triplets of `block` placed at (row0[k], col0[k]) for every k."""
        nr, nc = block.shape
        ii, jj = np.meshgrid(np.arange(nr), np.arange(nc), indexing="ij")
        rows.append((row0[:, None, None] + ii[None]).ravel())
        cols.append((col0[:, None, None] + jj[None]).ravel())
        vals.append(np.broadcast_to(block, (len(row0), nr, nc)).ravel())

    dense_blocks(-A, ts * n_x, ts * nz)
    dense_blocks(-B, ts * n_x, ts * nz + n_x)
    eye_r = (ts[:, None] * n_x + np.arange(n_x)[None]).ravel()
    rows.append(eye_r); cols.append(((ts[:, None] + 1) * nz + np.arange(n_x)[None]).ravel()); vals.append(np.ones(len(eye_r)))
    rows.append((T - 1) * n_x + np.arange(n_x)); cols.append(np.arange(n_x)); vals.append(np.ones(n_x))
    G = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(m, n))
    g0 = np.zeros(m)
    g0[(T - 1) * n_x:] = -xhat

    # --- stage costs
    q = np.zeros(n)
    rows, cols, vals = [], [], []
    for t in range(T):
        d = nz if t < T - 1 else n_x
        M = rng.standard_normal((d, d))
        Wt = M.T @ M / nz + 0.1 * np.eye(d)
        ii, jj = np.meshgrid(np.arange(d), np.arange(d), indexing="ij")
        rows.append((xs(t) + ii).ravel()); cols.append((xs(t) + jj).ravel()); vals.append(Wt.ravel())
        q[xs(t):xs(t) + d] = q_scale * rng.standard_normal(d)
    Q = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))

    # --- cones: box rows first (nonnegative), then SOC(3) blocks ordered by (stage, triple)
    n_box = 2 * n_u * (T - 1)
    socs = sorted((k % (T - 1), (k // (T - 1)) % max(n_u // 3, 1)) for k in range(n_soc))
    p = n_box + 3 * n_soc
    rows, cols, vals = [], [], []
    h0 = np.zeros(p)
    ucol = (ts[:, None] * nz + n_x + np.arange(n_u)[None]).ravel()
    r_lo = (2 * n_u * ts[:, None] + np.arange(n_u)[None]).ravel()
    rows += [r_lo, r_lo + n_u]; cols += [ucol, ucol]; vals += [np.ones(len(ucol)), -np.ones(len(ucol))]
    h0[:n_box] = u_max
    for k, (t, j) in enumerate(socs):
        r = n_box + 3 * k
        a = us(t) + 3 * j
        rows.append(np.array([r, r + 1, r + 2])); cols.append(np.array([a, a + 1, a + 2])); vals.append(np.array([mu, 1.0, 1.0]))
        h0[r] = c0
    C = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(p, n))

    # --- initial guess: states interpolate xhat -> 0, tiny actions
    x0 = np.zeros(n)
    for t in range(T):
        x0[xs(t):xs(t) + n_x] = xhat * (1.0 - t / (T - 1))
    for t in range(T - 1):
        x0[us(t):us(t) + n_u] = 1.0e-3 * rng.standard_normal(n_u)

    Wp, Wi, Wv = _csc(sp.triu(sp.csc_matrix(Q)))
    Gp, Gi, Gv = _csc(G)
    Cp, Ci, Cv = _csc(C)
    return ConicProblem(
        n=n, m=m, p=p, num_nonnegative=n_box, soc_dims=np.full(n_soc, 3, dtype=np.int32),
        W_colptr=Wp, W_rowval=Wi, G_colptr=Gp, G_rowval=Gi, C_colptr=Cp, C_rowval=Ci,
        W_val=Wv, G_val=Gv, C_val=Cv, q=q, g0=g0, h0=h0, x0=x0,
        meta=dict(family="LQC", T=T, n_x=n_x, n_u=n_u, n_soc=n_soc, seed=seed),
    )


# BASELINE.json configs 2-4 (cfg2 N=2142, cfg3 N=4584, cfg4 = 64 seeds of cfg3)
def cfg2(instance: int = 0) -> ConicProblem:
    return lqc(50, 12, 6, 20, 1000 * 2 + instance)


def cfg2_hard(instance: int = 0) -> ConicProblem:
    """cfg2 with more active cones (c0 = 5): refinement fails on a few iterations, exercising the fallback solve."""
    return lqc(50, 12, 6, 20, 1000 * 2 + instance, c0=5.0)


def cfg3(instance: int = 0) -> ConicProblem:
    return lqc(40, 36, 12, 100, 1000 * 3 + instance)


def tiny(instance: int = 0, T: int = 5, n_x: int = 4, n_u: int = 3, n_soc: int = 3) -> ConicProblem:
    """Small member of the family for fast tests."""
    return lqc(T, n_x, n_u, n_soc, 9000 + instance)


def quadruped_shape(instance: int = 0) -> ConicProblem:
    """An LQC instance with the dimensions of the reference's contact-implicit quadruped example
    (test/examples/quadruped_gait.jl:236-244,460-464: reduced KKT dimension 6954, stage width 76): T = 36, n_x = 36,
    n_u = 39, 100 SOC(3) cones -> N = 6987, stages of 75 variables.  Larger than BASELINE's cfg3 in N and wider than the
    48-column supernode cap: exercises the shared-memory plans for two / one resident CTA per SM."""
    return lqc(36, 36, 39, 100, 1000 * 5 + instance)
