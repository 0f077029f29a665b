"""Shape coverage for the supernodal LDL' paths (shared-memory supernodes with 1..n staging chunks, panels taller than
the CTA, widths that are not multiples of 8, one- and two-part panel streaming in the solves, roots without rows):
random quasi-definite matrices of varying size / density through the LinearSolver seam, checked against a direct
solve and against the QDLDL identity P K P' = L D L' (SURVEY.md Appendix B)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import backends
from calipso_b200.solver import LDLSolver
from test_oracle_qdldl import quasidefinite

CASES = [
    (30, 20, 0.5, 11),      # nearly dense: a few wide supernodes
    (90, 60, 0.25, 12),     # dense-ish, panels taller than 84 rows
    (200, 150, 0.08, 13),   # root supernodes taller than the CTA (generic panel path)
    (400, 300, 0.01, 14),   # sparse: many small supernodes
    (64, 0, 0.3, 15),       # positive definite, no second block
    (1, 40, 0.5, 16),       # bordered
]


def banded_kkt(T, nx, nu, seed):
    """Trajectory-optimisation shaped KKT (block tridiagonal with multipliers), widths not multiples of 8."""
    rng = np.random.default_rng(seed)
    nz = nx + nu
    n = T * nz
    m = (T - 1) * nx
    W = sp.block_diag([sp.csc_matrix((lambda M: M @ M.T + np.eye(nz))(rng.standard_normal((nz, nz)))) for _ in range(T)])
    G = sp.lil_matrix((m, n))
    for t in range(T - 1):
        G[t * nx:(t + 1) * nx, t * nz:(t + 1) * nz] = rng.standard_normal((nx, nz))
        G[t * nx:(t + 1) * nx, (t + 1) * nz:(t + 1) * nz + nx] = -np.eye(nx)
    return sp.bmat([[W, G.T], [G, -1e-3 * sp.eye(m)]]).tocsc()


def check(K, backend, batch=1):
    N = K.shape[0]
    s = LDLSolver(K, batch=batch, binding=backends.binding(backend))
    s.factorize(K)
    rng = np.random.default_rng(3)
    b = rng.standard_normal((batch, N))
    x = np.zeros_like(b)
    s.linear_solve(x, K, b)
    lu = spla.splu(sp.csc_matrix(K))
    for i in range(batch):
        xr = lu.solve(b[i])
        assert np.abs(x[i] - xr).max() <= 1e-8 * max(1.0, np.abs(xr).max())
    Lp, Li, Lx, D = s.factor(0)
    perm = s.symbolic()[0]
    L = sp.csc_matrix((Lx, Li, Lp), shape=(N, N)) + sp.eye(N)
    E = (L @ sp.diags(D) @ L.T - K[perm][:, perm]).tocoo()
    assert np.abs(E.data).max(initial=0.0) <= 1e-9 * max(1.0, np.abs(K.data).max())
    ev_pos = int((D > 0).sum())
    assert tuple(s.compute_inertia()[0]) == (ev_pos, N - ev_pos, 0)


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("n1,n2,density,seed", CASES)
def test_random_quasidefinite_shapes(backend, n1, n2, density, seed):
    check(quasidefinite(n1, n2, density, seed), backend)


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("T,nx,nu", [(6, 5, 3), (12, 10, 4), (5, 21, 7), (4, 40, 13)])
def test_trajopt_shaped_chains(backend, T, nx, nu):
    check(banded_kkt(T, nx, nu, seed=T * 100 + nx), backend, batch=2)


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("T,n_x,n_u,n_soc", [(7, 5, 6, 4), (4, 20, 9, 6), (9, 3, 3, 0), (3, 30, 12, 9)])
def test_lq_conic_shapes_solve_like_the_oracle(backend, T, n_x, n_u, n_soc):
    """Whole solve! on other LQ-conic shapes (stage widths 11 ... 42, with and without second-order cones)."""
    from calipso_b200 import lqc
    from calipso_b200.solver import BatchKKT
    from oracle import oracle as orc
    P = lqc.lqc(T, n_x, n_u, n_soc, seed=T * 1000 + n_x)
    k = BatchKKT(P, binding=backends.binding(backend))
    perm, _, _ = k.symbolic()
    k.load_lq(P)
    k.initialize(P.x0)
    k.lq_begin()
    r = k.lq_solve(max_steps=300, check_every=2)
    o = orc.from_problem(P, perm=perm)
    o.use_superlu_fallback()
    o.initialize(P.x0)
    rc = o.solve()
    assert (r["converged"] == 1) == (rc == 1)
    st = {kk: int(v[0]) for kk, v in k.stats().items()}
    assert st["total_iterations"] == o.stats["total_iterations"]
    assert np.abs(k.get("POINT")[0] - o.solution).max() <= 1e-6 * max(1.0, np.abs(o.solution).max())


def test_shared_memory_plan_follows_the_pattern():
    """cb200_path_info: the shared-memory budget is chosen per pattern (three, two or one resident CTA per SM) so that the
    solves keep x[N] in shared memory and every CTA-scope supernode is staged (host emulation: the plan is host-side)."""
    from calipso_b200 import lqc
    from calipso_b200.solver import BatchKKT
    expect = {"tiny": 3, "cfg2": 3, "cfg3": 3, "quadruped_shape": 3}      # (quadruped: only with the singleton leaves kept out of shared memory)
    for name, ctas in expect.items():
        k = BatchKKT(getattr(lqc, name)(), binding=backends.binding("emul"))
        p = k.paths()
        assert p["solve_in_shared_memory"] == 1 and p["cta_supernodes_generic"] == 0, (name, p)
        assert p["ctas_per_sm"] == ctas, (name, p)
        assert p["dynamic_smem_bytes"] <= {3: 74000, 2: 111616, 1: 224000}[ctas]
