"""LinearSolver seam (ldl_solver / factorize! / compute_inertia! / linear_solve!, src/solver/linear_solver.jl) and the
QDLDL integer contract: product vs oracle.

Integer work is compared bit-exactly by feeding the product's elimination order to the oracle through
qdldl(A; perm=p) (qdldl.jl:134-136); floating point (L, D, solutions) to 1e-10 relative.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import backends
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT, LDLSolver
from oracle import oracle as orc
from test_oracle_qdldl import quasidefinite


def kkt_like(seed, n1=60, n2=40, density=0.08):
    return quasidefinite(n1, n2, density, seed)


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("case", [(1, 0, 1.0, 0), (6, 4, 0.6, 1), (60, 40, 0.08, 2), (300, 200, 0.01, 3)])
def test_factor_matches_qdldl(backend, case):
    n1, n2, density, seed = case
    K = quasidefinite(n1, n2, density, seed)
    n = n1 + n2
    U = sp.triu(K).tocsc()
    U.sort_indices()
    s = LDLSolver(K, binding=backends.binding(backend))
    perm, etree, lnz = s.symbolic()
    F = orc.QDLDL(n, U.indptr, U.indices, U.data, perm=perm)
    # integer / permutation work: bit-exact
    assert np.array_equal(F.arr("perm"), perm)
    assert np.array_equal(F.arr("etree"), etree)
    assert np.array_equal(F.arr("Lnz"), lnz)
    Lp, Li, Lx, D = s.factor()
    assert np.array_equal(F.arr("Lp"), Lp)
    assert np.array_equal(F.arr("Li"), Li)
    # floating point
    assert np.allclose(D, F.arr("D"), rtol=1e-10, atol=0)
    assert np.allclose(Lx, F.arr("Lx"), rtol=1e-9, atol=1e-13)
    iner = s.compute_inertia()
    assert tuple(iner[0]) == (n1, n2, 0) == (F.positive_inertia, n - F.positive_inertia, 0)
    b = np.random.default_rng(seed).standard_normal(n)
    x = np.zeros(n)
    s.linear_solve(x, K, b)
    xo = F.solve(b)
    assert np.allclose(x, xo, rtol=1e-9, atol=1e-12)
    assert np.abs(K @ x - b).max() < 1e-9


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_refactor_and_batch(backend):
    """update_values! + refactor! (qdldl.jl:199-213,269-278) with new values; a batch of different matrices."""
    K = kkt_like(5)
    U = sp.triu(K).tocsc()
    U.sort_indices()
    n = K.shape[0]
    B = 3
    s = LDLSolver(K, batch=B, binding=backends.binding(backend))
    rng = np.random.default_rng(1)
    vals = np.stack([U.data * (1.0 + 0.1 * k) for k in range(B)])
    rhs = rng.standard_normal((B, n))
    x = np.zeros((B, n))
    s.linear_solve(x, vals, rhs)
    for k in range(B):
        assert np.abs((K * (1.0 + 0.1 * k)) @ x[k] - rhs[k]).max() < 1e-9
    iner = s.compute_inertia()
    assert np.all(iner == np.array([60, 40, 0]))
    # solve again without refactoring (fact=false)
    x2 = np.zeros((B, n))
    s.linear_solve(x2, None, 2.0 * rhs, fact=False)
    assert np.allclose(x2, 2.0 * x, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_indefinite_inertia_counts(backend):
    """compute_inertia!: positive = #(D>0), negative = #(D<=0) -- a matrix with the 'wrong' inertia is reported, not
    rejected (inertia.jl decides)."""
    A = sp.diags([2.0, -1.0, 3.0, -4.0, 5.0]).tocsc() + sp.csc_matrix(([0.1, 0.1], ([0, 4], [4, 0])), shape=(5, 5))
    s = LDLSolver(A, binding=backends.binding(backend))
    assert tuple(s.compute_inertia()[0]) == (3, 2, 0)


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_user_permutation_and_errors(backend):
    K = kkt_like(9, 20, 10, 0.2)
    n = 30
    p = np.random.default_rng(0).permutation(n).astype(np.int32)
    s = LDLSolver(K, perm=p, binding=backends.binding(backend))
    perm, _, _ = s.symbolic()
    assert sorted(perm.tolist()) == list(range(n))
    b = np.ones(n)
    x = np.zeros(n)
    s.linear_solve(x, K, b)
    assert np.abs(K @ x - b).max() < 1e-9
    with pytest.raises(Exception):
        LDLSolver(K, perm=np.zeros(n, dtype=np.int32), binding=backends.binding(backend))
    with pytest.raises(Exception):    # missing diagonal
        LDLSolver(sp.csc_matrix(([1.0], ([0], [1])), shape=(2, 2)), binding=backends.binding(backend))


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("make", [lqc.tiny, lqc.cfg2])
def test_kkt_symbolic_matches_oracle(backend, make):
    """Same reduced-matrix pattern as the oracle (column-major upper triangle in (x,y,z) order) and, for the
    product's ordering, bit-identical etree / column counts (QDLDL_etree!, qdldl.jl:358-395)."""
    P = make()
    k = BatchKKT(P, binding=backends.binding(backend))
    perm, etree, lnz = k.symbolic()
    o = orc.from_problem(P, perm=perm)
    F = o.ldl()
    assert np.array_equal(F.arr("perm"), perm)
    assert np.array_equal(F.arr("etree"), etree)
    assert np.array_equal(F.arr("Lnz"), lnz)
    info = k.info()
    assert info["nnzL"] == F.nnzL and info["nnzK"] == F.nnzA
    assert info["sum_lnz_sq"] == int((lnz.astype(np.int64) ** 2).sum())
