"""Pin the oracle's QDLDL restatement (oracle/qdldl.c <- src/solver/qdldl.jl) by its defining identities."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc


def quasidefinite(n1, n2, density, seed):
    rng = np.random.default_rng(seed)
    A = sp.random(n1, n1, density, random_state=rng.integers(1 << 30), format="csc")
    H = (A @ A.T + sp.eye(n1)).tocsc()
    B = sp.random(n2, n1, density, random_state=rng.integers(1 << 30), format="csc")
    K = sp.bmat([[H, B.T], [B, -sp.eye(n2) * 0.5]]).tocsc()
    return K


@pytest.mark.parametrize("n1,n2,density,seed", [(1, 0, 1.0, 0), (5, 3, 0.5, 1), (40, 25, 0.1, 2), (200, 120, 0.02, 3)])
def test_ldl_identity_and_solve(n1, n2, density, seed):
    K = quasidefinite(n1, n2, density, seed)
    n = n1 + n2
    U = sp.triu(K).tocsc()
    U.sort_indices()
    F = orc.QDLDL(n, U.indptr, U.indices, U.data)
    perm, iperm = F.arr("perm"), F.arr("iperm")
    assert sorted(perm.tolist()) == list(range(n))
    assert np.all(iperm[perm] == np.arange(n))                       # invperm, qdldl.jl:143
    Lp, Li, Lx, D = F.arr("Lp"), F.arr("Li"), F.arr("Lx"), F.arr("D")
    L = sp.csc_matrix((Lx, Li, Lp), shape=(n, n)) + sp.eye(n)
    PKP = K[perm][:, perm].toarray()
    assert np.allclose((L @ sp.diags(D) @ L.T).toarray(), PKP, atol=1e-10)   # P K P' = L D L'
    assert np.allclose(F.arr("Dinv"), 1.0 / D)
    assert F.positive_inertia == n1                                  # Sylvester: n1 positive pivots
    assert (D < 0).sum() == n2
    # columns of L hold ascending rows (Appendix B), Lp is the cumsum of Lnz
    assert np.all(np.diff(Lp) == F.arr("Lnz"))
    for j in range(n):
        assert np.all(np.diff(Li[Lp[j]:Lp[j + 1]]) > 0) and np.all(Li[Lp[j]:Lp[j + 1]] > j)
    # etree: parent = first sub-diagonal row of each column of L
    et = F.arr("etree")
    for j in range(n):
        assert et[j] == (Li[Lp[j]] if Lp[j + 1] > Lp[j] else -1)
    b = np.random.default_rng(seed).standard_normal(n)
    x = F.solve(b)
    assert np.allclose(K @ x, b, atol=1e-9)


def test_permute_symmetric_map_and_refactor():
    K = quasidefinite(30, 20, 0.15, 7)
    n = 50
    U = sp.triu(K).tocsc()
    U.sort_indices()
    perm = np.random.default_rng(0).permutation(n).astype(np.int32)
    F = orc.QDLDL(n, U.indptr, U.indices, U.data, perm=perm)           # qdldl(A; perm=p), qdldl.jl:134-136
    assert np.all(F.arr("perm") == perm)
    Pc, Pr, Pv, A2P = F.arr("triuA_colptr"), F.arr("triuA_rowval"), F.arr("triuA_nzval"), F.arr("AtoPAPt")
    iperm = F.arr("iperm")
    # every entry (r,c) of triu(A) lands at (min,max) of the permuted pair, values follow AtoPAPt (qdldl.jl:669-742)
    cols = np.repeat(np.arange(n), np.diff(U.indptr))
    colP = np.repeat(np.arange(n), np.diff(Pc))
    for k in range(U.nnz):
        r, c = iperm[U.indices[k]], iperm[cols[k]]
        assert Pr[A2P[k]] == min(r, c) and colP[A2P[k]] == max(r, c) and Pv[A2P[k]] == U.data[k]
    assert sorted(A2P.tolist()) == list(range(U.nnz))
    # update_values! + refactor!: new numeric values, same symbolic structure
    new = U.data * 1.5
    pos = F.refactor(new)
    assert pos == 30
    x = F.solve(np.ones(n))
    assert np.allclose((K * 1.5) @ x, np.ones(n), atol=1e-9)


def test_zero_pivot_aborts():
    # [[0,1],[1,0]] without permutation: first pivot is zero -> -1 (qdldl.jl:456)
    Ap = np.array([0, 1, 3], np.int32)
    Ai = np.array([0, 0, 1], np.int32)
    Ax = np.array([0.0, 1.0, 0.0])
    F = orc.QDLDL(2, Ap, Ai, Ax, perm=np.array([0, 1], np.int32))
    assert F.positive_inertia == -1


def test_empty_column_rejected():
    # "Input matrix is not upper triangular or has an empty column", qdldl.jl:70-73 / QDLDL_etree! :368
    Ap = np.array([0, 1, 1], np.int32)
    Ai = np.array([0], np.int32)
    with pytest.raises(ValueError):
        orc.QDLDL(2, Ap, Ai, np.ones(1), perm=np.array([0, 1], np.int32))


def test_min_degree_is_permutation_and_reduces_fill():
    K = quasidefinite(150, 100, 0.02, 11)
    n = 250
    U = sp.triu(K).tocsc()
    U.sort_indices()
    p = orc.min_degree(n, U.indptr, U.indices)
    assert sorted(p.tolist()) == list(range(n))
    F = orc.QDLDL(n, U.indptr, U.indices, U.data, perm=p)
    Fn = orc.QDLDL(n, U.indptr, U.indices, U.data, perm=np.arange(n, dtype=np.int32))
    assert F.nnzL <= Fn.nnzL
