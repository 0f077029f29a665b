"""BASELINE.json's full sizes (cfg3: LQC(40,36,12,100), N = 4584, total = 8496; cfg4: a batch of 64 cfg3 instances) on the GPU:

* against the ORACLE on the same seeded inputs -- complete solve! of cfg3 seeds and of a cfg4 batch (per-instance iteration,
  outer-iteration and fallback counts exact, solution to 1e-8; the per-step comparison of the pieces is
  tests/test_parity_kkt.py::test_newton_step_pieces_cfg3);
* through the size-independent properties the reference's own unit test pins (test/solver/problem.jl:100-211): expected
  inertia (n, m+p, 0) (inertia.jl:7-11), refinement drives the FULL Newton system to ||R - J step||_inf <= 1e-10 (:207-211),
  the reduced LDL' solve reproduces K x = b, and a complete solve! meets the four stopping criteria (e.g.
  test/solver/wachter.jl:36-45)."""
import numpy as np
import pytest
import scipy.sparse as sp

import backends
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT, LDLSolver
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-8


def _oracle_solve(P, perm):
    o = orc.from_problem(P, perm=perm)
    o.use_superlu_fallback()
    o.initialize(P.x0)
    assert o.solve() == 1
    return o


def _compare_batch_with_oracle(k, Ps):
    """Every instance of a solved batch against its own oracle solve!: integer outcomes exact, floating point 1e-8."""
    perm, _, _ = k.symbolic()
    W, lam, st, sc = k.get("POINT"), k.get("DUAL"), k.stats(), k.scalars()
    for i, P in enumerate(Ps):
        o = _oracle_solve(P, perm)
        assert st["total_iterations"][i] == o.stats["total_iterations"], i
        assert st["outer"][i] == o.stats["outer"], i
        assert st["fallbacks"][i] == o.stats["lu_fallbacks"], i
        assert np.abs(W[i] - o.solution).max() <= RTOL * np.abs(o.solution).max(), i
        assert np.abs(lam[i] - o.dual).max() <= RTOL * max(np.abs(o.dual).max(), 1.0), i
        osc = o.scalars()
        assert sc["kappa"][i] == pytest.approx(osc["kappa"], rel=1e-13) and sc["rho"][i] == pytest.approx(osc["rho"], rel=1e-13)


@pytest.mark.parametrize("first_seed", [0, 8])
def test_cfg3_solve_matches_oracle(first_seed):
    """Complete solve! of eight cfg3 seeds per case (one instance per CTA) against the oracle."""
    Ps = [lqc.cfg3(first_seed + i) for i in range(8)]
    k = BatchKKT(Ps[0], batch=8, binding=backends.binding("cuda"))
    k.load_lq(Ps)
    k.initialize(np.stack([P.x0 for P in Ps]))
    k.lq_begin()
    r = k.lq_solve(max_steps=400, check_every=400)
    assert r["converged"] == 8 and r["running"] == 0 and r["error"] == 0
    _compare_batch_with_oracle(k, Ps)


def test_cfg4_batch_matches_oracle():
    """BASELINE.json configs[4]: the batch of 64 independent cfg3 instances (seeds 0 .. 63, the blocks of 8 a rank owns at 8
    GPUs) solved in one handle (one instance per SM: the 512-thread kernels), every instance against its oracle solve!."""
    B = 64
    Ps = [lqc.cfg3(i) for i in range(B)]
    k = BatchKKT(Ps[0], batch=B, binding=backends.binding("cuda"))
    k.load_lq(Ps)
    k.initialize(np.stack([P.x0 for P in Ps]))
    k.lq_begin()
    r = k.lq_solve(max_steps=400, check_every=400)
    assert r["converged"] == B and r["running"] == 0 and r["error"] == 0
    _compare_batch_with_oracle(k, Ps)


def test_cfg3_search_direction_properties():
    Ps = [lqc.cfg3(i) for i in range(3)]
    k = BatchKKT(Ps[0], batch=3, binding=backends.binding("cuda"))
    info = k.info()
    assert (info["N"], info["total"]) == (4584, 8496)
    k.load_lq(Ps)
    k.initialize(np.stack([P.x0 for P in Ps]))
    k.lq_begin()
    k.lq_step(3)
    # one hand-driven Newton step at the current point through the reference-facing calls
    k.lq_evaluate(2 | 16 | 32)
    k.cone(barrier=True, barrier_gradient=True, product=True)
    k.residual()
    k.search_direction()
    st = k.stats()
    P = Ps[0]
    assert np.all(st["status"] == 0)
    assert np.all(st["inertia_pos"] == P.n) and np.all(st["inertia_neg"] == P.m + P.p) and np.all(st["inertia_zero"] == 0)
    assert np.all(st["n_refine"] >= 1)                                  # min_iterative_refinement = 1 (options.jl:16)
    R, step = k.get("RESIDUAL"), k.get("STEP")
    e = R - k.jacobian_times(step)
    assert np.abs(e).max(axis=1).max() <= 1e-10                          # iterative_refinement.jl:14-17
    # cone search: candidates stay strictly inside the cones with the fraction-to-boundary margin
    k.cone_search()
    st = k.stats()
    assert np.all(st["status"] == 0) and np.all(st["k_s"] <= 25) and np.all(st["k_t"] <= 25)
    cand, w = k.get("CANDIDATE"), k.get("POINT")
    tau = k.scalars()["tau"][:, None]
    q = P.num_nonnegative
    for sl in (k.is_, k.it):
        c, x = cand[:, sl], w[:, sl]
        assert np.all(c[:, :q] > (1 - tau) * x[:, :q])
        off = q
        for d in P.soc_dims:
            v = c[:, off:off + d] - (1 - tau) * x[:, off:off + d]
            assert np.all(v[:, 0] > np.linalg.norm(v[:, 1:], axis=1))
            off += d


def test_cfg3_sized_linear_solver_seam():
    """ldl_solver / factorize! / compute_inertia! / linear_solve! on a quasi-definite matrix with cfg3's reduced-KKT shape."""
    P = lqc.cfg3(0)
    n, m, p = P.n, P.m, P.p
    W, G, Cm = P.W_full(), P.G(), P.C()
    K = sp.bmat([[W + 1e-7 * sp.eye(n), G.T, Cm.T], [G, -(1.0 + 1e-7) * sp.eye(m), None],
                 [Cm, None, -0.7 * sp.eye(p)]]).tocsc()
    s = LDLSolver(K, batch=2, binding=backends.binding("cuda"))
    s.factorize(K)
    inertia = s.compute_inertia()
    assert tuple(inertia[0]) == (n, m + p, 0) and tuple(inertia[1]) == (n, m + p, 0)
    rng = np.random.default_rng(1)
    b = rng.standard_normal((2, n + m + p))
    x = np.zeros_like(b)
    s.linear_solve(x, K, b)
    for i in range(2):
        assert np.abs(K @ x[i] - b[i]).max() <= 1e-7 * np.abs(b[i]).max()
    # the factor reproduces P K P' = L D L' (QDLDL contract, Appendix B)
    Lp, Li, Lx, D = s.factor(0)
    perm = s.symbolic()[0]
    N = n + m + p
    L = sp.csc_matrix((Lx, Li, Lp), shape=(N, N)) + sp.eye(N)
    PKP = K[perm][:, perm]
    E = (L @ sp.diags(D) @ L.T - PKP).tocoo()
    assert np.abs(E.data).max() <= 1e-9 * np.abs(K.data).max()


def test_cfg3_batch_solve_meets_stopping_criteria():
    B = 8
    Ps = [lqc.cfg3(100 + i) for i in range(B)]
    k = BatchKKT(Ps[0], batch=B, binding=backends.binding("cuda"))
    k.load_lq(Ps)
    k.initialize(np.stack([P.x0 for P in Ps]))
    k.lq_begin()
    r = k.lq_solve(max_steps=400, check_every=4)
    assert r["converged"] == B and r["running"] == 0 and r["error"] == 0
    sc = k.scalars()
    for name in ("residual_violation", "slack_violation", "equality_violation", "cone_product_violation"):
        assert np.all(sc[name] <= 1e-4), name                           # options.jl tolerances
    # the converged points are feasible for the LQ problem itself
    W = k.get("POINT")
    for i, P in enumerate(Ps):
        x = W[i, k.ix]
        assert np.abs(P.G() @ x + P.g0).max() <= 1e-4
        h = P.C() @ x + P.h0
        assert h[:P.num_nonnegative].min() >= -1e-4


def test_narrow_and_wide_kernel_instantiations_agree():
    """Batches of at most one CTA per SM run the 512-thread instantiations of the heavy kernels, larger batches the
    256-thread ones (three CTAs per SM).  Same instances through both: same iteration counts, same solutions."""
    Ps = [lqc.tiny(i) for i in range(4)]
    big = 200                                   # > 148 SMs => narrow kernels
    k = BatchKKT(Ps[0], batch=big, binding=backends.binding("cuda"))
    k.load_lq([Ps[i % 4] for i in range(big)])
    k.initialize(np.stack([Ps[i % 4].x0 for i in range(big)]))
    k.lq_begin()
    r = k.lq_solve(max_steps=300, check_every=3)
    assert r["converged"] == big
    W, it = k.get("POINT"), k.stats()["total_iterations"]
    k1 = BatchKKT(Ps[0], batch=4, binding=backends.binding("cuda"))          # wide kernels
    k1.load_lq(Ps)
    k1.initialize(np.stack([P.x0 for P in Ps]))
    k1.lq_begin()
    assert k1.lq_solve(max_steps=300, check_every=3)["converged"] == 4
    W1, it1 = k1.get("POINT"), k1.stats()["total_iterations"]
    for i in range(big):
        assert it[i] == it1[i % 4]
        assert np.abs(W[i] - W1[i % 4]).max() <= 1e-9 * max(1.0, np.abs(W1[i % 4]).max())
    # and a cfg3-sized KKT solve unit through the narrow path: direction solves the full Newton system
    Pc = [lqc.cfg3(i) for i in range(2)]
    kc = BatchKKT(Pc[0], batch=150, binding=backends.binding("cuda"))
    kc.load_lq([Pc[i % 2] for i in range(150)])
    kc.initialize(np.stack([Pc[i % 2].x0 for i in range(150)]))
    kc.lq_begin()
    kc.lq_step(2)
    kc.lq_evaluate(2 | 16 | 32)
    kc.cone(barrier=True, barrier_gradient=True, product=True)
    kc.residual()
    kc.search_direction()
    st = kc.stats()
    assert np.all(st["status"] == 0) and np.all(st["inertia_pos"] == Pc[0].n)
    e = kc.get("RESIDUAL") - kc.jacobian_times(kc.get("STEP"))
    assert np.abs(e).max() <= 1e-10


def test_quadruped_shape_runs_on_the_shared_memory_path_and_matches_oracle():
    """A pattern beyond BASELINE's cfg3 -- the dimensions of the reference's quadruped example (N = 6987, stages of 75
    variables, test/examples/quadruped_gait.jl:236-244,460-464): the shared-memory plan is re-sized (leaves-first ordering:
    only the non-leaf unknowns stay in shared memory during the solves, which keeps three resident CTAs per SM), nothing falls
    back to the global-memory code, and the complete solve! agrees with the oracle."""
    Ps = [lqc.quadruped_shape(i) for i in range(3)]
    k = BatchKKT(Ps[0], batch=3, binding=backends.binding("cuda"))
    paths = k.paths()
    assert paths["solve_in_shared_memory"] == 1 and paths["cta_supernodes_generic"] == 0 and paths["ctas_per_sm"] == 3
    assert k.info()["N"] == 6987
    k.load_lq(Ps)
    k.initialize(np.stack([P.x0 for P in Ps]))
    k.lq_begin()
    r = k.lq_solve(max_steps=400, check_every=400)
    assert r["converged"] == 3 and r["error"] == 0
    _compare_batch_with_oracle(k, Ps)
