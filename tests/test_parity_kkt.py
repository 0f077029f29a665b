"""Newton/KKT hot path: product (C ABI) vs oracle on the same seeded inputs.

Tolerances: integer outcomes (inertia, regularisation trials, refinement passes, halving counts, iteration counts)
exact; floating point 1e-8 relative (BASELINE.json north_star) -- measured differences are ~1e-12.
"""
import itertools

import numpy as np
import pytest

import backends
import problems
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT, Options, Solver, initialize, solve
from oracle import oracle as orc

RTOL = 1e-8


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def oracle_at_iteration(P, iters, perm=None):
    """Run the oracle's solve! for `iters` Newton iterations and stop right before the next search direction."""
    o = orc.from_problem(P, perm=perm)
    o.use_superlu_fallback()
    o.initialize(P.x0)
    o.solve_begin()
    done = 0
    while done < iters:
        rc = o.newton_iteration()
        assert rc in (0, 2)
        if rc == 2:
            o.outer_update()
        else:
            done += 1
    return o


def push_state(k, o, b=0):
    """Copy the oracle's state (what evaluate! and the outer loop produced) into instance b of the handle."""
    sc = o.scalars()
    k.set("POINT", o.solution, first=b)
    k.set("DUAL", o.dual, first=b)
    s = k.get("SCALARS")
    for name in ("kappa", "tau", "rho", "eps_p_last"):
        s[b, backends._lib.S[name]] = sc[name]
    k.set("SCALARS", s)
    for name, arr in (("GRADIENT", o.gradient), ("EQ_DUAL_GRAD", o.eq_dual_grad), ("CONE_DUAL_GRAD", o.cone_dual_grad),
                      ("EQUALITY", o.equality), ("CONE", o.cone), ("W_VALUES", o.W_val), ("G_VALUES", o.G_val),
                      ("C_VALUES", o.C_val)):
        k.set(name, arr, first=b)


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("make,iters", [(lqc.tiny, 0), (lqc.tiny, 3), (lqc.tiny, 6), (lqc.cfg2, 2), (lqc.cfg2, 5)])
def test_newton_step_pieces(backend, make, iters):
    check_newton_step_pieces(backend, make(), iters)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,iters", [(s, it) for s in range(8) for it in ((2, 6) if s % 2 == 0 else (4,))])
def test_newton_step_pieces_cfg3(seed, iters):
    """BASELINE.json's headline configuration (cfg3: LQC(40,36,12,100), N = 4584) against the oracle, eight seeds: residual,
    reductions, inertia / trial / refinement / halving counts (exact), search direction and candidate (1e-8), L and D against
    QDLDL with the same ordering."""
    check_newton_step_pieces("cuda", lqc.cfg3(seed), iters)


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("make,iters", [(lqc.tiny, 3), (lqc.cfg2, 4)])
def test_newton_step_pieces_in_the_reference_ordering(backend, make, iters):
    """The same comparison with the reference's own elimination order, perm = amd(K) (qdldl.jl:135; restated ordering,
    tests/test_amd.py), handed to the product through the perm argument of cb200_create."""
    from test_amd import kkt_pattern
    P = make()
    K = kkt_pattern(P)
    check_newton_step_pieces(backend, P, iters, perm=orc.amd(K.shape[0], K.indptr, K.indices))


@pytest.mark.gpu
def test_newton_step_pieces_cfg3_in_the_reference_ordering():
    from test_amd import kkt_pattern
    P = lqc.cfg3(1)
    K = kkt_pattern(P)
    check_newton_step_pieces("cuda", P, 3, perm=orc.amd(K.shape[0], K.indptr, K.indices))


def check_newton_step_pieces(backend, P, iters, perm=None):
    k = BatchKKT(P, perm=perm, binding=backends.binding(backend))
    perm, _, _ = k.symbolic()
    o = oracle_at_iteration(P, iters, perm=perm)
    # the oracle now performs one more iteration's worth of hot path, piece by piece
    o.evaluate(2 | 16 | 32)
    o.cone_eval(barrier=True, barrier_gradient=True)
    o.merit_gradient_eval()
    o.residual_eval()
    o.evaluate(64 | 128 | 256)
    o.cone_eval(jacobian=True)
    push_state(k, o)
    # cone! + residual!
    k.cone(barrier=True, barrier_gradient=True, product=True)
    k.residual()
    # (near convergence the residual is a difference of O(1) terms: its error relative to its own magnitude reaches ~1e-11)
    assert rel(k.get("RESIDUAL")[0], o.residual) < 1e-10
    assert rel(k.get("BARRIER_GRADIENT")[0], o.barrier_gradient) < 1e-11
    assert rel(k.get("CONE_PRODUCT")[0], o.cone_product) < 1e-11
    sc = k.scalars()
    assert sc["barrier"][0] == pytest.approx(o.scalars()["barrier"], rel=1e-12)
    assert sc["residual_violation"][0] == pytest.approx(np.abs(o.residual).sum() / o.total, rel=1e-12)
    assert sc["optimality_violation"][0] == pytest.approx(o.optimality_error(), rel=1e-12)
    assert sc["theta"][0] == pytest.approx(o.constraint_violation(), rel=1e-12)
    # search_direction!
    rc = o.search_direction()
    k.search_direction()
    st = {kk: int(v[0]) for kk, v in k.stats().items()}
    ost = o.stats
    assert (st["inertia_pos"], st["inertia_neg"], st["inertia_zero"]) == o.inertia
    assert st["n_trials"] == ost["n_trials"]
    assert st["n_refine"] == ost["n_refine"]
    assert st["refine_ok"] == ost["refine_ok"]
    assert st["used_fallback"] == ost["used_lu"]
    assert st["status"] == 0 and rc == 0
    sc = k.scalars()
    osc = o.scalars()
    assert sc["eps_p"][0] == osc["eps_p"] and sc["eps_d"][0] == osc["eps_d"] and sc["eps_p_last"][0] == osc["eps_p_last"]
    step = k.get("STEP")[0]
    assert rel(step, o.step) < RTOL
    # the direction solves the full Newton system (refinement is part of the solve)
    e = o.residual - o.jacobian_times(step)
    assert np.abs(e).max() <= max(1e-10, 10 * sc["refine_norm"][0])
    # J v through the ABI
    v = np.random.default_rng(0).standard_normal(o.total)
    assert rel(k.jacobian_times(v)[0], o.jacobian_times(v)) < 1e-12
    # L and D against QDLDL with the same ordering
    F = o.ldl()
    Lp, Li, Lx, D = k.factor()
    assert np.array_equal(Lp, F.arr("Lp")) and np.array_equal(Li, F.arr("Li"))
    assert np.allclose(D, F.arr("D"), rtol=1e-9, atol=0)
    assert np.allclose(Lx, F.arr("Lx"), rtol=1e-8, atol=1e-12)
    # cone line search
    assert o.cone_search() == 0
    k.cone_search()
    st = {kk: int(v[0]) for kk, v in k.stats().items()}
    assert (st["k_s"], st["k_t"]) == (o.stats["k_s"], o.stats["k_t"])
    sc = k.scalars()
    assert sc["step_size"][0] == 0.5 ** st["k_s"] and sc["step_size_t"][0] == 0.5 ** st["k_t"]
    cand = k.get("CANDIDATE")[0]
    assert rel(cand[k.is_], o.candidate[o.is_]) < RTOL
    assert rel(cand[k.it], o.candidate[o.it]) < RTOL


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_inertia_correction_schedule_matches_oracle(backend):
    """Indefinite Hessian: same number of trials and the same eps_p sequence end point (inertia.jl:30-79)."""
    P = problems.maratos()
    k = BatchKKT(P, binding=backends.binding(backend))
    o = orc.Oracle(P.n, P.m, P.p, P.num_nonnegative, P.soc_dims, P.W_colptr, P.W_rowval, P.G_colptr, P.G_rowval,
                   P.C_colptr, P.C_rowval)
    o.set_callback(P.callback)
    o.initialize(np.array([0.5, 0.5]))
    o.solution[o.iy] = -10.0
    o.set_scalars(kappa=1.0, rho=1.0)
    o.evaluate(511)
    o.residual_eval()
    push_state(k, o)
    for trial in range(2):      # second call starts from eps_p_last / 3 and scales by 8
        assert o.search_direction() in (0, 2)
        k.search_direction()
        st = {kk: int(v[0]) for kk, v in k.stats().items()}
        assert st["n_trials"] == o.stats["n_trials"]
        assert (st["inertia_pos"], st["inertia_neg"], st["inertia_zero"]) == o.inertia
        sc, osc = k.scalars(), o.scalars()
        assert sc["eps_p"][0] == osc["eps_p"] and sc["eps_p_last"][0] == osc["eps_p_last"]


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("make", [lqc.tiny, lqc.cfg2, lqc.cfg2_hard, problems.rocket_landing])
def test_lq_solve_on_device_matches_oracle(backend, make):
    """Whole solve! on the device (LQ callbacks) vs the oracle's solve!: same iteration counts, same solution.
    (rocket_landing: the reference's own example test/examples/rocket_landing.jl, an LQ-conic problem.)"""
    P = make()
    k = BatchKKT(P, binding=backends.binding(backend))
    perm, _, _ = k.symbolic()
    k.load_lq(P)
    k.initialize(P.x0)
    k.lq_begin()
    r = k.lq_solve(max_steps=300, check_every=2)
    assert r["converged"] == 1 and r["running"] == 0
    o = orc.from_problem(P, perm=perm)
    o.use_superlu_fallback()
    o.initialize(P.x0)
    assert o.solve() == 1
    st = {kk: int(v[0]) for kk, v in k.stats().items()}
    assert st["total_iterations"] == o.stats["total_iterations"]
    assert st["outer"] == o.stats["outer"]
    assert st["fallbacks"] == o.stats["lu_fallbacks"]
    if make is lqc.cfg2_hard:
        assert st["fallbacks"] > 0          # the GMRES stand-in for `J \\ R` (search_direction.jl:22) was exercised
    w = k.get("POINT")[0]
    assert rel(w, o.solution) < RTOL          # whole solve!: measured 1e-15 .. 1e-13 (the refined steps agree to ~1e-12)
    sc = k.scalars()
    # kappa goes through pow(kappa, 1.5): device libm vs host libm may differ in the last ulp
    assert sc["kappa"][0] == pytest.approx(o.scalars()["kappa"], rel=1e-13)
    assert sc["rho"][0] == pytest.approx(o.scalars()["rho"], rel=1e-13)


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_batch_of_instances_is_independent(backend):
    """cfg4-style batch: different seeds share the pattern; each instance equals its own single solve."""
    Ps = [lqc.tiny(i) for i in range(5)]
    k = BatchKKT(Ps[0], batch=5, binding=backends.binding(backend))
    perm, _, _ = k.symbolic()
    k.load_lq(Ps)
    k.initialize(np.stack([P.x0 for P in Ps]))
    k.lq_begin()
    r = k.lq_solve(max_steps=300, check_every=3)
    assert r["converged"] == 5
    W = k.get("POINT")
    st = k.stats()
    for i, P in enumerate(Ps):
        o = orc.from_problem(P, perm=perm)
        o.use_superlu_fallback()
        o.initialize(P.x0)
        assert o.solve() == 1
        assert rel(W[i], o.solution) < RTOL
        assert st["total_iterations"][i] == o.stats["total_iterations"]


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("device_line_search", [False, True])
@pytest.mark.parametrize("name", ["wachter", "maratos", "knitro", "test1", "test2", "test3", "test4", "portfolio", "pendulum",
                                  "pendulum_overwrite", "qp_nonnegative"])
def test_reference_solver_cases_through_host_callbacks(backend, name, device_line_search):
    """test/solver/*.jl cases through Solver / initialize! / solve! with host callbacks and the GPU hot path:
    same stopping criteria (e.g. wachter.jl:36-45) and the oracle's solution.  device_line_search: the filter line search of
    solve.jl:224-306 through cb200_filter_search (candidates' callback outputs uploaded in blocks) instead of host Python."""
    P = getattr(problems, name)()
    s = Solver(P, P.callback, binding=backends.binding(backend), line_search_on_device=device_line_search)
    initialize(s, P.x0)
    assert solve(s) is True
    R = s.residual
    k = s.kkt
    assert np.abs(R).sum() / s.total < 1e-4
    assert max(np.abs(R[k.iy]).max(initial=0), np.abs(R[k.iz]).max(initial=0)) < 1e-4
    assert np.abs(s.out.equality).max(initial=0) <= 1e-4
    assert np.abs(s.cone_product).max(initial=0) <= 1e-4
    if P.x_star is not None:
        assert np.abs(s.solution[:P.n] - P.x_star).max() < 1e-3
    o = orc.Oracle(P.n, P.m, P.p, P.num_nonnegative, P.soc_dims, P.W_colptr, P.W_rowval, P.G_colptr, P.G_rowval,
                   P.C_colptr, P.C_rowval, perm=k.symbolic()[0])
    o.set_callback(P.callback)
    o.use_superlu_fallback()
    o.initialize(P.x0)
    assert o.solve() == 1
    assert s.iterations == o.stats["total_iterations"]
    assert rel(s.solution, o.solution) < RTOL


FRICTION_TABLE = list(itertools.product([[0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.0, 1.0, 1.0], [0.0, 10.0, 1.0]],
                                        [0.0, 0.5, 1.0], [0.0, 1.0]))
FRICTION_SUBSET = [([0.0, 1.0, 0.0], 0.5, 1.0), ([0.0, 1.0, 1.0], 1.0, 1.0), ([0.0, 10.0, 1.0], 0.5, 1.0), ([0.0, 0.0, 0.0], 0.0, 0.0)]


@pytest.mark.parametrize("backend,table", [("emul", FRICTION_SUBSET), pytest.param("cuda", FRICTION_TABLE, marks=pytest.mark.gpu)])
def test_friction_cone_table(backend, table):
    """test/solver/friction_cone.jl:19-63: all 30 (v, mu, gamma) combinations on the GPU (a subset through the host
    emulation, to keep the CPU suite short; the oracle test runs all of them too)."""
    for v, mu, gamma in table:
        P = problems.friction(v, mu, gamma, np.random.default_rng(3).standard_normal(3))
        s = Solver(P, P.callback, binding=backends.binding(backend))
        initialize(s, P.x0)
        assert solve(s) is True
        x = s.solution[:3]
        sl = s.solution[s.kkt.is_]
        assert sl[0] - np.linalg.norm(sl[1:]) > -1e-8
        if np.linalg.norm(v[1:]) > 0 and gamma > 0 and mu > 0:
            vdir = np.asarray(v[1:]) / np.linalg.norm(v[1:])
            bdir = x[1:] / np.linalg.norm(x[1:])
            assert np.abs(vdir + bdir).max() < 1e-3
            assert np.linalg.norm(x[1:]) <= mu * gamma + 1e-6


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_errors_are_reported(backend):
    P = lqc.tiny()
    b = backends.binding(backend)
    with pytest.raises(Exception):
        bad = lqc.tiny()
        bad.soc_dims = np.array([3, 3], dtype=np.int32)      # does not sum to p
        BatchKKT(bad, binding=b)
    k = BatchKKT(P, binding=b)
    with pytest.raises(Exception):
        k.get("RHS")                                           # LinearSolver-seam array on a KKT handle


@pytest.mark.gpu
@pytest.mark.parametrize("make,batch,check_every", [(lqc.tiny, 5, 3), (lqc.cfg2, 3, 4), (lqc.cfg2, 3, 400)])
def test_independent_launch_schedule_is_bitwise_the_lockstep_one(make, batch, check_every, monkeypatch):
    """cb200_lq_solve: a launch that carries several Newton iterations of an instance (default) computes exactly what
    one launch per iteration computes (CB200_LQ_LOCKSTEP=1) -- instances never interact, only the scheduling differs."""
    from calipso_b200.solver import BatchKKT
    Ps = [make(i) for i in range(batch)]
    out = []
    for lockstep in ("1", "0"):
        monkeypatch.setenv("CB200_LQ_LOCKSTEP", lockstep)
        k = BatchKKT(Ps[0], batch=batch, binding=backends.binding("cuda"))
        k.load_lq(Ps)
        k.initialize(np.stack([P.x0 for P in Ps]))
        k.lq_begin()
        r = k.lq_solve(max_steps=400, check_every=check_every)
        st = k.stats()
        out.append((r["converged"], k.get("POINT"), k.get("DUAL"), k.get("SCALARS"),
                    {n: st[n].copy() for n in ("total_iterations", "factorizations", "solves", "fallbacks")}))
        k.close()
    a, b = out
    assert a[0] == b[0] == batch
    for i in (1, 2, 3):
        assert np.array_equal(a[i], b[i])
    for n in a[4]:
        assert np.array_equal(a[4][n], b[4][n]), n
