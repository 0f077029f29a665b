"""Product (C ABI) against the committed golden fixtures (tests/golden/*.npz, generated from the oracle by
tests/golden/make_golden.py).  Needs neither the oracle nor /root/reference at run time.  Integer outcomes exact,
floating point 1e-8 relative (BASELINE.json north_star)."""
import glob
import os

import numpy as np
import pytest

import backends
from calipso_b200 import lqc
from calipso_b200.solver import BatchKKT

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEPS = sorted(glob.glob(os.path.join(GOLD, "step_*.npz")))
SOLVES = sorted(glob.glob(os.path.join(GOLD, "solve_*.npz")))
RTOL = 1e-8


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def parse(path):
    parts = os.path.basename(path)[:-4].split("_")
    return parts[1], [int(x) for x in parts[2:]]


def instance(name, seed):
    if name == "rocket":                      # test/examples/rocket_landing.jl as an LQ-conic instance
        import problems
        return problems.rocket_landing(seed)
    return getattr(lqc, name)(seed)


def test_fixtures_are_committed():
    assert len(STEPS) >= 4 and len(SOLVES) >= 3


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("path", STEPS, ids=os.path.basename)
def test_newton_step_matches_golden(backend, path):
    name, (seed, _) = parse(path)
    g = np.load(path)
    P = getattr(lqc, name)(seed)
    k = BatchKKT(P, perm=g["perm"], binding=backends.binding(backend))
    assert np.array_equal(k.symbolic()[0], g["perm"])
    k.set("POINT", g["point"])
    k.set("DUAL", g["dual"])
    s = k.get("SCALARS")
    for nm in ("kappa", "tau", "rho", "eps_p_last"):
        s[0, backends._lib.S[nm]] = float(g[nm])
    k.set("SCALARS", s)
    for arr, key in (("GRADIENT", "gradient"), ("EQ_DUAL_GRAD", "eq_dual_grad"), ("CONE_DUAL_GRAD", "cone_dual_grad"),
                     ("EQUALITY", "equality"), ("CONE", "cone"), ("W_VALUES", "W_val"), ("G_VALUES", "G_val"),
                     ("C_VALUES", "C_val")):
        k.set(arr, g[key])
    k.cone(barrier=True, barrier_gradient=True, product=True)
    k.residual()
    assert rel(k.get("RESIDUAL")[0], g["residual"]) < 1e-11
    k.search_direction()
    st = {kk: int(v[0]) for kk, v in k.stats().items()}
    assert (st["inertia_pos"], st["inertia_neg"], st["inertia_zero"]) == tuple(int(x) for x in g["inertia"])
    assert st["n_trials"] == int(g["n_trials"]) and st["n_refine"] == int(g["n_refine"])
    assert st["refine_ok"] == int(g["refine_ok"]) and st["used_fallback"] == int(g["used_lu"]) and st["status"] == 0
    sc = k.scalars()
    assert sc["eps_p"][0] == float(g["eps_p"]) and sc["eps_d"][0] == float(g["eps_d"])
    assert rel(k.get("STEP")[0], g["step"]) < RTOL
    k.cone_search()
    st = {kk: int(v[0]) for kk, v in k.stats().items()}
    assert (st["k_s"], st["k_t"]) == (int(g["k_s"]), int(g["k_t"]))
    cand = k.get("CANDIDATE")[0]
    assert rel(cand[k.is_], g["candidate"][k.is_]) < RTOL and rel(cand[k.it], g["candidate"][k.it]) < RTOL


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("path", SOLVES, ids=os.path.basename)
def test_solve_matches_golden(backend, path):
    name, (seed,) = parse(path)
    g = np.load(path)
    P = instance(name, seed)
    k = BatchKKT(P, perm=g["perm"], binding=backends.binding(backend))
    k.load_lq(P)
    k.initialize(P.x0)
    k.lq_begin()
    r = k.lq_solve(max_steps=300, check_every=2)
    assert r["converged"] == 1
    st = {kk: int(v[0]) for kk, v in k.stats().items()}
    assert st["total_iterations"] == int(g["total_iterations"]) and st["outer"] == int(g["outer"])
    assert st["fallbacks"] == int(g["lu_fallbacks"])
    assert rel(k.get("POINT")[0], g["solution"]) < 1e-6
