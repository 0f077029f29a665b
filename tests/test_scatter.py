"""evaluate!'s scatter of the flat derivative caches (src/solver/evaluate.jl:37-42,55-60,
73-78,95-100,109-114) through the C ABI (cb200_scatter_plan / cb200_scatter) against the oracle's literal restatement
(dense matrices written with `=` in cache order, then added: residual_jacobian_variables.jl:11-13).  Bit-exact: the
scatter moves values and adds at most three of them in the reference's order.

The pendulum case (BASELINE cfg1) uses the key order of the trajectory-optimisation front end (stage t of the dynamics
owns the keys of (x_t, u_t, x_{t+1}): trajectory_optimization/constraints.jl:285-301, methods.jl:26), where stage t and
stage t+1 both write (x_{t+1}, x_{t+1}) -- SURVEY.md Appendix A.17: the reference keeps the LAST write, not the sum."""
import numpy as np
import pytest

import backends
import problems
from calipso_b200 import _lib
from calipso_b200.solver import BatchKKT
from oracle import oracle as orc


def pendulum_caches(P, T, v, y):
    """Keys (0-based) and values of the three Hessian caches and of the equality-Jacobian cache at (v, y), in the
    reference's cache order (stage by stage)."""
    nx, nu, h = 2, 1, 0.05
    grav_l = 9.81 / 0.5
    ix = [t * (nx + nu) for t in range(T)]
    obj_keys = [(i, i) for i in range(P.n)]
    obj_vals = [0.2] * P.n
    eq_keys, eq_vals = [], []
    for t in range(T - 1):
        a, b = ix[t], ix[t + 1]                       # first state coordinate of x_t and of x_{t+1}
        d2 = -0.25 * h * grav_l * np.sin(0.5 * (v[a] + v[b])) * y[2 * t + 1]
        for key in ((a, a), (b, a), (a, b), (b, b)):  # column-major order of the stage block
            eq_keys.append(key)
            eq_vals.append(d2)
    J = P.jac_g(v)
    jac_keys = [(r, c) for r in range(P.m) for c in range(P.n) if J[r, c] != 0.0]
    jac_vals = [J[r, c] for r, c in jac_keys]
    return (obj_keys, eq_keys, []), (obj_vals, eq_vals, []), jac_keys, jac_vals


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_pendulum_last_write_wins(backend):
    T = 11
    P = problems.pendulum(0, T)
    B = 3
    k = BatchKKT(P, batch=B, binding=backends.binding(backend))
    rng = np.random.default_rng(3)
    pts = [(rng.standard_normal(P.n), rng.standard_normal(P.m)) for _ in range(B)]
    keys, _, jac_keys, _ = pendulum_caches(P, T, *pts[0])
    k.scatter_plan("W_VALUES", keys)
    k.scatter_plan("G_VALUES", [jac_keys])
    Wc, Gc = [], []
    for v, y in pts:
        _, vals, _, jv = pendulum_caches(P, T, v, y)
        Wc.append(np.concatenate([np.asarray(c, float) for c in vals]))
        Gc.append(np.asarray(jv, float))
    k.scatter("W_VALUES", np.stack(Wc))
    k.scatter("G_VALUES", np.stack(Gc))
    W, G = k.get("W_VALUES"), k.get("G_VALUES")
    differs = 0
    for b, (v, y) in enumerate(pts):
        keys, vals, jac_keys, jv = pendulum_caches(P, T, v, y)
        H = orc.lagrangian_hessian_dense(P.n, keys, vals)
        assert np.array_equal(W[b], orc.values_at_pattern(H, P.W_colptr, P.W_rowval))
        Jd = orc.scatter_dense((P.m, P.n), jac_keys, jv)
        assert np.array_equal(G[b], orc.values_at_pattern(Jd, P.G_colptr, P.G_rowval))
        assert np.array_equal(Jd, P.jac_g(v))
        # the summed Hessian (what the reference's own Hessian test builds with +=) differs on the shared states
        Hsum = P.hess(v) + P.hess_gy(v, y)
        differs += int(np.abs(H - Hsum).max() > 1e-6)
        inner = [t * 3 for t in range(1, T - 1)]
        for a in inner:
            t = a // 3
            d2_next = keys[1].index((a, a), 4 * t)          # the key written again by stage t (0-based stage index)
            assert H[a, a] == 0.2 + vals[1][d2_next]
    assert differs == B


def random_case(rng, n, m):
    """A random upper-triangular pattern with full diagonal, a random m x n pattern, and key lists with repeats."""
    dense = rng.random((n, n)) < 0.3
    dense = np.triu(dense) | np.eye(n, dtype=bool)
    Wp, Wi = [0], []
    for j in range(n):
        rows = np.nonzero(dense[:, j])[0]
        Wi += list(rows)
        Wp.append(len(Wi))
    gd = rng.random((m, n)) < 0.4
    gd[0, 0] = True
    Gp, Gi = [0], []
    for j in range(n):
        rows = np.nonzero(gd[:, j])[0]
        Gi += list(rows)
        Gp.append(len(Gi))
    wkeys = [(i, j) for j in range(n) for i in range(n) if dense[i, j]]
    gkeys = [(i, j) for j in range(n) for i in range(m) if gd[i, j]]

    def draw(keys, count, mirror):
        out = []
        for _ in range(count):
            i, j = keys[rng.integers(len(keys))]
            out.append((j, i) if mirror and rng.random() < 0.3 else (i, j))   # lower-triangle keys are dropped
        return out

    caches = [draw(wkeys, 3 * len(wkeys), True), draw(wkeys, len(wkeys) // 2, True), draw(wkeys, 5, False)]
    return (np.array(Wp), np.array(Wi), np.array(Gp), np.array(Gi)), caches, draw(gkeys, 2 * len(gkeys), False)


class _Pattern:
    def __init__(self, n, m, Wp, Wi, Gp, Gi):
        self.n, self.m, self.p, self.num_nonnegative, self.soc_dims = n, m, 0, 0, np.zeros(0, np.int32)
        self.W_colptr, self.W_rowval, self.G_colptr, self.G_rowval = Wp, Wi, Gp, Gi
        self.C_colptr, self.C_rowval = np.zeros(n + 1, np.int32), np.zeros(0, np.int32)


@pytest.mark.parametrize("backend", backends.BACKENDS)
@pytest.mark.parametrize("seed,n,m", [(0, 9, 4), (1, 40, 17), (2, 150, 60)])
def test_random_duplicates_exact(backend, seed, n, m):
    rng = np.random.default_rng(seed)
    (Wp, Wi, Gp, Gi), wc, gc = random_case(rng, n, m)
    B = 5
    k = BatchKKT(_Pattern(n, m, Wp, Wi, Gp, Gi), batch=B, binding=backends.binding(backend))
    k.scatter_plan("W_VALUES", wc)
    k.scatter_plan("G_VALUES", [gc])
    L = sum(len(c) for c in wc)
    vals = rng.standard_normal((B, L))
    gv = rng.standard_normal((B, len(gc)))
    sentinel = np.full((B, len(Wi)), 7.25)
    k.set("W_VALUES", sentinel)
    k.scatter("W_VALUES", vals[1:4], first=1)              # instances 1..3 only
    k.scatter("G_VALUES", gv)
    W, G = k.get("W_VALUES"), k.get("G_VALUES")
    assert np.array_equal(W[0], sentinel[0]) and np.array_equal(W[4], sentinel[4])
    offs = np.cumsum([0] + [len(c) for c in wc])
    for b in range(B):
        Gd = orc.scatter_dense((m, n), gc, gv[b])
        assert np.array_equal(G[b], orc.values_at_pattern(Gd, Gp, Gi))
        if 1 <= b <= 3:
            upper = [[(r, c) for (r, c) in ks if r <= c] for ks in wc]
            uvals = [[vals[b][offs[q] + i] for i, (r, c) in enumerate(ks) if r <= c] for q, ks in enumerate(wc)]
            H = orc.lagrangian_hessian_dense(n, upper, uvals)
            assert np.array_equal(W[b], orc.values_at_pattern(H, Wp, Wi))


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_scattered_values_reach_the_path(backend):
    """J v after a scatter equals J v after cb200_set_array of the same values (the row-ordered copies are refreshed)."""
    rng = np.random.default_rng(11)
    (Wp, Wi, Gp, Gi), wc, gc = random_case(rng, 12, 5)
    pat = _Pattern(12, 5, Wp, Wi, Gp, Gi)
    ka = BatchKKT(pat, batch=2, binding=backends.binding(backend))
    kb = BatchKKT(pat, batch=2, binding=backends.binding(backend))
    ka.scatter_plan("W_VALUES", wc[:2])
    ka.scatter_plan("G_VALUES", [gc])
    v = rng.standard_normal((2, ka.total))
    for rep in range(2):                                   # the second pass changes values after a first use
        wv = rng.standard_normal((2, len(wc[0]) + len(wc[1])))
        gv = rng.standard_normal((2, len(gc)))
        ka.scatter("W_VALUES", wv)
        ka.scatter("G_VALUES", gv)
        kb.set("W_VALUES", ka.get("W_VALUES"))
        kb.set("G_VALUES", ka.get("G_VALUES"))
        assert np.array_equal(ka.jacobian_times(v), kb.jacobian_times(v))


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_scatter_errors(backend):
    rng = np.random.default_rng(4)
    (Wp, Wi, Gp, Gi), wc, gc = random_case(rng, 9, 4)
    k = BatchKKT(_Pattern(9, 4, Wp, Wi, Gp, Gi), batch=1, binding=backends.binding(backend))
    with pytest.raises(_lib.CalipsoB200Error, match="no scatter plan"):
        k._scatter_len = {"W_VALUES": 1}
        k.scatter("W_VALUES", np.zeros((1, 1)))
    missing = next((i, j) for j in range(9) for i in range(j) if i not in Wi[Wp[j]:Wp[j + 1]])
    with pytest.raises(_lib.CalipsoB200Error, match="not in the pattern"):
        k.scatter_plan("W_VALUES", [[missing]])
    with pytest.raises(_lib.CalipsoB200Error, match="out of range"):
        k.scatter_plan("G_VALUES", [[(4, 0)]])
    with pytest.raises(_lib.CalipsoB200Error, match="one cache"):
        k.scatter_plan("G_VALUES", [gc, gc])
    with pytest.raises(_lib.CalipsoB200Error):
        k.scatter_plan("POINT", [[(0, 0)]])


@pytest.mark.parametrize("backend", backends.BACKENDS)
def test_a_new_plan_replaces_the_previous_one(backend):
    rng = np.random.default_rng(8)
    (Wp, Wi, Gp, Gi), wc, gc = random_case(rng, 10, 4)
    k = BatchKKT(_Pattern(10, 4, Wp, Wi, Gp, Gi), batch=2, binding=backends.binding(backend))
    for caches in (wc, wc[:1], [wc[1], wc[0]]):
        k.scatter_plan("W_VALUES", caches)
        vals = rng.standard_normal((2, sum(len(c) for c in caches)))
        k.scatter("W_VALUES", vals)
        W = k.get("W_VALUES")
        offs = np.cumsum([0] + [len(c) for c in caches])
        for b in range(2):
            upper = [[(r, c) for (r, c) in ks if r <= c] for ks in caches]
            uvals = [[vals[b][offs[q] + i] for i, (r, c) in enumerate(ks) if r <= c] for q, ks in enumerate(caches)]
            H = orc.lagrangian_hessian_dense(10, upper, uvals)
            assert np.array_equal(W[b], orc.values_at_pattern(H, Wp, Wi))
    with pytest.raises(_lib.CalipsoB200Error):                       # a failed plan leaves no plan behind
        k.scatter_plan("W_VALUES", [[(0, 10)]])
    with pytest.raises(_lib.CalipsoB200Error, match="no scatter plan"):
        k.scatter("W_VALUES", vals)
