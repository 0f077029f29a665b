"""Host-side mirror of the reference's solver interface over the B200 hot path.

Names follow src/solver of CALIPSO.jl (``Solver``, ``solve!`` -> :func:`solve`, ``initialize!`` -> :func:`initialize`,
``Options``; ``ldl_solver``/``factorize!``/``compute_inertia!``/``linear_solve!`` -> :class:`LDLSolver`).  The Julia
reference cannot run in this environment, so the host side is Python; the data-parallel work is done by the CUDA
library behind include/calipso_b200.h and nothing here computes a search direction, a factorisation or a cone test on
the CPU.  What stays on the host is what the reference also keeps in scalar Julia: user callbacks (``evaluate!``),
the filter (src/solver/filter.jl) and the outer-loop bookkeeping of solve.jl.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, fields

import numpy as np

from . import _lib
from ._lib import A, I, I_COUNT, S, S_COUNT, dp, f64, i32, ip


@dataclass
class Options:
    """src/solver/options.jl:6-59 (fields the hot path or solve! reads; same defaults)."""
    max_outer_iterations: int = 10
    max_residual_iterations: int = 100
    max_residual_line_search: int = 25
    max_cone_line_search: int = 25
    iterative_refinement: bool = True
    max_iterative_refinement: int = 10
    min_iterative_refinement: int = 1
    scaling_line_search: float = 0.5
    iterative_refinement_tolerance: float = 1.0e-10
    central_path_initial: float = 1.0
    central_path_update_tolerance: float = 10.0
    central_path_scaling: float = 0.2
    central_path_exponent: float = 1.5
    penalty_initial: float = 1.0
    penalty_scaling: float = 10.0
    dual_initial: float = 0.0
    residual_tolerance: float = 1.0e-4
    optimality_tolerance: float = 1.0e-4
    slack_tolerance: float = 1.0e-4
    equality_tolerance: float = 1.0e-4
    complementarity_tolerance: float = 1.0e-4
    min_regularization: float = 1.0e-20
    primal_regularization_initial: float = 1.0e-7
    dual_regularization_initial: float = 1.0e-7
    max_regularization: float = 1.0e40
    dual_regularization: float = 1.0e-8
    dual_regularization_exponent: float = 0.25
    scaling_regularization_initial: float = 100.0
    scaling_regularization: float = 8.0
    scaling_regularization_last: float = 1.0 / 3.0
    max_penalty: float = 1.0e8
    violation_tolerance: float = 1.0e-5
    violation_exponent: float = 1.1
    merit_tolerance: float = 1.0e-5
    merit_exponent: float = 2.3
    armijo_tolerance: float = 1.0e-4
    machine_tolerance: float = 1.0e-16
    max_filter: int = 1000
    gmres_restart: int = 30          # fallback for the reference's UMFPACK `J \ R` (search_direction.jl:22)
    gmres_max_cycles: int = 10
    warmstart: bool = False
    verbose: bool = False

    def to_c(self) -> _lib.COptions:
        o = _lib.COptions()
        names = {f[0] for f in _lib.COptions._fields_}
        for f in fields(self):
            if f.name in names:
                setattr(o, f.name, getattr(self, f.name))
        return o


class BatchKKT:
    """A batch of problem instances with one sparsity pattern on one GPU: thin wrapper of the C ABI handle."""

    def __init__(self, problem, batch: int = 1, perm=None, options: Options | None = None, device: int = 0,
                 binding: _lib.Binding | None = None):
        self.b = binding or _lib.default_binding()
        self.lib = self.b.lib
        self.problem = problem
        self.batch = batch
        self.options = options or Options()
        P = problem
        self.n, self.m, self.p = P.n, P.m, P.p
        self.N, self.total = P.n + P.m + P.p, P.n + 2 * P.m + 3 * P.p
        soc = i32(P.soc_dims)
        arrs = [i32(a) for a in (P.W_colptr, P.W_rowval, P.G_colptr, P.G_rowval, P.C_colptr, P.C_rowval)]
        pp = i32(perm) if perm is not None else None
        copt = self.options.to_c()
        self.h = self.lib.cb200_create(batch, P.n, P.m, P.p, P.num_nonnegative, len(soc), ip(soc), *[ip(a) for a in arrs],
                                       ip(pp) if pp is not None else None, C.byref(copt), device)
        if not self.h:
            raise _lib.CalipsoB200Error(self.lib.cb200_last_error().decode())
        n, m, p = self.n, self.m, self.p
        # slices of w = (x, r, s, y, z, t), src/solver/indices.jl:25-35
        self.ix, self.ir, self.is_ = slice(0, n), slice(n, n + m), slice(n + m, n + m + p)
        self.iy, self.iz = slice(n + m + p, n + 2 * m + p), slice(n + 2 * m + p, n + 2 * m + 2 * p)
        self.it = slice(n + 2 * m + 2 * p, n + 2 * m + 3 * p)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cb200_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # ---- data movement
    def length(self, name):
        return self.lib.cb200_array_length(self.h, A[name])

    def set(self, name, values, first=0):
        v = f64(values)
        ln = self.length(name)
        if ln <= 0:
            if v.size:
                raise _lib.CalipsoB200Error(f"array {name} is not available on this handle")
            return
        v = v.reshape(-1, ln)
        self.b.check(self.lib.cb200_set_array(self.h, A[name], dp(v), first, v.shape[0]))
        self.b.check(self.lib.cb200_synchronize(self.h))   # the source buffer may be a temporary

    def set_all(self, name, values):
        """Broadcast one instance's values to the whole batch."""
        v = f64(values).reshape(1, -1)
        self.set(name, np.repeat(v, self.batch, axis=0))

    def get(self, name, first=0, count=None):
        count = self.batch - first if count is None else count
        ln = self.length(name)
        if ln < 0 and name in ("RHS", "MATRIX_VALUES"):
            raise _lib.CalipsoB200Error(f"array {name} is not available on this handle")
        out = np.zeros((count, max(ln, 0)))
        if ln > 0:
            self.b.check(self.lib.cb200_get_array(self.h, A[name], dp(out), first, count))
        return out

    def scalars(self):
        s = self.get("SCALARS")
        return {k: s[:, i].copy() for k, i in S.items()}

    def set_scalars(self, **kw):
        s = self.get("SCALARS")
        for k, v in kw.items():
            s[:, S[k]] = v
        self.set("SCALARS", s)

    def stats(self):
        out = np.zeros((self.batch, I_COUNT), dtype=np.int32)
        self.b.check(self.lib.cb200_get_stats(self.h, ip(out), 0, self.batch))
        return {k: out[:, i].copy() for k, i in I.items()}

    def info(self):
        out = np.zeros(16, dtype=np.int64)
        self.lib.cb200_info(self.h, out.ctypes.data_as(_lib.c_llp))
        keys = ("N", "total", "nnzK", "nnzL", "supernodes", "levels", "phases", "max_width", "max_rows", "panel_total",
                "sum_lnz_sq", "batch", "n", "m", "p", "nnz_inputs")
        return dict(zip(keys, out.tolist()))

    def symbolic(self):
        perm, etree, lnz = (np.zeros(self.N, dtype=np.int32) for _ in range(3))
        self.lib.cb200_get_symbolic(self.h, ip(perm), ip(etree), ip(lnz))
        return perm, etree, lnz

    def paths(self):
        """Which code paths the pattern selected (cb200_path_info): speed, not results."""
        out = np.zeros(8, dtype=np.int64)
        self.lib.cb200_path_info(self.h, out.ctypes.data_as(_lib.c_llp))
        keys = ("solve_in_shared_memory", "ctas_per_sm", "dynamic_smem_bytes", "cta_supernodes", "cta_supernodes_generic",
                "chain_descriptors_cached", "schedule_cached", "threads_per_cta")
        return dict(zip(keys, out.tolist()))

    def factor(self, instance=0):
        nnzL = self.info()["nnzL"]
        Lp, Li = np.zeros(self.N + 1, dtype=np.int32), np.zeros(nnzL, dtype=np.int32)
        Lx, D = np.zeros(nnzL), np.zeros(self.N)
        self.b.check(self.lib.cb200_get_factor(self.h, instance, ip(Lp), ip(Li), dp(Lx), dp(D)))
        return Lp, Li, Lx, D

    def synchronize(self):
        self.b.check(self.lib.cb200_synchronize(self.h))

    def profile(self, reset=True):
        """Cycle counters of the device phases, summed over the batch (diagnostic)."""
        out = np.zeros((self.batch, _lib.PROF_COUNT), dtype=np.int64)
        self.b.check(self.lib.cb200_get_profile(self.h, out.ctypes.data_as(_lib.c_llp), int(reset)))
        return {k: int(out[:, i].sum()) for i, k in enumerate(_lib.PROFILE)}

    def set_options(self, options: Options):
        self.options = options
        self.b.check(self.lib.cb200_set_options(self.h, C.byref(options.to_c())))

    # ---- hot path (reference function names)
    def cone(self, barrier=False, barrier_gradient=False, product=False, at_candidate=False):
        """cone!(problem, methods, idx, solution; ...), cones/cone.jl:71-106"""
        flags = (1 if barrier else 0) | (2 if barrier_gradient else 0) | (4 if product else 0)
        self.b.check(self.lib.cb200_cone(self.h, flags, int(at_candidate)))

    def residual(self):
        """residual!(...), residual.jl:1-51 + the reductions of solve.jl:130-135"""
        self.b.check(self.lib.cb200_residual(self.h))

    def search_direction(self):
        """search_direction!(solver), search_direction.jl:1-23"""
        self.b.check(self.lib.cb200_search_direction(self.h))

    def cone_search(self):
        """cone line search, solve.jl:190-221"""
        self.b.check(self.lib.cb200_cone_search(self.h))

    def apply_step(self):
        """step update, solve.jl:309-333"""
        self.b.check(self.lib.cb200_apply_step(self.h))

    def kkt_factor_solve(self, nsolves=1):
        self.b.check(self.lib.cb200_kkt_factor_solve(self.h, nsolves))

    def differentiate(self, jacobian_parameters):
        """differentiate!(solver), src/solver/differentiate.jl:1-61: solution sensitivities dw/dtheta at the current point.
        jacobian_parameters: [batch, total, num_parameters] (or [total, num_parameters] for batch 1) = dR/dtheta
        (residual_jacobian_parameters!); returns solution_sensitivity in the same shape."""
        H = f64(jacobian_parameters)
        if H.ndim == 2:
            H = H[None]
        assert H.shape[0] == self.batch and H.shape[1] == self.total
        nparam = H.shape[2]
        Hc = np.ascontiguousarray(np.transpose(H, (0, 2, 1)))          # column i contiguous
        out = np.zeros_like(Hc)
        self.b.check(self.lib.cb200_differentiate(self.h, nparam, dp(Hc), dp(out)))
        S = np.transpose(out, (0, 2, 1))
        return S[0] if np.ndim(jacobian_parameters) == 2 else S

    def scatter_plan(self, name, sparsities):
        """The gather plan for evaluate!'s `=` scatter of the flat derivative caches (src/solver/evaluate.jl:37-42,55-60,
        73-78,95-100,109-114).  name: "W_VALUES" (sparsities = up to three key lists: objective, equality-dual,
        cone-dual Hessian caches), "G_VALUES" or "C_VALUES" (one key list).  A key list is a sequence of 0-based
        (row, col) pairs in cache order -- the reference's `*_sparsity` vectors; repeated keys: the last entry wins."""
        lens = i32([len(sp) for sp in sparsities])
        keys = [np.asarray(sp, dtype=np.int64).reshape(-1, 2) for sp in sparsities]
        allk = np.concatenate(keys, axis=0) if keys else np.zeros((0, 2), np.int64)
        rows, cols = i32(allk[:, 0]), i32(allk[:, 1])
        self.b.check(self.lib.cb200_scatter_plan(self.h, A[name], len(lens), ip(lens), ip(rows), ip(cols)))
        self._scatter_len = getattr(self, "_scatter_len", {})
        self._scatter_len[name] = int(lens.sum())

    def scatter(self, name, caches, first=0):
        """Scatter the flat caches ([count, sum of cache lengths], the caches of one instance concatenated in plan order)
        into the value array `name` of instances first .. first+count-1."""
        L = self._scatter_len[name]
        v = f64(caches).reshape(-1, L) if L > 0 else np.zeros((np.shape(caches)[0] if np.ndim(caches) > 1 else 1, 0))
        self.b.check(self.lib.cb200_scatter(self.h, A[name], dp(v) if L > 0 else None, first, v.shape[0]))
        self.b.check(self.lib.cb200_synchronize(self.h))   # the source buffer may be a temporary

    def stage_plan(self, name, indices, accumulate):
        """The gather plan of a stage-level scatter of the trajectory-optimisation front end (cb200_stage_plan;
        trajectory_optimization/evaluate.jl:15-28,77-136,206-241,297-327).  name: "GRADIENT", "EQ_DUAL_GRAD", "CONE_DUAL_GRAD"
        (accumulate=True: `gradient[idx] += cache[i]`) or "EQUALITY", "CONE" (accumulate=False: `violations[idx] .= cache`).
        indices: the stages' 0-based index lists in program order (a list of lists, or one flat list)."""
        flat = [np.asarray(ix, dtype=np.int64).ravel() for ix in indices] if len(indices) and np.ndim(indices[0]) > 0 else \
            [np.asarray(indices, dtype=np.int64).ravel()]
        dst = i32(np.concatenate(flat)) if flat else i32([])
        self.b.check(self.lib.cb200_stage_plan(self.h, A[name], int(bool(accumulate)), len(dst), ip(dst) if len(dst) else None))
        self._stage_len = getattr(self, "_stage_len", {})
        self._stage_len[name] = len(dst)

    def stage_scatter(self, name, caches, first=0):
        """Scatter the concatenated stage caches ([count, plan length]) into the vector `name` of instances
        first .. first+count-1."""
        L = self._stage_len[name]
        v = f64(caches).reshape(-1, L) if L > 0 else np.zeros((np.shape(caches)[0] if np.ndim(caches) > 1 else 1, 0))
        self.b.check(self.lib.cb200_stage_scatter(self.h, A[name], dp(v) if L > 0 else None, first, v.shape[0]))
        self.b.check(self.lib.cb200_synchronize(self.h))   # the source buffer may be a temporary

    def jacobian_times(self, v):
        v = f64(v).reshape(self.batch, self.total)
        out = np.zeros_like(v)
        self.b.check(self.lib.cb200_jacobian_times(self.h, dp(v), dp(out)))
        return out

    # ---- LQ-conic family, callbacks on the device
    def load_lq(self, problems):
        """problems: one ConicProblem (broadcast) or a list of `batch` problems sharing the pattern."""
        ps = problems if isinstance(problems, (list, tuple)) else [problems] * self.batch
        assert len(ps) == self.batch
        for name, attr in (("W_VALUES", "W_val"), ("G_VALUES", "G_val"), ("C_VALUES", "C_val"), ("LQ_Q", "q"),
                           ("LQ_G0", "g0"), ("LQ_H0", "h0")):
            self.set(name, np.stack([f64(getattr(P, attr)) for P in ps]))

    def initialize(self, guesses):
        """initialize!(solver, guess), initialize.jl:9-13, per instance"""
        g = f64(guesses).reshape(-1, self.n)
        if g.shape[0] == 1 and self.batch > 1:
            g = np.repeat(g, self.batch, axis=0)
        assert g.shape[0] == self.batch
        self.b.check(self.lib.cb200_initialize(self.h, dp(g), 0, self.batch))     # the primal block only
        self.b.check(self.lib.cb200_synchronize(self.h))   # the source buffer may be a temporary

    def lq_evaluate(self, flags, at_candidate=False):
        self.b.check(self.lib.cb200_lq_evaluate(self.h, flags, int(at_candidate)))

    def lq_begin(self, warmstart=False):
        self.b.check(self.lib.cb200_lq_begin(self.h, int(warmstart)))

    def lq_step(self, iterations=1):
        self.b.check(self.lib.cb200_lq_step(self.h, iterations))

    def filter_reset(self):
        """reset!(solver.filter), solve.jl:93,367 (the filter of cb200_filter_search lives on the device)"""
        self.b.check(self.lib.cb200_filter_reset(self.h))

    def filter_search(self, first, f, g, h):
        """Filter line search over candidates first .. first+count-1 (solve.jl:224-306): f [batch, count], g [batch, count, m],
        h [batch, count, p] = evaluate!'s outputs at w - alpha_cone 0.5^k step.  Returns the accepted index per instance or -1."""
        f = f64(f).reshape(self.batch, -1)
        count = f.shape[1]
        g = f64(g).reshape(self.batch, count, self.m)
        h = f64(h).reshape(self.batch, count, self.p)
        acc = np.zeros(self.batch, dtype=np.int32)
        self.b.check(self.lib.cb200_filter_search(self.h, first, count, dp(f), dp(g), dp(h), ip(acc)))
        return acc

    def lq_set_order(self, order=None):
        """Scheduling hint: start the instances in this order (a permutation of range(batch); None: identity).  Results do
        not depend on it; `np.argsort(-previous_iteration_counts, kind="stable")` packs the tail of a batched re-solve."""
        o = i32(order) if order is not None else None
        self.b.check(self.lib.cb200_lq_set_order(self.h, ip(o) if o is not None else None))

    def lq_solve(self, max_steps=1100, check_every=32):
        counts = np.zeros(4, dtype=np.int64)
        steps = C.c_int(0)
        self.b.check(self.lib.cb200_lq_solve(self.h, max_steps, check_every, counts.ctypes.data_as(_lib.c_llp),
                                             C.byref(steps)))
        return dict(running=int(counts[0]), converged=int(counts[1]), gave_up=int(counts[2]), error=int(counts[3]),
                    steps=steps.value)

    # ---- multi-GPU: the only exchange is the convergence-count all-reduce
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self.b.check(self.lib.cb200_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, rank, nranks, unique_id: bytes):
        self.b.check(self.lib.cb200_comm_init(self.h, rank, nranks, unique_id))

    def allreduce_counts(self):
        counts = np.zeros(4, dtype=np.int64)
        self.b.check(self.lib.cb200_allreduce_counts(self.h, counts.ctypes.data_as(_lib.c_llp)))
        return dict(running=int(counts[0]), converged=int(counts[1]), gave_up=int(counts[2]), error=int(counts[3]))


# ---------------------------------------------------------------------------------------------------------------------
class _EvalOut:
    pass


class Filter:
    """src/solver/filter.jl:1-89 (host scalar logic, as in the reference)."""

    def __init__(self, max_length=1000):
        self.pairs = [(1.0e8, 1.0e8)] * max_length
        self.index = 0

    def reset(self):
        for i in range(self.index):
            self.pairs[i] = (1.0e8, 1.0e8)
        self.index = 0

    def check(self, cv, merit):
        return all((cv < f[0] or merit < f[1]) for f in self.pairs)

    def augment(self, cv, merit):
        if self.index == 0:
            self.pairs[0] = (cv, merit)
            self.index = 1
        elif self.check(cv, merit):
            cache = self.pairs[:self.index]
            self.reset()
            self.index = 1
            self.pairs[0] = (cv, merit)
            for c in cache:
                if not (c[0] >= cv and c[1] >= merit):
                    self.pairs[self.index] = c
                    self.index += 1


class Solver:
    """Solver(methods, num_variables, num_parameters, num_equality, num_cone; ...), src/solver/solver.jl:46-150.

    ``problem`` carries the structural patterns (see calipso_b200.lqc.ConicProblem / tests/problems.DenseNLP);
    ``callback(flags, x, y, z, out)`` plays the role of ``ProblemMethods`` + ``evaluate!`` (src/solver/evaluate.jl).
    One instance per Solver (batch = 1), like the reference.
    """

    def __init__(self, problem, callback, options: Options | None = None, device: int = 0, perm=None, binding=None,
                 line_search_on_device: bool = False):
        self.options = options or Options()
        self.problem = problem
        self.callback = callback
        self.line_search_on_device = line_search_on_device   # filter line search through cb200_filter_search
        self.kkt = BatchKKT(problem, batch=1, perm=perm, options=self.options, device=device, binding=binding)
        k = self.kkt
        self.n, self.m, self.p, self.total = k.n, k.m, k.p, k.total
        self.solution = np.zeros(k.total)          # solver.solution.all (host mirror)
        self.candidate = np.zeros(k.total)
        self.step = np.zeros(k.total)
        self.residual = np.zeros(k.total)
        self.dual = np.zeros(k.m)                  # lambda
        self.central_path, self.fraction_to_boundary, self.penalty = 0.1, 0.99, 10.0   # solver.jl:81-86
        self.filter = Filter(self.options.max_filter)
        nW, nG, nC = len(problem.W_rowval), len(problem.G_rowval), len(problem.C_rowval)
        o = self.out = _EvalOut()
        o.objective, o.gradient = np.zeros(1), np.zeros(k.n)
        o.equality, o.cone = np.zeros(k.m), np.zeros(k.p)
        o.eq_dual_grad, o.cone_dual_grad = np.zeros(k.n), np.zeros(k.n)
        o.W_val, o.G_val, o.C_val = np.zeros(nW), np.zeros(nG), np.zeros(nC)
        self.cone_product = np.zeros(k.p)
        self.iterations = 0
        self.unrefined_steps = 0
        self.log = []

    # evaluate!(problem, methods, idx, point, parameters; flags...)
    def evaluate(self, flags, point):
        k = self.kkt
        self.callback(flags, point[k.ix], point[k.iy], point[k.iz], self.out)

    def _upload_first_order(self):
        k, o = self.kkt, self.out
        k.set("GRADIENT", o.gradient)
        k.set("EQ_DUAL_GRAD", o.eq_dual_grad)
        k.set("CONE_DUAL_GRAD", o.cone_dual_grad)
        k.set("EQUALITY", o.equality)
        k.set("CONE", o.cone)

    def _barrier(self, s):
        P = self.problem
        q = P.num_nonnegative
        phi = float(np.sum(np.log(s[:q]))) if q else 0.0
        off = q
        for d in P.soc_dims:
            if d > 0:
                phi += 0.5 * np.log(s[off] ** 2 - s[off + 1:off + d] @ s[off + 1:off + d])
            off += d
        return phi

    def _merit(self, f, r, phi):
        return f + self.dual @ r + 0.5 * self.penalty * (r @ r) - self.central_path * phi   # merit.jl:2-15

    def _theta(self, g, r, h, s):
        return (np.abs(g - r).sum() + np.abs(h - s).sum()) / (self.m + self.p)            # constraint_violation.jl


def initialize(solver: Solver, guess):
    """initialize!(solver, guess), src/solver/initialize.jl:9-13"""
    solver.solution[:solver.n] = guess


def solve(solver: Solver) -> bool:
    """solve!(solver), src/solver/solve.jl:8-377, with the data-parallel steps on the GPU."""
    s_, o, opt, k, P = solver, solver.out, solver.options, solver.kkt, solver.problem
    n, m, p = s_.n, s_.m, s_.p
    w = s_.solution
    ix, ir, isl, iy, iz, it = k.ix, k.ir, k.is_, k.iy, k.iz, k.it
    E = _lib
    if not opt.warmstart:                                           # initialize_slacks!/duals!, initialize.jl:15-36
        s_.evaluate(E.EV_EQUALITY | E.EV_CONE, w)
        w[ir] = o.equality
        w[iy] = 0.0
        w[iz] = 0.0
        init = np.ones(p)
        off = P.num_nonnegative
        for d in P.soc_dims:
            init[off + 1:off + d] = 0.1
            off += d
        w[isl] = init
        w[it] = init
    s_.central_path = opt.central_path_initial
    s_.fraction_to_boundary = max(0.99, 1.0 - s_.central_path)
    s_.penalty = opt.penalty_initial
    s_.dual[:] = opt.dual_initial
    s_.evaluate(E.EV_OBJECTIVE | E.EV_EQUALITY | E.EV_EQUALITY_JAC | E.EV_CONE, w)
    equality_violation = np.abs(o.equality).max(initial=0.0)
    cone_product_violation = np.abs(s_.cone_product).max(initial=0.0)   # before cone!(product), solve.jl:85-91
    k.set("POINT", w)
    k.cone(product=True)
    s_.filter.reset()
    if s_.line_search_on_device:
        k.filter_reset()
    total_iterations = 1
    for j in range(1, opt.max_outer_iterations + 1):
        for i in range(1, opt.max_residual_iterations + 1):
            s_.evaluate(E.EV_GRADIENT | E.EV_EQUALITY_DUAL_GRAD | E.EV_CONE_DUAL_GRAD, w)
            k.set("POINT", w)
            k.set("DUAL", s_.dual)
            k.set_scalars(kappa=s_.central_path, tau=s_.fraction_to_boundary, rho=s_.penalty)
            s_._upload_first_order()
            k.cone(barrier=True, barrier_gradient=True)
            k.residual()
            sc = {kk: v[0] for kk, v in k.scalars().items()}
            phi = sc["barrier"]
            bgrad = k.get("BARRIER_GRADIENT")[0]
            M = s_._merit(o.objective[0], w[ir], phi)
            merit_gradient = np.concatenate([o.gradient, s_.dual + s_.penalty * w[ir], -s_.central_path * bgrad])
            if (sc["residual_violation"] < opt.residual_tolerance and sc["slack_violation"] < opt.slack_tolerance and
                    equality_violation <= opt.equality_tolerance and
                    cone_product_violation <= opt.complementarity_tolerance):
                s_.residual = k.get("RESIDUAL")[0]
                s_.cone_product = k.get("CONE_PRODUCT")[0]
                s_.iterations = total_iterations
                return True
            elif sc["optimality_violation"] <= max(opt.central_path_update_tolerance * s_.central_path,
                                                   opt.optimality_tolerance):
                break
            theta = s_._theta(o.equality, w[ir], o.cone, w[isl])
            s_.evaluate(E.EV_HESSIAN | E.EV_EQUALITY_JAC | E.EV_CONE_JAC, w)
            k.set("W_VALUES", o.W_val)
            k.set("G_VALUES", o.G_val)
            k.set("C_VALUES", o.C_val)
            k.search_direction()
            k.cone_search()
            st = {kk: int(v[0]) for kk, v in k.stats().items()}
            if st["status"] == 1:
                raise RuntimeError("inertia correction failure")          # inertia.jl:72
            if st["status"] == 3:
                raise RuntimeError("cone search failure")                 # solve.jl:210,220
            if st["status"] == 2:
                # refinement and the fallback both missed the tolerance: the reference takes the `J \ R` step as it
                # comes (search_direction.jl:22); the step is taken here too, and the event is made visible
                import warnings
                warnings.warn("iterative refinement failure: search direction not refined to tolerance")
                s_.unrefined_steps += 1
            sc = {kk: v[0] for kk, v in k.scalars().items()}
            step_size = sc["step_size"]
            step = k.get("STEP")[0]
            cand = k.get("CANDIDATE")[0]
            if s_.line_search_on_device:
                # the filter line search on the device: evaluate! at blocks of candidate step sizes (1, 2, 4, ... at a time),
                # merit / violation / filter tests in cb200_filter_search; the host only runs the callbacks
                k.set_scalars(objective=o.objective[0])
                first, count, accepted = 0, 1, -1
                while accepted < 0:
                    fs, gs, hs = [], [], []
                    for kk in range(first, first + count):
                        xc = w[:n] - step_size * opt.scaling_line_search ** kk * step[:n]
                        s_.callback(E.EV_OBJECTIVE | E.EV_EQUALITY | E.EV_CONE, xc, w[iy], w[iz], o)
                        fs.append(o.objective[0]); gs.append(o.equality.copy()); hs.append(o.cone.copy())
                    accepted = int(k.filter_search(first, np.array(fs)[None], np.array(gs)[None] if m else np.zeros((1, count, 0)),
                                                   np.array(hs)[None] if p else np.zeros((1, count, 0)))[0])
                    if accepted >= 0:
                        o.objective[0] = fs[accepted - first]
                        o.equality[:] = gs[accepted - first]
                        o.cone[:] = hs[accepted - first]
                    first, count = first + count, 2 * count
                k.apply_step()
                w[:] = k.get("POINT")[0]
                sc = {kk: v[0] for kk, v in k.scalars().items()}
                equality_violation = sc["equality_violation"]
                cone_product_violation = sc["cone_product_violation"]
                s_.log.append(dict(st, iteration=total_iterations, outer=j, inner=i, step_size=sc["step_size"]))
                total_iterations += 1
                continue
            s_.evaluate(E.EV_OBJECTIVE | E.EV_EQUALITY | E.EV_CONE, cand)
            M_hat = s_._merit(o.objective[0], cand[ir], s_._barrier(cand[isl]))
            theta_hat = s_._theta(o.equality, cand[ir], o.cone, cand[isl])
            d = float(merit_gradient @ step[:n + m + p])
            switching = lambda a: d < 0.0 and a * (-d) ** opt.merit_exponent > theta ** opt.violation_exponent
            armijo = lambda a, Mc: Mc - M - 10.0 * opt.machine_tolerance * abs(M) <= opt.armijo_tolerance * a * d
            residual_iteration = 0
            while residual_iteration < opt.max_residual_line_search:
                if s_.filter.check(theta_hat, M_hat):
                    if theta <= opt.slack_tolerance and switching(step_size) and armijo(step_size, M_hat):
                        break
                    elif (theta_hat - 10.0 * opt.machine_tolerance * abs(theta) <= (1.0 - opt.violation_tolerance) * theta
                          or M_hat - 10.0 * opt.machine_tolerance * abs(M) <= M - opt.merit_tolerance * theta):
                        break
                step_size = opt.scaling_line_search * step_size
                cand[:n + m + p] = w[:n + m + p] - step_size * step[:n + m + p]
                s_.evaluate(E.EV_OBJECTIVE | E.EV_EQUALITY | E.EV_CONE, cand)
                M_hat = s_._merit(o.objective[0], cand[ir], s_._barrier(cand[isl]))
                theta_hat = s_._theta(o.equality, cand[ir], o.cone, cand[isl])
                residual_iteration += 1
            if not switching(step_size) or not armijo(step_size, M_hat):
                s_.filter.augment((1.0 - opt.violation_tolerance) * theta, M - opt.merit_tolerance * theta)
            k.set_scalars(step_size=step_size)
            k.set("EQUALITY", o.equality)          # g(x_hat): apply_step takes ||g||_inf from it
            k.apply_step()
            w[:] = k.get("POINT")[0]
            sc = {kk: v[0] for kk, v in k.scalars().items()}
            equality_violation = sc["equality_violation"]
            cone_product_violation = sc["cone_product_violation"]
            s_.log.append(dict(st, iteration=total_iterations, outer=j, inner=i, step_size=step_size))
            total_iterations += 1
        s_.central_path = max(opt.residual_tolerance / 10.0,
                              min(opt.central_path_scaling * s_.central_path, s_.central_path ** opt.central_path_exponent))
        s_.fraction_to_boundary = max(0.99, 1.0 - s_.central_path)
        s_.dual[:] = s_.dual + s_.penalty * w[ir]
        s_.penalty = min(max(opt.penalty_scaling * s_.penalty, 1.0 / s_.central_path), opt.max_penalty)
        s_.filter.reset()
        if s_.line_search_on_device:
            k.filter_reset()
    s_.iterations = total_iterations
    return False


def residual_jacobian_parameters(n, m, p, objective_xp, equality_dual_xp, cone_dual_xp, equality_p, cone_p):
    """residual_jacobian_parameters!(data, problem, idx), src/solver/residual_jacobian_parameters.jl:1-40: dR/dtheta
    (total x num_parameters) from the parameter derivatives evaluate! produces -- rows of x: objective + equality-dual +
    cone-dual cross derivatives (added in that order, :8-14), rows of y: equality Jacobian (:23-27), rows of z: cone
    Jacobian (:30-34); the r, s, t rows stay zero.  Pass None for a block that does not exist."""
    blocks = [b for b in (objective_xp, equality_dual_xp, cone_dual_xp, equality_p, cone_p) if b is not None]
    ntheta = np.shape(blocks[0])[1]
    H = np.zeros((n + 2 * m + 3 * p, ntheta))
    for b in (objective_xp, equality_dual_xp, cone_dual_xp):
        if b is not None:
            H[:n] += f64(b)
    if m and equality_p is not None:
        H[n + m + p:n + 2 * m + p] = f64(equality_p)
    if p and cone_p is not None:
        H[n + 2 * m + p:n + 2 * m + 2 * p] = f64(cone_p)
    return H


def differentiate(solver: Solver, jacobian_parameters):
    """differentiate!(solver), src/solver/differentiate.jl:1-61, at the point solve! stopped at: the reduced matrix is
    rebuilt from the second-order data of the last evaluate! (the reference does not re-evaluate it either, :13-20) and
    factored once; solution_sensitivity = -J^-1 dR/dtheta column by column (:35-57)."""
    k = solver.kkt
    k.set("POINT", solver.solution)
    solver.jacobian_parameters = f64(jacobian_parameters)
    solver.solution_sensitivity = k.differentiate(solver.jacobian_parameters)
    return solver.solution_sensitivity


# ---------------------------------------------------------------------------------------------------------------------
class Inertia:
    """src/solver/inertia.jl:1-5"""

    def __init__(self):
        self.positive = self.negative = self.zero = 0


class LDLSolver:
    """LinearSolver seam, src/solver/linear_solver.jl: ldl_solver(A) :46, factorize! :19, compute_inertia! :33,
    linear_solve! :52.  A is a scipy.sparse matrix (any triangle content; only the upper triangle is used, like
    triu! in the reference)."""

    def __init__(self, Amat, perm=None, batch=1, device=0, binding=None):
        import scipy.sparse as sp
        self.b = binding or _lib.default_binding()
        self.lib = self.b.lib
        U = sp.triu(sp.csc_matrix(Amat)).tocsc()
        U.sort_indices()
        self.N = U.shape[0]
        self.batch = batch
        self.colptr, self.rowval = i32(U.indptr), i32(U.indices)
        self.nnz = len(self.rowval)
        pp = i32(perm) if perm is not None else None
        self.h = self.lib.cb200_ldl_create(batch, self.N, ip(self.colptr), ip(self.rowval),
                                           ip(pp) if pp is not None else None, device)
        if not self.h:
            raise _lib.CalipsoB200Error(self.lib.cb200_last_error().decode())
        self.inertia = Inertia()
        self._values = np.tile(f64(U.data), (batch, 1))
        self.factorize(None)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cb200_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _triu_values(self, Amat):
        import scipy.sparse as sp
        if Amat is None:
            return self._values
        if isinstance(Amat, np.ndarray) and (Amat.shape == (self.nnz,) or Amat.shape == (self.batch, self.nnz)) \
                and Amat.shape != (self.N, self.N):
            return np.ascontiguousarray(np.broadcast_to(f64(Amat).reshape(-1, self.nnz), (self.batch, self.nnz)))   # packed values
        U = sp.triu(sp.csc_matrix(Amat)).tocsc()
        U.sort_indices()
        if not (np.array_equal(U.indptr, self.colptr) and np.array_equal(U.indices, self.rowval)):
            raise ValueError("sparsity pattern differs from the one given to ldl_solver")
        return np.tile(f64(U.data), (self.batch, 1))

    def factorize(self, Amat=None, update=True):
        """factorize!(s, A; update)"""
        self._values = self._triu_values(Amat)
        self.b.check(self.lib.cb200_set_array(self.h, A["MATRIX_VALUES"], dp(self._values), 0, self.batch))
        self.b.check(self.lib.cb200_ldl_factorize(self.h))
        self.b.check(self.lib.cb200_synchronize(self.h))

    def compute_inertia(self):
        """compute_inertia!(s); returns the per-instance (positive, negative, zero) table too"""
        out = np.zeros((self.batch, 3), dtype=np.int32)
        self.b.check(self.lib.cb200_ldl_inertia(self.h, ip(out)))
        self.inertia.positive, self.inertia.negative, self.inertia.zero = (int(v) for v in out[0])
        return out

    def linear_solve(self, x, Amat, b, fact=True, update=True):
        """linear_solve!(s, x, A, b; fact, update): x <- A^-1 b in place"""
        bb = f64(b).reshape(self.batch, self.N)
        out = np.zeros_like(bb)
        if fact:
            self._values = self._triu_values(Amat)
        self.b.check(self.lib.cb200_ldl_linear_solve(self.h, dp(self._values), dp(bb), dp(out), int(fact)))
        x[...] = out.reshape(np.shape(x))
        return x

    def symbolic(self):
        perm, etree, lnz = (np.zeros(self.N, dtype=np.int32) for _ in range(3))
        self.lib.cb200_get_symbolic(self.h, ip(perm), ip(etree), ip(lnz))
        return perm, etree, lnz

    def factor(self, instance=0):
        out = np.zeros(16, dtype=np.int64)
        self.lib.cb200_info(self.h, out.ctypes.data_as(_lib.c_llp))
        nnzL = int(out[3])
        Lp, Li = np.zeros(self.N + 1, dtype=np.int32), np.zeros(nnzL, dtype=np.int32)
        Lx, D = np.zeros(nnzL), np.zeros(self.N)
        self.b.check(self.lib.cb200_get_factor(self.h, instance, ip(Lp), ip(Li), dp(Lx), dp(D)))
        return Lp, Li, Lx, D


def ldl_solver(Amat, **kw) -> LDLSolver:
    """ldl_solver(A), src/solver/linear_solver.jl:46-48"""
    return LDLSolver(Amat, **kw)
