"""ctypes binding of the CPU ORACLE (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import this
module; the product package (calipso_b200/) never does.  See oracle/oracle.h for what is restated and for the
parity statement (Julia reference not runnable here; AMD boundary unpinned).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class EvalOut(C.Structure):
    _fields_ = [(k, c_dp) for k in ("objective", "gradient", "equality", "cone", "eq_dual_grad", "cone_dual_grad",
                                    "W_val", "G_val", "C_val")]


class Options(C.Structure):
    """orc_options, mirrors src/solver/options.jl:6-59 (hot-path subset)."""
    _fields_ = ([(k, C.c_int) for k in ("max_outer_iterations", "max_residual_iterations", "max_residual_line_search",
                                        "max_cone_line_search", "iterative_refinement", "max_iterative_refinement",
                                        "min_iterative_refinement")] +
                [(k, C.c_double) for k in (
                    "scaling_line_search", "iterative_refinement_tolerance", "central_path_initial",
                    "central_path_update_tolerance", "central_path_scaling", "central_path_exponent",
                    "penalty_initial", "penalty_scaling", "dual_initial", "residual_tolerance", "optimality_tolerance",
                    "slack_tolerance", "equality_tolerance", "complementarity_tolerance", "min_regularization",
                    "primal_regularization_initial", "dual_regularization_initial", "max_regularization",
                    "dual_regularization", "dual_regularization_exponent", "scaling_regularization_initial",
                    "scaling_regularization", "scaling_regularization_last", "max_penalty", "violation_tolerance",
                    "violation_exponent", "merit_tolerance", "merit_exponent", "armijo_tolerance",
                    "machine_tolerance")] +
                [(k, C.c_int) for k in ("max_filter", "warmstart", "reference_schedule")])


EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, c_dp, c_dp, c_dp, C.POINTER(EvalOut))
LU_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_dp, c_dp)

EV_OBJECTIVE, EV_GRADIENT, EV_EQUALITY, EV_CONE = 1, 2, 4, 8
EV_EQUALITY_DUAL_GRAD, EV_CONE_DUAL_GRAD, EV_HESSIAN, EV_EQUALITY_JAC, EV_CONE_JAC = 16, 32, 64, 128, 256


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("qdldl.c", "solver.c", "amd.c", "oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.orc_qdldl_new.restype = vp
        L.orc_qdldl_new.argtypes = [C.c_int, c_ip, c_ip, c_dp, c_ip]
        L.orc_qdldl_free.argtypes = [vp]
        L.orc_qdldl_refactor.argtypes = [vp, c_dp]
        L.orc_qdldl_solve.argtypes = [vp, c_dp]
        for name in ("n", "nnzL", "nnzA", "positive_inertia"):
            getattr(L, "orc_qdldl_" + name).argtypes = [vp]
            getattr(L, "orc_qdldl_" + name).restype = C.c_int
        L.orc_qdldl_factor_count.argtypes = [vp]
        L.orc_qdldl_factor_count.restype = C.c_longlong
        for name in ("perm", "iperm", "etree", "Lnz", "Lp", "Li", "triuA_colptr", "triuA_rowval", "AtoPAPt"):
            getattr(L, "orc_qdldl_" + name).argtypes = [vp]
            getattr(L, "orc_qdldl_" + name).restype = c_ip
        for name in ("Lx", "D", "Dinv", "triuA_nzval"):
            getattr(L, "orc_qdldl_" + name).argtypes = [vp]
            getattr(L, "orc_qdldl_" + name).restype = c_dp
        L.orc_min_degree.argtypes = [C.c_int, c_ip, c_ip, c_ip]
        L.orc_amd_order.argtypes = [C.c_int, c_ip, c_ip, c_ip, C.c_double, C.c_int]
        L.orc_options_default.argtypes = [C.POINTER(Options)]
        L.orc_solver_new.restype = vp
        L.orc_solver_new.argtypes = [C.c_int] * 5 + [c_ip] * 8 + [C.POINTER(Options)]
        L.orc_solver_free.argtypes = [vp]
        L.orc_solver_set_callback.argtypes = [vp, EVAL_FN, vp]
        L.orc_solver_set_lq.argtypes = [vp] + [c_dp] * 6
        for name in ("solution", "candidate", "step", "residual", "residual_symmetric_vec", "step_symmetric", "dual",
                     "scalars", "cone_product", "cone_target", "barrier_gradient", "merit_gradient"):
            getattr(L, "orc_" + name).argtypes = [vp]
            getattr(L, "orc_" + name).restype = c_dp
        L.orc_problem.argtypes = [vp]
        L.orc_problem.restype = C.POINTER(EvalOut)
        for name in ("inertia", "stats", "K_colptr", "K_rowval"):
            getattr(L, "orc_" + name).argtypes = [vp]
            getattr(L, "orc_" + name).restype = c_ip
        L.orc_K_nzval.argtypes = [vp]
        L.orc_K_nzval.restype = c_dp
        L.orc_K_nnz.argtypes = [vp]
        L.orc_linear_solver.argtypes = [vp]
        L.orc_linear_solver.restype = vp
        L.orc_evaluate.argtypes = [vp, C.c_int, C.c_int]
        L.orc_cone.argtypes = [vp] + [C.c_int] * 6
        for name in ("residual_eval", "residual_jacobian_variables", "residual_jacobian_variables_symmetric",
                     "merit_gradient_eval", "solve_begin", "outer_update"):
            getattr(L, "orc_" + name).argtypes = [vp]
            getattr(L, "orc_" + name).restype = None
        L.orc_residual_symmetric.argtypes = [vp, c_dp]
        for name in ("factorize", "inertia_correction", "search_direction", "cone_search", "solve",
                     "newton_iteration"):
            getattr(L, "orc_" + name).argtypes = [vp]
            getattr(L, "orc_" + name).restype = C.c_int
        L.orc_search_direction_symmetric.argtypes = [vp, c_dp, c_dp, C.c_int]
        L.orc_iterative_refinement.argtypes = [vp, c_dp]
        L.orc_cone_violation.argtypes = [vp, c_dp, c_dp, C.c_double]
        L.orc_jacobian_times.argtypes = [vp, c_dp, c_dp]
        L.orc_dense_jacobian.argtypes = [vp, c_dp]
        L.orc_dense_symmetric.argtypes = [vp, c_dp]
        L.orc_jacobian_coo.argtypes = [vp, c_ip, c_ip, c_dp]
        L.orc_solver_set_lu_fallback.argtypes = [vp, LU_FN, vp]
        L.orc_merit.argtypes = [vp, C.c_int]
        L.orc_merit.restype = C.c_double
        L.orc_constraint_violation.argtypes = [vp, C.c_int]
        L.orc_constraint_violation.restype = C.c_double
        L.orc_optimality_error.argtypes = [vp]
        L.orc_optimality_error.restype = C.c_double
        L.orc_initialize.argtypes = [vp, c_dp]
        _LIB = L
    return _LIB


def _ip(a):
    return a.ctypes.data_as(c_ip)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _view(ptr, n, dtype=np.float64):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,))


class QDLDL:
    """qdldl(A; perm) on an upper-triangular CSC matrix (src/solver/qdldl.jl)."""

    def __init__(self, n, Ap, Ai, Ax, perm=None):
        self.L = lib()
        Ap, Ai, Ax = _i32(Ap), _i32(Ai), _f64(Ax)
        pp = _i32(perm) if perm is not None else None
        self.h = self.L.orc_qdldl_new(n, _ip(Ap), _ip(Ai), _dp(Ax), _ip(pp) if pp is not None else None)
        if not self.h:
            raise ValueError("Input matrix is not upper triangular or has an empty column")
        self.n = n
        self.nnzA = len(Ai)
        self.nnzL = self.L.orc_qdldl_nnzL(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_qdldl_free(self.h)
            self.h = None

    def refactor(self, Ax):
        Ax = _f64(Ax)
        return self.L.orc_qdldl_refactor(self.h, _dp(Ax))

    def solve(self, b):
        x = _f64(b).copy()
        self.L.orc_qdldl_solve(self.h, _dp(x))
        return x

    def arr(self, name):
        sizes = dict(perm=self.n, iperm=self.n, etree=self.n, Lnz=self.n, Lp=self.n + 1, Li=self.nnzL, Lx=self.nnzL,
                     D=self.n, Dinv=self.n, triuA_colptr=self.n + 1, triuA_rowval=self.nnzA, triuA_nzval=self.nnzA,
                     AtoPAPt=self.nnzA)
        return _view(getattr(self.L, "orc_qdldl_" + name)(self.h), sizes[name]).copy()

    @property
    def positive_inertia(self):
        return self.L.orc_qdldl_positive_inertia(self.h)


def min_degree(n, Ap, Ai):
    Ap, Ai = _i32(Ap), _i32(Ai)
    perm = np.zeros(n, dtype=np.int32)
    lib().orc_min_degree(n, _ip(Ap), _ip(Ai), _ip(perm))
    return perm


def amd(n, Ap, Ai, dense=10.0, aggressive=True):
    """amd(A), src/solver/qdldl.jl:135: the restated SuiteSparse AMD ordering (oracle/amd.c) of a CSC pattern."""
    Ap, Ai = _i32(Ap), _i32(Ai)
    perm = np.zeros(n, dtype=np.int32)
    if lib().orc_amd_order(n, _ip(Ap), _ip(Ai), _ip(perm), float(dense), int(aggressive)) != 0:
        raise MemoryError("orc_amd_order failed")
    return perm


class Oracle:
    """Restated ``Solver`` (src/solver/solver.jl) over a structural pattern; see oracle.h."""

    def __init__(self, n, m, p, num_nonnegative, soc_dims, Wp, Wi, Gp, Gi, Cp, Ci, perm=None, options=None):
        self.L = lib()
        self.n, self.m, self.p = n, m, p
        self.N, self.total = n + m + p, n + 2 * m + 3 * p
        self.opt = Options()
        self.L.orc_options_default(C.byref(self.opt))
        for k, v in (options or {}).items():
            setattr(self.opt, k, v)
        soc = _i32(soc_dims)
        arrs = [_i32(a) for a in (Wp, Wi, Gp, Gi, Cp, Ci)]
        pp = _i32(perm) if perm is not None else None
        self.h = self.L.orc_solver_new(n, m, p, num_nonnegative, len(soc), _ip(soc), *[_ip(a) for a in arrs],
                                       _ip(pp) if pp is not None else None, C.byref(self.opt))
        if not self.h:
            raise ValueError("invalid problem description (W must be upper triangular with a full diagonal; "
                             "cone dims must sum to p)")
        self.nnzW, self.nnzG, self.nnzC = len(arrs[1]), len(arrs[3]), len(arrs[5])
        self._keep = None
        L, h = self.L, self.h
        t, N = self.total, self.N
        self.solution = _view(L.orc_solution(h), t)
        self.candidate = _view(L.orc_candidate(h), t)
        self.step = _view(L.orc_step(h), t)
        self.residual = _view(L.orc_residual(h), t)
        self.residual_symmetric = _view(L.orc_residual_symmetric_vec(h), N)
        self.step_symmetric = _view(L.orc_step_symmetric(h), N)
        self.dual = _view(L.orc_dual(h), m)
        self.cone_product = _view(L.orc_cone_product(h), p)
        self.cone_target = _view(L.orc_cone_target(h), p)
        self.barrier_gradient = _view(L.orc_barrier_gradient(h), p)
        self.merit_gradient = _view(L.orc_merit_gradient(h), N)
        pd = L.orc_problem(h).contents
        self.objective = _view(pd.objective, 1)
        self.gradient = _view(pd.gradient, n)
        self.equality = _view(pd.equality, m)
        self.cone = _view(pd.cone, p)
        self.eq_dual_grad = _view(pd.eq_dual_grad, n)
        self.cone_dual_grad = _view(pd.cone_dual_grad, n)
        self.W_val = _view(pd.W_val, self.nnzW)
        self.G_val = _view(pd.G_val, self.nnzG)
        self.C_val = _view(pd.C_val, self.nnzC)
        self._scal = _view(L.orc_scalars(h), 8)
        # index ranges of w = (x, r, s, y, z, t), indices.jl:25-35
        self.ix = slice(0, n)
        self.ir = slice(n, n + m)
        self.is_ = slice(n + m, n + m + p)
        self.iy = slice(n + m + p, n + 2 * m + p)
        self.iz = slice(n + 2 * m + p, n + 2 * m + 2 * p)
        self.it = slice(n + 2 * m + 2 * p, n + 2 * m + 3 * p)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_solver_free(self.h)
            self.h = None

    # --- scalars: kappa, tau, rho, eps_p, eps_d, eps_p_last, objective, barrier
    def scalars(self):
        self.L.orc_scalars(self.h)
        return dict(zip(("kappa", "tau", "rho", "eps_p", "eps_d", "eps_p_last", "objective", "barrier"),
                        self._scal.tolist()))

    def set_scalars(self, **kw):
        names = ("kappa", "tau", "rho", "eps_p", "eps_d", "eps_p_last")
        for k, v in kw.items():
            self._scal[names.index(k)] = v

    @property
    def inertia(self):
        return tuple(_view(self.L.orc_inertia(self.h), 3).tolist())

    @property
    def stats(self):
        v = _view(self.L.orc_stats(self.h), 12).tolist()
        return dict(zip(("n_trials", "n_refine", "refine_ok", "k_s", "k_t", "total_iterations", "outer", "status",
                         "lu_fallbacks", "used_lu"), v))

    def set_lq(self, W_val, G_val, C_val, q, g0, h0):
        a = [_f64(x) for x in (W_val, G_val, C_val, q, g0, h0)]
        self.L.orc_solver_set_lq(self.h, *[_dp(x) for x in a])

    def set_callback(self, pyfunc):
        """pyfunc(flags, x, y, z, out) where out has numpy views named like EvalOut fields."""
        n, m, p = self.n, self.m, self.p
        sizes = dict(objective=1, gradient=n, equality=m, cone=p, eq_dual_grad=n, cone_dual_grad=n,
                     W_val=self.nnzW, G_val=self.nnzG, C_val=self.nnzC)

        class Out:
            pass

        def tramp(user, flags, x, y, z, out):
            o = Out()
            for k, sz in sizes.items():
                setattr(o, k, _view(getattr(out.contents, k), sz))
            pyfunc(flags, _view(x, n), _view(y, m), _view(z, p), o)

        self._keep = EVAL_FN(tramp)
        self.L.orc_solver_set_callback(self.h, self._keep, None)

    # --- reference functions
    def evaluate(self, flags, at_candidate=False):
        self.L.orc_evaluate(self.h, flags, int(at_candidate))

    def cone_eval(self, at_candidate=False, barrier=False, barrier_gradient=False, product=False, jacobian=False,
                  target=False):
        self.L.orc_cone(self.h, int(at_candidate), int(barrier), int(barrier_gradient), int(product), int(jacobian),
                        int(target))

    def residual_eval(self):
        self.L.orc_residual_eval(self.h)

    def residual_jacobian_variables(self):
        self.L.orc_residual_jacobian_variables(self.h)

    def residual_jacobian_variables_symmetric(self):
        self.L.orc_residual_jacobian_variables_symmetric(self.h)

    def residual_symmetric_eval(self, residual=None):
        r = self.residual if residual is None else _f64(residual)
        self.L.orc_residual_symmetric(self.h, _dp(r))

    def factorize(self):
        return self.L.orc_factorize(self.h)

    def inertia_correction(self):
        return self.L.orc_inertia_correction(self.h)

    def search_direction_symmetric(self, step=None, residual=None, factorize=True):
        st = self.step if step is None else step
        r = self.residual if residual is None else residual
        self.L.orc_search_direction_symmetric(self.h, _dp(st), _dp(r), int(factorize))
        return st

    def differentiate(self, jacobian_parameters):
        """differentiate!(solver), src/solver/differentiate.jl:1-61, given dR/dtheta (total x num_parameters,
        residual_jacobian_parameters.jl:1-40): rebuild and factor the reduced matrix at the current point and
        regularisation (:13-20), one search_direction_symmetric! per parameter (it re-factors every time, :35-46;
        same matrix), solution_sensitivity[:, i] = -result (:55-57)."""
        H = np.asarray(jacobian_parameters, dtype=np.float64)
        self.residual_jacobian_variables()
        self.residual_jacobian_variables_symmetric()
        self.factorize()
        S = np.zeros_like(H)
        for i in range(H.shape[1]):
            col = np.ascontiguousarray(H[:, i])
            st = np.zeros(self.total)
            self.search_direction_symmetric(step=st, residual=col, factorize=True)
            S[:, i] = -1.0 * st
        return S

    def iterative_refinement(self, step=None):
        st = self.step if step is None else step
        return bool(self.L.orc_iterative_refinement(self.h, _dp(st)))

    def search_direction(self):
        return self.L.orc_search_direction(self.h)

    def cone_violation(self, xhat, x, tau):
        xhat, x = _f64(xhat), _f64(x)
        return bool(self.L.orc_cone_violation(self.h, _dp(xhat), _dp(x), float(tau)))

    def cone_search(self):
        return self.L.orc_cone_search(self.h)

    def jacobian_times(self, v):
        v = _f64(v)
        out = np.zeros(self.total)
        self.L.orc_jacobian_times(self.h, _dp(v), _dp(out))
        return out

    def dense_jacobian(self):
        J = np.zeros((self.total, self.total), order="F")
        self.L.orc_dense_jacobian(self.h, _dp(J))
        return J

    def dense_symmetric(self):
        K = np.zeros((self.N, self.N), order="F")
        self.L.orc_dense_symmetric(self.h, _dp(K))
        return K

    def jacobian_coo(self):
        cnt = self.L.orc_jacobian_coo(self.h, None, None, None)
        r, cc, v = np.zeros(cnt, np.int32), np.zeros(cnt, np.int32), np.zeros(cnt)
        self.L.orc_jacobian_coo(self.h, _ip(r), _ip(cc), _dp(v))
        return r, cc, v

    def use_superlu_fallback(self):
        """Stand-in for the reference's UMFPACK `J \\ R` fallback (search_direction.jl:22,113): SciPy SuperLU."""
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla

        def lu(user, total, res, step):
            r, cc, v = self.jacobian_coo()
            J = sp.csc_matrix((v, (r, cc)), shape=(total, total))
            try:
                x = spla.splu(J).solve(_view(res, total).copy())
            except RuntimeError:
                return 1
            _view(step, total)[:] = x
            return 0

        self._keep_lu = LU_FN(lu)
        self.L.orc_solver_set_lu_fallback(self.h, self._keep_lu, None)

    def K_csc(self):
        nnz = self.L.orc_K_nnz(self.h)
        return (_view(self.L.orc_K_colptr(self.h), self.N + 1).copy(), _view(self.L.orc_K_rowval(self.h), nnz).copy(),
                _view(self.L.orc_K_nzval(self.h), nnz).copy())

    def merit(self, at_candidate=False):
        return self.L.orc_merit(self.h, int(at_candidate))

    def merit_gradient_eval(self):
        self.L.orc_merit_gradient_eval(self.h)

    def constraint_violation(self, at_candidate=False):
        return self.L.orc_constraint_violation(self.h, int(at_candidate))

    def optimality_error(self):
        return self.L.orc_optimality_error(self.h)

    def initialize(self, guess):
        g = _f64(guess)
        self.L.orc_initialize(self.h, _dp(g))

    def solve_begin(self):
        self.L.orc_solve_begin(self.h)

    def newton_iteration(self):
        return self.L.orc_newton_iteration(self.h)

    def outer_update(self):
        self.L.orc_outer_update(self.h)

    def solve(self):
        return self.L.orc_solve(self.h)

    def ldl(self):
        """Non-owning view of the solver's LDLSolver (linear_solver.jl:3-8)."""
        q = QDLDL.__new__(QDLDL)
        q.L = self.L
        q.h = self.L.orc_linear_solver(self.h)
        q.n = self.N
        q.nnzA = self.L.orc_qdldl_nnzA(q.h)
        q.nnzL = self.L.orc_qdldl_nnzL(q.h)
        q.__class__ = _BorrowedQDLDL
        return q


class _BorrowedQDLDL(QDLDL):
    def __del__(self):
        self.h = None


def from_problem(P, perm=None, options=None) -> Oracle:
    """Build an Oracle from any object with the ConicProblem attribute names (duck-typed; no product import)."""
    o = Oracle(P.n, P.m, P.p, P.num_nonnegative, P.soc_dims, P.W_colptr, P.W_rowval, P.G_colptr, P.G_rowval,
               P.C_colptr, P.C_rowval, perm=perm, options=options)
    if getattr(P, "W_val", None) is not None:
        o.set_lq(P.W_val, P.G_val, P.C_val, P.q, P.g0, P.h0)
    return o


# ---------------------------------------------------------------------------------------------------- evaluate!'s scatter
def scatter_dense(shape, sparsity, cache):
    """src/solver/evaluate.jl:39-41 (and :57-59, 75-77, 97-99, 111-113), restated literally: a dense matrix that starts at
    zero (problem_data.jl) and receives `M[idx...] = cache[i]` in cache order -- a repeated key keeps the LAST value.
    sparsity: 0-based (row, col) pairs."""
    M = np.zeros(shape)
    for i, (r, c) in enumerate(sparsity):
        M[r, c] = cache[i]
    return M


def stage_scatter(n_out, indices, caches, accumulate):
    """The stage loops of the trajectory-optimisation front end, restated literally: `fill!(out, 0.0)` (evaluate.jl:16,78,112,
    207,298), then stage by stage in program order either `out[idx...] += cache[i]` (constraint_dual_jacobian_variables!,
    dynamics.jl:172-179, constraints.jl:203-212; gradient_variables!, costs.jl:115-120) or `out[indices[t]] .= cache`
    (constraints!, dynamics.jl:143-148, constraints.jl:169-176).  indices / caches: per-stage lists (0-based indices)."""
    out = np.zeros(n_out)
    for idx, cache in zip(indices, caches):
        for i, j in enumerate(idx):
            if accumulate:
                out[j] += cache[i]
            else:
                out[j] = cache[i]
    return out


def lagrangian_hessian_dense(n, sparsities, caches):
    """residual_jacobian_variables.jl:9-15: H[i, j] = objective_xx[i, j]; H[i, j] += equality_dual_xx[i, j];
    H[i, j] += cone_dual_xx[i, j] on the dense matrices scatter_dense produced (missing caches count as zero matrices)."""
    H = scatter_dense((n, n), sparsities[0], caches[0])
    for sp, ca in zip(sparsities[1:], caches[1:]):
        H = H + scatter_dense((n, n), sp, ca)
    return H


def values_at_pattern(M, colptr, rowval):
    """The entries of dense M at a CSC pattern (what the hot path reads)."""
    out = np.zeros(len(rowval))
    for j in range(len(colptr) - 1):
        for k in range(colptr[j], colptr[j + 1]):
            out[k] = M[rowval[k], j]
    return out
