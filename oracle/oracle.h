/*
 * oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the CALIPSO.jl Newton/KKT hot path.
 *
 * This is a plain-C restatement of the reference's algorithm, function by function, used ONLY as the checker in
 * tests/, in __graft_entry__.smoke() and as bench.py's cpu_baseline / --impl reference arm.  Nothing under
 * calipso_b200/ (the product) may import, link or call it.
 *
 * Parity pinning: the Julia reference cannot run in this environment (no Julia toolchain) and its tests hold no
 * golden vectors (all inputs are unseeded randn); the oracle is therefore pinned against the *properties* the
 * reference's own tests assert (test/solver/problem.jl:100-211 block identities, LDL-vs-LU direction to 1e-6,
 * refinement to 1e-10; test/solver/{wachter,friction_cone,portfolio,maratos,knitro}.jl known answers and
 * stopping criteria) -- see tests/test_oracle_*.py.  The fill-reducing ordering (AMD.jl -> SuiteSparse libamd,
 * Project.toml:19, call site src/solver/qdldl.jl:135) lives outside the reference tree and is absent here:
 * PARITY IS UNPINNED AT THE AMD BOUNDARY.  qdldl(A; perm=p) (qdldl.jl:134-136) accepts a caller permutation, which
 * is how the tests feed the product's permutation through this oracle for the bit-exact integer comparisons.
 *
 * All indices are 0-based int32 here (the reference is 1-based Int64).  Arithmetic is double, no FMA contraction
 * (compile with -ffp-contract=off) to match Julia's separate multiply/subtract.
 */
#ifndef CALIPSO_ORACLE_H
#define CALIPSO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- QDLDL (src/solver/qdldl.jl) */
typedef struct orc_qdldl orc_qdldl;

/* qdldl(A; perm) ctor, qdldl.jl:134-188.  A = upper-triangular CSC (sorted rows); perm==NULL -> oracle's own
 * minimum-degree ordering (stand-in for AMD.amd, see header).  Performs the first numeric factorisation. */
orc_qdldl *orc_qdldl_new(int n, const int *Ap, const int *Ai, const double *Ax, const int *perm);
void orc_qdldl_free(orc_qdldl *F);
/* update_values!(F, 1:nnz, values) qdldl.jl:199-213 followed by refactor! :269-278.  Returns positive-inertia
 * count (or -1 on a zero pivot, qdldl.jl:456,579). */
int orc_qdldl_refactor(orc_qdldl *F, const double *Ax);
/* solve!(F, b) in place, qdldl.jl:330-351 */
void orc_qdldl_solve(orc_qdldl *F, double *b);
/* accessors for the Appendix-B data contract (all 0-based) */
int orc_qdldl_n(const orc_qdldl *F);
int orc_qdldl_nnzL(const orc_qdldl *F);
int orc_qdldl_nnzA(const orc_qdldl *F);
const int *orc_qdldl_perm(const orc_qdldl *F);
const int *orc_qdldl_iperm(const orc_qdldl *F);
const int *orc_qdldl_etree(const orc_qdldl *F);
const int *orc_qdldl_Lnz(const orc_qdldl *F);
const int *orc_qdldl_Lp(const orc_qdldl *F);
const int *orc_qdldl_Li(const orc_qdldl *F);
const double *orc_qdldl_Lx(const orc_qdldl *F);
const double *orc_qdldl_D(const orc_qdldl *F);
const double *orc_qdldl_Dinv(const orc_qdldl *F);
const int *orc_qdldl_triuA_colptr(const orc_qdldl *F);
const int *orc_qdldl_triuA_rowval(const orc_qdldl *F);
const double *orc_qdldl_triuA_nzval(const orc_qdldl *F);
const int *orc_qdldl_AtoPAPt(const orc_qdldl *F);
int orc_qdldl_positive_inertia(const orc_qdldl *F);
long long orc_qdldl_factor_count(const orc_qdldl *F); /* numeric factorisations performed so far */

/* stand-in ordering: exact minimum degree on the pattern of A+A' (ties -> lowest index) */
void orc_min_degree(int n, const int *Ap, const int *Ai, int *perm);
/* amd(A) of src/solver/qdldl.jl:135 (AMD.jl -> SuiteSparse amd_l_order, defaults dense = 10, aggressive = 1): restated in
 * amd.c from the published algorithm.  Any triangle content; P[k] = index eliminated k-th.  Returns 0 on success. */
int orc_amd_order(int n, const int *Ap, const int *Ai, int *P, double dense, int aggressive);

/* ---------------------------------------------------------------- solver (src/solver/ *.jl) */

/* evaluate! flags, src/solver/evaluate.jl:1-124 */
enum {
    ORC_EV_OBJECTIVE = 1,        /* f                               -> objective            */
    ORC_EV_GRADIENT = 2,         /* grad f                          -> objective_gradient   */
    ORC_EV_EQUALITY = 4,         /* g(x)                            -> equality_constraint  */
    ORC_EV_CONE = 8,             /* h(x)                            -> cone_constraint      */
    ORC_EV_EQUALITY_DUAL_GRAD = 16, /* (g'y)_x = G'y                -> equality_dual_jacobian_variables */
    ORC_EV_CONE_DUAL_GRAD = 32,  /* (h'z)_x = C'z                   -> cone_dual_jacobian_variables     */
    ORC_EV_HESSIAN = 64,         /* upper triangle of f_xx + (g'y)_xx + (h'z)_xx at the pattern -> W_val */
    ORC_EV_EQUALITY_JAC = 128,   /* G values at the pattern         -> G_val                */
    ORC_EV_CONE_JAC = 256        /* C values at the pattern         -> C_val                */
};

typedef struct {
    double *objective;   /* [1] */
    double *gradient;    /* [n] */
    double *equality;    /* [m] */
    double *cone;        /* [p] */
    double *eq_dual_grad;   /* [n] */
    double *cone_dual_grad; /* [n] */
    double *W_val;       /* [nnz(W upper)] */
    double *G_val;       /* [nnz(G)] */
    double *C_val;       /* [nnz(C)] */
} orc_eval_out;

typedef void (*orc_eval_fn)(void *user, int flags, const double *x, const double *y, const double *z,
                            orc_eval_out *out);

/* Options, src/solver/options.jl:6-59 (hot-path relevant subset; same defaults) */
typedef struct {
    int max_outer_iterations, max_residual_iterations, max_residual_line_search, max_cone_line_search;
    int iterative_refinement, max_iterative_refinement, min_iterative_refinement;
    double scaling_line_search, iterative_refinement_tolerance;
    double central_path_initial, central_path_update_tolerance, central_path_scaling, central_path_exponent;
    double penalty_initial, penalty_scaling, dual_initial;
    double residual_tolerance, optimality_tolerance, slack_tolerance, equality_tolerance, complementarity_tolerance;
    double min_regularization, primal_regularization_initial, dual_regularization_initial, max_regularization;
    double dual_regularization, dual_regularization_exponent;
    double scaling_regularization_initial, scaling_regularization, scaling_regularization_last;
    double max_penalty;
    double violation_tolerance, violation_exponent, merit_tolerance, merit_exponent, armijo_tolerance,
        machine_tolerance;
    int max_filter;
    int warmstart;
    int reference_schedule; /* 1: re-factor where the reference does (linear_solver.jl:56,
                               iterative_refinement.jl:21-25); 0: one factorisation per regularisation trial */
} orc_options;

void orc_options_default(orc_options *o);

typedef struct orc_solver orc_solver;

/* Solver(...) ctor, src/solver/solver.jl:46-150 (pattern given structurally instead of discovered at a random
 * point).  W = upper triangle incl. diagonal (CSC, sorted), G (m x n), C (p x n).
 * Cone rows [0,num_nonnegative) nonnegative, then SOC blocks of soc_dims[k] contiguous rows.  perm may be NULL. */
orc_solver *orc_solver_new(int n, int m, int p, int num_nonnegative, int num_soc, const int *soc_dims,
                           const int *Wp, const int *Wi, const int *Gp, const int *Gi, const int *Cp, const int *Ci,
                           const int *perm, const orc_options *opts);
void orc_solver_free(orc_solver *s);
void orc_solver_set_callback(orc_solver *s, orc_eval_fn fn, void *user);
/* built-in LQ evaluator (calipso_b200/lqc.py family): f = 1/2 x'Qx + q'x, g = Gx + g0, h = Cx + h0 */
void orc_solver_set_lq(orc_solver *s, const double *W_val, const double *G_val, const double *C_val,
                       const double *q, const double *g0, const double *h0);

/* raw views into the solver state (lengths in comments); valid until orc_solver_free */
double *orc_solution(orc_solver *s);        /* [total]  w = (x,r,s,y,z,t), indices.jl:25-35 */
double *orc_candidate(orc_solver *s);       /* [total] */
double *orc_step(orc_solver *s);            /* [total] */
double *orc_residual(orc_solver *s);        /* [total] */
double *orc_residual_symmetric_vec(orc_solver *s); /* [n+m+p] */
double *orc_step_symmetric(orc_solver *s);  /* [n+m+p] */
double *orc_dual(orc_solver *s);            /* [m] lambda */
double *orc_scalars(orc_solver *s);         /* [8]: kappa, tau, rho, eps_p, eps_d, eps_p_last, objective, barrier */
orc_eval_out *orc_problem(orc_solver *s);   /* problem data vectors */
double *orc_cone_product(orc_solver *s);    /* [p] */
double *orc_cone_target(orc_solver *s);     /* [p] */
double *orc_barrier_gradient(orc_solver *s);/* [p] */
double *orc_merit_gradient(orc_solver *s);  /* [n+m+p] */
const int *orc_inertia(orc_solver *s);      /* [3] positive, negative, zero */
const int *orc_stats(orc_solver *s);        /* [12]: last n_trials, last n_refine, refine_ok, k_s, k_t, total_iterations, outer,
                                               status, lu_fallbacks (cumulative), used_lu (last), reserved x2 */
orc_qdldl *orc_linear_solver(orc_solver *s);
int orc_K_nnz(orc_solver *s);
const int *orc_K_colptr(orc_solver *s);
const int *orc_K_rowval(orc_solver *s);
const double *orc_K_nzval(orc_solver *s);

/* hot path, one function per reference function */
void orc_evaluate(orc_solver *s, int flags, int at_candidate);                 /* evaluate!  evaluate.jl:1 */
void orc_cone(orc_solver *s, int at_candidate, int barrier, int barrier_gradient, int product, int jacobian,
              int target);                                                      /* cone!  cones/cone.jl:71 */
void orc_residual_eval(orc_solver *s);                                          /* residual!  residual.jl:1 */
void orc_residual_jacobian_variables(orc_solver *s);                            /* residual_jacobian_variables.jl:1 */
void orc_residual_jacobian_variables_symmetric(orc_solver *s);                  /* :110 */
void orc_residual_symmetric(orc_solver *s, const double *residual);             /* residual.jl:53 */
int orc_factorize(orc_solver *s);                                               /* factorize!+compute_inertia! */
int orc_inertia_correction(orc_solver *s);      /* inertia.jl:30; 0 ok, 1 = "inertia correction failure" */
void orc_search_direction_symmetric(orc_solver *s, double *step, const double *residual, int factorize);
int orc_iterative_refinement(orc_solver *s, double *step);                      /* iterative_refinement.jl:1; 1=true */
int orc_search_direction(orc_solver *s);        /* search_direction.jl:1; 0 ok, 1 inertia failure, 2 refinement failure */
int orc_cone_violation(orc_solver *s, const double *xhat, const double *x, double tau); /* cone.jl:62 */
int orc_cone_search(orc_solver *s);             /* solve.jl:190-221; 0 ok, 3 = "cone search failure" */
void orc_jacobian_times(orc_solver *s, const double *v, double *out);           /* out = J v (matrix-free mul!) */
void orc_dense_jacobian(orc_solver *s, double *J /* total x total, column-major */);
/* J as COO triplets (0-based); returns the count; rows/cols/vals may be NULL to query the count */
int orc_jacobian_coo(orc_solver *s, int *rows, int *cols, double *vals);
/* search_direction_nonsymmetric!(step, J, R, lu) = `J \ R` with UMFPACK (search_direction.jl:106-119), reached when
 * iterative refinement fails (:22).  UMFPACK is a Julia-stdlib binary dependency that is absent here; the stand-in is
 * supplied by the caller (tests use SciPy's SuperLU on orc_jacobian_coo()).  Without a callback a dense
 * partial-pivoting LU is used when total <= 2500, otherwise the failure is only recorded.
 * fn(user, total, residual, step) must write step = J^-1 residual and return 0. */
typedef int (*orc_lu_fn)(void *user, int total, const double *residual, double *step);
void orc_solver_set_lu_fallback(orc_solver *s, orc_lu_fn fn, void *user);
void orc_dense_symmetric(orc_solver *s, double *K /* N x N column-major, full symmetric from upper triangle */);

/* reductions */
double orc_merit(orc_solver *s, int at_candidate);                              /* merit.jl:2 */
void orc_merit_gradient_eval(orc_solver *s);                                    /* merit.jl:17 */
double orc_constraint_violation(orc_solver *s, int at_candidate);               /* constraint_violation.jl:1 */
double orc_optimality_error(orc_solver *s);                                     /* optimality_error.jl:1 */

/* solve!  solve.jl:8-377.  Returns 1 (true) on convergence, 0 otherwise; <0 on error (-1 inertia, -3 cone search) */
int orc_initialize(orc_solver *s, const double *guess);                         /* initialize!  initialize.jl:9 */
int orc_solve(orc_solver *s);
/* one inner Newton iteration of solve! (solve.jl:98-350) assuming the state a preceding iteration left behind;
 * returns 0 continue, 1 outer-converged, 2 inner-converged(break), <0 error */
int orc_newton_iteration(orc_solver *s);
void orc_solve_begin(orc_solver *s);     /* solve.jl:8-95 */
void orc_outer_update(orc_solver *s);    /* solve.jl:356-368 */

#ifdef __cplusplus
}
#endif
#endif
