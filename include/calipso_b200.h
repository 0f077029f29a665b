/*
 * calipso_b200.h -- C ABI of libcalipso_b200.so: CALIPSO.jl's per-iteration Newton/KKT hot path on B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's `src/solver` seam.  Every entry point names the reference
 * interface it replaces (file:line relative to the CALIPSO.jl tree); INTEGRATION.md shows the Julia `ccall` glue.
 * Conventions:
 *   - plain C: opaque handle, ints, doubles, caller-owned HOST pointers unless a name ends in `_dev`;
 *   - a handle owns `batch` independent problem instances with ONE sparsity pattern;
 *     per-instance arrays are instance-major: element i of instance b lives at [b * len + i];
 *   - indices are 0-based int32 (the reference: 1-based Int64); values are Float64, column-major;
 *   - the point is w = (x, r, s, y, z, t) with offsets (0, n, n+m, n+m+p, n+2m+p, n+2m+2p), total = n+2m+3p
 *     (src/solver/indices.jl:25-35); the reduced system is ordered (x, y, z), N = n+m+p (:32-35);
 *   - cone rows [0, num_nonnegative) are the nonnegative orthant, then num_soc second-order blocks of soc_dims[k]
 *     contiguous rows (cones/cone.jl:27-59 concatenation order);
 *   - every call is asynchronous on the handle's CUDA stream unless it returns data to the host;
 *   - return value 0 = ok, negative = CUDA/NCCL/argument error (text via cb200_last_error()); solver outcomes
 *     (inertia failure, cone-search failure, ...) are per-instance status codes, never process errors;
 *   - there is no CPU fallback: without a CUDA device cb200_create fails.
 * A handle is not thread-safe (one host thread per handle, like one workspace per Solver in the reference).
 */
#ifndef CALIPSO_B200_H
#define CALIPSO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb200_handle cb200_handle;

/* Options: src/solver/options.jl:6-59 (hot-path subset, same names and defaults) + the GMRES fallback knobs */
typedef struct cb200_options {
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
    int gmres_restart, gmres_max_cycles; /* fallback replacing `J \ R` (UMFPACK), search_direction.jl:22,113 */
} cb200_options;

void cb200_options_default(cb200_options *o);
const char *cb200_last_error(void);
int cb200_device_count(void);

/* per-instance status codes (I_STATUS) -- the reference's error()/@warn sites */
enum {
    CB200_OK = 0,
    CB200_INERTIA_FAILURE = 1,     /* error("inertia correction failure"), inertia.jl:72 */
    CB200_REFINEMENT_FAILURE = 2,  /* @warn "iterative refinement failure" and the fallback failed too, :50 */
    CB200_CONE_SEARCH_FAILURE = 3, /* error("cone search failure"), solve.jl:210,220 */
    CB200_ZERO_PIVOT = 4           /* @warn "Zero entry in D", qdldl.jl:309-311 */
};

/* arrays addressable through cb200_set_array / cb200_get_array / cb200_device_ptr (per-instance length) */
enum {
    CB200_POINT = 0,        /* w                      [total]  solver.solution.all */
    CB200_CANDIDATE,        /* candidate              [total]  solver.candidate.all */
    CB200_STEP,             /* search direction       [total]  solver.data.step.all */
    CB200_RESIDUAL,         /* R(w)                   [total]  solver.data.residual.all */
    CB200_GRADIENT,         /* grad f                 [n]      problem.objective_gradient_variables */
    CB200_EQ_DUAL_GRAD,     /* (g'y)_x                [n]      problem.equality_dual_jacobian_variables */
    CB200_CONE_DUAL_GRAD,   /* (h'z)_x                [n]      problem.cone_dual_jacobian_variables */
    CB200_EQUALITY,         /* g(x)                   [m]      problem.equality_constraint */
    CB200_CONE,             /* h(x)                   [p]      problem.cone_constraint */
    CB200_W_VALUES,         /* upper(f_xx+(g'y)_xx+(h'z)_xx) at the pattern [nnz(W)] */
    CB200_G_VALUES,         /* g_x at the pattern     [nnz(G)] problem.equality_jacobian_variables */
    CB200_C_VALUES,         /* h_x at the pattern     [nnz(C)] problem.cone_jacobian_variables */
    CB200_CONE_PRODUCT,     /* s o t                  [p]      problem.cone_product */
    CB200_BARRIER_GRADIENT, /* grad Phi(s)            [p]      problem.barrier_gradient */
    CB200_DUAL,             /* lambda                 [m]      solver.dual */
    CB200_LQ_Q,             /* q of the LQ family     [n] */
    CB200_LQ_G0,            /* g0                     [m] */
    CB200_LQ_H0,            /* h0                     [p] */
    CB200_SCALARS,          /* CB200_S_* slots        [24] */
    CB200_MERIT_GRADIENT,   /* merit gradient         [N]      solver.data.merit_gradient */
    CB200_RESIDUAL_SYMMETRIC, /* reduced rhs          [N]      solver.data.residual_symmetric.all */
    CB200_STEP_SYMMETRIC,   /* reduced solution       [N]      solver.data.step_symmetric.all */
    CB200_PIVOTS,           /* D in elimination order [N]      F.workspace.D */
    CB200_MATRIX_VALUES,    /* LinearSolver seam: upper-triangle values of A [nnz(A)] */
    CB200_RHS,              /* LinearSolver seam: b -> x in place            [N] */
    CB200_PANELS,           /* supernodal factor panels [panel_total] (debug / tests) */
    CB200_NUM_ARRAYS
};

/* scalar slots of CB200_SCALARS */
enum {
    CB200_S_KAPPA = 0, CB200_S_TAU, CB200_S_RHO, CB200_S_EPSP, CB200_S_EPSD, CB200_S_EPSP_LAST, CB200_S_OBJECTIVE,
    CB200_S_BARRIER, CB200_S_RESIDUAL_VIOLATION, CB200_S_OPTIMALITY_VIOLATION, CB200_S_SLACK_VIOLATION,
    CB200_S_THETA, CB200_S_MERIT, CB200_S_STEP_SIZE, CB200_S_STEP_SIZE_T, CB200_S_EQUALITY_VIOLATION,
    CB200_S_CONE_PRODUCT_VIOLATION, CB200_S_REFINE_NORM, CB200_S_REFINE_NORM_INITIAL, CB200_S_MERIT_CANDIDATE,
    CB200_S_THETA_CANDIDATE, CB200_S_MERIT_SLOPE, CB200_S_STEP_SIZE_CONE, CB200_S_COUNT = 24
};
/* integer slots returned by cb200_get_stats */
enum {
    CB200_I_INERTIA_POS = 0, CB200_I_INERTIA_NEG, CB200_I_INERTIA_ZERO, CB200_I_TRIALS, CB200_I_REFINE,
    CB200_I_REFINE_OK, CB200_I_KS, CB200_I_KT, CB200_I_STATUS, CB200_I_USED_FALLBACK, CB200_I_FALLBACKS,
    CB200_I_TOTAL_ITERATIONS, CB200_I_OUTER, CB200_I_LINE_SEARCH, CB200_I_CONVERGED, CB200_I_GMRES_ITERS,
    CB200_I_FILTER_INDEX, CB200_I_INNER, CB200_I_FACTORIZATIONS, CB200_I_SOLVES, CB200_I_UNREFINED_STEPS, CB200_I_COUNT = 24
};

/* evaluate! flags (src/solver/evaluate.jl keyword arguments) for cb200_lq_evaluate */
enum {
    CB200_EV_OBJECTIVE = 1, CB200_EV_GRADIENT = 2, CB200_EV_EQUALITY = 4, CB200_EV_CONE = 8,
    CB200_EV_EQUALITY_DUAL_GRAD = 16, CB200_EV_CONE_DUAL_GRAD = 32
};
/* cone! flags (cones/cone.jl:71-77 keyword arguments) */
enum { CB200_CONE_BARRIER = 1, CB200_CONE_BARRIER_GRADIENT = 2, CB200_CONE_PRODUCT_FLAG = 4 };

/* ------------------------------------------------------------------------------------------------ construction
 * Solver(methods, num_variables, num_parameters, num_equality, num_cone; nonnegative_indices, second_order_indices,
 * options), src/solver/solver.jl:46-150: builds Indices/Dimensions (:54-61), the reduced-matrix pattern and the
 * symbolic factorisation (ldl_solver -> qdldl(A): ordering + permute_symmetric + QDLDL_etree!, :122, qdldl.jl:134-188).
 * W = UPPER triangle (with every diagonal entry) of the Lagrangian Hessian pattern, G = g_x (m x n), C = h_x (p x n), CSC.
 * perm (nullable) = caller ordering of the N reduced unknowns, as qdldl(A; perm=p), qdldl.jl:134-136.
 * Returns NULL on failure (see cb200_last_error). */
cb200_handle *cb200_create(int batch, int n, int m, int p, int num_nonnegative, int num_soc, const int *soc_dims,
                           const int *Wp, const int *Wi, const int *Gp, const int *Gi, const int *Cp, const int *Ci,
                           const int *perm, const cb200_options *options, int device);
/* ldl_solver(A::SparseMatrixCSC), src/solver/linear_solver.jl:46-48: LinearSolver seam on a generic symmetric
 * matrix given by its upper triangle (CSC, sorted, full diagonal). */
cb200_handle *cb200_ldl_create(int batch, int N, const int *Ap, const int *Ai, const int *perm, int device);
void cb200_destroy(cb200_handle *h); /* Julia finalizer */

/* out[16]: N, total, nnz(K upper), nnz(L) (true fill), supernodes, levels, phases, max supernode width,
 * max panel rows, panel_total (doubles), sum Lnz^2, batch, n, m, p, nnz(W)+nnz(G)+nnz(C) */
int cb200_info(const cb200_handle *h, long long *out);
/* Which code paths the handle's pattern selected (they differ in speed, not in results).  out[8]: 1 if the solves keep
 * x[N] in shared memory and stream the chain panels by TMA (else the global-memory solve); resident CTAs per SM the plan was
 * sized for (3, 2 or 1: chosen from N and the widest supernode against the 228 KB of shared memory of an SM; if the
 * occupancy query at creation grants fewer than that, the granted number is reported instead); dynamic
 * shared memory per CTA in bytes; CTA-scope supernodes; those of them that do not fit the shared-memory staging and run
 * the global-memory code; 1 if the chain descriptors / the phase schedule are cached in shared memory (<= 64 chain supernodes,
 * <= 48 phases; otherwise they are read from global memory); threads per CTA of the heavy kernels (256 or 512). */
int cb200_path_info(const cb200_handle *h, long long *out);
/* amd(A) of src/solver/qdldl.jl:135 (AMD.jl / SuiteSparse AMD, default controls) for an N x N CSC pattern with sorted rows
 * (any triangle content; the ordering is that of A + A'): perm[k] = 0-based index eliminated k-th.  Host-only, no handle,
 * no GPU.  Pass the result as `perm` to cb200_create / cb200_ldl_create to factor in the reference's own elimination
 * order (the library's default ordering is a minimum-degree variant better suited to the device). */
int cb200_amd_order(int N, const int *Ap, const int *Ai, int *perm);
/* Appendix-B integer contract for parity tests: perm, etree, Lnz (each [N], 0-based, elimination order) */
int cb200_get_symbolic(const cb200_handle *h, int *perm, int *etree, int *Lnz);
/* L in QDLDL's CSC form (strictly lower, unit diagonal implied) and D of instance b, extracted from the supernodal
 * panels on the host: Lp[N+1], Li[nnzL], Lx[nnzL], D[N] (F.workspace.{Lp,Li,Lx,D}, qdldl.jl:21-29) */
int cb200_get_factor(cb200_handle *h, int instance, int *Lp, int *Li, double *Lx, double *D);

/* ------------------------------------------------------------------------------------------------ data movement */
int cb200_set_array(cb200_handle *h, int which, const double *host, int first_instance, int count); /* H2D, async */
int cb200_get_array(cb200_handle *h, int which, double *host, int first_instance, int count);       /* D2H + sync */
/* initialize!(solver, guess), src/solver/initialize.jl:9-13: solution.variables .= guess for instances first .. first+count-1
 * (guess_host = [count][n]; H2D of the primal block only, asynchronous; slacks and duals are set by the caller or by
 * cb200_lq_begin as initialize_slacks! / initialize_duals! do, initialize.jl:15-36) */
int cb200_initialize(cb200_handle *h, const double *guess_host, int first_instance, int count);
int cb200_get_stats(cb200_handle *h, int *host /* [count][CB200_I_COUNT] */, int first_instance, int count);
/* the same without the final stream synchronisation (host buffers should be pinned; call cb200_synchronize before reading
 * them): lets several handles, each on its own stream, overlap their copies with each other's kernels */
int cb200_get_array_async(cb200_handle *h, int which, double *host, int first_instance, int count);
int cb200_get_stats_async(cb200_handle *h, int *host, int first_instance, int count);
int cb200_array_length(const cb200_handle *h, int which);
/* per-instance cycle counters of the factor/solve phases, [batch][24] (thread 0 of each CTA, clock64); reset != 0
 * zeroes them afterwards.  Diagnostic only: the counters are off until this function is called for the first time
 * (that call returns zeros) because they cost a global read-modify-write per phase. */
int cb200_get_profile(cb200_handle *h, long long *host, int reset);
void *cb200_device_ptr(cb200_handle *h, int which);   /* base of the instance-major device array */
/* The library keeps row-ordered copies of the W and G values (refreshed lazily after cb200_set_array).  A caller that
 * writes CB200_W_VALUES / CB200_G_VALUES through cb200_device_ptr must announce it with this call. */
int cb200_values_changed(cb200_handle *h);
void *cb200_stream(cb200_handle *h);                  /* cudaStream_t of the handle */
int cb200_synchronize(cb200_handle *h);
int cb200_set_options(cb200_handle *h, const cb200_options *options);

/* ------------------------------------------------------------------------------------------------ hot path (batched) */
/* cone!(problem, methods, idx, solution; barrier, barrier_gradient, product), cones/cone.jl:71-106, at the point
 * (at_candidate = 0) or the candidate (1).  The arrow-matrix Jacobians (jacobian=true) are never materialised: every
 * consumer rebuilds them from (s, t). */
int cb200_cone(cb200_handle *h, int flags, int at_candidate);
/* residual!(data, problem, idx, solution, kappa, rho, lambda), residual.jl:1-51, plus the scalar reductions solve!
 * takes right after it (solve.jl:130-135, optimality_error.jl:1-27) into the CB200_S_* slots */
int cb200_residual(cb200_handle *h);
/* search_direction!(solver), search_direction.jl:1-23: inertia_correction! (inertia.jl:30-79; each trial =
 * residual_jacobian_variables! + _symmetric! + factorize! + compute_inertia!), search_direction_symmetric!
 * (:25-104), iterative_refinement! (iterative_refinement.jl:1-53) and, if that fails, a GMRES solve of J step = R
 * standing in for the reference's UMFPACK fallback (:22).  Outcome per instance in the stats. */
int cb200_search_direction(cb200_handle *h);
/* cone line search, solve.jl:190-221 (cone_violation, cones/cone.jl:62-68): fills candidate s and t, step sizes in
 * CB200_S_STEP_SIZE / _T and halving counts in CB200_I_KS / _KT */
int cb200_cone_search(cb200_handle *h);
/* step update, solve.jl:309-333: w <- candidate (x, r, s), y,z -= alpha * step, t <- candidate t, cone!(product),
 * equality / complementarity violations.  alpha = CB200_S_STEP_SIZE of each instance. */
int cb200_apply_step(cb200_handle *h);
/* One regularisation trial without the inertia loop, for measurement: assemble K at the current (eps_p, eps_d), factor,
 * then `nsolves` x (residual_symmetric! + solve! + recovery) of the current residual -- the "KKT solve" unit of
 * SURVEY.md section 8(d) */
int cb200_kkt_factor_solve(cb200_handle *h, int nsolves);
/* differentiate!(solver), src/solver/differentiate.jl:1-61: ONE factorisation of the
 * reduced matrix at the current point and the current (eps_p, eps_d) (:13-20), then per parameter one
 * search_direction_symmetric! (reduced rhs + LDL' solve + recovery, no refinement, :35-46) on column i of
 * residual_jacobian_parameters! (residual_jacobian_parameters.jl:1-40), and solution_sensitivity[:, i] = -result (:55-57).
 * (The reference re-factors for every parameter; the factor is reused here.)  Host buffers, instance-major:
 * jacobian_parameters[b][i][total] (column i contiguous), sensitivity[b][i][total]. */
int cb200_differentiate(cb200_handle *h, int num_parameters, const double *jacobian_parameters_host,
                        double *sensitivity_host);
/* evaluate!'s scatter of the flat derivative caches, src/solver/evaluate.jl:37-42,
 * 73-78, 109-114 (Hessian caches) and :55-60, 95-100 (Jacobian caches): the reference writes cache entry i into a
 * dense matrix at key sparsity[i] with `=` -- where keys repeat (overlapping stage sparsities of the trajectory
 * optimisation front end, trajectory_optimization/methods.jl:26,41) the LAST entry wins, absent
 * keys stay 0.0 -- and, for the Hessian, adds the objective, equality-dual and cone-dual matrices in that order
 * (residual_jacobian_variables.jl:11-13).  cb200_scatter_plan turns that into a gather list, once per sparsity:
 *   which = CB200_W_VALUES: ncaches = 1..3 caches (objective, equality-dual, cone-dual); keys below the diagonal are
 *           dropped (the path reads the upper triangle, linear_solver.jl:23);
 *   which = CB200_G_VALUES / CB200_C_VALUES: one cache (equality / cone Jacobian).
 * cache_len[ncaches]; rows/cols = the keys of all caches, concatenated, 0-based.  A key outside the pattern given to
 * cb200_create is an error.  cb200_scatter copies caches_host ([count][sum cache_len], instance-major; NULL = the
 * caller already wrote the device buffer cb200_scatter_buffer(h, which), laid out [batch][sum cache_len]) and fills
 * the value array `which` of instances first .. first+count-1 on the handle's stream (asynchronous). */
int cb200_scatter_plan(cb200_handle *h, int which, int ncaches, const int *cache_len, const int *rows, const int *cols);
int cb200_scatter(cb200_handle *h, int which, const double *caches_host, int first_instance, int count);
void *cb200_scatter_buffer(cb200_handle *h, int which);
/* Stage-level scatter of the trajectory-optimisation front end (SURVEY.md section 8(f) row N2 proper): the callbacks the
 * front end hands to the solver (trajectory_optimization/methods.jl:2-44) clear a vector of the problem data and then loop
 * over the stages, each stage evaluating its generated function into a small cache and writing the cache through an index
 * list:
 *   accumulate = 1   `gradient[indices[t]] .+= cache` / `gradient[idx...] += cache[i]` -- objective gradient
 *                    (evaluate.jl:15-28, costs.jl:115-120), (g'y)_x (evaluate.jl:206-241, dynamics.jl:172-179,
 *                    constraints.jl:203-212) and (h'z)_x (evaluate.jl:297-327): neighbouring stages overlap (dynamics t
 *                    and t+1 both touch x_{t+1}), the sums are formed in program order: dynamics, stage constraints,
 *                    general constraints, each stage by stage;
 *   accumulate = 0   `violations[indices[t]] .= cache` -- g(x) (evaluate.jl:77-109, dynamics.jl:143-148,
 *                    constraints.jl:169-176) and h(x) (evaluate.jl:111-136): the last write to an entry wins.
 * Entries no stage writes are 0.0 (the fill!).  cb200_stage_plan takes the concatenation of the stages' index lists in that
 * program order (dst[k], 0-based, = the entry of `which` that entry k of the concatenated caches goes to) and turns it into
 * per-entry gather lists that keep the order, so the result is bit-identical to the reference's loops.  which =
 * CB200_GRADIENT, CB200_EQ_DUAL_GRAD, CB200_CONE_DUAL_GRAD (accumulate = 1) or CB200_EQUALITY, CB200_CONE (accumulate = 0).
 * cb200_stage_scatter copies caches_host ([count][length], instance-major; NULL = the caller already wrote the device buffer
 * cb200_stage_buffer(h, which), laid out [batch][length]) and fills `which` for instances first .. first+count-1 on the
 * handle's stream (asynchronous).  The flat second-derivative and Jacobian caches of the same callbacks
 * (`jacobians[sparsity + count] = v`, dynamics.jl:150-170,181-205, constraints.jl:178-201,214-239) are plain concatenations of
 * the stage caches: they go through cb200_scatter unchanged. */
int cb200_stage_plan(cb200_handle *h, int which, int accumulate, int length, const int *dst);
int cb200_stage_scatter(cb200_handle *h, int which, const double *caches_host, int first_instance, int count);
void *cb200_stage_buffer(cb200_handle *h, int which);
/* out = J v for instance-major v (mul! with jacobian_variables, iterative_refinement.jl:9) -- tests / glue */
int cb200_jacobian_times(cb200_handle *h, const double *v_host, double *out_host);

/* LQ-conic family with on-device callbacks: the whole of solve! stays on the GPU */
int cb200_lq_evaluate(cb200_handle *h, int flags, int at_candidate);   /* evaluate!, evaluate.jl:1-124 */
int cb200_lq_begin(cb200_handle *h, int warmstart);                    /* solve.jl:8-95 */
/* up to `iterations` passes of solve.jl:98-368 per instance in ONE launch: instances are independent, a converged (or
 * failed) instance stops early and frees its place on the SM.  (CB200_LQ_LOCKSTEP=1 in the environment: one launch per
 * pass, all instances in lock-step -- bitwise the same results, for A/B timing.) */
int cb200_lq_step(cb200_handle *h, int iterations);
/* The filter line search of solve! (src/solver/solve.jl:224-306 with filter.jl:43-89 and line_search.jl:2-15) for callers
 * whose callbacks run outside the library: the candidates are w - alpha_k * step, alpha_k = alpha_cone * 0.5^k, k = first ..
 * first+count-1 (alpha_cone = the step size cb200_cone_search left), and the caller hands over what evaluate! returned at
 * each of them -- objective f_host[batch][count], equality constraint g_host[batch][count][m], cone constraint
 * h_host[batch][count][p] (host pointers).  Per instance the kernel forms the candidate slacks, barrier, merit and constraint
 * violation of every candidate, walks them in order through the filter / switching / Armijo / sufficient-progress tests
 * (first == 0 also computes merit, merit gradient and slope at the current point from CB200_GRADIENT, CB200_DUAL,
 * CB200_EQUALITY, CB200_CONE, CB200_S_OBJECTIVE and the barrier of the last cb200_cone call) and stops at the first
 * acceptable one: accepted_host[b] = its index k (then CB200_S_STEP_SIZE, _MERIT_CANDIDATE, _THETA_CANDIDATE and
 * CB200_I_LINE_SEARCH are set, the filter is augmented as solve.jl:298-306 does, CB200_EQUALITY / CB200_CONE hold the
 * accepted candidate's values and cb200_apply_step may follow) or -1 (call again with first += count).  The filter itself
 * lives on the device; cb200_filter_reset empties it (solve.jl:93, :367). */
int cb200_filter_reset(cb200_handle *h);
int cb200_filter_search(cb200_handle *h, int first, int count, const double *f_host, const double *g_host, const double *h_host,
                        int *accepted_host);
/* Scheduling hint for cb200_lq_step / cb200_lq_solve: the order in which the instances of the batch are started (order[k] =
 * instance started k-th; a permutation of 0 .. batch-1; NULL restores the identity).  Instances are independent, so the
 * order does not change any result; starting the instances that will need the most Newton iterations first (e.g. by the
 * iteration counts of the previous solve in a warm-started MPC loop, examples/autotuning/cartpole.jl:146-226) packs the last
 * wave of a batched solve. */
int cb200_lq_set_order(cb200_handle *h, const int *order);
/* run until every instance converged / gave up or max_steps passes; with an NCCL communicator attached the
 * termination test is the all-reduced count over ranks.  counts[4] = running, converged, gave up, error (global).
 * check_every = Newton iterations an instance may run inside one launch before the host looks at the counters (32-byte
 * read-back + the all-reduce): every instance leaves its launch as soon as it has converged, so a large value costs
 * nothing but the granularity of max_steps; small values put a host synchronisation into the loop (measured on B200:
 * 290 / 282 / 280 ms per batched solve for 4 / 8 / until convergence, profiles/r1_lq_schedule_ab.txt). */
int cb200_lq_solve(cb200_handle *h, int max_steps, int check_every, long long *counts, int *steps_done);

/* ------------------------------------------------------------------------------------------------ LinearSolver seam
 * factorize!(s, A; update=true) (linear_solver.jl:19-31: triu! + update_values! + refactor!) for the values in
 * CB200_MATRIX_VALUES; compute_inertia!(s) (:33-44); linear_solve!(s, x, A, b) (:52-60) on CB200_RHS in place. */
int cb200_ldl_factorize(cb200_handle *h);
int cb200_ldl_inertia(cb200_handle *h, int *out /* [batch][3] positive, negative, zero */);
int cb200_ldl_solve(cb200_handle *h);
/* host-buffer convenience: values in, solution out (H2D + factor + solve + D2H) = linear_solve!(fact=true) */
int cb200_ldl_linear_solve(cb200_handle *h, const double *Ax_host, const double *b_host, double *x_host, int factorize);

/* ------------------------------------------------------------------------------------------------ multi-GPU (NCCL)
 * One process per GPU, instances sharded across ranks, the only exchange is the convergence flag all-reduce. */
int cb200_nccl_unique_id(char *out128);
int cb200_comm_init(cb200_handle *h, int rank, int nranks, const char *unique_id128);
/* counts[4] (running, converged, gave up, error) of this rank's instances summed over ranks, in-stream */
int cb200_allreduce_counts(cb200_handle *h, long long *counts);

#ifdef __cplusplus
}
#endif
#endif
