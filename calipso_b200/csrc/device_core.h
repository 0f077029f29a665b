// device_core.h -- CTA-cooperative building blocks of the B200 Newton/KKT path.
//
// One CTA owns one problem instance.  Every routine below is written as "parallel-for over an index range, then
// barrier": under nvcc the index ranges are strided over the threads of the cooperating scope (whole CTA or one
// warp) and barriers are __syncthreads()/__syncwarp(); when this header is compiled by a host compiler with
// CB200_HOST_EMULATION (tests only -- never linked into the product library) the same source runs the index ranges
// sequentially so that the numerical logic can be checked on a machine without a GPU.
//
// Reference map (function -> reference file:line it replaces):
//   cone_eval            cones/cone.jl:71-106, nonnegative.jl:11-15, second_order.jl:13-17
//   residual_eval        residual.jl:1-51 + norms solve.jl:130-135, optimality_error.jl:1-27
//   kkt_entries          residual_jacobian_variables.jl:1-167 (full J never materialised; K gathered straight into
//                        the supernodal panels = triu! + update_values!, linear_solver.jl:23-24, qdldl.jl:199-213)
//   ldl_factor           refactor!/QDLDL_factor!, qdldl.jl:269-278,400-589 + compute_inertia!, linear_solver.jl:33-44
//   ldl_solve            solve!/QDLDL_solve!, qdldl.jl:330-351,592-640
//   reduced_rhs          residual_symmetric!, residual.jl:53-101
//   recover_step         search_direction_symmetric! recovery, search_direction.jl:45-101
//   jacobian_times       mul!(.., jacobian_variables, ..), iterative_refinement.jl:9,39
//   search_direction     search_direction.jl:1-23 + inertia.jl:30-79 + iterative_refinement.jl:1-53
//   cone_search          solve.jl:190-221, cones/cone.jl:62-68
#pragma once

#include <math.h>

#include "symbolic.h"

#if defined(__CUDACC__) && !defined(CB200_HOST_EMULATION)
#define CB_DEV __device__ __forceinline__
#define CB_DEVN __device__ __noinline__
#define CB_ON_DEVICE 1
#else
#define CB_DEV inline
#define CB_DEVN inline
#define CB_ON_DEVICE 0
#endif

namespace cb200 {

// ------------------------------------------------------------------------------------------------ shared data
struct Csr {          // row-wise view of a CSC matrix: entry k of row i is value[src[k]] at column col[k]
    const int *ptr, *col, *src;
};

struct DevProblem {   // everything shared by the instances of a batch (device pointers)
    int n, m, p, N, total;
    int q_nn, nsoc, tri_total;
    int nnzW, nnzG, nnzC, nnzWf;               // nnzWf: entries of the symmetric W by rows (both triangles)
    const int *soc_off, *soc_dims, *soc_tri;  // soc_tri[k]: offset of cone k's upper-triangle entries in the kx block
    // patterns
    const int *Wp, *Wi, *Wdiag;               // upper triangle CSC, Wdiag[j] = position of (j,j)
    Csr Wfull;                                 // symmetric W by rows (both triangles), src -> position in W values
    const int *Gp, *Gi, *Cp, *Ci;              // CSC
    Csr Grow, Crow;
    // symbolic factorisation
    int ns, nphases, max_w, max_nrow;
    long long panel_total;
    const int *perm;
    const int *sn_start, *rows_ptr, *rows;
    const long long *panel_off;
    const int *upd_ptr;
    const UpdateEntry *upd;
    const int *rel;
    const int *order;
    const Phase *phases;
    const int *fwd_ptr;                        // per permuted column: range of (descendant, row) pairs in fwd
    const FwdEntry *fwd;                       // (non-leaf descendants)
    const int *lcsr_ptr, *lcsr_col, *leaf_csr_pos;   // row-ordered copy of the singleton-leaf columns
    const int *lcsr_cols;                            // the columns that have singleton-leaf descendants
    int lcsr_ncols;
    const int *lcsr_rowinfo, *leaf_info;             // packed (16-byte) per-row / per-leaf descriptors of the bulk solve passes
    // fused assembly (value codes: array id << 30 | offset, see KSrc)
    const int *leaf_e_src, *leaf_piv_src, *basm_src, *basm_dst, *gasm_src, *gasm_dst, *gasm_zero;
    int n_gasm, n_gasm_zero;
    long long lcsr_total;
    const int *leaf_e_off, *leaf_e_col, *leaf_e_pos;   // flat below-diagonal entries of the singleton leaves
    const int *leaf_e4;                                // the same, packed: {panel offset, pivot column, Lcsr position, value code}
    const int *big_index;                      // [ns] -> big[] (shared-memory path) or -1
    const BigTarget *big;
    const ChainDesc *bdesc;                    // [nbig] packed descriptors (copied to shared memory when nbig <= CB_MAX_CHAIN)
    const YChunk *ychunks;
    const int *ystage_src, *ystage_dst, *ypiv;
    const unsigned *ymask;
    const int *yb_row, *yb_ptr, *yb_col;       // below-pivot entries of the staged columns, grouped by target row
    const int *dg_dst, *dg_ptr, *dg_src, *dg_piv;   // diagonal updates from single-entry descendant columns
    long long kx_total;
    const int *big_seq, *big_seq_bwd;          // shared-memory supernodes in forward / backward schedule order
    int nbig, max_sb_doubles, solve_smem;
    int nleaf;                                 // > 0: the first nleaf permuted columns are the singleton leaves and the shared-memory solve keeps x[nleaf .. N) only
    const FwdEntry *pfwd;                      // per-phase bulk pulls into the pivot columns of the shared-memory supernodes
    const int *prow, *pphase_ptr;
    int max_big_nR;
    const int *parts_fwd, *parts_bwd;          // (panel offset, doubles) of the TMA copies of the solves, in issue order
    int nparts_fwd, nparts_bwd;
    int nnzA;
};

struct Options {      // src/solver/options.jl:6-59, hot-path subset (same defaults, see api.cu)
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
    int gmres_restart, gmres_max_cycles;   // fallback when refinement fails (replaces the reference's UMFPACK J\R)
};

// cycle counters (thread 0 of each CTA, clock64) -- where a Newton iteration spends its time
enum {
    PROF_ASSEMBLE = 0, PROF_FACTOR_LEAVES, PROF_FACTOR_SMALL, PROF_FACTOR_BIG_STAGE, PROF_FACTOR_BIG_GEMM,
    PROF_FACTOR_BIG_PANEL, PROF_FACTOR_BIG_GENERIC, PROF_SOLVE_FWD, PROF_SOLVE_BWD, PROF_RHS_RECOVER, PROF_JTIMES,
    PROF_EVAL_LINESEARCH, PROF_CONE_RESIDUAL, PROF_INERTIA, PROF_TOTAL,
    // sub-phases of the shared-memory solve
    PROF_SF_BULK, PROF_SF_PULL, PROF_SF_SWEEP, PROF_SF_PUSH, PROF_SF_OTHER, PROF_SB_GATHER, PROF_SB_SWEEP, PROF_SB_OTHER,
    PROF_SB_LEAVES,
    // chain runs of the shared-memory solve, as seen by the sweep warp (forward / backward): waiting for a TMA part, sweeping,
    // waiting for the other warps
    PROF_CF_TMA, PROF_CF_SWEEP, PROF_CF_WAIT, PROF_CB_TMA, PROF_CB_WAIT, PROF_CB_SWEEP, PROF_SPARE0, PROF_SPARE1, PROF_COUNT = 32
};

// per-instance scalar slots
enum {
    S_KAPPA = 0, S_TAU, S_RHO, S_EPSP, S_EPSD, S_EPSP_LAST, S_OBJECTIVE, S_BARRIER,
    S_RESIDUAL_VIOLATION, S_OPTIMALITY_VIOLATION, S_SLACK_VIOLATION, S_THETA, S_MERIT,
    S_STEP_SIZE, S_STEP_SIZE_T, S_EQUALITY_VIOLATION, S_CONE_PRODUCT_VIOLATION,
    S_REFINE_NORM, S_REFINE_NORM_INITIAL, S_MERIT_CANDIDATE, S_THETA_CANDIDATE, S_MERIT_SLOPE, S_STEP_SIZE_CONE, S_COUNT = 24
};
enum {
    I_INERTIA_POS = 0, I_INERTIA_NEG, I_INERTIA_ZERO, I_TRIALS, I_REFINE, I_REFINE_OK, I_KS, I_KT, I_STATUS,
    I_USED_FALLBACK, I_FALLBACKS, I_TOTAL_ITERATIONS, I_OUTER, I_LINE_SEARCH, I_CONVERGED, I_GMRES_ITERS,
    I_FILTER_INDEX, I_INNER, I_FACTORIZATIONS, I_SOLVES, I_UNREFINED_STEPS, I_COUNT = 24
};
// status codes (also the C ABI's, include/calipso_b200.h)
enum { ST_OK = 0, ST_INERTIA_FAILURE = 1, ST_REFINEMENT_FAILURE = 2, ST_CONE_SEARCH_FAILURE = 3, ST_ZERO_PIVOT = 4 };

struct Inst {         // device pointers of ONE instance
    double *w, *cand, *step, *res, *err, *corr, *tmp;  // [total]
    double *grad, *gyx, *hzx;                          // [n]
    double *g, *h;                                     // [m], [p]
    double *Wv, *Gv, *Cv;                              // values at the patterns
    double *Wf, *Gr;                                   // row-ordered copies (expand_values): symmetric W by rows, G by rows
    double *prod, *bgrad;                              // [p]
    double *lambda;                                    // [m]
    double *panels, *D, *Dinv, *kx, *Lcsr;           // factor (+ M blocks of the big supernodes, leaf CSR copy)
    long long *prof;                                   // [PROF_COUNT] cycle counters (may be null)
    double *xs, *rs, *xp;                              // [N] reduced solution / rhs / permuted scratch
    double *mgrad;                                     // [N]
    double *q, *g0, *h0;                               // LQ data (may be null)
    double *filter;                                    // [2*max_filter]
    double *krylov;                                    // [(restart+1) * total] GMRES basis (may be null)
    double *scal;                                      // [S_COUNT]
    int *istat;                                        // [I_COUNT]
};

// The value arrays the assembly codes refer to.  KKT handles: W, G, C values at their patterns and the computed
// entries kx (kkt_entries); LinearSolver-seam handles: a0 = the caller's matrix values.
struct KSrc {
    const double *a0, *a1, *a2, *a3;
};
CB_DEV double ksrc_load(const KSrc &K, int code)
{
    const unsigned id = (unsigned)code >> 30;
    const double *base = id == 0 ? K.a0 : (id == 1 ? K.a1 : (id == 2 ? K.a2 : K.a3));
    return base[code & 0x3fffffff];
}

// ------------------------------------------------------------------------------------------------ cooperation scope
struct Ctx {
    int tid, nthr;     // thread index / count inside the cooperating scope
    int warp_scope;    // 1: the scope is a single warp
    double *red;       // CTA scratch for reductions (>= 34 doubles), unused in warp scope
    double *scratch;   // shared-memory work area of the CTA (P.scratch_doubles doubles); heap in host emulation
    unsigned long long *bars;   // two mbarriers in static shared memory, initialised once per kernel (device only)
    unsigned *bar_uses;         // number of TMA copies issued so far in this kernel (selects slot and phase parity)
    CB_DEV void sync() const
    {
#if CB_ON_DEVICE
        if (warp_scope) __syncwarp(); else __syncthreads();
#endif
    }
};

#if CB_ON_DEVICE
#define PAR_FOR(i, count) for (int i = ctx.tid; i < (count); i += ctx.nthr)
// The CTA's shared memory is named directly (not through the generic pointers of Ctx) so that the compiler emits
// LDS/STS and knows that shared and global accesses cannot alias.
extern __shared__ __align__(16) double cb_dyn_smem[];
__shared__ double cb_red[34];
__shared__ __align__(8) unsigned long long cb_bars[2];
__shared__ unsigned cb_bar_uses;
__shared__ long long cb_prof[32];
#define CB_MAX_CHAIN 64
__shared__ int4 cb_chain[2 * CB_MAX_CHAIN];   // ChainDesc of the shared-memory supernodes (filled once per kernel)
__shared__ int cb_chain_n;
#define CB_MAX_PHASES 48
__shared__ int cb_phase[6 * CB_MAX_PHASES];     // the schedule (Phase records) and the bulk-pull ranges, cached per kernel
__shared__ int cb_pphase[CB_MAX_PHASES + 1];
__shared__ int cb_phase_n;                     // 0: read the schedule from global memory                     // number of valid entries (0: read descriptors from global memory)      // phase counters of this CTA, flushed to the instance's slots when the kernel ends
#define CB_SCRATCH(ctx) (cb_dyn_smem)
#define CB_RED(ctx) (cb_red)
#else
#define PAR_FOR(i, count) for (int i = 0; i < (count); i++)
#define CB_SCRATCH(ctx) ((ctx).scratch)
#define CB_RED(ctx) ((ctx).red)
#endif

// Reductions over f(i), i in [0,n): every thread of the scope returns the same value.  Fixed association order
// (thread-strided partials, shuffle tree, warp partials in warp order) => run-to-run deterministic.
template <class F> CB_DEV double scope_sum(const Ctx &ctx, int n, F f)
{
    double a = 0.0;
    PAR_FOR(i, n) a += f(i);
#if CB_ON_DEVICE
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (ctx.warp_scope) return a;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) CB_RED(ctx)[wid] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < nw; k++) t += CB_RED(ctx)[k];
        CB_RED(ctx)[33] = t;
    }
    __syncthreads();
    a = CB_RED(ctx)[33];
#endif
    return a;
}

template <class F> CB_DEV double scope_max(const Ctx &ctx, int n, F f)
{
    double a = 0.0;  // all uses are maxima of absolute values
    PAR_FOR(i, n) { double v = f(i); a = v > a ? v : a; }
#if CB_ON_DEVICE
    for (int o = 16; o > 0; o >>= 1) { double b = __shfl_xor_sync(0xffffffffu, a, o); a = b > a ? b : a; }
    if (ctx.warp_scope) return a;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) CB_RED(ctx)[wid] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < nw; k++) t = CB_RED(ctx)[k] > t ? CB_RED(ctx)[k] : t;
        CB_RED(ctx)[33] = t;
    }
    __syncthreads();
    a = CB_RED(ctx)[33];
#endif
    return a;
}

// integer "any" over the scope (used for violation tests)
template <class F> CB_DEV int scope_any(const Ctx &ctx, int n, F f)
{
    return scope_max(ctx, n, [&](int i) { return f(i) ? 1.0 : 0.0; }) > 0.5;
}

// Sparse rows handled by groups of G lanes (G = 4: 64 rows of a 256-thread CTA in flight, short dependent-load chains):
// part(row, sub, G) returns the lane's partial sum over the entries sub, sub + G, ... of the row, fin(row, sum) runs on
// the group's first lane.  Fixed association order.  Host emulation: one "lane" per row.
template <int G, class Part, class Fin> CB_DEV void grouped_rows(const Ctx &ctx, int nrows, Part part, Fin fin)
{
#if CB_ON_DEVICE
    if (!ctx.warp_scope) {
        const int sub = threadIdx.x & (G - 1), grp = threadIdx.x / G, ngrp = blockDim.x / G;
        const int padded = (nrows + ngrp - 1) / ngrp * ngrp;      // every warp runs every shuffle
        for (int r = grp; r < padded; r += ngrp) {
            double acc = r < nrows ? part(r, sub, G) : 0.0;
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (sub == 0 && r < nrows) fin(r, acc);
        }
        return;
    }
#endif
    PAR_FOR(r, nrows) fin(r, part(r, 0, 1));
}

// ------------------------------------------------------------------------------------------------ second-order cone algebra
// arrow(u)^-1 x with the reference's closed form and operation order (second_order.jl:50-61); u, x given as
// element functors so that no temporary vectors are needed; out(i, value) receives the result.
template <class U, class X, class O> CB_DEV void soc_arrow_inverse(int d, U u, X x, O out)
{
    double u1 = u(0);
    double uu = 0.0;
    for (int i = 1; i < d; i++) uu += u(i) * u(i);
    double alpha = -1.0 / (u1 * u1) * uu;
    double beta = 1.0 / (1.0 + alpha);
    double acc = 0.0;
    for (int i = 1; i < d; i++) acc += (u(i) / u1) * x(i);
    double x0_1 = x(0) - acc;
    double acc2 = 0.0;
    for (int i = 1; i < d; i++) {
        double x1_i = x(i) - beta * ((u(i) / u1) * x0_1);
        acc2 += (u(i) / u1) * x1_i;
        out(i, 1.0 / u1 * x1_i);
    }
    out(0, 1.0 / u1 * (x(0) - acc2));
}

// [arrow(a) v]_i given dot = a . v
CB_DEV double arrow_apply(const double *a, const double *v, int i, double dot)
{
    return i == 0 ? dot : a[0] * v[i] + v[0] * a[i];
}

// ------------------------------------------------------------------------------------------------ cone!
CB_DEVN void cone_eval(const Ctx &ctx, const DevProblem &P, const Inst &I, const double *w, int barrier,
                       int barrier_gradient, int product)
{
    const double *s = w + P.n + P.m, *t = w + P.n + 2 * P.m + 2 * P.p;
    if (barrier) {
        double nn = scope_sum(ctx, P.q_nn, [&](int i) { return log(s[i]); });
        double so = scope_sum(ctx, P.nsoc, [&](int k) {
            const double *x = s + P.soc_off[k];
            int d = P.soc_dims[k];
            if (d <= 0) return 0.0;
            double tt = 0.0;
            for (int i = 1; i < d; i++) tt += x[i] * x[i];
            return 0.5 * log(x[0] * x[0] - tt);
        });
        if (ctx.tid == 0) I.scal[S_BARRIER] = nn + so;
    }
    if (barrier_gradient) {
        PAR_FOR(i, P.q_nn) I.bgrad[i] = 1.0 / s[i];
        PAR_FOR(k, P.nsoc) {
            const double *x = s + P.soc_off[k];
            double *g = I.bgrad + P.soc_off[k];
            int d = P.soc_dims[k];
            if (d > 0) {
                double tt = 0.0;
                for (int i = 1; i < d; i++) tt += x[i] * x[i];
                double c = 1.0 / (x[0] * x[0] - tt);
                g[0] = c * x[0];
                for (int i = 1; i < d; i++) g[i] = c * (-x[i]);
            }
        }
    }
    if (product) {
        PAR_FOR(i, P.q_nn) I.prod[i] = s[i] * t[i];
        PAR_FOR(k, P.nsoc) {
            const double *a = s + P.soc_off[k], *b = t + P.soc_off[k];
            double *o = I.prod + P.soc_off[k];
            int d = P.soc_dims[k];
            if (d > 0) {
                double dot = 0.0;
                for (int i = 0; i < d; i++) dot += a[i] * b[i];
                o[0] = dot;
                for (int i = 1; i < d; i++) o[i] = a[0] * b[i] + b[0] * a[i];
            }
        }
    }
    ctx.sync();
}

// ------------------------------------------------------------------------------------------------ residual! + norms
CB_DEVN void residual_eval(const Ctx &ctx, const DevProblem &P, const Inst &I)
{
    const int n = P.n, m = P.m, p = P.p;
    const double *w = I.w;
    const double *r = w + n, *s = w + n + m, *y = w + n + m + p, *z = w + n + 2 * m + p, *t = w + n + 2 * m + 2 * p;
    double *res = I.res;
    const double kappa = I.scal[S_KAPPA], rho = I.scal[S_RHO];
    PAR_FOR(i, n) res[i] = I.grad[i] + I.gyx[i] + I.hzx[i];
    PAR_FOR(i, m) {
        res[n + i] = I.lambda[i] + rho * r[i] - y[i];
        res[n + m + p + i] = I.g[i] - r[i];
    }
    PAR_FOR(i, p) {
        res[n + m + i] = -z[i] - t[i];
        res[n + 2 * m + p + i] = I.h[i] - s[i];
    }
    PAR_FOR(i, P.q_nn) res[n + 2 * m + 2 * p + i] = I.prod[i] - kappa;
    PAR_FOR(k, P.nsoc) {
        int c0 = P.soc_off[k], d = P.soc_dims[k];
        for (int i = 0; i < d; i++) res[n + 2 * m + 2 * p + c0 + i] = I.prod[c0 + i] - (i == 0 ? kappa : 0.0);
    }
    ctx.sync();
    // violations (solve.jl:130-135, optimality_error.jl:1-27)
    double r1 = scope_sum(ctx, P.total, [&](int i) { return fabs(res[i]); });
    double lag = scope_max(ctx, n + m + p, [&](int i) { return fabs(res[i]); });
    double eq = scope_max(ctx, m, [&](int i) { return fabs(res[n + m + p + i]); });
    double cn = scope_max(ctx, p, [&](int i) { return fabs(res[n + 2 * m + p + i]); });
    double cp = scope_max(ctx, p, [&](int i) { return fabs(res[n + 2 * m + 2 * p + i]); });
    double y1 = scope_sum(ctx, m, [&](int i) { return fabs(y[i]); });
    double z1 = scope_sum(ctx, p, [&](int i) { return fabs(z[i]); });
    double t1 = scope_sum(ctx, p, [&](int i) { return fabs(t[i]); });
    if (ctx.tid == 0) {
        double sd = (m + p) > 0 ? fmax(100.0, (y1 + z1) / (double)(m + p)) / 100.0 : 1.0;
        double sc = p > 0 ? fmax(100.0, t1 / (double)p) / 100.0 : 1.0;
        I.scal[S_RESIDUAL_VIOLATION] = r1 / (double)P.total;
        I.scal[S_OPTIMALITY_VIOLATION] = fmax(fmax(lag / sd, eq), fmax(cn, cp / sc));
        I.scal[S_SLACK_VIOLATION] = fmax(eq, cn);
    }
    ctx.sync();
}

// theta = ||[g - r; h - s]||_1 / (m + p) at point w (constraint_violation.jl:1-13)
CB_DEV double constraint_violation(const Ctx &ctx, const DevProblem &P, const Inst &I, const double *w)
{
    const double *r = w + P.n, *s = w + P.n + P.m;
    double a = scope_sum(ctx, P.m, [&](int i) { return fabs(I.g[i] - r[i]); });
    double b = scope_sum(ctx, P.p, [&](int i) { return fabs(I.h[i] - s[i]); });
    return (a + b) / (double)(P.m + P.p);
}

// M = f + lambda'r + rho/2 r'r - kappa Phi at point w (merit.jl:2-15)
CB_DEV double merit_value(const Ctx &ctx, const DevProblem &P, const Inst &I, const double *w)
{
    const double *r = w + P.n;
    double lr = scope_sum(ctx, P.m, [&](int i) { return I.lambda[i] * r[i]; });
    double rr = scope_sum(ctx, P.m, [&](int i) { return r[i] * r[i]; });
    double M = 0.0;
    M += I.scal[S_OBJECTIVE];
    M += lr + 0.5 * I.scal[S_RHO] * rr;
    M -= I.scal[S_KAPPA] * I.scal[S_BARRIER];
    return M;
}

CB_DEV void merit_gradient(const Ctx &ctx, const DevProblem &P, const Inst &I)
{   // merit.jl:17-31
    const double *r = I.w + P.n;
    const double rho = I.scal[S_RHO], kappa = I.scal[S_KAPPA];
    PAR_FOR(i, P.n) I.mgrad[i] = I.grad[i];
    PAR_FOR(i, P.m) I.mgrad[P.n + i] = I.lambda[i] + rho * r[i];
    PAR_FOR(i, P.p) I.mgrad[P.n + P.m + i] = -1.0 * kappa * I.bgrad[i];
    ctx.sync();
}

// ------------------------------------------------------------------------------------------------ KKT assembly
// Row-ordered copies of the callback outputs, refreshed whenever W or G values change: Wf = symmetric W by rows (both
// triangles), Gr = G by rows.  Every row-wise consumer (J v, the LQ callbacks, the multiplier leaves of the
// factorisation, whose columns are rows of G) then reads contiguous values instead of gathering through the CSC storage.
CB_DEVN void expand_values(const Ctx &ctx, const DevProblem &P, const Inst &I)
{
    PAR_FOR(k, P.nnzWf) I.Wf[k] = I.Wv[P.Wfull.src[k]];
    PAR_FOR(k, P.nnzG) I.Gr[k] = I.Gv[P.Grow.src[k]];
    ctx.sync();
}

// evaluate!'s cache scatter as a gather (plan: host_setup.h ScatterPlan): out[e] = sum over the caches, in order, of the
// last cache entry written to pattern entry e (0.0 where a cache has none) -- the dense-matrix `=` scatter of
// src/solver/evaluate.jl:39-41,75-77,111-113 followed by the sum of residual_jacobian_variables.jl:11-13.
CB_DEV void scatter_caches(const Ctx &ctx, int nnz, int ncaches, const int *__restrict__ idx,
                           const double *__restrict__ cache, double *__restrict__ out, int first, int stride)
{
    for (int e = first + ctx.tid; e < nnz; e += stride) {
        const int i0 = idx[e];
        double v = i0 >= 0 ? cache[i0] : 0.0;
        for (int c = 1; c < ncaches; c++) {
            const int i = idx[(long long)c * nnz + e];
            v += i >= 0 ? cache[i] : 0.0;
        }
        out[e] = v;
    }
}

// Stage-level scatter of the trajectory-optimisation front end as a gather (plan: host_setup.h StagePlan): output entry i
// receives, after fill!(out, 0.0), the entries src[ptr[i] .. ptr[i + 1]) of the concatenated per-stage caches in program order
// -- added one by one (`gradient[idx...] += cache[i]`, trajectory_optimization/dynamics.jl:172-179, constraints.jl:203-212,
// `gradient[indices[t]] .+= cache`, costs.jl:115-120) or only the last of them (`violations[indices[t]] .= cache`,
// dynamics.jl:143-148, constraints.jl:169-176).
CB_DEV void stage_gather(const Ctx &ctx, int nout, int accumulate, const int *__restrict__ ptr, const int *__restrict__ src,
                         const double *__restrict__ cache, double *__restrict__ out, int first, int stride)
{
    for (int i = first + ctx.tid; i < nout; i += stride) {
        const int k0 = ptr[i], k1 = ptr[i + 1];
        double v = 0.0;
        if (accumulate) {
            for (int k = k0; k < k1; k++) v += cache[src[k]];
        } else if (k1 > k0) {
            v = cache[src[k1 - 1]];
        }
        out[i] = v;
    }
}

// The entries of the reduced matrix K that are not plain copies of W, G, C values:
// kx = [W diagonal + eps_p | y diagonal | nonnegative z diagonal | upper triangles of the second-order z blocks].
// eps_p, eps_d, rho enter exactly where residual_jacobian_variables.jl:83-105,131,143-164 puts them.  The matrix itself
// is never stored: the factorisation gathers W, G, C values and these entries straight into its panels (triu! +
// update_values!, linear_solver.jl:23-24, qdldl.jl:199-213, fused with refactor!).
CB_DEVN void kkt_entries(const Ctx &ctx, const DevProblem &P, const Inst &I)
{
    const double ep = I.scal[S_EPSP], ed = I.scal[S_EPSD], rho = I.scal[S_RHO];
    const double *s = I.w + P.n + P.m, *t = I.w + P.n + 2 * P.m + 2 * P.p;
    double *kw = I.kx, *ky = kw + P.n, *kz = ky + P.m, *ks = kz + P.q_nn;
    const double Jrr = rho + ep, Jyy = -ed, Jzz = -ed, Jss = ep;
    PAR_FOR(j, P.n) kw[j] = I.Wv[P.Wdiag[j]] + ep;
    PAR_FOR(i, P.m) ky[i] = -1.0 / Jrr + Jyy;
    PAR_FOR(i, P.q_nn) {
        double Sb = s[i] - ed, Ti = t[i];
        kz[i] = -1.0 * Sb / (Ti + Sb * Jss) + Jzz;
    }
    // SOC blocks: column j (rows i <= j) of -(arrow(u))^-1 Sbar + D, u = first row of T + Sbar*P
    PAR_FOR(k, P.nsoc) {
        int d = P.soc_dims[k];
        const double *sk = s + P.soc_off[k], *tk = t + P.soc_off[k];
        double *dst = ks + P.soc_tri[k];
        auto u = [&](int i) { return i == 0 ? tk[0] + (sk[0] - ed) * Jss : tk[i] + sk[i] * Jss; };
        int tri = 0;
        for (int j = 0; j < d; j++) {
            auto x = [&](int i) {   // column j of Sbar = arrow(s) - eps_d I
                if (j == 0) return i == 0 ? sk[0] - ed : sk[i];
                return i == 0 ? sk[j] : (i == j ? sk[0] - ed : 0.0);
            };
            soc_arrow_inverse(d, u, x, [&](int i, double v) {
                if (i <= j) {
                    double e = 0.0;
                    e -= v;
                    if (i == j) e += Jzz;
                    dst[tri + i] = e;
                }
            });
            tri += j + 1;
        }
    }
    ctx.sync();
}

// ------------------------------------------------------------------------------------------------ profiling
struct ProfTimer {       // thread 0 of the CTA accumulates clock64() deltas into per-instance slots
    long long *slots;
    long long t0;
    CB_DEV void start()
    {
#if CB_ON_DEVICE
        if (slots && threadIdx.x == 0) t0 = clock64();
#endif
    }
    CB_DEV void stop(int slot)
    {
#if CB_ON_DEVICE
        if (slots && threadIdx.x == 0) { long long t1 = clock64(); cb_prof[slot] += t1 - t0; t0 = t1; }
#endif
    }
};

// ------------------------------------------------------------------------------------------------ supernodal LDL^T
// In-place LDL' of a column-major block: `rows` x `w`, leading dimension ld, pivots on the diagonal of the top w x w
// part.  Unscaled columns are kept until the end so that each pivot step needs one barrier; the final pass divides by
// the pivots (L = A D^-1).  Works on shared or global memory.
CB_DEV void panel_factor(const Ctx &ctx, double *Ps, int rows, int w, int ld)
{
    for (int k = 0; k < w; k++) {
        ctx.sync();
        const double dk = Ps[k + (long long)k * ld];
        const double dinv = dk != 0.0 ? 1.0 / dk : 0.0;
        const int ncol = w - 1 - k;
        PAR_FOR(e, ncol * rows) {
            int j = k + 1 + e / rows, i = e % rows;
            if (i >= j) Ps[i + (long long)j * ld] -= Ps[i + (long long)k * ld] * (Ps[j + (long long)k * ld] * dinv);
        }
    }
    ctx.sync();
    PAR_FOR(e, w * rows) {
        int k = e / rows, i = e % rows;
        if (i > k) {
            double dk = Ps[k + (long long)k * ld];
            Ps[i + (long long)k * ld] *= (dk != 0.0 ? 1.0 / dk : 0.0);
        }
    }
    ctx.sync();
}

// Generic supernode (any scope, panel in global memory): pull the updates of the descendants one by one, then
// factor the (nrow x w) panel in place.
CB_DEV void factor_supernode(const Ctx &ctx, const DevProblem &P, double *pan, double *D, double *Dinv, int s)
{
    const int c0 = P.sn_start[s], w = P.sn_start[s + 1] - c0;
    const int nrow = w + (P.rows_ptr[s + 1] - P.rows_ptr[s]);
    double *Ps = pan + P.panel_off[s];
    for (int q = P.upd_ptr[s]; q < P.upd_ptr[s + 1]; q++) {
        const UpdateEntry u = P.upd[q];
        const int cd0 = P.sn_start[u.d], wd = P.sn_start[u.d + 1] - cd0;
        const int nRd = P.rows_ptr[u.d + 1] - P.rows_ptr[u.d], nrowd = wd + nRd;
        const double *Pd = pan + P.panel_off[u.d] + wd + u.a;   // first contributing row of d
        const double *Dd = D + cd0;
        const int *rel = P.rel + u.rel;
        const int nt = nRd - u.a, nb = u.b - u.a;
        PAR_FOR(e, nt * nb) {
            int i = e % nt, j = e / nt;
            if (i >= j) {
                double acc = 0.0;
                for (int k = 0; k < wd; k++) acc += Pd[i + (long long)k * nrowd] * (Pd[j + (long long)k * nrowd] * Dd[k]);
                Ps[rel[i] + (long long)rel[j] * nrow] -= acc;
            }
        }
        ctx.sync();
    }
    panel_factor(ctx, Ps, nrow, w, nrow);
    PAR_FOR(k, w) {
        double dk = Ps[k + (long long)k * nrow];
        D[c0 + k] = dk;
        Dinv[c0 + k] = dk != 0.0 ? 1.0 / dk : 0.0;
    }
    ctx.sync();
}

#if CB_ON_DEVICE
// ---- asynchronous copies: TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

#if CB_ON_DEVICE
// FP64 tensor-core tile: C(8x8) += A(8x4) B(4x8), mma.sync m8n8k4 (SASS DMMA).  Lane l holds A[l/4][l%4], B[l%4][l/4],
// C[l/4][2(l%4)] and C[l/4][2(l%4)+1].
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// Blocked right-looking LDL' of the (rows x w) column-major panel S in shared memory (leading dimension ld = 4 mod 8,
// every row below `rows` up to the next multiple of 8 readable and finite).  Eight columns at a time:
//   1. one thread per row keeps its 8 entries in registers; the warp that owns the 8 pivot rows eliminates them with
//      shuffles (no CTA barrier inside the block) and publishes the unscaled pivot rows U and the reciprocals;
//   2. after one barrier the other warps eliminate their rows against U, every row writes L back, the rows that are
//      also columns of the trailing part write their unscaled multipliers L*D to LD;
//   3. the trailing columns get their rank-8 update S -= L (L D)' on the FP64 tensor cores, one warp per 8-row strip.
// Same arithmetic as panel_factor (L = A D^-1, zero pivot => zero column).  dd[] receives the pivots.
__device__ __forceinline__ void panel_factor_tc(double *__restrict__ S, int rows, int w, int ld, double *__restrict__ U,
                                                double *__restrict__ dinvs, double *__restrict__ dd,
                                                double *__restrict__ LD, int ldw)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int i = tid;                       // this thread's row (rows <= blockDim.x)
    for (int kb = 0; kb < w; kb += 8) {
        const int nb = min(8, w - kb);
        const bool active = i >= kb && i < rows;
        double a[8], u[8];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = (active && c < nb) ? S[i + (kb + c) * ld] : 0.0;
        const int pw = kb >> 5, base = kb & 31;
        if (warp == pw) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k < nb) {
                    double pr[8];
#pragma unroll
                    for (int cc = 0; cc < 8; cc++)
                        if (cc >= k) pr[cc] = __shfl_sync(0xffffffffu, a[k], base + cc);   // A[kb+cc, kb+k], unscaled
                    const double d = pr[k];
                    const double dinv = d != 0.0 ? 1.0 / d : 0.0;
                    if (lane == 0) {
                        dinvs[k] = dinv;
                        dd[kb + k] = d;
#pragma unroll
                        for (int cc = 0; cc < 8; cc++)
                            if (cc > k) U[k * 8 + cc] = pr[cc];
                    }
                    u[k] = a[k];
                    if (active && i > kb + k) {
                        const double lik = a[k] * dinv;
#pragma unroll
                        for (int cc = 0; cc < 8; cc++)
                            if (cc > k) a[cc] -= lik * pr[cc];
                        a[k] = lik;
                    }
                }
            }
        }
        __syncthreads();
        if (warp != pw && active) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k < nb) {
                    u[k] = a[k];
                    const double lik = a[k] * dinvs[k];
#pragma unroll
                    for (int cc = 0; cc < 8; cc++)
                        if (cc > k && cc < nb) a[cc] -= lik * U[k * 8 + cc];
                    a[k] = lik;
                }
            }
        }
        const int j0 = kb + nb;
        if (active) {
#pragma unroll
            for (int c = 0; c < 8; c++)
                if (c < nb && i > kb + c) S[i + (kb + c) * ld] = a[c];
            if (j0 < w && i >= j0 && i < w) {
#pragma unroll
                for (int c = 0; c < 8; c++) LD[c * ldw + i] = u[c];
            }
        }
        __syncthreads();
        if (j0 < w) {      // nb == 8 here
            const int ntI = (rows + 7) >> 3, ntJ = (w + 7) >> 3, tj0 = j0 >> 3;
            for (int ti = tj0 + warp; ti < ntI; ti += nwarps) {
                const double a0 = -S[8 * ti + gid + (kb + tig) * ld], a1 = -S[8 * ti + gid + (kb + 4 + tig) * ld];
                const int r = 8 * ti + gid;
                const int tjmax = min(ti, ntJ - 1);
                for (int tj = tj0; tj <= tjmax; tj++) {
                    const double b0 = LD[tig * ldw + 8 * tj + gid], b1 = LD[(4 + tig) * ldw + 8 * tj + gid];
                    const int col = 8 * tj + 2 * tig;
                    double c0 = col < w ? S[r + col * ld] : 0.0, c1 = col + 1 < w ? S[r + (col + 1) * ld] : 0.0;
                    dmma_8x8x4(c0, c1, a0, b0);
                    dmma_8x8x4(c0, c1, a1, b1);
                    if (r < rows) {
                        if (col < w) S[r + col * ld] = c0;
                        if (col + 1 < w) S[r + (col + 1) * ld] = c1;
                    }
                }
            }
            __syncthreads();
        }
    }
}
#endif

// CTA-scope supernode on the shared-memory path.  All descendant columns that touch the target are staged as the
// columns of a dense matrix Y (rows = target rows) and applied as ONE GEMM  S -= Y diag(D) Y_top' instead of one
// barrier per descendant: on the device on the FP64 tensor cores (mma.sync m8n8k4), 8x8 output tiles, skipping the
// 8x4 blocks of Y that are structurally zero (P.ymask).  Layout of the work area: S[ldp x w] | Y[ldy x kc4] | Dy[kc4];
// the staging lists carry the pivots of the Y columns as extra entries (src < 0 -> D[-1 - src]); the Y area is reused
// by the panel factorisation (U, reciprocals, pivots, unscaled multipliers).
CB_DEV void factor_supernode_big(const Ctx &ctx, const DevProblem &P, double *pan, double *D, double *Dinv,
                                 const KSrc &K, int s, const BigTarget bt, ProfTimer &pt)
{
    const int c0 = P.sn_start[s], w = P.sn_start[s + 1] - c0;
    const int nR = P.rows_ptr[s + 1] - P.rows_ptr[s], nrow = w + nR;
    const int ldp = bt.ldp, ldy = bt.ldy;
    double *Ps = pan + P.panel_off[s];
    double *S = CB_SCRATCH(ctx);
    double *Y = S + (long long)ldp * w;
    pt.start();
    // fused assembly: clear the work area and gather the supernode's input entries (W values, computed entries)
#if CB_ON_DEVICE
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    {
        const int na = bt.asm_end - bt.asm_begin, nthr = ctx.nthr;
        const int *__restrict__ asrc = P.basm_src + bt.asm_begin, *__restrict__ adst = P.basm_dst + bt.asm_begin;
        int sc[4], dc[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {          // first batch of indices in flight while S is cleared
            const int ee = ctx.tid + u * nthr;
            sc[u] = ee < na ? asrc[ee] : 0;
            dc[u] = ee < na ? adst[ee] : -1;
        }
        double2 *S2 = reinterpret_cast<double2 *>(S);
        const int n2 = (ldp * w) >> 1;          // ldp is a multiple of 4
        for (int e = ctx.tid; e < n2; e += nthr) S2[e] = make_double2(0.0, 0.0);
        __syncthreads();
        for (int e = ctx.tid;;) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = dc[u] >= 0 ? ksrc_load(K, sc[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (dc[u] >= 0) S[dc[u]] = v[u];
            e += 4 * nthr;
            if (e >= na) break;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int ee = e + u * nthr;
                sc[u] = ee < na ? asrc[ee] : 0;
                dc[u] = ee < na ? adst[ee] : -1;
            }
        }
    }
#else
    PAR_FOR(e, ldp * w) S[e] = 0.0;
    PAR_FOR(e, bt.asm_end - bt.asm_begin) S[P.basm_dst[bt.asm_begin + e]] = ksrc_load(K, P.basm_src[bt.asm_begin + e]);
#endif
    ctx.sync();
    for (int ci = bt.chunk_begin; ci < bt.chunk_end; ci++) {
        const YChunk ch = P.ychunks[ci];
        const int kc = ch.col_end - ch.col_begin, kc4 = (kc + 3) & ~3;
        const double *Dy = Y + (long long)ldy * kc4;
        const int nst = ch.stage_end - ch.stage_begin;
#if CB_ON_DEVICE
        {
            const int *__restrict__ src = P.ystage_src + ch.stage_begin, *__restrict__ dst = P.ystage_dst + ch.stage_begin;
            const double *__restrict__ pang = pan, *__restrict__ Dg = D;
            int sidx[8], didx[8];
            const int nthr = ctx.nthr;
#pragma unroll
            for (int u = 0; u < 8; u++) {     // indices of the first batch are in flight while Y is cleared
                const int ee = ctx.tid + u * nthr;
                sidx[u] = ee < nst ? src[ee] : 0x7fffffff;
                didx[u] = ee < nst ? dst[ee] : 0;
            }
            double2 *Y2 = reinterpret_cast<double2 *>(Y);
            const int n2 = ((ldy + 1) * kc4) >> 1;     // Y and Dy; kc4 is a multiple of 4
            for (int e = ctx.tid; e < n2; e += nthr) Y2[e] = make_double2(0.0, 0.0);
            __syncthreads();
            for (int e = ctx.tid;;) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    v[u] = sidx[u] == 0x7fffffff ? 0.0 : (sidx[u] >= 0 ? pang[sidx[u]] : Dg[-1 - sidx[u]]);
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (sidx[u] != 0x7fffffff) Y[didx[u]] = v[u];
                e += 8 * nthr;
                if (e >= nst) break;
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int ee = e + u * nthr;
                    sidx[u] = ee < nst ? src[ee] : 0x7fffffff;
                    didx[u] = ee < nst ? dst[ee] : 0;
                }
            }
        }
#else
        PAR_FOR(e, (ldy + 1) * kc4) Y[e] = 0.0;
        PAR_FOR(e, nst) {
            const int sidx = P.ystage_src[ch.stage_begin + e];
            Y[P.ystage_dst[ch.stage_begin + e]] = sidx >= 0 ? pan[sidx] : D[-1 - sidx];
        }
#endif
        ctx.sync();
        pt.stop(PROF_FACTOR_BIG_STAGE);
#if CB_ON_DEVICE
        {   // S[i, j] -= sum_c Y[i, c] Dy[c] Y[j, c] for the 8x8 tiles that touch the lower triangle
            const int gid = lane >> 2, tig = lane & 3;
            const int ntI = (w + 7) >> 3, ntJ = ntI;          // Y holds the pivot rows only
            const unsigned *__restrict__ msk = P.ymask + ch.mask_begin;
            const bool wide = ntI > 32;                                // (then the masks are read from memory per tile)
            const unsigned mymask = lane < ntI ? msk[lane] : 0u;
            int ti = wid, tj = 0;                                      // tiles in column-major order, warp-strided
            while (ti >= ntI) { ti -= ntI; tj++; }
            while (tj < ntJ) {
                unsigned m = wide ? (msk[ti] & msk[tj])
                                  : (__shfl_sync(0xffffffffu, mymask, ti) & __shfl_sync(0xffffffffu, mymask, tj));
                if (ti >= tj && m) {
                    const double *ya = Y + 8 * ti + gid + tig * ldy, *yb = Y + 8 * tj + gid + tig * ldy;
                    const double *dy = Dy + tig;
                    // dense loop over the span of column groups present in both tile rows (block-structured sparsity:
                    // the span is tight); regular addressing, two accumulators
                    const int g0 = __ffs(m) - 1, g1 = 32 - __clz(m);
                    ya += 4 * g0 * ldy; yb += 4 * g0 * ldy; dy += 4 * g0;
                    double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
                    int g = g0;
#pragma unroll 2
                    for (; g + 1 < g1; g += 2) {
                        const double a0 = ya[0], b0 = yb[0] * dy[0], a1 = ya[4 * ldy], b1 = yb[4 * ldy] * dy[4];
                        dmma_8x8x4(e0, e1, a0, b0);
                        dmma_8x8x4(f0, f1, a1, b1);
                        ya += 8 * ldy; yb += 8 * ldy; dy += 8;
                    }
                    if (g < g1) dmma_8x8x4(e0, e1, ya[0], yb[0] * dy[0]);
                    const int r = 8 * ti + gid, col = 8 * tj + 2 * tig;
                    if (r < w) {
                        if (col < w) S[r + col * ldp] -= e0 + f0;
                        if (col + 1 < w) S[r + (col + 1) * ldp] -= e1 + f1;
                    }
                }
                ti += nw;
                while (ti >= ntI) { ti -= ntI; tj++; }
            }
        }
#else
        PAR_FOR(e, w * w) {
            const int j = e / w, i = e % w;
            if (i >= j) {
                double acc = 0.0;
                for (int cc = 0; cc < kc; cc++) acc += Y[i + (long long)cc * ldy] * (Y[j + (long long)cc * ldy] * Dy[cc]);
                S[i + (long long)j * ldp] -= acc;
            }
        }
#endif
        {   // rows below the pivots: S[r, 0:w] -= sum_e l_e Dy[c_e] Y[0:w, c_e] (few entries; scalars staged after Dy)
            const int ng = ch.bg_end - ch.bg_begin, e_first = P.yb_ptr[ch.bg_begin];
            const double *Bl = Dy + kc4;
            PAR_FOR(it, ng * w) {
                const int g = it / w, j = it - g * w;
                double acc = 0.0;
                for (int e = P.yb_ptr[ch.bg_begin + g]; e < P.yb_ptr[ch.bg_begin + g + 1]; e++) {
                    const int cc = P.yb_col[e];
                    acc += (Bl[e - e_first] * Dy[cc]) * Y[j + (long long)cc * ldy];
                }
                S[P.yb_row[ch.bg_begin + g] + (long long)j * ldp] -= acc;
            }
        }
        ctx.sync();
        pt.stop(PROF_FACTOR_BIG_GEMM);
    }
    {   // descendant columns with a single entry in this target: S[r, r] -= sum l^2 d (grouped by destination)
        const int ng = bt.dg_end - bt.dg_begin;
        PAR_FOR(g, ng) {
            double acc = 0.0;
            for (int e = P.dg_ptr[bt.dg_begin + g]; e < P.dg_ptr[bt.dg_begin + g + 1]; e++) {
                const double l = pan[P.dg_src[e]];
                acc += l * (l * D[P.dg_piv[e]]);
            }
            S[P.dg_dst[bt.dg_begin + g]] -= acc;
        }
        if (ng > 0) ctx.sync();
    }
#ifdef CB_PROF_PANEL_SPLIT
    pt.stop(PROF_SPARE0);
#endif
    double *dd = Y + 72;                 // w pivots
#if CB_ON_DEVICE
    if (nrow <= (int)blockDim.x) {
        const int ldw = ((w + 7) & ~7) + 4;
        panel_factor_tc(S, nrow, w, ldp, Y, Y + 64, dd, Y + 72 + ((w + 3) & ~3), ldw);
    } else {
        panel_factor(ctx, S, nrow, w, ldp);
        PAR_FOR(k, w) dd[k] = S[k + (long long)k * ldp];
        ctx.sync();
    }
#ifdef CB_PROF_PANEL_SPLIT
    __syncthreads();
    pt.stop(PROF_FACTOR_BIG_PANEL);
#endif
    // write back the factor panel; the diagonal and the upper triangle of the pivot block are stored as zeros so
    // that the triangular sweeps of the solves need no predicates
    for (int k = wid; k < w; k += nw) {
        const double *src = S + k * ldp;
        for (int i = lane; i < nrow; i += 32) Ps[i + k * nrow] = i <= k ? 0.0 : src[i];   // strictly lower part only
    }
#else
    panel_factor(ctx, S, nrow, w, ldp);
    PAR_FOR(k, w) dd[k] = S[k + (long long)k * ldp];
    ctx.sync();
    PAR_FOR(e, nrow * w) {
        int k = e / nrow, i = e % nrow;
        Ps[e] = i <= k ? 0.0 : S[i + (long long)k * ldp];
    }
#endif
    PAR_FOR(k, w) {
        const double dk = dd[k];
        D[c0 + k] = dk;
        Dinv[c0 + k] = dk != 0.0 ? 1.0 / dk : 0.0;
    }
    ctx.sync();
#ifdef CB_PROF_PANEL_SPLIT
    pt.stop(PROF_SPARE1);
#else
    pt.stop(PROF_FACTOR_BIG_PANEL);
#endif
}

#if CB_ON_DEVICE
#define CB_WARP_ID (threadIdx.x >> 5)
#define CB_NUM_WARPS ((blockDim.x + 31) >> 5)
#define CB_CTA_SYNC() __syncthreads()
#else
#define CB_WARP_ID 0
#define CB_NUM_WARPS 1
#define CB_CTA_SYNC()
#endif

// Run the level schedule (forward = leaves first): f(scope, supernode) for warp-/CTA-scope supernodes and
// g(cta, begin, end) for a phase of singleton leaves (range in P.order).
template <class F, class G, class E>
CB_DEV void for_each_supernode(const Ctx &cta, const DevProblem &P, bool forward, F f, G g, E phase_end)
{
    Ctx wctx = cta;
#if CB_ON_DEVICE
    wctx.tid = threadIdx.x & 31;
    wctx.nthr = 32;
    wctx.warp_scope = 1;
#endif
    for (int pi = 0; pi < P.nphases; pi++) {
        const int pidx = forward ? pi : P.nphases - 1 - pi;
#if CB_ON_DEVICE
        Phase ph;
        if (cb_phase_n > 0) {
            const int *pp = cb_phase + 6 * pidx;
            ph.mode = pp[0]; ph.begin = pp[1]; ph.end = pp[2]; ph.ebegin = pp[3]; ph.eend = pp[4]; ph.first_big = pp[5];
        } else {
            ph = P.phases[pidx];
        }
#else
        const Phase ph = P.phases[pidx];
#endif
        if (ph.mode == 2) {
            g(cta, ph.begin, ph.end, ph.ebegin, ph.eend);
        } else if (ph.mode == 1) {
            for (int q = ph.begin; q < ph.end; q++) {
                const int hint = q == ph.begin ? ph.first_big : -2;
#if CB_ON_DEVICE
                const int s = (hint >= 0 && hint < cb_chain_n) ? cb_chain[2 * hint].x : P.order[q];
#else
                const int s = P.order[q];
#endif
                f(cta, s, hint);
            }
        } else {
            for (int q = ph.begin + CB_WARP_ID; q < ph.end; q += CB_NUM_WARPS) f(wctx, P.order[q], -1);
        }
        CB_CTA_SYNC();
        phase_end(forward ? pi : P.nphases - 1 - pi, ph.mode);
    }
}
template <class F, class G> CB_DEV void for_each_supernode(const Ctx &cta, const DevProblem &P, bool forward, F f, G g)
{
    for_each_supernode(cta, P, forward, f, g, [](int, int) {});
}

// numeric factorisation + inertia (positive = #(D>0), negative = #(D<=0), zero = #(D==0); linear_solver.jl:33-44)
CB_DEVN void ldl_factor(const Ctx &ctx, const DevProblem &P, double *pan, double *D, double *Dinv, const KSrc &K,
                        double *Lcsr, int *istat, long long *prof)
{
    ProfTimer pt{prof, 0};
    pt.start();
    // supernodes off the shared-memory path are assembled in the factor storage: clear their panels, gather their entries
    PAR_FOR(e, P.n_gasm_zero) pan[P.gasm_zero[e]] = 0.0;
    ctx.sync();
    PAR_FOR(e, P.n_gasm) pan[P.gasm_dst[e]] = ksrc_load(K, P.gasm_src[e]);
    ctx.sync();
    pt.stop(PROF_ASSEMBLE);
    for_each_supernode(
        ctx, P, true,
        [&](const Ctx &c, int s, int hint) {
            const int bi = (c.warp_scope || !P.big_index) ? -1 : (hint != -2 ? hint : P.big_index[s]);
            if (bi >= 0) {
                factor_supernode_big(c, P, pan, D, Dinv, K, s, P.big[bi], pt);
            } else {
                factor_supernode(c, P, pan, D, Dinv, s);
                if (!c.warp_scope) pt.stop(PROF_FACTOR_BIG_GENERIC);
            }
        },
        [&](const Ctx &ctx, int begin, int end, int ebegin, int eend) {   // singleton leaves: L = a / d, straight from the input values
            pt.stop(PROF_FACTOR_SMALL);
            PAR_FOR(q, end - begin) {       // pivots
                const int s = P.order[begin + q], c0 = P.sn_start[s];
                const double dk = ksrc_load(K, P.leaf_piv_src[begin + q]);
                pan[P.panel_off[s]] = dk;
                D[c0] = dk;
                Dinv[c0] = dk != 0.0 ? 1.0 / dk : 0.0;
            }
            ctx.sync();
            {                               // entries, leaf by leaf => coalesced panel accesses
                const int cnt = eend - ebegin;
#if CB_ON_DEVICE
                const int4 *__restrict__ e4 = reinterpret_cast<const int4 *>(P.leaf_e4) + ebegin;
                int e = ctx.tid;
                const int st = ctx.nthr;
                for (; e + 7 * st < cnt; e += 8 * st) {      // eight entries in flight per thread
                    int4 d[8];
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) d[u] = e4[e + u * st];
#pragma unroll
                    for (int u = 0; u < 8; u++) v[u] = ksrc_load(K, d[u].w) * Dinv[d[u].y];
#pragma unroll
                    for (int u = 0; u < 8; u++) { pan[d[u].x] = v[u]; Lcsr[d[u].z] = v[u]; }
                }
                for (; e < cnt; e += st) {
                    const int4 d = e4[e];
                    const double l = ksrc_load(K, d.w) * Dinv[d.y];
                    pan[d.x] = l;
                    Lcsr[d.z] = l;
                }
#else
                const int *eo = P.leaf_e_off + ebegin, *ec = P.leaf_e_col + ebegin, *ep = P.leaf_e_pos + ebegin,
                          *es = P.leaf_e_src + ebegin;
                for (int e = 0; e < cnt; e++) {
                    const double l = ksrc_load(K, es[e]) * Dinv[ec[e]];
                    pan[eo[e]] = l;
                    Lcsr[ep[e]] = l;     // row-ordered copy for the bulk forward pass
                }
#endif
            }
            ctx.sync();
            pt.stop(PROF_FACTOR_LEAVES);
        });
    pt.stop(PROF_FACTOR_SMALL);
    // one reduction for the three counts (exact in a double: N < 2^17): negatives are counted explicitly so that a NaN
    // pivot is neither positive nor negative and fails the inertia test, as in the reference (linear_solver.jl:33-44)
    long long c;
    if (P.N < 131072) {
        c = (long long)scope_sum(ctx, P.N, [&](int i) {
            const double d = D[i];
            return (d > 0.0 ? 1.0 : 0.0) + (d <= 0.0 ? 131072.0 : 0.0) + (d == 0.0 ? 17179869184.0 : 0.0);
        });
    } else {
        const long long cp = (long long)scope_sum(ctx, P.N, [&](int i) { return D[i] > 0.0 ? 1.0 : 0.0; });
        const long long cn = (long long)scope_sum(ctx, P.N, [&](int i) { return D[i] <= 0.0 ? 1.0 : 0.0; });
        const long long cz = (long long)scope_sum(ctx, P.N, [&](int i) { return D[i] == 0.0 ? 1.0 : 0.0; });
        c = -1;
        if (ctx.tid == 0) { istat[I_INERTIA_POS] = (int)cp; istat[I_INERTIA_NEG] = (int)cn; istat[I_INERTIA_ZERO] = (int)cz; }
    }
    if (ctx.tid == 0) {
        if (c >= 0) {
            istat[I_INERTIA_POS] = (int)(c & 131071);
            istat[I_INERTIA_NEG] = (int)((c >> 17) & 131071);
            istat[I_INERTIA_ZERO] = (int)(c >> 34);
        }
        istat[I_FACTORIZATIONS]++;
    }
    ctx.sync();
    pt.stop(PROF_INERTIA);
}


#if CB_ON_DEVICE
// ---- fast path of ldl_solve (same arithmetic): shared-memory solve with warp-specialised chain runs ---------------
// The permuted vector lives in shared memory for the whole solve.  The shared-memory ("chain") supernodes stream their
// factor panels, in one or two column parts, through two shared-memory slots filled by TMA bulk copies (mbarrier
// complete_tx) up to two parts ahead.  A *chain run* (symbolic.cpp: consecutive CTA-scope phases on the shared-memory
// path with no bulk pull in between; for a trajectory-optimisation KKT matrix: the whole stage chain) is processed
// without CTA-wide barriers:
//   * warp 0 runs the triangular sweeps, eight pivots at a time: the eight current values are broadcast with shuffles,
//     every lane solves the 8x8 unit-triangular block redundantly in registers (no shuffle / FMA ping-pong per pivot) and
//     then updates its own rows (forward) or columns (backward) -- the same FMA sequence per entry as a pivot-by-pivot
//     sweep, so the results do not depend on the blocking;
//   * the other warps apply the rectangular parts (forward: x[R] -= L_R y, backward: t = L_R' x[R] + the pivots solved
//     before), four threads per row / column, and issue the TMA copies;
//   * the two sides meet at named barriers (bar.arrive / bar.sync, producer-consumer): forward, the sweep of the next
//     column part overlaps the rectangular update of the previous one.
// Singleton leaves are handled in bulk by four-lane groups (coalesced); supernodes off the shared-memory path keep the
// generic code.  Work area: x[N] | slot 0 | slot 1 | y[64] | zero cell | x[R] window.
enum { CB_BAR_A0 = 1, CB_BAR_A1 = 2, CB_BAR_B0 = 3, CB_BAR_B1 = 4, CB_BAR_REST = 5 };   // named barriers (0 = __syncthreads)
// (immediate barrier numbers: with a register operand ptxas reserves all 16 hardware barriers for the CTA)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    switch (id) {
    case 1: asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); break;
    case 2: asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); break;
    case 3: asm volatile("bar.sync 3, %0;" ::"r"(nthreads) : "memory"); break;
    case 4: asm volatile("bar.sync 4, %0;" ::"r"(nthreads) : "memory"); break;
    default: asm volatile("bar.sync 5, %0;" ::"r"(nthreads) : "memory"); break;
    }
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads)
{
    switch (id) {
    case 1: asm volatile("bar.arrive 1, %0;" ::"r"(nthreads) : "memory"); break;
    case 2: asm volatile("bar.arrive 2, %0;" ::"r"(nthreads) : "memory"); break;
    case 3: asm volatile("bar.arrive 3, %0;" ::"r"(nthreads) : "memory"); break;
    default: asm volatile("bar.arrive 4, %0;" ::"r"(nthreads) : "memory"); break;
    }
}

struct ChainTask { int s, c0, w, nR, rows_off, panel_off, h1; };
__device__ __forceinline__ ChainTask chain_task(const DevProblem &P, int bi)
{
    int4 d0, d1;
    if (bi < cb_chain_n) { d0 = cb_chain[2 * bi]; d1 = cb_chain[2 * bi + 1]; }       // cached: no global-memory chase
    else { const int4 *g = reinterpret_cast<const int4 *>(P.bdesc) + 2 * bi; d0 = g[0]; d1 = g[1]; }
    return ChainTask{d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z};
}
__device__ __forceinline__ Phase load_phase(const DevProblem &P, int pidx)
{
    Phase ph;
    if (cb_phase_n > 0) {
        const int *pp = cb_phase + 6 * pidx;
        ph.mode = pp[0]; ph.begin = pp[1]; ph.end = pp[2]; ph.ebegin = pp[3]; ph.eend = pp[4]; ph.first_big = pp[5];
    } else {
        ph = P.phases[pidx];
    }
    return ph;
}

// The TMA part stream of one solve direction.  Every thread tracks the (uniform) counters; part g lands in slot
// (base + g) & 1 and completes phase ((base + g) >> 1) & 1 of that slot's mbarrier.
struct PartStream {
    const double *pan;
    const int *parts;       // (panel offset, doubles) pairs in issue order
    int nparts, issued, done;
    unsigned base;          // mbarrier uses before this direction started
    int buf_off, buf_len;
};
__device__ __forceinline__ void part_issue(const PartStream &st, int g, int off, int doubles)      // ONE thread
{
    const unsigned u = st.base + (unsigned)g;
    const unsigned bytes = (unsigned)doubles * 8u;
    fence_proxy_async();
    mbar_expect_tx(&cb_bars[u & 1], bytes);
    tma_bulk_g2s(cb_dyn_smem + st.buf_off + (int)(u & 1) * st.buf_len, st.pan + off, bytes, &cb_bars[u & 1]);
}
__device__ __forceinline__ const double *part_wait(const PartStream &st, int g)      // returns the slot's base
{
    const unsigned u = st.base + (unsigned)g;
    mbar_wait(&cb_bars[u & 1], (u >> 1) & 1);
    return cb_dyn_smem + st.buf_off + (int)(u & 1) * st.buf_len;
}
// (all threads, after a CTA barrier) make sure the next two parts are in flight
__device__ __forceinline__ void part_prefetch(PartStream &st, int leader)
{
    const int want = min(st.done + 2, st.nparts);
    if ((int)threadIdx.x == leader)
        for (int g = st.issued; g < want; g++) part_issue(st, g, st.parts[2 * g], st.parts[2 * g + 1]);
    if (st.issued < want) st.issued = want;
}

// Forward sweep of columns [k0, k1) of a unit-lower-triangular pivot block (zeros stored on and above the diagonal):
// lane holds the unknowns of rows lane (y0) and lane + 32 (y1); L + k * nrow is column k of the panel.  Four pivots per
// step; every shared-memory load of a step is issued before the shuffles (the addresses do not depend on the data), so
// the dependent chain of a step is one shuffle and four FMAs.  (Measured on B200, tools/microbench/sweep_blocked.cu: 28
// cycles per pivot against 55 pivot by pivot; inside the solve kernel the single sweep warp is bound by its own
// instruction issue, ~70 cycles per pivot either way.)
__device__ __forceinline__ void sweep_forward_blocked(const double *__restrict__ L, int nrow, int k0, int k1, int lane,
                                                      double &y0, double &y1)
{
    const int r0 = min(lane, nrow - 1), r1 = min(lane + 32, nrow - 1);
    int kb = k0;
    for (; kb + 4 <= k1; kb += 4) {
        const double *c0 = L + kb * nrow, *c1 = c0 + nrow, *c2 = c1 + nrow, *c3 = c2 + nrow;
        const double l10 = c0[kb + 1], l20 = c0[kb + 2], l30 = c0[kb + 3], l21 = c1[kb + 2], l31 = c1[kb + 3], l32 = c2[kb + 3];
        const double a00 = c0[r0], a01 = c1[r0], a02 = c2[r0], a03 = c3[r0];
        const double a10 = c0[r1], a11 = c1[r1], a12 = c2[r1], a13 = c3[r1];
        const double v0 = __shfl_sync(0xffffffffu, kb < 32 ? y0 : y1, kb & 31);
        double v1 = __shfl_sync(0xffffffffu, kb + 1 < 32 ? y0 : y1, (kb + 1) & 31);
        double v2 = __shfl_sync(0xffffffffu, kb + 2 < 32 ? y0 : y1, (kb + 2) & 31);
        double v3 = __shfl_sync(0xffffffffu, kb + 3 < 32 ? y0 : y1, (kb + 3) & 31);
        v1 -= l10 * v0;
        v2 -= l20 * v0;
        v3 -= l30 * v0;
        y0 -= a00 * v0;
        y1 -= a10 * v0;
        v2 -= l21 * v1;
        v3 -= l31 * v1;
        y0 -= a01 * v1;
        y1 -= a11 * v1;
        v3 -= l32 * v2;
        y0 -= a02 * v2;
        y1 -= a12 * v2;
        y0 -= a03 * v3;
        y1 -= a13 * v3;
    }
    for (; kb < k1; kb++) {                           // remainder: pivot by pivot
        const double l0 = L[kb * nrow + r0], l1 = L[kb * nrow + r1];
        const double yk = __shfl_sync(0xffffffffu, kb < 32 ? y0 : y1, kb & 31);
        y0 -= l0 * yk;
        y1 -= l1 * yk;
    }
}
// Backward sweep (L' z = v) restricted to columns [k0, k1): lane holds columns lane (z0) and lane + 32 (z1); lanes whose
// column is outside the part read a zero cell with stride 0.  Four pivots per step, from the last column down.
__device__ __forceinline__ void sweep_backward_blocked(const double *__restrict__ L, int nrow, int k0, int k1, int lane,
                                                       const double *zero_cell, double &z0, double &z1)
{
    const bool in0 = lane >= k0 && lane < k1, in1 = lane + 32 >= k0 && lane + 32 < k1;
    const double *p0 = in0 ? L + lane * nrow : zero_cell, *p1 = in1 ? L + (lane + 32) * nrow : zero_cell;
    const int s0 = in0 ? 1 : 0, s1 = in1 ? 1 : 0;
    int kt = k1;
    for (; kt - 4 >= k0; kt -= 4) {
        const int kb = kt - 4;
        // rows kb+1 .. kb+3 of the block (row r of column c at L[c * nrow + r])
        const double *c0 = L + kb * nrow, *c1 = c0 + nrow, *c2 = c1 + nrow;
        const double l10 = c0[kb + 1], l20 = c0[kb + 2], l30 = c0[kb + 3], l21 = c1[kb + 2], l31 = c1[kb + 3], l32 = c2[kb + 3];
        const double a00 = p0[kb * s0], a01 = p0[(kb + 1) * s0], a02 = p0[(kb + 2) * s0], a03 = p0[(kb + 3) * s0];
        const double a10 = p1[kb * s1], a11 = p1[(kb + 1) * s1], a12 = p1[(kb + 2) * s1], a13 = p1[(kb + 3) * s1];
        double v0 = __shfl_sync(0xffffffffu, kb < 32 ? z0 : z1, kb & 31);
        double v1 = __shfl_sync(0xffffffffu, kb + 1 < 32 ? z0 : z1, (kb + 1) & 31);
        double v2 = __shfl_sync(0xffffffffu, kb + 2 < 32 ? z0 : z1, (kb + 2) & 31);
        const double v3 = __shfl_sync(0xffffffffu, kb + 3 < 32 ? z0 : z1, (kb + 3) & 31);
        v2 -= l32 * v3;
        v1 -= l31 * v3;
        v0 -= l30 * v3;
        z0 -= a03 * v3;
        z1 -= a13 * v3;
        v1 -= l21 * v2;
        v0 -= l20 * v2;
        z0 -= a02 * v2;
        z1 -= a12 * v2;
        v0 -= l10 * v1;
        z0 -= a01 * v1;
        z1 -= a11 * v1;
        z0 -= a00 * v0;
        z1 -= a10 * v0;
    }
    for (int k = kt - 1; k > k0; k--) {               // remainder: pivot by pivot
        const double l0 = p0[k * s0], l1 = p1[k * s1];
        const double zk = __shfl_sync(0xffffffffu, k < 32 ? z0 : z1, k & 31);
        z0 -= l0 * zk;
        z1 -= l1 * zk;
    }
}

// Forward substitution through the chain supernodes bi0 .. bi0 + ntasks - 1 (see the header comment).  Ends with a CTA barrier.
__device__ __forceinline__ void chain_run_forward(const DevProblem &P, PartStream &st, int bi0, int ntasks, double *xs, double *yv,
                                                  long long *prof)
{
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int leader = 32;
    part_prefetch(st, leader);
    int g = st.done;
    if (wid == 0) {
        ProfTimer pc{prof, 0};
        pc.start();
        int prev_parts = 0;
        for (int t = 0; t < ntasks; t++) {
            const ChainTask d = chain_task(P, bi0 + t);
            const int nrow = d.w + d.nR, np = d.h1 < d.w ? 2 : 1;
            for (int q = prev_parts; q > 0; q--) named_bar_sync(CB_BAR_B0 + (int)((st.base + g - q) & 1), nthr);   // pushes into these pivots
            pc.stop(PROF_CF_WAIT);
            double y0 = lane < d.w ? xs[d.c0 + lane] : 0.0, y1 = lane + 32 < d.w ? xs[d.c0 + lane + 32] : 0.0;
            for (int a = 0; a < np; a++, g++) {
                const int k0 = a == 0 ? 0 : d.h1, k1 = (a == 0 && np == 2) ? d.h1 : d.w;
                const double *L = part_wait(st, g) - k0 * nrow;
                pc.stop(PROF_CF_TMA);
                sweep_forward_blocked(L, nrow, k0, k1, lane, y0, y1);
                if (lane >= k0 && lane < k1) { yv[lane] = y0; xs[d.c0 + lane] = y0; }
                if (lane + 32 >= k0 && lane + 32 < k1) { yv[lane + 32] = y1; xs[d.c0 + lane + 32] = y1; }
                named_bar_arrive(CB_BAR_A0 + (int)((st.base + g) & 1), nthr);      // y[k0 .. k1) published
                pc.stop(PROF_CF_SWEEP);
            }
            prev_parts = np;
        }
        for (int q = prev_parts; q > 0; q--) named_bar_sync(CB_BAR_B0 + (int)((st.base + g - q) & 1), nthr);
        pc.stop(PROF_CF_WAIT);
    } else {
        const int part = tid & 3, grp = (tid - 32) >> 2, ngrp = (nthr - 32) >> 2;
        for (int t = 0; t < ntasks; t++) {
            const ChainTask d = chain_task(P, bi0 + t);
            const int nrow = d.w + d.nR, np = d.h1 < d.w ? 2 : 1, nR = d.nR, w = d.w;
            const int *__restrict__ R = P.rows + d.rows_off;
            const int rI = grp < nR ? R[grp] : 0;       // row index of this thread's first push row, in flight early
            for (int a = 0; a < np; a++, g++) {
                const int k0 = a == 0 ? 0 : d.h1, k1 = (a == 0 && np == 2) ? d.h1 : w;
                int noff = 0, nlen = 0;
                const bool refill = tid == leader && g + 2 < st.nparts;
                if (refill) { noff = st.parts[2 * (g + 2)]; nlen = st.parts[2 * (g + 2) + 1]; }
                const double *L = part_wait(st, g) - k0 * nrow;
                named_bar_sync(CB_BAR_A0 + (int)((st.base + g) & 1), nthr);
                for (int base = 0; base < nR; base += ngrp) {      // x[R] -= L_R[:, k0:k1) y[k0:k1), four threads per row
                    const int i = base + grp;
                    double a0 = 0.0, a1 = 0.0;
                    if (i < nR) {
                        const double *Lr = L + w + i + (k0 + part) * nrow;
                        const double *yp = yv + k0 + part;
                        const int cnt = (k1 - k0 - part + 3) >> 2;
                        int j = 0;
                        for (; j + 1 < cnt; j += 2) {
                            a0 += Lr[0] * yp[0];
                            a1 += Lr[4 * nrow] * yp[4];
                            Lr += 8 * nrow;
                            yp += 8;
                        }
                        if (j < cnt) a0 += Lr[0] * yp[0];
                    }
                    double acc = a0 + a1;
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    if (part == 0 && i < nR) xs[base == 0 ? rI : R[i]] -= acc;
                }
                named_bar_sync(CB_BAR_REST, nthr - 32);            // every reader of the slot is done
                named_bar_arrive(CB_BAR_B0 + (int)((st.base + g) & 1), nthr);      // pushes of this part are in xs
                if (refill) part_issue(st, g + 2, noff, nlen);      // (after the arrival: off the sweep warp's critical path)
            }
        }
    }
    st.done = g;
    st.issued = max(st.issued, min(g + 2, st.nparts));
    __syncthreads();
}

// Backward substitution through the chain supernodes bi0 + ntasks - 1 .. bi0 (in that order).  Ends with a CTA barrier.
__device__ __forceinline__ void chain_run_backward(const DevProblem &P, PartStream &st, int bi0, int ntasks, double *xs, double *yv,
                                                   double *xr, const double *zero_cell, long long *prof)
{
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int leader = 32;
    part_prefetch(st, leader);
    const int g_first = st.done;
    int g = g_first;
    if (wid == 0) {
        ProfTimer pc{prof, 0};
        pc.start();
        for (int t = ntasks - 1; t >= 0; t--) {
            const ChainTask d = chain_task(P, bi0 + t);
            const int nrow = d.w + d.nR, np = d.h1 < d.w ? 2 : 1;
            double z0 = lane < d.w ? xs[d.c0 + lane] : 0.0, z1 = lane + 32 < d.w ? xs[d.c0 + lane + 32] : 0.0;
            for (int a = np - 1; a >= 0; a--, g++) {
                const int k0 = a == 0 ? 0 : d.h1, k1 = (a == 0 && np == 2) ? d.h1 : d.w;
                const double *L = part_wait(st, g) - k0 * nrow;
                pc.stop(PROF_CB_TMA);
                named_bar_sync(CB_BAR_A0 + (int)((st.base + g) & 1), nthr);        // t[k0 .. k1) gathered
                pc.stop(PROF_CB_WAIT);
                if (lane >= k0 && lane < k1) z0 -= yv[lane];
                if (lane + 32 >= k0 && lane + 32 < k1) z1 -= yv[lane + 32];
                sweep_backward_blocked(L, nrow, k0, k1, lane, zero_cell, z0, z1);
                if (lane >= k0 && lane < k1) xs[d.c0 + lane] = z0;
                if (lane + 32 >= k0 && lane + 32 < k1) xs[d.c0 + lane + 32] = z1;
                named_bar_arrive(CB_BAR_B0 + (int)((st.base + g) & 1), nthr);      // x[c0 + k0 .. c0 + k1) final
                pc.stop(PROF_CB_SWEEP);
            }
        }
    } else {
        const int part = tid & 3, grp = (tid - 32) >> 2, ngrp = (nthr - 32) >> 2, nrest = nthr - 32;
        for (int t = ntasks - 1; t >= 0; t--) {
            const ChainTask d = chain_task(P, bi0 + t);
            const int nrow = d.w + d.nR, np = d.h1 < d.w ? 2 : 1, nR = d.nR, w = d.w;
            const int *__restrict__ R = P.rows + d.rows_off;
            for (int a = np - 1; a >= 0; a--, g++) {
                const int k0 = a == 0 ? 0 : d.h1, k1 = (a == 0 && np == 2) ? d.h1 : w;
                if (g > g_first) {
                    // the previous part's sweep is done: its unknowns are final and its slot is free for part g + 1
                    int noff = 0, nlen = 0;
                    const bool refill = tid == leader && g + 1 < st.nparts;
                    if (refill) { noff = st.parts[2 * (g + 1)]; nlen = st.parts[2 * (g + 1) + 1]; }
                    named_bar_sync(CB_BAR_B0 + (int)((st.base + g - 1) & 1), nthr);
                    if (refill) part_issue(st, g + 1, noff, nlen);
                }
                if (a == np - 1) {          // x[R] (final), staged once for both parts
                    for (int i = tid - 32; i < nR; i += nrest) xr[i] = xs[R[i]];
                    named_bar_sync(CB_BAR_REST, nrest);
                }
                const double *L = part_wait(st, g) - k0 * nrow;
                // t_k = sum_{r >= k1} L[r, k] x_r for the columns k of this part: rows k1 .. w-1 are the pivots of the part
                // solved before (final in xs), rows w .. are R; four threads per column
                for (int base = k0; base < k1; base += ngrp) {
                    const int k = base + grp;
                    double a0 = 0.0, a1 = 0.0;
                    if (k < k1) {
                        const double *Lc = L + k * nrow;
                        const double *xp = xs + d.c0;
                        int r = k1 + part;
                        for (; r + 4 < w; r += 8) { a0 += Lc[r] * xp[r]; a1 += Lc[r + 4] * xp[r + 4]; }
                        if (r < w) { a0 += Lc[r] * xp[r]; r += 4; }
                        const double *xq = xr - w;
                        for (; r + 4 < nrow; r += 8) { a0 += Lc[r] * xq[r]; a1 += Lc[r + 4] * xq[r + 4]; }
                        if (r < nrow) a0 += Lc[r] * xq[r];
                    }
                    double acc = a0 + a1;
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    if (part == 0 && k < k1) yv[k] = acc;
                }
                named_bar_arrive(CB_BAR_A0 + (int)((st.base + g) & 1), nthr);
            }
        }
        if (g > g_first) named_bar_sync(CB_BAR_B0 + (int)((st.base + g - 1) & 1), nthr);      // pairs the last sweep's arrival
    }
    st.done = g;
    st.issued = max(st.issued, min(g + 1, st.nparts));
    __syncthreads();
}

__device__ __noinline__ void ldl_solve_smem(const Ctx &ctx, const DevProblem &P, const double *__restrict__ pan,
                                            const double *D, const double *__restrict__ Dinv,
                                            const double *__restrict__ Lcsr, const double *b, double *x, double *xg,
                                            int *istat, long long *prof)
{
    ProfTimer pt{prof, 0}, ps{prof, 0};
    pt.start();
    ps.start();
    // With the leaves-first ordering (P.nleaf > 0) only x[nleaf .. N) lives in shared memory: a singleton leaf is final from
    // the start in the forward sweep (y_leaf = b_leaf, read from the global scratch xg) and is written straight to the
    // result in the backward sweep.  xs is biased so that xs[k] addresses column k for k >= nleaf.
    const int N = P.N, nl = P.nleaf, Npad = (N - nl + 1) & ~1;
    const int tid = threadIdx.x, nthr = blockDim.x, wid = tid >> 5, nw = nthr >> 5;
    double *xs = cb_dyn_smem - nl;
    const int buf_off = Npad, buf_len = P.max_sb_doubles;      // slot s at cb_dyn_smem + buf_off + s * buf_len
    double *yv = cb_dyn_smem + buf_off + 2 * buf_len;            // (named from cb_dyn_smem so that accesses stay LDS/STS)
    double *zero_cell = yv + 64;
    double *xr = yv + 80;                                         // x[R] of the current chain supernode (backward)
    const int leader = 32;
    if (tid == 0) *zero_cell = 0.0;
    for (int k = tid; k < N; k += nthr) {
        const double v = b[P.perm[k]];
        if (k >= nl) xs[k] = v; else xg[k] = v;
    }
    const double *__restrict__ xleaf = nl > 0 ? xg : xs;         // where the forward pass finds y_leaf
    // the mbarriers live for the whole kernel: continue the use numbering where the previous solve stopped
    PartStream st{pan, P.parts_fwd, P.nparts_fwd, 0, 0, cb_bar_uses, buf_off, buf_len};
    __syncthreads();
    ps.stop(PROF_SF_PULL);                                        // (counter: permutation gather)
    part_prefetch(st, leader);
    // bulk pass: every column pulls the contributions of its singleton-leaf descendants (x_leaf = b_leaf is final);
    // four lanes per column, columns without leaf descendants are not visited
    {
        const int4 *__restrict__ info = reinterpret_cast<const int4 *>(P.lcsr_rowinfo);
        const int *__restrict__ lcol = P.lcsr_col;
        const int sub = tid & 3, grp = tid >> 2, ngrp = nthr >> 2, nrows = P.lcsr_ncols;
        const int padded = (nrows + ngrp - 1) / ngrp * ngrp;          // every warp runs every shuffle
        int4 ri = grp < nrows ? info[grp] : make_int4(0, 0, 0, 0);
        for (int r = grp; r < padded; r += ngrp) {
            const int rn = r + ngrp;
            const int4 rin = rn < nrows ? info[rn] : make_int4(0, 0, 0, 0);   // next row's descriptor in flight
            double acc = 0.0;
            for (int q0 = ri.y + sub; q0 < ri.z; q0 += 32) {      // eight entries per lane per round, loads batched
                double l[8];
                int cidx[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int q = q0 + 4 * u;
                    const bool ok = q < ri.z;
                    l[u] = ok ? Lcsr[q] : 0.0;
                    cidx[u] = ok ? lcol[q] : -1;
                }
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (cidx[u] >= 0) acc += l[u] * xleaf[cidx[u]];
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            if (sub == 0 && r < nrows) xs[ri.x] -= acc;
            ri = rin;
        }
    }
    __syncthreads();
    ps.stop(PROF_SF_BULK);
    Ctx wctx = ctx;
    wctx.tid = tid & 31;
    wctx.nthr = 32;
    wctx.warp_scope = 1;
    // generic forward step of a supernode that is not on the shared-memory path (scope: warp or CTA)
    auto forward_generic = [&](const Ctx &c, int s) {
        const Ctx &ctx = c;
        const int c0 = P.sn_start[s], w = P.sn_start[s + 1] - c0;
        const int nrow = w + (P.rows_ptr[s + 1] - P.rows_ptr[s]);
        PAR_FOR(j, w) {   // pull from small (non-leaf, non-shared-memory) descendants
            double acc = 0.0;
            for (int q = P.fwd_ptr[c0 + j]; q < P.fwd_ptr[c0 + j + 1]; q++) {
                const FwdEntry fe = P.fwd[q];
                const double *Ld = pan + fe.off;
                for (int k = 0; k < fe.width; k++) acc += Ld[(long long)k * fe.stride] * xs[fe.col0 + k];
            }
            if (acc != 0.0) xs[c0 + j] -= acc;
        }
        const double *Ps = pan + P.panel_off[s];
        for (int k = 0; k + 1 < w; k++) {
            ctx.sync();
            const double xk = xs[c0 + k];
            PAR_FOR(i, w - 1 - k) xs[c0 + k + 1 + i] -= Ps[(k + 1 + i) + (long long)k * nrow] * xk;
        }
        ctx.sync();
    };
    const int nph = P.nphases;
    for (int pi = 0; pi < nph; pi++) {
        const Phase ph = load_phase(P, pi);
        int last = pi;
        if (ph.mode == 1) {
            if (ph.ebegin == pi + 1) {                            // a chain run starts here
                last = ph.eend - 1;
                const Phase pl = load_phase(P, last);
                chain_run_forward(P, st, ph.first_big, pl.first_big + (pl.end - pl.begin) - ph.first_big, xs, yv, prof);
            } else {
                for (int q = ph.begin; q < ph.end; q++) {
                    const int s = P.order[q], bi = P.big_index[s];
                    if (bi >= 0) chain_run_forward(P, st, bi, 1, xs, yv, prof);
                    else { forward_generic(ctx, s); __syncthreads(); }
                }
            }
            ps.stop(PROF_SF_SWEEP);                               // (counter: chain supernodes)
        } else if (ph.mode == 0) {
            for (int q = ph.begin + wid; q < ph.end; q += nw) forward_generic(wctx, P.order[q]);
            __syncthreads();
            ps.stop(PROF_SF_OTHER);                               // (counter: small supernodes)
        }
        pi = last;
        // bulk pull: the finished phase's small supernodes -> pivot columns of chain supernodes
        const int r0 = cb_phase_n > 0 ? cb_pphase[pi] : P.pphase_ptr[pi], r1 = cb_phase_n > 0 ? cb_pphase[pi + 1] : P.pphase_ptr[pi + 1];
        if (ph.mode != 2 && r1 > r0) {
            const int4 *__restrict__ rowi = reinterpret_cast<const int4 *>(P.prow);
            for (int r = r0 + tid; r < r1; r += nthr) {
                const int4 ri = rowi[r];
                double acc = 0.0;
                for (int q = ri.y; q < ri.z; q++) {
                    const FwdEntry fe = P.pfwd[q];
                    const double *Ld = pan + fe.off;
                    for (int k = 0; k < fe.width; k++) acc += Ld[(long long)k * fe.stride] * xs[fe.col0 + k];
                }
                xs[ri.x] -= acc;
            }
            __syncthreads();
        }
    }
    for (int k = nl + tid; k < N; k += nthr) xs[k] *= Dinv[k];
    __syncthreads();
    pt.stop(PROF_SOLVE_FWD);
    ps.stop(PROF_SF_PUSH);                                        // (counter: D^-1 scaling and bulk pulls after the last phase)
    st = PartStream{pan, P.parts_bwd, P.nparts_bwd, 0, 0, st.base + (unsigned)st.nparts, buf_off, buf_len};
    part_prefetch(st, leader);
    auto backward_generic = [&](const Ctx &c, int s) {
        const Ctx &ctx = c;
        const int c0 = P.sn_start[s], w = P.sn_start[s + 1] - c0;
        const int rows_off = P.rows_ptr[s], nR = P.rows_ptr[s + 1] - rows_off, nrow = w + nR;
        const int *__restrict__ R = P.rows + rows_off;
        const double *Ps = pan + P.panel_off[s];
        PAR_FOR(k, w) {
            double acc = 0.0;
            const double *col = Ps + (long long)k * nrow + w;
            for (int i = 0; i < nR; i++) acc += col[i] * xs[R[i]];
            xs[c0 + k] -= acc;
        }
        for (int k = w - 1; k > 0; k--) {
            ctx.sync();
            const double xk = xs[c0 + k];
            PAR_FOR(i, k) xs[c0 + i] -= Ps[k + (long long)i * nrow] * xk;
        }
        ctx.sync();
    };
    for (int pi = nph - 1; pi >= 0; pi--) {
        const Phase ph = load_phase(P, pi);
        if (ph.mode == 1) {
            if (ph.eend == pi + 1) {                              // the last phase of a chain run: the run is walked backwards
                const int first = ph.ebegin - 1;
                const Phase pf = load_phase(P, first);
                chain_run_backward(P, st, pf.first_big, ph.first_big + (ph.end - ph.begin) - pf.first_big, xs, yv, xr, zero_cell, prof);
                pi = first;
            } else {
                for (int q = ph.end - 1; q >= ph.begin; q--) {
                    const int s = P.order[q], bi = P.big_index[s];
                    if (bi >= 0) chain_run_backward(P, st, bi, 1, xs, yv, xr, zero_cell, prof);
                    else { backward_generic(ctx, s); __syncthreads(); }
                }
            }
            ps.stop(PROF_SB_SWEEP);                               // (counter: chain supernodes)
        } else if (ph.mode == 0) {
            for (int q = ph.begin + wid; q < ph.end; q += nw) backward_generic(wctx, P.order[q]);
            __syncthreads();
            ps.stop(PROF_SB_GATHER);                              // (counter: small supernodes)
        } else {      // singleton leaves: four lanes per leaf
            const int begin = ph.begin, end = ph.end;
            const int4 *__restrict__ info = reinterpret_cast<const int4 *>(P.leaf_info) + begin;
            const int *__restrict__ rows = P.rows;
            const int sub = tid & 3, grp = tid >> 2, ngrp = nthr >> 2, cnt = end - begin;
            const int padded = (cnt + ngrp - 1) / ngrp * ngrp;
            int4 li = grp < cnt ? info[grp] : make_int4(0, 0, 0, 0);    // pivot column, |R|, panel offset, rows offset
            for (int q = grp; q < padded; q += ngrp) {
                const int qn = q + ngrp;
                const int4 lin = qn < cnt ? info[qn] : make_int4(0, 0, 0, 0);
                const double *__restrict__ col = pan + li.z;
                const int *__restrict__ R = rows + li.w;
                double acc = 0.0;
                for (int i0 = sub; i0 < li.y; i0 += 32) {
                    double l[8];
                    int ridx[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int i = i0 + 4 * u;
                        const bool ok = i < li.y;
                        l[u] = ok ? col[i] : 0.0;
                        ridx[u] = ok ? R[i] : -1;
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (ridx[u] >= 0) acc += l[u] * xs[ridx[u]];      // (no stray reads of entries other groups write)
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                if (sub == 0 && q < cnt) {
                    if (nl > 0) x[P.perm[li.x]] = xg[li.x] * Dinv[li.x] - acc;      // straight to the result
                    else xs[li.x] -= acc;
                }
                li = lin;
            }
            __syncthreads();
            ps.stop(PROF_SB_LEAVES);
        }
    }
    for (int k = nl + tid; k < N; k += nthr) x[P.perm[k]] = xs[k];
    if (tid == 0 && istat) istat[I_SOLVES]++;
    __syncthreads();
    if (tid == 0) cb_bar_uses = st.base + (unsigned)st.nparts;
    __syncthreads();
    pt.stop(PROF_SOLVE_BWD);
    ps.stop(PROF_SB_OTHER);                                       // (counter: permutation scatter)
}
#endif

// x = P' L^-T D^-1 L^-1 P b; b and x are in natural order (may alias), xp is an N-vector of scratch.
// Supernodes on the shared-memory path use M = L_tt^-T D_t^-1 (stored by the factorisation) so that their triangular
// solves are mat-vecs: forward y_t = D_t M' v, backward x_t = M (D_t v).
CB_DEVN void ldl_solve(const Ctx &ctx, const DevProblem &P, const double *pan, const double *D, const double *Dinv,
                       const double *kx, const double *Lcsr, const double *b, double *x, double *xp, int *istat,
                       long long *prof)
{
#if CB_ON_DEVICE
    if (P.solve_smem && ctx.scratch && blockDim.x >= 128) {
        ldl_solve_smem(ctx, P, pan, D, Dinv, Lcsr, b, x, xp, istat, prof);
        return;
    }
#endif
    ProfTimer pt{prof, 0};
    pt.start();
    PAR_FOR(k, P.N) xp[k] = b[P.perm[k]];
    ctx.sync();
    // bulk pass: every column pulls the contributions of its singleton-leaf descendants (x_leaf = b_leaf is final)
    PAR_FOR(cidx, P.N) {
        const int q0 = P.lcsr_ptr[cidx], q1 = P.lcsr_ptr[cidx + 1];
        if (q1 > q0) {
            double acc = 0.0;
            for (int q = q0; q < q1; q++) acc += Lcsr[q] * xp[P.lcsr_col[q]];
            xp[cidx] -= acc;
        }
    }
    ctx.sync();
    // forward: pull from descendants through the per-column row lists, then the unit-lower diagonal block
    for_each_supernode(
        ctx, P, true,
        [&](const Ctx &c, int s, int hint) {
            const Ctx &ctx = c;
            const int c0 = P.sn_start[s], w = P.sn_start[s + 1] - c0;
            const int nrow = w + (P.rows_ptr[s + 1] - P.rows_ptr[s]);
            const double *Ps = pan + P.panel_off[s];
            const int bi = P.big_index ? P.big_index[s] : -1;
            PAR_FOR(j, w) {
                double acc = 0.0;
                for (int q = P.fwd_ptr[c0 + j]; q < P.fwd_ptr[c0 + j + 1]; q++) {
                    const FwdEntry fe = P.fwd[q];
                    const double *Ld = pan + fe.off;
                    for (int k = 0; k < fe.width; k++) acc += Ld[(long long)k * fe.stride] * xp[fe.col0 + k];
                }
                xp[c0 + j] -= acc;
            }
            for (int k = 0; k + 1 < w; k++) {
                ctx.sync();
                const double xk = xp[c0 + k];
                PAR_FOR(i, w - 1 - k) xp[c0 + k + 1 + i] -= Ps[(k + 1 + i) + (long long)k * nrow] * xk;
            }
            ctx.sync();
            if (bi >= 0) {      // shared-memory supernodes are not in the pull lists of their ancestors: PUSH x[R] -= L_R y
                const int nR = nrow - w;
                const int *R = P.rows + P.rows_ptr[s];
                PAR_FOR(i, nR) {
                    double acc = 0.0;
                    for (int k = 0; k < w; k++) acc += Ps[(w + i) + (long long)k * nrow] * xp[c0 + k];
                    xp[R[i]] -= acc;
                }
                ctx.sync();
            }
        },
        [&](const Ctx &, int, int, int, int) {});   // singleton leaves have nothing to pull and no diagonal block
    PAR_FOR(k, P.N) xp[k] *= Dinv[k];
    ctx.sync();
    pt.stop(PROF_SOLVE_FWD);
    // backward: gather from the ancestors' (already final) entries, then the unit-upper diagonal block
    for_each_supernode(
        ctx, P, false,
        [&](const Ctx &c, int s, int hint) {
            const Ctx &ctx = c;
            const int c0 = P.sn_start[s], w = P.sn_start[s + 1] - c0;
            const int nR = P.rows_ptr[s + 1] - P.rows_ptr[s], nrow = w + nR;
            const double *Ps = pan + P.panel_off[s];
            const int *R = P.rows + P.rows_ptr[s];
            PAR_FOR(k, w) {
                double acc = 0.0;
                const double *col = Ps + (long long)k * nrow + w;
                for (int i = 0; i < nR; i++) acc += col[i] * xp[R[i]];
                xp[c0 + k] -= acc;
            }
            for (int k = w - 1; k > 0; k--) {
                ctx.sync();
                const double xk = xp[c0 + k];
                PAR_FOR(i, k) xp[c0 + i] -= Ps[k + (long long)i * nrow] * xk;
            }
            ctx.sync();
        },
        [&](const Ctx &ctx, int begin, int end, int, int) {
            PAR_FOR(q, end - begin) {
                const int s = P.order[begin + q], c0 = P.sn_start[s];
                const int nR = P.rows_ptr[s + 1] - P.rows_ptr[s];
                const double *col = pan + P.panel_off[s] + 1;
                const int *R = P.rows + P.rows_ptr[s];
                double acc = 0.0;
                for (int i = 0; i < nR; i++) acc += col[i] * xp[R[i]];
                xp[c0] -= acc;
            }
            ctx.sync();
        });
    PAR_FOR(k, P.N) x[P.perm[k]] = xp[k];
    if (ctx.tid == 0 && istat) istat[I_SOLVES]++;
    ctx.sync();
    pt.stop(PROF_SOLVE_BWD);
}

// ------------------------------------------------------------------------------------------------ reduced rhs / recovery
CB_DEVN void reduced_rhs(const Ctx &ctx, const DevProblem &P, const Inst &I, const double *res, double *rs)
{
    const int n = P.n, m = P.m, p = P.p;
    const double ep = I.scal[S_EPSP], ed = I.scal[S_EPSD], rho = I.scal[S_RHO];
    const double Jrr = rho + ep, Jss = ep;
    const double *s = I.w + n + m, *t = I.w + n + 2 * m + 2 * p;
    const double *rr = res + n, *rsl = res + n + m, *ry = res + n + m + p, *rz = res + n + 2 * m + p,
                 *rt = res + n + 2 * m + 2 * p;
    PAR_FOR(i, n) rs[i] = res[i];
    PAR_FOR(i, m) rs[n + i] = ry[i] + rr[i] / Jrr;
    PAR_FOR(i, P.q_nn) {
        double Sb = s[i] - ed, Ti = t[i];
        rs[n + m + i] = rz[i] + (rt[i] + Sb * rsl[i]) / (Ti + Sb * Jss);
    }
    PAR_FOR(k, P.nsoc) {
        int d = P.soc_dims[k], c0 = P.soc_off[k];
        if (d > 0) {
            const double *sk = s + c0, *tk = t + c0, *a = rsl + c0, *b = rt + c0;
            auto u = [&](int i) { return i == 0 ? tk[0] + (sk[0] - ed) * Jss : tk[i] + sk[i] * Jss; };
            double dot = 0.0;   // first row of Sbar times rs
            for (int i = 0; i < d; i++) dot += (i == 0 ? sk[0] - ed : sk[i]) * a[i];
            auto x = [&](int i) { return (i == 0 ? dot : (sk[0] - ed) * a[i] + a[0] * sk[i]) + b[i]; };
            soc_arrow_inverse(d, u, x, [&](int i, double v) { rs[n + m + c0 + i] = rz[c0 + i] + v; });
        }
    }
    ctx.sync();
}

CB_DEVN void recover_step(const Ctx &ctx, const DevProblem &P, const Inst &I, const double *res, const double *xs,
                          double *step)
{
    const int n = P.n, m = P.m, p = P.p;
    const double ep = I.scal[S_EPSP], ed = I.scal[S_EPSD], rho = I.scal[S_RHO];
    const double Jrr = rho + ep, Jss = ep;
    const double *s = I.w + n + m, *t = I.w + n + 2 * m + 2 * p;
    const double *rr = res + n, *rsl = res + n + m, *rt = res + n + 2 * m + 2 * p;
    const double *dy = xs + n, *dz = xs + n + m;
    double *dr = step + n, *ds = step + n + m, *dt = step + n + 2 * m + 2 * p;
    PAR_FOR(i, n) step[i] = xs[i];
    PAR_FOR(i, m) {
        step[n + m + p + i] = dy[i];
        dr[i] = (rr[i] + dy[i]) / Jrr;
    }
    PAR_FOR(i, p) step[n + 2 * m + p + i] = dz[i];
    PAR_FOR(i, P.q_nn) {
        double Sb = s[i] - ed, Ti = t[i];
        double dsi = (rt[i] + Sb * (rsl[i] + dz[i])) / (Ti + Sb * Jss);
        ds[i] = dsi;
        dt[i] = (rt[i] - Ti * dsi) / Sb;
    }
    PAR_FOR(k, P.nsoc) {
        int d = P.soc_dims[k], c0 = P.soc_off[k];
        if (d > 0) {
            const double *sk = s + c0, *tk = t + c0, *a = rsl + c0, *b = rt + c0, *z = dz + c0;
            double *dsk = ds + c0, *dtk = dt + c0;
            auto sb = [&](int i) { return i == 0 ? sk[0] - ed : sk[i]; };           // first row of Sbar
            auto u = [&](int i) { return i == 0 ? tk[0] + (sk[0] - ed) * Jss : tk[i] + sk[i] * Jss; };
            double dot = 0.0;
            for (int i = 0; i < d; i++) dot += sb(i) * (a[i] + z[i]);
            auto x = [&](int i) {
                double v = i == 0 ? dot : (sk[0] - ed) * (a[i] + z[i]) + (a[0] + z[0]) * sk[i];
                return b[i] + v;
            };
            soc_arrow_inverse(d, u, x, [&](int i, double v) { dsk[i] = v; });
            double dot2 = 0.0;   // first row of arrow(t) times ds
            for (int i = 0; i < d; i++) dot2 += tk[i] * dsk[i];
            auto x2 = [&](int i) { return b[i] - (i == 0 ? dot2 : tk[0] * dsk[i] + dsk[0] * tk[i]); };
            soc_arrow_inverse(d, sb, x2, [&](int i, double v) { dtk[i] = v; });
        }
    }
    ctx.sync();
}

// ------------------------------------------------------------------------------------------------ J * v (matrix-free)
#if CB_ON_DEVICE
#ifndef CB_JV_LANES
#define CB_JV_LANES 4      // lanes per sparse row in J v (tunable: 4 or 8; measured in profiles/)
#endif
// lane `sub` of a G-lane group: sum over k = k0 + sub, k0 + sub + G, ... < k1 of vals[src[k]] * vec[col[k]] (src == nullptr:
// vals[k]); eight entries per round with the index loads, then the value loads, batched
template <int G>
__device__ __forceinline__ double sparse_dot(const double *__restrict__ vals, const int *__restrict__ src,
                                             const int *__restrict__ col, const double *vec, int k0, int k1, int sub)
{
    double acc = 0.0;
    for (int q0 = k0 + sub; q0 < k1; q0 += 8 * G) {
        int si[8], ci[8];
        double vv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int q = q0 + G * u;
            const bool ok = q < k1;
            ci[u] = ok ? col[q] : 0;
            si[u] = ok ? (src ? src[q] : q) : -1;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) vv[u] = si[u] >= 0 ? vals[si[u]] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++) acc += vv[u] * vec[ci[u]];
    }
    return acc;
}
__device__ __forceinline__ double sparse_dot4(const double *__restrict__ vals, const int *__restrict__ src,
                                              const int *__restrict__ col, const double *vec, int k0, int k1, int sub)
{
    return sparse_dot<4>(vals, src, col, vec, k0, k1, sub);
}
template <int G> __device__ __forceinline__ double group_sum(double a)
{
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    return a;
}
#endif
CB_DEVN void jacobian_times(const Ctx &ctx, const DevProblem &P, const Inst &I, const double *v, double *out)
{
    const int n = P.n, m = P.m, p = P.p;
    const double ep = I.scal[S_EPSP], ed = I.scal[S_EPSD], rho = I.scal[S_RHO];
    const double *s = I.w + n + m, *t = I.w + n + 2 * m + 2 * p;
    const double *vx = v, *vr = v + n, *vs = v + n + m, *vy = v + n + m + p, *vz = v + n + 2 * m + p,
                 *vt = v + n + 2 * m + 2 * p;
    {
        const int *__restrict__ wp = P.Wfull.ptr, *__restrict__ wc = P.Wfull.col;
        const int *__restrict__ gp = P.Gp, *__restrict__ gi = P.Gi, *__restrict__ cp = P.Cp, *__restrict__ ci = P.Ci;
        const double *__restrict__ Gv = I.Gv, *__restrict__ Cv = I.Cv;
        const int *__restrict__ grp = P.Grow.ptr, *__restrict__ grc = P.Grow.col;
#if CB_ON_DEVICE
        unsigned dyn_bytes;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
        if (!ctx.warp_scope && dyn_bytes >= (unsigned)P.N * 8u) {
            // the gathered parts of v (x, y, z blocks) are staged in shared memory; four lanes per sparse row with the
            // row pointers of the next row in flight
            double *sx = cb_dyn_smem, *sy = sx + n, *sz = sy + m;
            for (int k = ctx.tid; k < n; k += ctx.nthr) sx[k] = vx[k];
            for (int k = ctx.tid; k < m; k += ctx.nthr) sy[k] = vy[k];
            for (int k = ctx.tid; k < p; k += ctx.nthr) sz[k] = vz[k];
            __syncthreads();
            constexpr int JG = CB_JV_LANES;
            const int sub = threadIdx.x & (JG - 1), gidx = threadIdx.x / JG, ngrp = blockDim.x / JG;
            {
                const int padded = (n + ngrp - 1) / ngrp * ngrp;
                int i = gidx;
                int w0 = i < n ? wp[i] : 0, w1 = i < n ? wp[i + 1] : 0, g0 = i < n ? gp[i] : 0, g1 = i < n ? gp[i + 1] : 0,
                    c0 = i < n ? cp[i] : 0, c1 = i < n ? cp[i + 1] : 0;
                for (; i < padded; i += ngrp) {
                    const int in = i + ngrp;
                    const bool okn = in < n;
                    const int nw0 = okn ? wp[in] : 0, nw1 = okn ? wp[in + 1] : 0, ng0 = okn ? gp[in] : 0,
                              ng1 = okn ? gp[in + 1] : 0, nc0 = okn ? cp[in] : 0, nc1 = okn ? cp[in + 1] : 0;
                    double a = sparse_dot<JG>(I.Wf, nullptr, wc, sx, w0, w1, sub) + sparse_dot<JG>(Gv, nullptr, gi, sy, g0, g1, sub);
                    for (int k = c0 + sub; k < c1; k += JG) a += Cv[k] * sz[ci[k]];
                    a = group_sum<JG>(a);
                    if (sub == 0 && i < n) out[i] = ep * sx[i] + a;
                    w0 = nw0; w1 = nw1; g0 = ng0; g1 = ng1; c0 = nc0; c1 = nc1;
                }
            }
            {
                const int padded = (m + ngrp - 1) / ngrp * ngrp;
                int i = gidx;
                int a0 = i < m ? P.Grow.ptr[i] : 0, a1 = i < m ? P.Grow.ptr[i + 1] : 0;
                for (; i < padded; i += ngrp) {
                    const int in = i + ngrp;
                    const int na0 = in < m ? P.Grow.ptr[in] : 0, na1 = in < m ? P.Grow.ptr[in + 1] : 0;
                    double a = sparse_dot<JG>(I.Gr, nullptr, grc, sx, a0, a1, sub);
                    a = group_sum<JG>(a);
                    if (sub == 0 && i < m) {
                        out[n + i] = (rho + ep) * vr[i] - sy[i];
                        out[n + m + p + i] = a - vr[i] - ed * sy[i];
                    }
                    a0 = na0; a1 = na1;
                }
            }
            __syncthreads();
        } else
#endif
        {
            grouped_rows<4>(
                ctx, n,
                [&](int i, int sub, int st) {
                    double a = 0.0;
                    for (int k = wp[i] + sub; k < wp[i + 1]; k += st) a += I.Wf[k] * vx[wc[k]];
                    for (int k = gp[i] + sub; k < gp[i + 1]; k += st) a += Gv[k] * vy[gi[k]];
                    for (int k = cp[i] + sub; k < cp[i + 1]; k += st) a += Cv[k] * vz[ci[k]];
                    return a;
                },
                [&](int i, double a) { out[i] = ep * vx[i] + a; });
            grouped_rows<4>(
                ctx, m,
                [&](int i, int sub, int st) {
                    double a = 0.0;
                    for (int k = grp[i] + sub; k < grp[i + 1]; k += st) a += I.Gr[k] * vx[grc[k]];
                    return a;
                },
                [&](int i, double a) {
                    out[n + i] = (rho + ep) * vr[i] - vy[i];
                    out[n + m + p + i] = a - vr[i] - ed * vy[i];
                });
        }
    }
    PAR_FOR(i, p) {
        out[n + m + i] = ep * vs[i] - vz[i] - vt[i];
        double a = 0.0;
        for (int k = P.Crow.ptr[i]; k < P.Crow.ptr[i + 1]; k++) a += I.Cv[P.Crow.src[k]] * vx[P.Crow.col[k]];
        out[n + 2 * m + p + i] = a - vs[i] - ed * vz[i];
    }
    PAR_FOR(i, P.q_nn) out[n + 2 * m + 2 * p + i] = t[i] * vs[i] + (s[i] - ed) * vt[i];
    PAR_FOR(k, P.nsoc) {
        int d = P.soc_dims[k], c0 = P.soc_off[k];
        const double *sk = s + c0, *tk = t + c0, *a = vs + c0, *b = vt + c0;
        double dt_ = 0.0, dsb = 0.0;
        for (int i = 0; i < d; i++) { dt_ += tk[i] * a[i]; dsb += sk[i] * b[i]; }
        for (int i = 0; i < d; i++)
            out[n + 2 * m + 2 * p + c0 + i] = arrow_apply(tk, a, i, dt_) + arrow_apply(sk, b, i, dsb) - ed * b[i];
    }
    ctx.sync();
}

// ------------------------------------------------------------------------------------------------ search direction
// factorize_regularized_residual_jacobian_variables!  inertia.jl:13-28
CB_DEV bool factorize_regularized(const Ctx &ctx, const DevProblem &P, const Inst &I)
{
    ProfTimer pt{I.prof, 0};
    pt.start();
    kkt_entries(ctx, P, I);
    pt.stop(PROF_ASSEMBLE);
    ldl_factor(ctx, P, I.panels, I.D, I.Dinv, KSrc{I.Wv, I.Gr, I.Cv, I.kx}, I.Lcsr, I.istat, I.prof);
    bool ok = I.istat[I_INERTIA_POS] == P.n && I.istat[I_INERTIA_NEG] == P.m + P.p && I.istat[I_INERTIA_ZERO] == 0;
    ctx.sync();
    if (ctx.tid == 0) I.istat[I_TRIALS]++;
    ctx.sync();
    return ok;
}

// inertia_correction!  inertia.jl:30-79.  All threads follow the same (uniform) control flow.
CB_DEVN int inertia_correction(const Ctx &ctx, const DevProblem &P, const Inst &I, const Options &o)
{
    if (ctx.tid == 0) {
        I.scal[S_EPSP] = o.primal_regularization_initial;
        I.scal[S_EPSD] = o.dual_regularization_initial;
        I.istat[I_TRIALS] = 0;
    }
    ctx.sync();
    if (factorize_regularized(ctx, P, I)) return ST_OK;   // IC-1
    if (ctx.tid == 0) {
        if (I.istat[I_INERTIA_ZERO] != 0)                  // IC-2
            I.scal[S_EPSD] = o.dual_regularization * pow(I.scal[S_KAPPA], o.dual_regularization_exponent);
        // IC-3: the reference's `primal_regularization_last == 0.0` compares a Vector with a Float64 (always false)
        double v = o.scaling_regularization_last * I.scal[S_EPSP_LAST];
        I.scal[S_EPSP] = o.min_regularization > v ? o.min_regularization : v;
    }
    ctx.sync();
    for (;;) {
        if (factorize_regularized(ctx, P, I)) break;      // IC-4
        double ep = I.scal[S_EPSP];
        ep = (I.scal[S_EPSP_LAST] == 0.0 ? o.scaling_regularization_initial : o.scaling_regularization) * ep;   // IC-5
        ctx.sync();
        if (ctx.tid == 0) I.scal[S_EPSP] = ep;
        ctx.sync();
        if (ep > o.max_regularization) return ST_INERTIA_FAILURE;   // IC-6
    }
    if (ctx.tid == 0) I.scal[S_EPSP_LAST] = I.scal[S_EPSP];
    ctx.sync();
    return ST_OK;
}

// search_direction_symmetric!  search_direction.jl:25-104 (the redundant re-factorisation of linear_solve! is skipped,
// SURVEY.md Appendix A.7)
CB_DEV void direction_symmetric(const Ctx &ctx, const DevProblem &P, const Inst &I, const double *res, double *step)
{
    ProfTimer pt{I.prof, 0};
    pt.start();
    reduced_rhs(ctx, P, I, res, I.rs);
    pt.stop(PROF_RHS_RECOVER);
    ldl_solve(ctx, P, I.panels, I.D, I.Dinv, I.kx, I.Lcsr, I.rs, I.xs, I.xp, I.istat, I.prof);
    pt.start();
    recover_step(ctx, P, I, res, I.xs, step);
    pt.stop(PROF_RHS_RECOVER);
}

CB_DEV double residual_error(const Ctx &ctx, const DevProblem &P, const Inst &I, const double *step)
{   // err = R - J step, returns its infinity norm
    ProfTimer pt{I.prof, 0};
    pt.start();
    jacobian_times(ctx, P, I, step, I.tmp);
    // one pass: the difference is stored and its magnitude enters the (order-independent) maximum
    double r = scope_max(ctx, P.total, [&](int i) {
        const double e = I.res[i] - I.tmp[i];
        I.err[i] = e;
        return fabs(e);
    });
    pt.stop(PROF_JTIMES);
    return r;
}

// iterative_refinement!  iterative_refinement.jl:1-53
CB_DEVN bool iterative_refinement(const Ctx &ctx, const DevProblem &P, const Inst &I, const Options &o)
{
    int iteration = 0;
    double rn = residual_error(ctx, P, I, I.step);
    const double rn0 = rn;
    bool done = false;
    while (iteration <= o.max_iterative_refinement) {
        if (rn <= o.iterative_refinement_tolerance && iteration >= o.min_iterative_refinement) { done = true; break; }
        direction_symmetric(ctx, P, I, I.err, I.corr);
#if CB_ON_DEVICE
        for (int i0 = ctx.tid; i0 < P.total; i0 += 4 * ctx.nthr) {      // four independent load pairs in flight
            double a[4], c[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * ctx.nthr;
                a[u] = i < P.total ? I.step[i] : 0.0;
                c[u] = i < P.total ? I.corr[i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * ctx.nthr;
                if (i < P.total) I.step[i] = a[u] + c[u];
            }
        }
#else
        PAR_FOR(i, P.total) I.step[i] += I.corr[i];
#endif
        ctx.sync();
        rn = residual_error(ctx, P, I, I.step);
        iteration++;
    }
    if (ctx.tid == 0) {
        I.istat[I_REFINE] = iteration;
        I.scal[S_REFINE_NORM] = rn;
        I.scal[S_REFINE_NORM_INITIAL] = rn0;
    }
    ctx.sync();
    return done || rn <= rn0;
}

// Fallback when refinement fails.  The reference re-solves J step = R with UMFPACK (search_direction.jl:22,113); here
// the same system is solved by restarted GMRES on the matrix-free J, right-preconditioned with the reduced LDL^T
// solve + recovery (the linear map the refinement also uses).  Starts from zero like a direct solve.
CB_DEVN bool gmres_fallback(const Ctx &ctx, const DevProblem &P, const Inst &I, const Options &o)
{
    const int T = P.total, mr = o.gmres_restart;
    double *V = I.krylov;            // (mr+1) x T
    double *H = V + (long long)(mr + 1) * T;   // (mr+1) x mr column-major Hessenberg, then cs[mr], sn[mr], g[mr+1], y[mr]
    double *cs = H + (mr + 1) * mr, *sn = cs + mr, *g = sn + mr, *y = g + mr + 1;
    PAR_FOR(i, T) I.step[i] = 0.0;
    ctx.sync();
    int total_it = 0;
    bool ok = false;
    double rn = 0.0;
    for (int cycle = 0; cycle < o.gmres_max_cycles; cycle++) {
        rn = residual_error(ctx, P, I, I.step);     // err = R - J step
        if (rn <= o.iterative_refinement_tolerance) { ok = true; break; }
        double beta = sqrt(scope_sum(ctx, T, [&](int i) { return I.err[i] * I.err[i]; }));
        PAR_FOR(i, T) V[i] = I.err[i] / beta;
        if (ctx.tid == 0) { for (int k = 0; k <= mr; k++) g[k] = 0.0; g[0] = beta; }
        ctx.sync();
        int j = 0;
        for (; j < mr; j++) {
            double *vj = V + (long long)j * T, *wv = V + (long long)(j + 1) * T;
            direction_symmetric(ctx, P, I, vj, I.corr);      // z = M^-1 v_j
            jacobian_times(ctx, P, I, I.corr, wv);           // w = J z
            for (int i = 0; i <= j; i++) {                   // modified Gram-Schmidt
                const double *vi = V + (long long)i * T;
                double h = scope_sum(ctx, T, [&](int k) { return wv[k] * vi[k]; });
                PAR_FOR(k, T) wv[k] -= h * vi[k];
                if (ctx.tid == 0) H[i + j * (mr + 1)] = h;
                ctx.sync();
            }
            double hn = sqrt(scope_sum(ctx, T, [&](int k) { return wv[k] * wv[k]; }));
            PAR_FOR(k, T) wv[k] = hn > 0.0 ? wv[k] / hn : 0.0;
            if (ctx.tid == 0) {
                H[j + 1 + j * (mr + 1)] = hn;
                for (int i = 0; i < j; i++) {                // apply previous rotations
                    double a = H[i + j * (mr + 1)], b = H[i + 1 + j * (mr + 1)];
                    H[i + j * (mr + 1)] = cs[i] * a + sn[i] * b;
                    H[i + 1 + j * (mr + 1)] = -sn[i] * a + cs[i] * b;
                }
                double a = H[j + j * (mr + 1)], b = H[j + 1 + j * (mr + 1)];
                double r = sqrt(a * a + b * b);
                cs[j] = r > 0.0 ? a / r : 1.0;
                sn[j] = r > 0.0 ? b / r : 0.0;
                H[j + j * (mr + 1)] = r;
                H[j + 1 + j * (mr + 1)] = 0.0;
                g[j + 1] = -sn[j] * g[j];
                g[j] = cs[j] * g[j];
            }
            ctx.sync();
            total_it++;
            double est = fabs(g[j + 1]);
            ctx.sync();
            if (est <= 0.1 * o.iterative_refinement_tolerance || hn == 0.0) { j++; break; }
        }
        if (ctx.tid == 0) {                                   // back substitution
            for (int i = j - 1; i >= 0; i--) {
                double a = g[i];
                for (int k = i + 1; k < j; k++) a -= H[i + k * (mr + 1)] * y[k];
                y[i] = a / H[i + i * (mr + 1)];
            }
        }
        ctx.sync();
        PAR_FOR(k, T) {                                       // tmp = V y
            double a = 0.0;
            for (int i = 0; i < j; i++) a += V[(long long)i * T + k] * y[i];
            I.tmp[k] = a;
        }
        ctx.sync();
        PAR_FOR(k, T) I.err[k] = I.tmp[k];
        ctx.sync();
        direction_symmetric(ctx, P, I, I.err, I.corr);        // step += M^-1 (V y)
        PAR_FOR(k, T) I.step[k] += I.corr[k];
        ctx.sync();
    }
    if (!ok) {
        rn = residual_error(ctx, P, I, I.step);
        ok = rn <= o.iterative_refinement_tolerance;
    }
    if (ctx.tid == 0) { I.istat[I_GMRES_ITERS] = total_it; I.scal[S_REFINE_NORM] = rn; }
    ctx.sync();
    return ok;
}

// search_direction!  search_direction.jl:1-23
CB_DEVN int search_direction(const Ctx &ctx, const DevProblem &P, const Inst &I, const Options &o)
{
    int st = inertia_correction(ctx, P, I, o);
    if (st != ST_OK) return st;
    direction_symmetric(ctx, P, I, I.res, I.step);
    if (ctx.tid == 0) { I.istat[I_REFINE_OK] = 1; I.istat[I_USED_FALLBACK] = 0; I.istat[I_REFINE] = 0; }
    ctx.sync();
    if (o.iterative_refinement) {
        bool ok = iterative_refinement(ctx, P, I, o);
        if (ctx.tid == 0) I.istat[I_REFINE_OK] = ok ? 1 : 0;
        ctx.sync();
        if (!ok) {
            if (I.krylov == nullptr) {
                if (ctx.tid == 0) I.istat[I_UNREFINED_STEPS]++;
                ctx.sync();
                return ST_REFINEMENT_FAILURE;
            }
            bool gok = gmres_fallback(ctx, P, I, o);
            if (ctx.tid == 0) { I.istat[I_USED_FALLBACK] = 1; I.istat[I_FALLBACKS]++; }
            ctx.sync();
            if (!gok) {      // neither refinement nor the fallback reached the tolerance: counted, reported as status 2
                if (ctx.tid == 0) I.istat[I_UNREFINED_STEPS]++;
                ctx.sync();
                return ST_REFINEMENT_FAILURE;
            }
        }
    }
    return ST_OK;
}

// ------------------------------------------------------------------------------------------------ cone line search
// cone_violation(xhat, x, tau), cones/cone.jl:62-68 with xhat = x - alpha * dx formed on the fly
CB_DEV int cone_violation_step(const Ctx &ctx, const DevProblem &P, const double *x, const double *dx, double alpha,
                               double tau, double *xhat)
{
    PAR_FOR(i, P.p) xhat[i] = x[i] - alpha * dx[i];
    ctx.sync();
    int nn = scope_any(ctx, P.q_nn, [&](int i) { return xhat[i] <= (1.0 - tau) * x[i]; });
    int so = scope_any(ctx, P.nsoc, [&](int k) {
        int d = P.soc_dims[k], c0 = P.soc_off[k];
        if (d <= 0) return false;
        double acc = 0.0;
        for (int i = 1; i < d; i++) {
            double v = xhat[c0 + i] - (1.0 - tau) * x[c0 + i];
            acc += v * v;
        }
        return xhat[c0] - (1.0 - tau) * x[c0] <= sqrt(acc);
    });
    return nn | so;
}

// solve.jl:190-221: independent halving searches for s (step_size) and t (step_size_cone_slack_dual)
CB_DEVN int cone_search(const Ctx &ctx, const DevProblem &P, const Inst &I, const Options &o)
{
    const int os = P.n + P.m, ot = P.n + 2 * P.m + 2 * P.p;
    const double tau = I.scal[S_TAU];
    double a = 1.0, at = 1.0;
    int it = 0, status = ST_OK;
    while (cone_violation_step(ctx, P, I.w + os, I.step + os, a, tau, I.cand + os)) {
        a = o.scaling_line_search * a;
        it++;
        if (it > o.max_cone_line_search) { status = ST_CONE_SEARCH_FAILURE; break; }
    }
    int ks = it;
    it = 0;
    if (status == ST_OK)
        while (cone_violation_step(ctx, P, I.w + ot, I.step + ot, at, tau, I.cand + ot)) {
            at = o.scaling_line_search * at;
            it++;
            if (it > o.max_cone_line_search) { status = ST_CONE_SEARCH_FAILURE; break; }
        }
    if (ctx.tid == 0) {
        I.scal[S_STEP_SIZE] = a;
        I.scal[S_STEP_SIZE_T] = at;
        I.istat[I_KS] = ks;
        I.istat[I_KT] = it;
    }
    ctx.sync();
    return status;
}

}  // namespace cb200
