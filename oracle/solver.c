/*
 * solver.c -- CPU ORACLE (test infrastructure only): restatement of the reference's per-iteration Newton/KKT
 * path and of the solve! loop that drives it (src/solver/ *.jl).  See oracle.h for scope and parity statement.
 *
 * Differences from the reference that do not change results:
 *   - the sparsity pattern is structural (given) rather than "non-zero at a random point" (A.1);
 *   - the full Jacobian J is kept as blocks and applied matrix-free (mul! with a SparseMatrixCSC in the reference,
 *     iterative_refinement.jl:9,39); orc_dense_jacobian() materialises it for the block-identity tests;
 *   - the three Hessian pieces (objective, equality tensor, cone tensor; residual_jacobian_variables.jl:11-13) are
 *     delivered already summed by the evaluation callback;
 *   - assembly is O(nnz) through index maps instead of all-pairs scalar indexing.
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double theta, merit;
} filter_pair;

struct orc_solver {
    int n, m, p, N, total;
    int q_nn, nsoc;
    int *soc_dims, *soc_off, *soc_blk; /* soc_off: first cone row; soc_blk: offset of d*d block storage */
    int soc_blk_total;
    int oxr, os, oy, oz, ot; /* offsets of r,s,y,z,t in w (x at 0), indices.jl:25-35 */
    /* patterns */
    int *Wp, *Wi, *Wdiag; /* W upper CSC; Wdiag[j] = position of (j,j) */
    int *Gp, *Gi, *Cp, *Ci;
    int nnzW, nnzG, nnzC;
    int *Grp, *Gcj, *Gsrc; /* CSR view of G: row pointer, column, position in G_val */
    int *Crp, *Ccj, *Csrc;
    /* problem data (problem_data.jl:33-100) */
    orc_eval_out pd;
    double objective[1];
    double *cone_product, *cone_target, *barrier_gradient;
    double barrier[1];
    double *jac_primal_nn, *jac_dual_nn;   /* diagonal parts of cone_product_jacobian_primal/dual */
    double *jac_primal_soc, *jac_dual_soc; /* dense d x d column-major blocks */
    /* full Jacobian J in block form (residual_jacobian_variables.jl:1-108) */
    double *Jxx;                  /* W upper values + eps_p on the diagonal */
    double Jrr, Jss, Jyy, Jzz;    /* rho+eps_p, eps_p, -eps_d, -eps_d */
    double *Jts_nn, *Jtt_nn, *Jts_soc, *Jtt_soc;
    /* reduced matrix K (upper triangle, natural order, sorted rows) */
    int *Kp, *Ki;
    double *Kx;
    int nnzK;
    int *KfromW, *Kydiag, *Kzdiag; /* maps */
    int *KfromG, *KfromC;          /* position in Kx of each G / C entry */
    int *Kzblk;                    /* per SOC: positions of the d*(d+1)/2 upper entries, column by column */
    orc_qdldl *ldl;
    int inertia[3];
    /* solver data (solver_data.jl:26-92) */
    double *solution, *candidate, *step, *residual, *residual_error, *step_correction;
    double *residual_symmetric, *step_symmetric, *merit_gradient, *constraint_violation;
    double *dual; /* lambda */
    double scal[8]; /* kappa, tau, rho, eps_p, eps_d, eps_p_last, objective(copy), barrier(copy) */
    filter_pair *fpairs, *fcache;
    int findex;
    int stats[12];
    orc_lu_fn lu_fn;
    void *lu_user;
    orc_options opt;
    orc_eval_fn fn;
    void *user;
    /* LQ evaluator data */
    double *lqW, *lqG, *lqC, *lqq, *lqg0, *lqh0;
    double *tmp_p, *tmp_p2, *tmp_tot;
    /* violations carried between iterations (solve.jl:85-86,332-333) */
    double equality_violation, cone_product_violation;
};

#define KAPPA s->scal[0]
#define TAU s->scal[1]
#define RHO s->scal[2]
#define EPSP s->scal[3]
#define EPSD s->scal[4]
#define EPSP_LAST s->scal[5]

void orc_options_default(orc_options *o)
{ /* options.jl:6-59 */
    o->max_outer_iterations = 10;
    o->max_residual_iterations = 100;
    o->scaling_line_search = 0.5;
    o->max_residual_line_search = 25;
    o->max_cone_line_search = 25;
    o->iterative_refinement = 1;
    o->max_iterative_refinement = 10;
    o->min_iterative_refinement = 1;
    o->iterative_refinement_tolerance = 1.0e-10;
    o->central_path_initial = 1.0;
    o->central_path_update_tolerance = 10.0;
    o->central_path_scaling = 0.2;
    o->central_path_exponent = 1.5;
    o->penalty_initial = 1.0;
    o->penalty_scaling = 10.0;
    o->dual_initial = 0.0;
    o->residual_tolerance = 1.0e-4;
    o->optimality_tolerance = 1.0e-4;
    o->slack_tolerance = 1.0e-4;
    o->equality_tolerance = 1.0e-4;
    o->complementarity_tolerance = 1.0e-4;
    o->min_regularization = 1.0e-20;
    o->primal_regularization_initial = 1.0e-7;
    o->dual_regularization_initial = 1.0e-7;
    o->max_regularization = 1.0e40;
    o->dual_regularization = 1.0e-8;
    o->dual_regularization_exponent = 0.25;
    o->scaling_regularization_initial = 100.0;
    o->scaling_regularization = 8.0;
    o->scaling_regularization_last = 1.0 / 3.0;
    o->max_penalty = 1.0e8;
    o->violation_tolerance = 1.0e-5;
    o->violation_exponent = 1.1;
    o->merit_tolerance = 1.0e-5;
    o->merit_exponent = 2.3;
    o->armijo_tolerance = 1.0e-4;
    o->machine_tolerance = 1.0e-16;
    o->max_filter = 1000;
    o->warmstart = 0;
    o->reference_schedule = 0;
}

static double *dalloc(size_t k) { return (double *)calloc(k > 0 ? k : 1, sizeof(double)); }
static int *ialloc(size_t k) { return (int *)calloc(k > 0 ? k : 1, sizeof(int)); }
static int *icopy(const int *a, size_t k)
{
    int *r = ialloc(k);
    memcpy(r, a, k * sizeof(int));
    return r;
}

static void csr_view(int nrows, int ncols, const int *cp, const int *ri, int **rp_out, int **cj_out, int **src_out)
{
    int nnz = cp[ncols];
    int *rp = ialloc((size_t)nrows + 1), *cj = ialloc((size_t)nnz), *src = ialloc((size_t)nnz);
    for (int k = 0; k < nnz; k++) rp[ri[k] + 1]++;
    for (int i = 0; i < nrows; i++) rp[i + 1] += rp[i];
    int *next = icopy(rp, (size_t)nrows + 1);
    for (int j = 0; j < ncols; j++)
        for (int k = cp[j]; k < cp[j + 1]; k++) {
            int d = next[ri[k]]++;
            cj[d] = j;
            src[d] = k;
        }
    free(next);
    *rp_out = rp;
    *cj_out = cj;
    *src_out = src;
}

orc_solver *orc_solver_new(int n, int m, int p, int q_nn, int nsoc, const int *soc_dims, const int *Wp,
                           const int *Wi, const int *Gp, const int *Gi, const int *Cp, const int *Ci,
                           const int *perm, const orc_options *opts)
{
    orc_solver *s = (orc_solver *)calloc(1, sizeof(orc_solver));
    s->n = n; s->m = m; s->p = p; s->N = n + m + p; s->total = n + 2 * m + 3 * p;
    s->q_nn = q_nn; s->nsoc = nsoc;
    s->oxr = n; s->os = n + m; s->oy = n + m + p; s->oz = n + 2 * m + p; s->ot = n + 2 * m + 2 * p;
    if (opts) s->opt = *opts; else orc_options_default(&s->opt);
    s->soc_dims = ialloc((size_t)nsoc); s->soc_off = ialloc((size_t)nsoc); s->soc_blk = ialloc((size_t)nsoc);
    int off = q_nn, blk = 0;
    for (int k = 0; k < nsoc; k++) {
        s->soc_dims[k] = soc_dims[k];
        s->soc_off[k] = off;
        s->soc_blk[k] = blk;
        off += soc_dims[k];
        blk += soc_dims[k] * soc_dims[k];
    }
    s->soc_blk_total = blk;
    if (off != p) { free(s); return NULL; }
    s->nnzW = Wp[n]; s->nnzG = Gp[n]; s->nnzC = Cp[n];
    s->Wp = icopy(Wp, (size_t)n + 1); s->Wi = icopy(Wi, (size_t)s->nnzW);
    s->Gp = icopy(Gp, (size_t)n + 1); s->Gi = icopy(Gi, (size_t)s->nnzG);
    s->Cp = icopy(Cp, (size_t)n + 1); s->Ci = icopy(Ci, (size_t)s->nnzC);
    s->Wdiag = ialloc((size_t)n);
    for (int j = 0; j < n; j++) {
        s->Wdiag[j] = -1;
        for (int k = Wp[j]; k < Wp[j + 1]; k++) {
            if (Wi[k] > j) { free(s); return NULL; } /* must be upper triangular */
            if (Wi[k] == j) s->Wdiag[j] = k;
        }
        if (s->Wdiag[j] < 0) { free(s); return NULL; } /* diagonal must be structurally present */
    }
    csr_view(m, n, Gp, Gi, &s->Grp, &s->Gcj, &s->Gsrc);
    csr_view(p, n, Cp, Ci, &s->Crp, &s->Ccj, &s->Csrc);

    s->pd.objective = s->objective;
    s->pd.gradient = dalloc((size_t)n); s->pd.equality = dalloc((size_t)m); s->pd.cone = dalloc((size_t)p);
    s->pd.eq_dual_grad = dalloc((size_t)n); s->pd.cone_dual_grad = dalloc((size_t)n);
    s->pd.W_val = dalloc((size_t)s->nnzW); s->pd.G_val = dalloc((size_t)s->nnzG); s->pd.C_val = dalloc((size_t)s->nnzC);
    s->cone_product = dalloc((size_t)p); s->cone_target = dalloc((size_t)p); s->barrier_gradient = dalloc((size_t)p);
    s->jac_primal_nn = dalloc((size_t)q_nn); s->jac_dual_nn = dalloc((size_t)q_nn);
    s->jac_primal_soc = dalloc((size_t)blk); s->jac_dual_soc = dalloc((size_t)blk);
    s->Jxx = dalloc((size_t)s->nnzW);
    s->Jts_nn = dalloc((size_t)q_nn); s->Jtt_nn = dalloc((size_t)q_nn);
    s->Jts_soc = dalloc((size_t)blk); s->Jtt_soc = dalloc((size_t)blk);

    /* ---- pattern of K (upper triangle, natural (x,y,z) order, rows sorted), residual_jacobian_variables.jl:110-167 */
    int N = s->N;
    int nnzK = s->nnzW + s->nnzG + m + s->nnzC + q_nn;
    for (int k = 0; k < nsoc; k++) nnzK += soc_dims[k] * (soc_dims[k] + 1) / 2;
    s->nnzK = nnzK;
    s->Kp = ialloc((size_t)N + 1); s->Ki = ialloc((size_t)nnzK); s->Kx = dalloc((size_t)nnzK);
    s->KfromW = ialloc((size_t)s->nnzW); s->KfromG = ialloc((size_t)s->nnzG); s->KfromC = ialloc((size_t)s->nnzC);
    s->Kydiag = ialloc((size_t)m); s->Kzdiag = ialloc((size_t)p);
    int tri_total = 0;
    for (int k = 0; k < nsoc; k++) tri_total += soc_dims[k] * (soc_dims[k] + 1) / 2;
    s->Kzblk = ialloc((size_t)tri_total);
    int pos = 0;
    for (int j = 0; j < n; j++) { /* x columns: W upper */
        s->Kp[j] = pos;
        for (int k = Wp[j]; k < Wp[j + 1]; k++) { s->Ki[pos] = Wi[k]; s->KfromW[k] = pos; pos++; }
    }
    for (int i = 0; i < m; i++) { /* y columns: row i of G, then the diagonal */
        s->Kp[n + i] = pos;
        for (int k = s->Grp[i]; k < s->Grp[i + 1]; k++) { s->Ki[pos] = s->Gcj[k]; s->KfromG[s->Gsrc[k]] = pos; pos++; }
        s->Ki[pos] = n + i; s->Kydiag[i] = pos; pos++;
    }
    int tri = 0, soc_k = 0;
    for (int i = 0; i < p; i++) { /* z columns: row i of C, SOC block rows above the diagonal, the diagonal */
        s->Kp[n + m + i] = pos;
        for (int k = s->Crp[i]; k < s->Crp[i + 1]; k++) { s->Ki[pos] = s->Ccj[k]; s->KfromC[s->Csrc[k]] = pos; pos++; }
        if (i >= q_nn) {
            while (soc_k < nsoc && i >= s->soc_off[soc_k] + s->soc_dims[soc_k]) soc_k++;
            int c0 = s->soc_off[soc_k];
            for (int r = c0; r <= i; r++) { s->Ki[pos] = n + m + r; s->Kzblk[tri++] = pos; if (r == i) s->Kzdiag[i] = pos; pos++; }
        } else {
            s->Ki[pos] = n + m + i; s->Kzdiag[i] = pos; pos++;
        }
    }
    s->Kp[N] = pos;
    if (pos != nnzK) { fprintf(stderr, "oracle: K pattern count mismatch %d != %d\n", pos, nnzK); }

    /* symbolic + first numeric factorisation on a well-posed stand-in matrix (solver.jl:88-122 uses a random point;
     * only the pattern matters here) */
    for (int k = 0; k < nnzK; k++) s->Kx[k] = 0.0;
    for (int j = 0; j < n; j++) s->Kx[s->KfromW[s->Wdiag[j]]] = 1.0;
    for (int i = 0; i < m; i++) s->Kx[s->Kydiag[i]] = -1.0;
    for (int i = 0; i < p; i++) s->Kx[s->Kzdiag[i]] = -1.0;
    s->ldl = N > 0 ? orc_qdldl_new(N, s->Kp, s->Ki, s->Kx, perm) : NULL;

    int tot = s->total;
    s->solution = dalloc((size_t)tot); s->candidate = dalloc((size_t)tot); s->step = dalloc((size_t)tot);
    s->residual = dalloc((size_t)tot); s->residual_error = dalloc((size_t)tot); s->step_correction = dalloc((size_t)tot);
    s->residual_symmetric = dalloc((size_t)N); s->step_symmetric = dalloc((size_t)N);
    s->merit_gradient = dalloc((size_t)N); s->constraint_violation = dalloc((size_t)(m + p));
    s->dual = dalloc((size_t)m);
    s->tmp_p = dalloc((size_t)p); s->tmp_p2 = dalloc((size_t)p); s->tmp_tot = dalloc((size_t)tot);
    s->fpairs = (filter_pair *)malloc(sizeof(filter_pair) * (size_t)s->opt.max_filter);
    s->fcache = (filter_pair *)malloc(sizeof(filter_pair) * (size_t)s->opt.max_filter);
    for (int i = 0; i < s->opt.max_filter; i++) { s->fpairs[i].theta = s->fpairs[i].merit = 1.0e8; s->fcache[i] = s->fpairs[i]; }
    s->findex = 0;
    KAPPA = 0.1; TAU = 0.99; RHO = 10.0; /* solver.jl:81-86 */
    EPSP = 0.0; EPSD = 0.0; EPSP_LAST = 0.0; /* solver.jl:125-127 */
    return s;
}

void orc_solver_free(orc_solver *s)
{
    if (!s) return;
    free(s->soc_dims); free(s->soc_off); free(s->soc_blk);
    free(s->Wp); free(s->Wi); free(s->Wdiag); free(s->Gp); free(s->Gi); free(s->Cp); free(s->Ci);
    free(s->Grp); free(s->Gcj); free(s->Gsrc); free(s->Crp); free(s->Ccj); free(s->Csrc);
    free(s->pd.gradient); free(s->pd.equality); free(s->pd.cone); free(s->pd.eq_dual_grad); free(s->pd.cone_dual_grad);
    free(s->pd.W_val); free(s->pd.G_val); free(s->pd.C_val);
    free(s->cone_product); free(s->cone_target); free(s->barrier_gradient);
    free(s->jac_primal_nn); free(s->jac_dual_nn); free(s->jac_primal_soc); free(s->jac_dual_soc);
    free(s->Jxx); free(s->Jts_nn); free(s->Jtt_nn); free(s->Jts_soc); free(s->Jtt_soc);
    free(s->Kp); free(s->Ki); free(s->Kx); free(s->KfromW); free(s->KfromG); free(s->KfromC);
    free(s->Kydiag); free(s->Kzdiag); free(s->Kzblk);
    orc_qdldl_free(s->ldl);
    free(s->solution); free(s->candidate); free(s->step); free(s->residual); free(s->residual_error);
    free(s->step_correction); free(s->residual_symmetric); free(s->step_symmetric); free(s->merit_gradient);
    free(s->constraint_violation); free(s->dual); free(s->tmp_p); free(s->tmp_p2); free(s->tmp_tot);
    free(s->fpairs); free(s->fcache);
    free(s->lqW); free(s->lqG); free(s->lqC); free(s->lqq); free(s->lqg0); free(s->lqh0);
    free(s);
}

/* -------------------------------------------------------------------- accessors */
double *orc_solution(orc_solver *s) { return s->solution; }
double *orc_candidate(orc_solver *s) { return s->candidate; }
double *orc_step(orc_solver *s) { return s->step; }
double *orc_residual(orc_solver *s) { return s->residual; }
double *orc_residual_symmetric_vec(orc_solver *s) { return s->residual_symmetric; }
double *orc_step_symmetric(orc_solver *s) { return s->step_symmetric; }
double *orc_dual(orc_solver *s) { return s->dual; }
double *orc_scalars(orc_solver *s) { s->scal[6] = s->objective[0]; s->scal[7] = s->barrier[0]; return s->scal; }
orc_eval_out *orc_problem(orc_solver *s) { return &s->pd; }
double *orc_cone_product(orc_solver *s) { return s->cone_product; }
double *orc_cone_target(orc_solver *s) { return s->cone_target; }
double *orc_barrier_gradient(orc_solver *s) { return s->barrier_gradient; }
double *orc_merit_gradient(orc_solver *s) { return s->merit_gradient; }
const int *orc_inertia(orc_solver *s) { return s->inertia; }
const int *orc_stats(orc_solver *s) { return s->stats; }
orc_qdldl *orc_linear_solver(orc_solver *s) { return s->ldl; }
int orc_K_nnz(orc_solver *s) { return s->nnzK; }
const int *orc_K_colptr(orc_solver *s) { return s->Kp; }
const int *orc_K_rowval(orc_solver *s) { return s->Ki; }
const double *orc_K_nzval(orc_solver *s) { return s->Kx; }

void orc_solver_set_callback(orc_solver *s, orc_eval_fn fn, void *user) { s->fn = fn; s->user = user; }

/* -------------------------------------------------------------------- LQ evaluator (calipso_b200/lqc.py family) */
static void lq_eval(void *user, int flags, const double *x, const double *y, const double *z, orc_eval_out *o)
{
    orc_solver *s = (orc_solver *)user;
    int n = s->n, m = s->m, p = s->p;
    if (flags & (ORC_EV_OBJECTIVE | ORC_EV_GRADIENT)) {
        double *Qx = s->tmp_tot; /* n <= total */
        for (int j = 0; j < n; j++) Qx[j] = 0.0;
        for (int j = 0; j < n; j++)
            for (int k = s->Wp[j]; k < s->Wp[j + 1]; k++) {
                int i = s->Wi[k];
                Qx[i] += s->lqW[k] * x[j];
                if (i != j) Qx[j] += s->lqW[k] * x[i];
            }
        if (flags & ORC_EV_OBJECTIVE) {
            double f = 0.0;
            for (int j = 0; j < n; j++) f += x[j] * (0.5 * Qx[j] + s->lqq[j]);
            o->objective[0] = f;
        }
        if (flags & ORC_EV_GRADIENT)
            for (int j = 0; j < n; j++) o->gradient[j] = Qx[j] + s->lqq[j];
    }
    if (flags & ORC_EV_EQUALITY)
        for (int i = 0; i < m; i++) {
            double a = s->lqg0[i];
            for (int k = s->Grp[i]; k < s->Grp[i + 1]; k++) a += s->lqG[s->Gsrc[k]] * x[s->Gcj[k]];
            o->equality[i] = a;
        }
    if (flags & ORC_EV_CONE)
        for (int i = 0; i < p; i++) {
            double a = s->lqh0[i];
            for (int k = s->Crp[i]; k < s->Crp[i + 1]; k++) a += s->lqC[s->Csrc[k]] * x[s->Ccj[k]];
            o->cone[i] = a;
        }
    if (flags & ORC_EV_EQUALITY_DUAL_GRAD)
        for (int j = 0; j < n; j++) {
            double a = 0.0;
            for (int k = s->Gp[j]; k < s->Gp[j + 1]; k++) a += s->lqG[k] * y[s->Gi[k]];
            o->eq_dual_grad[j] = a;
        }
    if (flags & ORC_EV_CONE_DUAL_GRAD)
        for (int j = 0; j < n; j++) {
            double a = 0.0;
            for (int k = s->Cp[j]; k < s->Cp[j + 1]; k++) a += s->lqC[k] * z[s->Ci[k]];
            o->cone_dual_grad[j] = a;
        }
    if (flags & ORC_EV_HESSIAN) memcpy(o->W_val, s->lqW, sizeof(double) * (size_t)s->nnzW);
    if (flags & ORC_EV_EQUALITY_JAC) memcpy(o->G_val, s->lqG, sizeof(double) * (size_t)s->nnzG);
    if (flags & ORC_EV_CONE_JAC) memcpy(o->C_val, s->lqC, sizeof(double) * (size_t)s->nnzC);
}

static double *dcopy(const double *a, size_t k)
{
    double *r = dalloc(k);
    if (k) memcpy(r, a, k * sizeof(double));
    return r;
}

void orc_solver_set_lq(orc_solver *s, const double *W, const double *G, const double *C, const double *q,
                       const double *g0, const double *h0)
{
    free(s->lqW); free(s->lqG); free(s->lqC); free(s->lqq); free(s->lqg0); free(s->lqh0);
    s->lqW = dcopy(W, (size_t)s->nnzW); s->lqG = dcopy(G, (size_t)s->nnzG); s->lqC = dcopy(C, (size_t)s->nnzC);
    s->lqq = dcopy(q, (size_t)s->n); s->lqg0 = dcopy(g0, (size_t)s->m); s->lqh0 = dcopy(h0, (size_t)s->p);
    s->fn = lq_eval;
    s->user = s;
}

/* -------------------------------------------------------------------- evaluate!  evaluate.jl:1-124 */
void orc_evaluate(orc_solver *s, int flags, int at_candidate)
{
    const double *w = at_candidate ? s->candidate : s->solution;
    /* the reference passes (x, y, z) of the given point: evaluate.jl:24-27 */
    s->fn(s->user, flags, w, w + s->oy, w + s->oz, &s->pd);
}

/* -------------------------------------------------------------------- cones (cones/{nonnegative,second_order,cone}.jl) */
static double dotv(const double *a, const double *b, int k)
{
    double r = 0.0;
    for (int i = 0; i < k; i++) r += a[i] * b[i];
    return r;
}

void orc_cone(orc_solver *s, int at_candidate, int barrier, int barrier_gradient, int product, int jacobian,
              int target)
{ /* cone!  cones/cone.jl:71-106 */
    const double *w = at_candidate ? s->candidate : s->solution;
    const double *sv = w + s->os, *tv = w + s->ot;
    int p = s->p;
    if (barrier) { /* cone_barrier cone.jl:7-25: nonnegative_barrier = sum(log.(x)), second_order_barrier :13 */
        double phi = 0.0;
        if (s->q_nn > 0) {
            double a = 0.0;
            for (int i = 0; i < s->q_nn; i++) a += log(sv[i]);
            phi += a;
        }
        for (int k = 0; k < s->nsoc; k++) {
            int d = s->soc_dims[k];
            if (d > 0) {
                const double *x = sv + s->soc_off[k];
                phi += 0.5 * log(x[0] * x[0] - dotv(x + 1, x + 1, d - 1));
            }
        }
        s->barrier[0] = phi;
    }
    if (barrier_gradient) { /* nonnegative.jl:12, second_order.jl:14 */
        for (int i = 0; i < s->q_nn; i++) s->barrier_gradient[i] = 1.0 / sv[i];
        for (int k = 0; k < s->nsoc; k++) {
            int d = s->soc_dims[k];
            if (d > 0) {
                const double *x = sv + s->soc_off[k];
                double *g = s->barrier_gradient + s->soc_off[k];
                double c = 1.0 / (x[0] * x[0] - dotv(x + 1, x + 1, d - 1));
                g[0] = c * x[0];
                for (int i = 1; i < d; i++) g[i] = c * (-x[i]);
            }
        }
    }
    if (product && p > 0) { /* nonnegative.jl:15, second_order.jl:17 */
        for (int i = 0; i < s->q_nn; i++) s->cone_product[i] = sv[i] * tv[i];
        for (int k = 0; k < s->nsoc; k++) {
            int d = s->soc_dims[k];
            if (d > 0) {
                const double *a = sv + s->soc_off[k], *b = tv + s->soc_off[k];
                double *o = s->cone_product + s->soc_off[k];
                o[0] = dotv(a, b, d);
                for (int i = 1; i < d; i++) o[i] = a[0] * b[i] + b[0] * a[i];
            }
        }
    }
    if (jacobian && p > 0) { /* product_jacobian(s,t) = arrow(t) -> primal; (t,s) = arrow(s) -> dual; cone.jl:91-102 */
        for (int i = 0; i < s->q_nn; i++) { s->jac_primal_nn[i] = tv[i]; s->jac_dual_nn[i] = sv[i]; }
        for (int k = 0; k < s->nsoc; k++) {
            int d = s->soc_dims[k];
            const double *a = sv + s->soc_off[k], *b = tv + s->soc_off[k];
            double *P = s->jac_primal_soc + s->soc_blk[k], *Dm = s->jac_dual_soc + s->soc_blk[k];
            for (int i = 0; i < d * d; i++) { P[i] = 0.0; Dm[i] = 0.0; }
            for (int i = 0; i < d; i++) { P[i + i * d] = b[0]; Dm[i + i * d] = a[0]; }
            for (int i = 1; i < d; i++) {
                P[0 + i * d] = b[i]; P[i + 0 * d] = b[i];
                Dm[0 + i * d] = a[i]; Dm[i + 0 * d] = a[i];
            }
        }
    }
    if (target && p > 0) { /* nonnegative.jl:26, second_order.jl:42 */
        for (int i = 0; i < s->q_nn; i++) s->cone_target[i] = 1.0;
        for (int k = 0; k < s->nsoc; k++)
            for (int i = 0; i < s->soc_dims[k]; i++) s->cone_target[s->soc_off[k] + i] = (i == 0) ? 1.0 : 0.0;
    }
}

/* second_order_vector_inverse(u, x), second_order.jl:50-61 (operation order kept) */
static void soc_vector_inverse(int d, const double *u, const double *x, double *out)
{
    double u1 = u[0];
    double alpha = -1.0 / (u1 * u1) * dotv(u + 1, u + 1, d - 1);
    double beta = 1.0 / (1.0 + alpha);
    /* us = u[2:end] / u[1] */
    double acc = 0.0;
    for (int i = 1; i < d; i++) acc += (u[i] / u1) * x[i];
    double x0_1 = x[0] - acc;                       /* x0 = x - [us' x[2:end]; 0] */
    double acc2 = 0.0;
    for (int i = 1; i < d; i++) {
        double x1_i = x[i] - beta * ((u[i] / u1) * x0_1); /* x1 = x - beta [0; us x0[1]] */
        out[i] = x1_i;
        acc2 += (u[i] / u1) * x1_i;
    }
    double x2_1 = x[0] - acc2;                      /* x2 = x1 - [us' x1[2:end]; 0] */
    out[0] = 1.0 / u1 * x2_1;
    for (int i = 1; i < d; i++) out[i] = 1.0 / u1 * out[i];
}

/* -------------------------------------------------------------------- residual!  residual.jl:1-51 */
void orc_residual_eval(orc_solver *s)
{
    int n = s->n, m = s->m, p = s->p;
    const double *w = s->solution;
    const double *r = w + s->oxr, *sv = w + s->os, *y = w + s->oy, *z = w + s->oz, *t = w + s->ot;
    double *res = s->residual;
    for (int i = 0; i < s->total; i++) res[i] = 0.0;
    for (int i = 0; i < n; i++) {
        res[i] = s->pd.gradient[i];
        res[i] += s->pd.eq_dual_grad[i];
        res[i] += s->pd.cone_dual_grad[i];
    }
    for (int i = 0; i < m; i++) res[s->oxr + i] = s->dual[i] + RHO * r[i] - y[i];
    for (int i = 0; i < p; i++) res[s->os + i] = -z[i] - t[i];
    for (int i = 0; i < m; i++) res[s->oy + i] = s->pd.equality[i] - r[i];
    for (int i = 0; i < p; i++) res[s->oz + i] = s->pd.cone[i] - sv[i];
    for (int i = 0; i < p; i++) res[s->ot + i] = s->cone_product[i] - KAPPA * s->cone_target[i];
}

/* -------------------------------------------------------------------- residual_jacobian_variables!  :1-108 */
void orc_residual_jacobian_variables(orc_solver *s)
{
    for (int k = 0; k < s->nnzW; k++) s->Jxx[k] = s->pd.W_val[k];
    for (int j = 0; j < s->n; j++) s->Jxx[s->Wdiag[j]] += EPSP;                       /* :83-85 */
    s->Jrr = RHO; s->Jrr += EPSP;                                                     /* :66-68, :87-89 */
    s->Jss = 0.0; s->Jss += EPSP;                                                     /* :91-93 */
    s->Jyy = 0.0; s->Jyy -= EPSD;                                                     /* :95-97 */
    s->Jzz = 0.0; s->Jzz -= EPSD;                                                     /* :99-101 */
    for (int i = 0; i < s->q_nn; i++) {                                               /* :71-74, :103-105 */
        s->Jts_nn[i] = s->jac_primal_nn[i];
        s->Jtt_nn[i] = s->jac_dual_nn[i] - EPSD;
    }
    for (int k = 0; k < s->nsoc; k++) {                                               /* :77-86 */
        int d = s->soc_dims[k];
        const double *P = s->jac_primal_soc + s->soc_blk[k], *Dm = s->jac_dual_soc + s->soc_blk[k];
        double *Ts = s->Jts_soc + s->soc_blk[k], *Tt = s->Jtt_soc + s->soc_blk[k];
        for (int i = 0; i < d * d; i++) { Ts[i] = P[i]; Tt[i] = Dm[i]; }
        for (int i = 0; i < d; i++) Tt[i + i * d] -= EPSD;
    }
}

/* -------------------------------------------------------------------- residual_jacobian_variables_symmetric!  :110-167 */
void orc_residual_jacobian_variables_symmetric(orc_solver *s)
{
    for (int k = 0; k < s->nnzK; k++) s->Kx[k] = 0.0;
    for (int k = 0; k < s->nnzW; k++) s->Kx[s->KfromW[k]] = s->Jxx[k];                 /* :115-119 */
    for (int k = 0; k < s->nnzG; k++) s->Kx[s->KfromG[k]] = s->pd.G_val[k];            /* :122-127 */
    for (int i = 0; i < s->m; i++) s->Kx[s->Kydiag[i]] = -1.0 / s->Jrr + s->Jyy;       /* :130-132 */
    for (int k = 0; k < s->nnzC; k++) s->Kx[s->KfromC[k]] = s->pd.C_val[k];            /* :135-140 */
    for (int i = 0; i < s->q_nn; i++) {                                                /* :143-149 */
        double Sb = s->Jtt_nn[i], Ti = s->Jts_nn[i], Pi = s->Jss, Di = s->Jzz;
        s->Kx[s->Kzdiag[i]] += -1.0 * Sb / (Ti + Sb * Pi) + Di;
    }
    int tri = 0;
    double u[64], col[64], out[64];
    for (int k = 0; k < s->nsoc; k++) {                                                /* :152-164 */
        int d = s->soc_dims[k];
        if (d <= 0) continue;
        const double *Cs = s->Jts_soc + s->soc_blk[k], *Ct = s->Jtt_soc + s->soc_blk[k];
        double *uu = d <= 64 ? u : dalloc((size_t)d), *cc = d <= 64 ? col : dalloc((size_t)d),
               *oo = d <= 64 ? out : dalloc((size_t)d);
        for (int j = 0; j < d; j++) uu[j] = Cs[0 + j * d] + Ct[0 + j * d] * s->Jss; /* first row of Cs + Ct*P, second_order.jl:63-65 */
        for (int j = 0; j < d; j++) { /* column j of the block; only rows i <= j survive triu! (linear_solver.jl:23) */
            for (int i = 0; i < d; i++) cc[i] = Ct[i + j * d];
            soc_vector_inverse(d, uu, cc, oo);
            for (int i = 0; i <= j; i++) {
                double v = 0.0;
                v -= oo[i];
                if (i == j) v += s->Jzz; /* += D, :162 */
                s->Kx[s->Kzblk[tri++]] = v;
            }
        }
        if (d > 64) { free(uu); free(cc); free(oo); }
    }
}

/* -------------------------------------------------------------------- factorize! + compute_inertia!  linear_solver.jl:19-44 */
int orc_factorize(orc_solver *s)
{
    int pos = orc_qdldl_refactor(s->ldl, s->Kx);
    const double *D = orc_qdldl_D(s->ldl);
    int neg = 0, zero = 0;
    for (int i = 0; i < s->N; i++) {
        if (D[i] <= 0.0) neg++;
        if (D[i] == 0.0) zero++;
    }
    s->inertia[0] = pos; s->inertia[1] = neg; s->inertia[2] = zero;
    return pos;
}

static int inertia_ok(orc_solver *s)
{ /* inertia.jl:7-11 */
    return s->inertia[0] == s->n && s->inertia[1] == s->m + s->p && s->inertia[2] == 0;
}

static void factorize_regularized(orc_solver *s)
{ /* inertia.jl:13-28 */
    orc_residual_jacobian_variables(s);
    orc_residual_jacobian_variables_symmetric(s);
    orc_factorize(s);
    s->stats[0]++;
}

int orc_inertia_correction(orc_solver *s)
{ /* inertia.jl:30-79 */
    const orc_options *o = &s->opt;
    s->stats[0] = 0;
    EPSP = o->primal_regularization_initial;
    EPSD = o->dual_regularization_initial;
    factorize_regularized(s); /* IC-1 */
    if (inertia_ok(s)) return 0;
    if (s->inertia[2] != 0) EPSD = o->dual_regularization * pow(KAPPA, o->dual_regularization_exponent); /* IC-2 */
    /* IC-3: the reference tests `primal_regularization_last == 0.0` on a Vector, which is always false (:48) */
    {
        double v = o->scaling_regularization_last * EPSP_LAST;
        EPSP = o->min_regularization > v ? o->min_regularization : v;
    }
    while (!inertia_ok(s)) {
        factorize_regularized(s); /* IC-4 */
        if (inertia_ok(s)) break;
        if (EPSP_LAST == 0.0) EPSP = o->scaling_regularization_initial * EPSP; /* IC-5 */
        else EPSP = o->scaling_regularization * EPSP;
        if (EPSP > o->max_regularization) return 1; /* IC-6: error("inertia correction failure") */
    }
    EPSP_LAST = EPSP; /* :76 */
    return 0;
}

/* -------------------------------------------------------------------- residual_symmetric!  residual.jl:53-101 */
void orc_residual_symmetric(orc_solver *s, const double *res)
{
    int n = s->n, m = s->m, p = s->p;
    const double *rx = res, *rr = res + s->oxr, *rs = res + s->os, *ry = res + s->oy, *rz = res + s->oz,
                 *rt = res + s->ot;
    double *o = s->residual_symmetric;
    for (int i = 0; i < s->N; i++) o[i] = 0.0;
    for (int i = 0; i < n; i++) o[i] = rx[i];
    for (int i = 0; i < m; i++) o[n + i] = ry[i];
    for (int i = 0; i < p; i++) o[n + m + i] = rz[i];
    for (int i = 0; i < m; i++) o[n + i] += rr[i] / s->Jrr;
    for (int i = 0; i < s->q_nn; i++) {
        double Sb = s->Jtt_nn[i], Ti = s->Jts_nn[i], Pi = s->Jss;
        o[n + m + i] += (rt[i] + Sb * rs[i]) / (Ti + Sb * Pi);
    }
    double ub[64], vb[64], ob[64];
    for (int k = 0; k < s->nsoc; k++) {
        int d = s->soc_dims[k], c0 = s->soc_off[k];
        if (d <= 0) continue;
        const double *Cs = s->Jts_soc + s->soc_blk[k], *Ct = s->Jtt_soc + s->soc_blk[k];
        double *u = d <= 64 ? ub : dalloc((size_t)d), *v = d <= 64 ? vb : dalloc((size_t)d),
               *oo = d <= 64 ? ob : dalloc((size_t)d);
        for (int j = 0; j < d; j++) u[j] = Cs[0 + j * d] + Ct[0 + j * d] * s->Jss;
        for (int i = 0; i < d; i++) { /* Ct * rs_soc + rt_soc */
            double a = 0.0;
            for (int j = 0; j < d; j++) a += Ct[i + j * d] * rs[c0 + j];
            v[i] = a + rt[c0 + i];
        }
        soc_vector_inverse(d, u, v, oo);
        for (int i = 0; i < d; i++) o[n + m + c0 + i] += oo[i];
        if (d > 64) { free(u); free(v); free(oo); }
    }
}

/* -------------------------------------------------------------------- search_direction_symmetric!  search_direction.jl:25-104 */
void orc_search_direction_symmetric(orc_solver *s, double *step, const double *res, int factorize)
{
    int n = s->n, m = s->m, p = s->p;
    orc_residual_symmetric(s, res);
    if (factorize) orc_factorize(s); /* linear_solve!(fact=true), linear_solver.jl:56 -- same matrix, a no-op numerically */
    memcpy(s->step_symmetric, s->residual_symmetric, sizeof(double) * (size_t)s->N);
    orc_qdldl_solve(s->ldl, s->step_symmetric);
    const double *dx = s->step_symmetric, *dy = s->step_symmetric + n, *dz = s->step_symmetric + n + m;
    for (int i = 0; i < n; i++) step[i] = dx[i];
    for (int i = 0; i < m; i++) step[s->oy + i] = dy[i];
    for (int i = 0; i < p; i++) step[s->oz + i] = dz[i];
    double *dr = step + s->oxr, *ds = step + s->os, *dt = step + s->ot;
    const double *rr = res + s->oxr, *rs = res + s->os, *rt = res + s->ot;
    for (int i = 0; i < m; i++) dr[i] = (rr[i] + dy[i]) / s->Jrr;
    for (int i = 0; i < s->q_nn; i++) {
        double Sb = s->Jtt_nn[i], Ti = s->Jts_nn[i], Pi = s->Jss;
        ds[i] = (rt[i] + Sb * (rs[i] + dz[i])) / (Ti + Sb * Pi);
        dt[i] = (rt[i] - Ti * ds[i]) / Sb;
    }
    double ub[64], vb[64], ob[64];
    for (int k = 0; k < s->nsoc; k++) {
        int d = s->soc_dims[k], c0 = s->soc_off[k];
        if (d <= 0) continue;
        const double *Cs = s->Jts_soc + s->soc_blk[k], *Ct = s->Jtt_soc + s->soc_blk[k];
        double *u = d <= 64 ? ub : dalloc((size_t)d), *v = d <= 64 ? vb : dalloc((size_t)d),
               *oo = d <= 64 ? ob : dalloc((size_t)d);
        /* ds = inv(Cs + Ct*P, rt + Ct*(rs + dz)) */
        for (int j = 0; j < d; j++) u[j] = Cs[0 + j * d] + Ct[0 + j * d] * s->Jss;
        for (int i = 0; i < d; i++) {
            double a = 0.0;
            for (int j = 0; j < d; j++) a += Ct[i + j * d] * (rs[c0 + j] + dz[c0 + j]);
            v[i] = rt[c0 + i] + a;
        }
        soc_vector_inverse(d, u, v, oo);
        for (int i = 0; i < d; i++) ds[c0 + i] = oo[i];
        /* dt = inv(Ct, rt - Cs*ds) */
        for (int j = 0; j < d; j++) u[j] = Ct[0 + j * d];
        for (int i = 0; i < d; i++) {
            double a = 0.0;
            for (int j = 0; j < d; j++) a += Cs[i + j * d] * ds[c0 + j];
            v[i] = rt[c0 + i] - a;
        }
        soc_vector_inverse(d, u, v, oo);
        for (int i = 0; i < d; i++) dt[c0 + i] = oo[i];
        if (d > 64) { free(u); free(v); free(oo); }
    }
}

/* -------------------------------------------------------------------- J * v (mul!), blocks per SURVEY.md section 3.3 */
void orc_jacobian_times(orc_solver *s, const double *v, double *out)
{
    int n = s->n, m = s->m, p = s->p;
    const double *vx = v, *vr = v + s->oxr, *vs = v + s->os, *vy = v + s->oy, *vz = v + s->oz, *vt = v + s->ot;
    double *ox = out, *orr = out + s->oxr, *oss = out + s->os, *oyy = out + s->oy, *ozz = out + s->oz, *ott = out + s->ot;
    for (int i = 0; i < n; i++) ox[i] = 0.0;
    for (int j = 0; j < n; j++) {
        for (int k = s->Wp[j]; k < s->Wp[j + 1]; k++) {
            int i = s->Wi[k];
            ox[i] += s->Jxx[k] * vx[j];
            if (i != j) ox[j] += s->Jxx[k] * vx[i];
        }
        for (int k = s->Gp[j]; k < s->Gp[j + 1]; k++) ox[j] += s->pd.G_val[k] * vy[s->Gi[k]];
        for (int k = s->Cp[j]; k < s->Cp[j + 1]; k++) ox[j] += s->pd.C_val[k] * vz[s->Ci[k]];
    }
    for (int i = 0; i < m; i++) orr[i] = s->Jrr * vr[i] - vy[i];
    for (int i = 0; i < p; i++) oss[i] = s->Jss * vs[i] - vz[i] - vt[i];
    for (int i = 0; i < m; i++) {
        double a = 0.0;
        for (int k = s->Grp[i]; k < s->Grp[i + 1]; k++) a += s->pd.G_val[s->Gsrc[k]] * vx[s->Gcj[k]];
        oyy[i] = a - vr[i] + s->Jyy * vy[i];
    }
    for (int i = 0; i < p; i++) {
        double a = 0.0;
        for (int k = s->Crp[i]; k < s->Crp[i + 1]; k++) a += s->pd.C_val[s->Csrc[k]] * vx[s->Ccj[k]];
        ozz[i] = a - vs[i] + s->Jzz * vz[i];
    }
    for (int i = 0; i < s->q_nn; i++) ott[i] = s->Jts_nn[i] * vs[i] + s->Jtt_nn[i] * vt[i];
    for (int k = 0; k < s->nsoc; k++) {
        int d = s->soc_dims[k], c0 = s->soc_off[k];
        const double *Cs = s->Jts_soc + s->soc_blk[k], *Ct = s->Jtt_soc + s->soc_blk[k];
        for (int i = 0; i < d; i++) {
            double a = 0.0;
            for (int j = 0; j < d; j++) a += Cs[i + j * d] * vs[c0 + j] + Ct[i + j * d] * vt[c0 + j];
            ott[c0 + i] = a;
        }
    }
}

void orc_dense_jacobian(orc_solver *s, double *J)
{
    int tot = s->total;
    double *e = dalloc((size_t)tot);
    for (int j = 0; j < tot; j++) {
        e[j] = 1.0;
        orc_jacobian_times(s, e, J + (size_t)j * (size_t)tot);
        e[j] = 0.0;
    }
    free(e);
}

int orc_jacobian_coo(orc_solver *s, int *rows, int *cols, double *vals)
{
    int n = s->n, m = s->m, p = s->p, cnt = 0;
#define EMIT(i, j, v) do { if (rows) { rows[cnt] = (i); cols[cnt] = (j); vals[cnt] = (v); } cnt++; } while (0)
    for (int j = 0; j < n; j++) {
        for (int k = s->Wp[j]; k < s->Wp[j + 1]; k++) {
            int i = s->Wi[k];
            EMIT(i, j, s->Jxx[k]);
            if (i != j) EMIT(j, i, s->Jxx[k]);
        }
        for (int k = s->Gp[j]; k < s->Gp[j + 1]; k++) { EMIT(s->oy + s->Gi[k], j, s->pd.G_val[k]); EMIT(j, s->oy + s->Gi[k], s->pd.G_val[k]); }
        for (int k = s->Cp[j]; k < s->Cp[j + 1]; k++) { EMIT(s->oz + s->Ci[k], j, s->pd.C_val[k]); EMIT(j, s->oz + s->Ci[k], s->pd.C_val[k]); }
    }
    for (int i = 0; i < m; i++) {
        EMIT(s->oxr + i, s->oxr + i, s->Jrr); EMIT(s->oxr + i, s->oy + i, -1.0); EMIT(s->oy + i, s->oxr + i, -1.0);
        EMIT(s->oy + i, s->oy + i, s->Jyy);
    }
    for (int i = 0; i < p; i++) {
        EMIT(s->os + i, s->os + i, s->Jss); EMIT(s->os + i, s->oz + i, -1.0); EMIT(s->oz + i, s->os + i, -1.0);
        EMIT(s->os + i, s->ot + i, -1.0); EMIT(s->oz + i, s->oz + i, s->Jzz);
    }
    for (int i = 0; i < s->q_nn; i++) { EMIT(s->ot + i, s->os + i, s->Jts_nn[i]); EMIT(s->ot + i, s->ot + i, s->Jtt_nn[i]); }
    for (int k = 0; k < s->nsoc; k++) {
        int d = s->soc_dims[k], c0 = s->soc_off[k];
        const double *Cs = s->Jts_soc + s->soc_blk[k], *Ct = s->Jtt_soc + s->soc_blk[k];
        for (int j = 0; j < d; j++)
            for (int i = 0; i < d; i++) {
                if (i == j || i == 0 || j == 0) { /* arrow pattern, cones/codegen.jl:15 sparse(pa) */
                    EMIT(s->ot + c0 + i, s->os + c0 + j, Cs[i + j * d]);
                    EMIT(s->ot + c0 + i, s->ot + c0 + j, Ct[i + j * d]);
                }
            }
    }
#undef EMIT
    return cnt;
}

void orc_solver_set_lu_fallback(orc_solver *s, orc_lu_fn fn, void *user) { s->lu_fn = fn; s->lu_user = user; }

/* dense partial-pivoting LU stand-in for `J \ R` on small systems */
static int dense_lu_solve(orc_solver *s, const double *rhs, double *x)
{
    int t = s->total;
    double *A = dalloc((size_t)t * (size_t)t);
    int *piv = ialloc((size_t)t);
    orc_dense_jacobian(s, A);
    for (int i = 0; i < t; i++) x[i] = rhs[i];
    for (int k = 0; k < t; k++) {
        int pk = k;
        double best = fabs(A[k + (size_t)k * t]);
        for (int i = k + 1; i < t; i++) if (fabs(A[i + (size_t)k * t]) > best) { best = fabs(A[i + (size_t)k * t]); pk = i; }
        piv[k] = pk;
        if (best == 0.0) { free(A); free(piv); return 1; }
        if (pk != k) {
            for (int j = 0; j < t; j++) { double tmp = A[k + (size_t)j * t]; A[k + (size_t)j * t] = A[pk + (size_t)j * t]; A[pk + (size_t)j * t] = tmp; }
            double tmp = x[k]; x[k] = x[pk]; x[pk] = tmp;
        }
        double inv = 1.0 / A[k + (size_t)k * t];
        for (int i = k + 1; i < t; i++) A[i + (size_t)k * t] *= inv;
        for (int j = k + 1; j < t; j++) {
            double akj = A[k + (size_t)j * t];
            if (akj != 0.0) for (int i = k + 1; i < t; i++) A[i + (size_t)j * t] -= A[i + (size_t)k * t] * akj;
        }
        for (int i = k + 1; i < t; i++) x[i] -= A[i + (size_t)k * t] * x[k];
    }
    for (int k = t - 1; k >= 0; k--) {
        x[k] /= A[k + (size_t)k * t];
        for (int i = 0; i < k; i++) x[i] -= A[i + (size_t)k * t] * x[k];
    }
    free(A); free(piv);
    return 0;
}

void orc_dense_symmetric(orc_solver *s, double *K)
{
    int N = s->N;
    for (size_t i = 0; i < (size_t)N * (size_t)N; i++) K[i] = 0.0;
    for (int j = 0; j < N; j++)
        for (int k = s->Kp[j]; k < s->Kp[j + 1]; k++) {
            int i = s->Ki[k];
            K[(size_t)i + (size_t)j * N] = s->Kx[k];
            K[(size_t)j + (size_t)i * N] = s->Kx[k];
        }
}

static double norm_inf(const double *a, int k)
{
    double r = 0.0;
    for (int i = 0; i < k; i++) { double v = fabs(a[i]); if (v > r) r = v; }
    return r;
}
static double norm_one(const double *a, int k)
{
    double r = 0.0;
    for (int i = 0; i < k; i++) r += fabs(a[i]);
    return r;
}

/* -------------------------------------------------------------------- iterative_refinement!  iterative_refinement.jl:1-53 */
int orc_iterative_refinement(orc_solver *s, double *step)
{
    const orc_options *o = &s->opt;
    int tot = s->total;
    double *err = s->residual_error, *corr = s->step_correction, *Jv = s->tmp_tot;
    for (int i = 0; i < tot; i++) { corr[i] = 0.0; err[i] = 0.0; }
    int iteration = 0;
    orc_jacobian_times(s, step, Jv);
    for (int i = 0; i < tot; i++) err[i] = s->residual[i] - Jv[i];
    double rn = norm_inf(err, tot), rn0 = rn;
    s->stats[1] = 0;
    while (iteration <= o->max_iterative_refinement) {
        if (rn <= o->iterative_refinement_tolerance && iteration >= o->min_iterative_refinement) return 1;
        /* search_direction_symmetric!(correction, residual_error, ...) with default update=true: re-factors (:21-25) */
        orc_search_direction_symmetric(s, corr, err, s->opt.reference_schedule);
        for (int i = 0; i < tot; i++) step[i] += corr[i];
        orc_jacobian_times(s, step, Jv);
        for (int i = 0; i < tot; i++) err[i] = s->residual[i] - Jv[i];
        rn = norm_inf(err, tot);
        iteration++;
        s->stats[1] = iteration;
    }
    return rn <= rn0 ? 1 : 0;
}

/* -------------------------------------------------------------------- search_direction!  search_direction.jl:1-23 */
int orc_search_direction(orc_solver *s)
{
    if (orc_inertia_correction(s)) return 1;
    orc_search_direction_symmetric(s, s->step, s->residual, s->opt.reference_schedule);
    s->stats[2] = 1;
    if (s->opt.iterative_refinement) {
        int ok = orc_iterative_refinement(s, s->step);
        s->stats[2] = ok;
        s->stats[9] = 0;
        if (!ok) { /* search_direction_nonsymmetric!(step, J, R, lu; update_factorization=false): step = J \ R (:22,:113) */
            int rc;
            if (s->lu_fn) rc = s->lu_fn(s->lu_user, s->total, s->residual, s->step);
            else if (s->total <= 2500) rc = dense_lu_solve(s, s->residual, s->step);
            else rc = 1;
            if (rc) return 2;
            s->stats[8]++;
            s->stats[9] = 1;
        }
    }
    return 0;
}

/* -------------------------------------------------------------------- cone_violation  cone.jl:62-68 */
int orc_cone_violation(orc_solver *s, const double *xh, const double *x, double tau)
{
    for (int i = 0; i < s->q_nn; i++) /* nonnegative.jl:29-34 */
        if (xh[i] <= (1.0 - tau) * x[i]) return 1;
    for (int k = 0; k < s->nsoc; k++) { /* second_order.jl:45-47 */
        int d = s->soc_dims[k], c0 = s->soc_off[k];
        if (d <= 0) continue;
        double acc = 0.0;
        for (int i = 1; i < d; i++) {
            double v = xh[c0 + i] - (1.0 - tau) * x[c0 + i];
            acc += v * v;
        }
        if (xh[c0] - (1.0 - tau) * x[c0] <= sqrt(acc)) return 1;
    }
    return 0;
}

/* cone line search, solve.jl:190-221.  Leaves step sizes in stats[3], stats[4] as halving counts. */
static double g_step_size, g_step_size_t; /* not thread safe: oracle is single-threaded like the reference */
int orc_cone_search(orc_solver *s)
{
    const orc_options *o = &s->opt;
    int p = s->p;
    const double *sv = s->solution + s->os, *tv = s->solution + s->ot;
    double *sh = s->candidate + s->os, *th = s->candidate + s->ot;
    const double *ds = s->step + s->os, *dt = s->step + s->ot;
    double a = 1.0, at = 1.0;
    for (int i = 0; i < p; i++) sh[i] = sv[i] - a * ds[i];
    for (int i = 0; i < p; i++) th[i] = tv[i] - at * dt[i];
    int it = 0;
    while (orc_cone_violation(s, sh, sv, TAU)) {
        a = o->scaling_line_search * a;
        for (int i = 0; i < p; i++) sh[i] = sv[i] - a * ds[i];
        it++;
        if (it > o->max_cone_line_search) return 3; /* error("cone search failure") */
    }
    s->stats[3] = it;
    it = 0;
    while (orc_cone_violation(s, th, tv, TAU)) {
        at = o->scaling_line_search * at;
        for (int i = 0; i < p; i++) th[i] = tv[i] - at * dt[i];
        it++;
        if (it > o->max_cone_line_search) return 3;
    }
    s->stats[4] = it;
    g_step_size = a;
    g_step_size_t = at;
    return 0;
}

/* -------------------------------------------------------------------- merit.jl, constraint_violation.jl, optimality_error.jl */
double orc_merit(orc_solver *s, int at_candidate)
{ /* merit.jl:2-15 */
    const double *r = (at_candidate ? s->candidate : s->solution) + s->oxr;
    double M = 0.0;
    M += s->objective[0];
    M += dotv(s->dual, r, s->m) + 0.5 * RHO * dotv(r, r, s->m);
    M -= KAPPA * s->barrier[0];
    return M;
}

void orc_merit_gradient_eval(orc_solver *s)
{ /* merit.jl:17-31 */
    const double *r = s->solution + s->oxr;
    for (int i = 0; i < s->n; i++) s->merit_gradient[i] = s->pd.gradient[i];
    for (int i = 0; i < s->m; i++) s->merit_gradient[s->n + i] = s->dual[i] + RHO * r[i];
    for (int i = 0; i < s->p; i++) s->merit_gradient[s->n + s->m + i] = -1.0 * KAPPA * s->barrier_gradient[i];
}

double orc_constraint_violation(orc_solver *s, int at_candidate)
{ /* constraint_violation.jl:1-13, norm_type = 1 */
    const double *w = at_candidate ? s->candidate : s->solution;
    const double *r = w + s->oxr, *sv = w + s->os;
    double *c = s->constraint_violation;
    for (int i = 0; i < s->m; i++) c[i] = s->pd.equality[i] - r[i];
    for (int j = 0; j < s->p; j++) c[s->m + j] = s->pd.cone[j] - sv[j];
    int len = s->m + s->p;
    return norm_one(c, len) / (double)len; /* NaN for m+p == 0, like the reference (0/0) */
}

double orc_optimality_error(orc_solver *s)
{ /* optimality_error.jl:1-27 */
    int m = s->m, p = s->p;
    const double *y = s->solution + s->oy, *z = s->solution + s->oz, *t = s->solution + s->ot;
    double sd = (m + p) > 0 ? fmax(100.0, (norm_one(y, m) + norm_one(z, p)) / (double)(m + p)) / 100.0 : 1.0;
    double sc = p > 0 ? fmax(100.0, norm_one(t, p) / (double)p) / 100.0 : 1.0;
    double a = norm_inf(s->residual, s->n + m + p) / sd; /* residual.primals = (x, r, s) blocks, point.jl:20 */
    double b = norm_inf(s->residual + s->oy, m);
    double c = norm_inf(s->residual + s->oz, p);
    double d = norm_inf(s->residual + s->ot, p) / sc;
    return fmax(fmax(a, b), fmax(c, d));
}

/* -------------------------------------------------------------------- line_search.jl, filter.jl */
static int switching_condition(double step_size, const double *dir, const double *mg, int len, double merit_exponent,
                               double violation, double violation_exponent, double regularization)
{
    double d = dotv(mg, dir, len);
    return d < 0.0 && step_size * pow(-d, merit_exponent) > regularization * pow(violation, violation_exponent);
}
static int sufficient_progress(double v, double vc, double M, double Mc, double vt, double mt, double mach)
{
    return (vc - 10.0 * mach * fabs(v) <= (1.0 - vt) * v) || (Mc - 10.0 * mach * fabs(M) <= M - mt * v);
}
static int armijo(double M, double Mc, const double *mg, const double *dir, int len, double step_size, double at,
                  double mach)
{
    double d = dotv(mg, dir, len);
    return Mc - M - 10.0 * mach * fabs(M) <= at * step_size * d;
}
static void filter_reset(orc_solver *s)
{ /* reset!, filter.jl:37-41 */
    for (int i = 0; i < s->findex; i++) { s->fcache[i].theta = 1.0e8; s->fcache[i].merit = 1.0e8; }
    for (int i = 0; i < s->findex; i++) { s->fpairs[i].theta = 1.0e8; s->fpairs[i].merit = 1.0e8; }
    s->findex = 0;
}
static int check_filter(orc_solver *s, double cv, double merit)
{ /* filter.jl:43-50: loops over every stored pair, including the (1e8,1e8) placeholders */
    for (int i = 0; i < s->opt.max_filter; i++)
        if (!(cv < s->fpairs[i].theta || merit < s->fpairs[i].merit)) return 0;
    return 1;
}
static void augment_filter_pair(orc_solver *s, double cv, double merit)
{ /* filter.jl:52-79 */
    if (s->findex == 0) {
        s->fpairs[0].theta = cv; s->fpairs[0].merit = merit;
        s->findex++;
        return;
    } else if (check_filter(s, cv, merit)) {
        int nf = s->findex;
        for (int i = 0; i < nf; i++) s->fcache[i] = s->fpairs[i];
        for (int i = 0; i < nf; i++) { s->fpairs[i].theta = 1.0e8; s->fpairs[i].merit = 1.0e8; }
        s->findex = 0;
        s->findex++;
        s->fpairs[s->findex - 1].theta = cv; s->fpairs[s->findex - 1].merit = merit;
        for (int i = 0; i < nf; i++)
            if (!(s->fcache[i].theta >= cv && s->fcache[i].merit >= merit)) {
                s->findex++;
                s->fpairs[s->findex - 1] = s->fcache[i];
            }
    }
}

/* -------------------------------------------------------------------- initialize.jl */
int orc_initialize(orc_solver *s, const double *guess)
{
    memcpy(s->solution, guess, sizeof(double) * (size_t)s->n);
    return 0;
}
static void initialize_cone(orc_solver *s, double *x)
{ /* initialize_cone! cone.jl:1-4 */
    for (int i = 0; i < s->q_nn; i++) x[i] = 1.0;
    for (int k = 0; k < s->nsoc; k++)
        for (int i = 0; i < s->soc_dims[k]; i++) x[s->soc_off[k] + i] = (i == 0) ? 1.0 : 0.1;
}

void orc_solve_begin(orc_solver *s)
{ /* solve.jl:8-95 */
    const orc_options *o = &s->opt;
    if (!o->warmstart) {
        orc_evaluate(s, ORC_EV_EQUALITY | ORC_EV_CONE, 0);                    /* initialize_slacks! */
        for (int i = 0; i < s->m; i++) s->solution[s->oxr + i] = s->pd.equality[i];
        initialize_cone(s, s->solution + s->os);
        for (int i = 0; i < s->m; i++) s->solution[s->oy + i] = 0.0;          /* initialize_duals! */
        for (int i = 0; i < s->p; i++) s->solution[s->oz + i] = 0.0;
        initialize_cone(s, s->solution + s->ot);
    }
    KAPPA = o->central_path_initial;
    TAU = fmax(0.99, 1.0 - KAPPA);
    RHO = o->penalty_initial;
    for (int i = 0; i < s->m; i++) s->dual[i] = o->dual_initial;
    orc_evaluate(s, ORC_EV_OBJECTIVE | ORC_EV_EQUALITY | ORC_EV_EQUALITY_JAC | ORC_EV_CONE, 0);
    s->equality_violation = norm_inf(s->pd.equality, s->m);
    s->cone_product_violation = norm_inf(s->cone_product, s->p); /* computed BEFORE cone!(product), solve.jl:86-91 */
    orc_cone(s, 0, 0, 0, 1, 0, 1);
    filter_reset(s);
    s->stats[5] = 1; /* total_iterations */
    s->stats[6] = 1; /* outer iteration j */
    s->stats[7] = 0;
}

void orc_outer_update(orc_solver *s)
{ /* solve.jl:356-368 */
    const orc_options *o = &s->opt;
    KAPPA = fmax(o->residual_tolerance / 10.0, fmin(o->central_path_scaling * KAPPA, pow(KAPPA, o->central_path_exponent)));
    TAU = fmax(0.99, 1.0 - KAPPA);
    for (int i = 0; i < s->m; i++) s->dual[i] = s->dual[i] + RHO * s->solution[s->oxr + i];
    RHO = fmin(fmax(o->penalty_scaling * RHO, 1.0 / KAPPA), o->max_penalty);
    filter_reset(s);
    s->stats[6]++;
}

int orc_newton_iteration(orc_solver *s)
{ /* one pass of the inner loop body, solve.jl:98-350 */
    const orc_options *o = &s->opt;
    int n = s->n, m = s->m, p = s->p;
    orc_evaluate(s, ORC_EV_GRADIENT | ORC_EV_EQUALITY_DUAL_GRAD | ORC_EV_CONE_DUAL_GRAD, 0);
    orc_cone(s, 0, 1, 1, 0, 0, 0);
    double M = orc_merit(s, 0);
    orc_merit_gradient_eval(s);
    orc_residual_eval(s);
    double residual_violation = norm_one(s->residual, s->total) / (double)s->total;
    double optimality_violation = orc_optimality_error(s);
    double slack_violation = fmax(norm_inf(s->residual + s->oy, m), norm_inf(s->residual + s->oz, p));
    if (residual_violation < o->residual_tolerance && slack_violation < o->slack_tolerance &&
        s->equality_violation <= o->equality_tolerance && s->cone_product_violation <= o->complementarity_tolerance)
        return 1;
    else if (optimality_violation <= fmax(o->central_path_update_tolerance * KAPPA, o->optimality_tolerance))
        return 2;
    double theta = orc_constraint_violation(s, 0);
    orc_evaluate(s, ORC_EV_HESSIAN | ORC_EV_EQUALITY_JAC | ORC_EV_CONE_JAC, 0);
    orc_cone(s, 0, 0, 0, 0, 1, 0);
    int sd = orc_search_direction(s);
    if (sd == 1) return -1;
    /* sd == 2: refinement failed and no LU stand-in was available; the unrefined direction is kept (recorded). */
    if (orc_cone_search(s)) return -3;
    double step_size = g_step_size;
    const double *x = s->solution, *r = s->solution + s->oxr, *sv = s->solution + s->os;
    double *xh = s->candidate, *rh = s->candidate + s->oxr, *sh = s->candidate + s->os;
    const double *dx = s->step, *dr = s->step + s->oxr, *ds = s->step + s->os;
    for (int i = 0; i < n; i++) xh[i] = x[i] - step_size * dx[i];
    for (int i = 0; i < m; i++) rh[i] = r[i] - step_size * dr[i];
    orc_evaluate(s, ORC_EV_OBJECTIVE | ORC_EV_EQUALITY | ORC_EV_CONE, 1);
    orc_cone(s, 1, 1, 1, 0, 0, 0);
    double Mh = orc_merit(s, 1);
    double theta_h = orc_constraint_violation(s, 1);
    int residual_iteration = 0;
    int len = n + m + p; /* step.primals */
    while (residual_iteration < o->max_residual_line_search) {
        if (check_filter(s, theta_h, Mh)) {
            if (theta <= o->slack_tolerance &&
                switching_condition(step_size, s->step, s->merit_gradient, len, o->merit_exponent, theta,
                                    o->violation_exponent, 1.0) &&
                armijo(M, Mh, s->merit_gradient, s->step, len, step_size, o->armijo_tolerance, o->machine_tolerance))
                break;
            else if (sufficient_progress(theta, theta_h, M, Mh, o->violation_tolerance, o->merit_tolerance,
                                         o->machine_tolerance))
                break;
        }
        step_size = o->scaling_line_search * step_size;
        for (int i = 0; i < n; i++) xh[i] = x[i] - step_size * dx[i];
        for (int i = 0; i < m; i++) rh[i] = r[i] - step_size * dr[i];
        for (int i = 0; i < p; i++) sh[i] = sv[i] - step_size * ds[i];
        orc_evaluate(s, ORC_EV_OBJECTIVE | ORC_EV_EQUALITY | ORC_EV_CONE, 1);
        orc_cone(s, 1, 1, 1, 0, 0, 0);
        Mh = orc_merit(s, 1);
        theta_h = orc_constraint_violation(s, 1);
        residual_iteration++;
    }
    /* augment_filter!(solver, M, M_hat, merit_gradient, theta, step_size, dp), filter.jl:81-89 */
    if (!switching_condition(step_size, s->step, s->merit_gradient, len, o->merit_exponent, theta,
                             o->violation_exponent, 1.0) ||
        !armijo(M, Mh, s->merit_gradient, s->step, len, step_size, o->armijo_tolerance, o->machine_tolerance))
        augment_filter_pair(s, (1.0 - o->violation_tolerance) * theta, M - o->merit_tolerance * theta);
    /* update, solve.jl:309-326 */
    double *w = s->solution;
    for (int i = 0; i < n; i++) w[i] = xh[i];
    for (int i = 0; i < m; i++) w[s->oxr + i] = rh[i];
    for (int i = 0; i < p; i++) w[s->os + i] = sh[i];
    for (int i = 0; i < m; i++) w[s->oy + i] = w[s->oy + i] - step_size * s->step[s->oy + i];
    for (int i = 0; i < p; i++) w[s->oz + i] = w[s->oz + i] - step_size * s->step[s->oz + i];
    for (int i = 0; i < p; i++) w[s->ot + i] = s->candidate[s->ot + i];
    orc_cone(s, 0, 0, 0, 1, 0, 0);
    s->equality_violation = norm_inf(s->pd.equality, m);
    s->cone_product_violation = norm_inf(s->cone_product, p);
    s->stats[5]++;
    s->tmp_p[0] = step_size; /* last accepted step size, for inspection */
    return 0;
}

int orc_solve(orc_solver *s)
{ /* solve!  solve.jl:8-377 */
    const orc_options *o = &s->opt;
    orc_solve_begin(s);
    for (int j = 1; j <= o->max_outer_iterations; j++) {
        for (int i = 1; i <= o->max_residual_iterations; i++) {
            int rc = orc_newton_iteration(s);
            if (rc == 1) { s->stats[7] = 1; return 1; }
            if (rc == 2) break;
            if (rc < 0) { s->stats[7] = rc; return rc; }
        }
        orc_outer_update(s);
    }
    s->stats[7] = 0;
    return 0;
}
