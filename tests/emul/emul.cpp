// emul.cpp -- TEST INFRASTRUCTURE ONLY: the device headers of calipso_b200/csrc compiled for the host
// (CB200_HOST_EMULATION: parallel-for ranges run sequentially, barriers are no-ops) behind the same C ABI as
// libcalipso_b200.so, so that the numerical logic of the kernels can be checked against the oracle on a machine
// without a GPU.  Never shipped, never loaded by the product package (calipso_b200/_lib.py only loads the CUDA library).
#define CB200_HOST_EMULATION 1
#include <cstring>
#include <string>
#include <vector>

#include "../../calipso_b200/csrc/device_newton.h"
#include "../../calipso_b200/csrc/host_setup.h"
#include "../../include/calipso_b200.h"

using namespace cb200;

static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return -1; }

struct cb200_handle {
    int batch = 0;
    bool generic = false;
    HostProblem hp;
    Symbolic gsym;
    DevProblem P{};
    Options opt{};
    std::vector<std::vector<double>> arr;   // [CB200_NUM_ARRAYS + extras], instance-major
    std::vector<long long> len;
    std::vector<int> istat;
    std::vector<double> scratch;
    long long ksize = 0;
    const Symbolic &sym() const { return generic ? gsym : hp.sym; }
    struct Scatter { ScatterPlan plan; std::vector<double> caches; bool set = false; } scatter[3];
    struct Stage { StagePlan plan; std::vector<double> caches; bool set = false; } stage[5];
};

enum { X_ERR = CB200_NUM_ARRAYS, X_CORR, X_TMP, X_DINV, X_XP, X_FILTER, X_KRYLOV, X_KX, X_LCSR, X_WF, X_GR, X_COUNT };

extern "C" const char *cb200_last_error(void) { return g_err.c_str(); }
extern "C" int cb200_device_count(void) { return 0; }

extern "C" void cb200_options_default(cb200_options *o)
{
    memset(o, 0, sizeof(*o));
    o->max_outer_iterations = 10; o->max_residual_iterations = 100; o->max_residual_line_search = 25;
    o->max_cone_line_search = 25; o->iterative_refinement = 1; o->max_iterative_refinement = 10;
    o->min_iterative_refinement = 1; o->scaling_line_search = 0.5; o->iterative_refinement_tolerance = 1.0e-10;
    o->central_path_initial = 1.0; o->central_path_update_tolerance = 10.0; o->central_path_scaling = 0.2;
    o->central_path_exponent = 1.5; o->penalty_initial = 1.0; o->penalty_scaling = 10.0; o->dual_initial = 0.0;
    o->residual_tolerance = 1.0e-4; o->optimality_tolerance = 1.0e-4; o->slack_tolerance = 1.0e-4;
    o->equality_tolerance = 1.0e-4; o->complementarity_tolerance = 1.0e-4; o->min_regularization = 1.0e-20;
    o->primal_regularization_initial = 1.0e-7; o->dual_regularization_initial = 1.0e-7;
    o->max_regularization = 1.0e40; o->dual_regularization = 1.0e-8; o->dual_regularization_exponent = 0.25;
    o->scaling_regularization_initial = 100.0; o->scaling_regularization = 8.0;
    o->scaling_regularization_last = 1.0 / 3.0; o->max_penalty = 1.0e8; o->violation_tolerance = 1.0e-5;
    o->violation_exponent = 1.1; o->merit_tolerance = 1.0e-5; o->merit_exponent = 2.3; o->armijo_tolerance = 1.0e-4;
    o->machine_tolerance = 1.0e-16; o->max_filter = 1000; o->gmres_restart = 30; o->gmres_max_cycles = 10;
}

static void alloc(cb200_handle *h, int which, long long per)
{
    h->len[which] = per;
    h->arr[which].assign((size_t)std::max<long long>(per * h->batch, 1), 0.0);
}

static Inst inst(cb200_handle *h, int b)
{
    Inst I{};
    auto at = [&](int which) { return h->arr[which].data() + (long long)b * h->len[which]; };
    I.w = at(CB200_POINT); I.cand = at(CB200_CANDIDATE); I.step = at(CB200_STEP); I.res = at(CB200_RESIDUAL);
    I.err = at(X_ERR); I.corr = at(X_CORR); I.tmp = at(X_TMP);
    I.grad = at(CB200_GRADIENT); I.gyx = at(CB200_EQ_DUAL_GRAD); I.hzx = at(CB200_CONE_DUAL_GRAD);
    I.g = at(CB200_EQUALITY); I.h = at(CB200_CONE);
    I.Wv = at(CB200_W_VALUES); I.Gv = at(CB200_G_VALUES); I.Cv = at(CB200_C_VALUES);
    I.prod = at(CB200_CONE_PRODUCT); I.bgrad = at(CB200_BARRIER_GRADIENT); I.lambda = at(CB200_DUAL);
    I.panels = at(CB200_PANELS); I.D = at(CB200_PIVOTS); I.Dinv = at(X_DINV); I.kx = at(X_KX); I.Lcsr = at(X_LCSR); I.prof = nullptr;
    I.Wf = at(X_WF); I.Gr = at(X_GR);
    I.xs = at(CB200_STEP_SYMMETRIC); I.rs = at(CB200_RESIDUAL_SYMMETRIC); I.xp = at(X_XP);
    I.mgrad = at(CB200_MERIT_GRADIENT); I.q = at(CB200_LQ_Q); I.g0 = at(CB200_LQ_G0); I.h0 = at(CB200_LQ_H0);
    I.filter = at(X_FILTER); I.krylov = h->ksize ? at(X_KRYLOV) : nullptr;
    I.scal = at(CB200_SCALARS);
    I.istat = h->istat.data() + (long long)b * I_COUNT;
    return I;
}

static const int BIG_TASK_THRESHOLD = 3000;

extern "C" cb200_handle *cb200_create(int batch, int n, int m, int p, int q_nn, int nsoc, const int *soc_dims,
                                      const int *Wp, const int *Wi, const int *Gp, const int *Gi, const int *Cp,
                                      const int *Ci, const int *perm, const cb200_options *options, int device)
{
    if (batch <= 0) { fail("invalid dimensions"); return nullptr; }
    cb200_handle *h = new cb200_handle();
    h->batch = batch;
    cb200_options d;
    cb200_options_default(&d);
    memcpy(&h->opt, options ? options : &d, sizeof(Options));
    std::string msg = h->hp.build(n, m, p, q_nn, nsoc, soc_dims, Wp, Wi, Gp, Gi, Cp, Ci, perm, BIG_TASK_THRESHOLD);
    if (!msg.empty()) { fail(msg); delete h; return nullptr; }
    fill_problem(h->P, h->hp, [](const auto &v) { return v.data(); });
    const DevProblem &P = h->P;
    h->arr.resize(X_COUNT);
    h->len.assign(X_COUNT, 0);
    const long long T = P.total, N = P.N;
    for (int w : {(int)CB200_POINT, (int)CB200_CANDIDATE, (int)CB200_STEP, (int)CB200_RESIDUAL, (int)X_ERR, (int)X_CORR, (int)X_TMP}) alloc(h, w, T);
    for (int w : {(int)CB200_GRADIENT, (int)CB200_EQ_DUAL_GRAD, (int)CB200_CONE_DUAL_GRAD, (int)CB200_LQ_Q}) alloc(h, w, n);
    for (int w : {(int)CB200_EQUALITY, (int)CB200_DUAL, (int)CB200_LQ_G0}) alloc(h, w, m);
    for (int w : {(int)CB200_CONE, (int)CB200_CONE_PRODUCT, (int)CB200_BARRIER_GRADIENT, (int)CB200_LQ_H0}) alloc(h, w, p);
    alloc(h, CB200_W_VALUES, P.nnzW); alloc(h, CB200_G_VALUES, P.nnzG); alloc(h, CB200_C_VALUES, P.nnzC);
    alloc(h, CB200_SCALARS, S_COUNT);
    for (int w : {(int)CB200_MERIT_GRADIENT, (int)CB200_RESIDUAL_SYMMETRIC, (int)CB200_STEP_SYMMETRIC, (int)CB200_PIVOTS, (int)X_DINV, (int)X_XP}) alloc(h, w, N);
    alloc(h, CB200_PANELS, P.panel_total);
    alloc(h, X_KX, P.kx_total);
    alloc(h, X_WF, P.nnzWf); alloc(h, X_GR, P.nnzG);
    alloc(h, X_LCSR, P.lcsr_total);
    h->scratch.assign((size_t)h->hp.sym.scratch_doubles + 8, 0.0);
    alloc(h, X_FILTER, 4LL * h->opt.max_filter);
    const int mr = h->opt.gmres_restart;
    h->ksize = mr > 0 ? (long long)(mr + 1) * T + (long long)(mr + 1) * mr + 4LL * mr + 8 : 0;
    if (h->ksize) alloc(h, X_KRYLOV, h->ksize);
    h->istat.assign((size_t)batch * I_COUNT, 0);
    for (int b = 0; b < batch; b++) {
        double *s = h->arr[CB200_SCALARS].data() + (size_t)b * S_COUNT;
        s[S_KAPPA] = 0.1; s[S_TAU] = 0.99; s[S_RHO] = 10.0;
    }
    (void)device;
    return h;
}

extern "C" cb200_handle *cb200_ldl_create(int batch, int N, const int *Ap, const int *Ai, const int *perm, int device)
{
    cb200_handle *h = new cb200_handle();
    h->batch = batch;
    h->generic = true;
    cb200_options d;
    cb200_options_default(&d);
    memcpy(&h->opt, &d, sizeof(Options));
    const char *msg = h->gsym.analyze_auto(N, Ap, Ai, perm, BIG_TASK_THRESHOLD);
    if (msg[0]) { fail(msg); delete h; return nullptr; }
    fill_symbolic(h->P, h->gsym, [](const auto &v) { return v.data(); });
    h->P.nnzA = Ap[N];
    h->arr.resize(X_COUNT);
    h->len.assign(X_COUNT, 0);
    for (int w : {(int)CB200_PIVOTS, (int)X_DINV, (int)X_XP, (int)CB200_RHS}) alloc(h, w, N);
    alloc(h, CB200_MATRIX_VALUES, Ap[N]);
    alloc(h, CB200_PANELS, h->P.panel_total);
    alloc(h, X_KX, h->P.kx_total);
    alloc(h, X_LCSR, h->P.lcsr_total);
    h->scratch.assign((size_t)h->gsym.scratch_doubles + 8, 0.0);
    h->istat.assign((size_t)batch * I_COUNT, 0);
    (void)device;
    return h;
}

extern "C" void cb200_destroy(cb200_handle *h) { delete h; }

extern "C" int cb200_info(const cb200_handle *h, long long *out)
{
    const Symbolic &S = h->sym();
    out[0] = S.N; out[1] = h->P.total; out[2] = S.nnzA; out[3] = S.nnzL; out[4] = S.ns; out[5] = S.nlevels;
    out[6] = (long long)S.phases.size(); out[7] = S.max_w; out[8] = S.max_nrow; out[9] = S.panel_total;
    out[10] = S.flops; out[11] = h->batch; out[12] = h->P.n; out[13] = h->P.m; out[14] = h->P.p;
    out[15] = (long long)h->P.nnzW + h->P.nnzG + h->P.nnzC;
    return 0;
}

extern "C" int cb200_path_info(const cb200_handle *h, long long *out)
{
    const Symbolic &S = h->sym();
    out[0] = S.solve_smem; out[1] = S.ctas_per_sm; out[2] = (long long)S.scratch_doubles * 8; out[3] = S.n_cta_tasks;
    out[4] = S.n_generic_cta_tasks; out[5] = 1; out[6] = 1; out[7] = 1;      // (one emulated thread per instance)
    return 0;
}

extern "C" int cb200_amd_order(int N, const int *Ap, const int *Ai, int *perm)
{
    std::vector<int> p;
    amd_order(N, Ap, Ai, p);
    std::copy(p.begin(), p.end(), perm);
    return 0;
}

extern "C" int cb200_get_symbolic(const cb200_handle *h, int *perm, int *etree, int *Lnz)
{
    const Symbolic &S = h->sym();
    if (perm) memcpy(perm, S.perm.data(), sizeof(int) * S.N);
    if (etree) memcpy(etree, S.etree.data(), sizeof(int) * S.N);
    if (Lnz) memcpy(Lnz, S.Lnz.data(), sizeof(int) * S.N);
    return 0;
}

extern "C" int cb200_get_factor(cb200_handle *h, int instance, int *Lp, int *Li, double *Lx, double *D)
{
    const Symbolic &S = h->sym();
    const double *pan = h->arr[CB200_PANELS].data() + (long long)instance * S.panel_total;
    memcpy(D, h->arr[CB200_PIVOTS].data() + (long long)instance * S.N, sizeof(double) * S.N);
    extract_factor(S, pan, Lp, Li, Lx);
    return 0;
}

static int check(cb200_handle *h, int which, int first, int count)
{
    if (which < 0 || which >= CB200_NUM_ARRAYS || h->len[which] == 0) return fail("array not available on this handle");
    if (first < 0 || count < 0 || first + count > h->batch) return fail("instance range out of bounds");
    return 0;
}
extern "C" int cb200_set_array(cb200_handle *h, int which, const double *host, int first, int count)
{
    if (check(h, which, first, count)) return -1;
    memcpy(h->arr[which].data() + (long long)first * h->len[which], host, sizeof(double) * h->len[which] * count);
    return 0;
}
extern "C" int cb200_initialize(cb200_handle *h, const double *guess, int first, int count)
{
    if (h->generic) return fail("not available on a LinearSolver-seam handle");
    if (check(h, CB200_POINT, first, count)) return -1;
    const long long len = h->len[CB200_POINT];
    for (int b = 0; b < count; b++)
        memcpy(h->arr[CB200_POINT].data() + (first + b) * len, guess + (long long)b * h->hp.n, sizeof(double) * h->hp.n);
    return 0;
}
extern "C" int cb200_get_array(cb200_handle *h, int which, double *host, int first, int count)
{
    if (check(h, which, first, count)) return -1;
    memcpy(host, h->arr[which].data() + (long long)first * h->len[which], sizeof(double) * h->len[which] * count);
    return 0;
}
extern "C" int cb200_get_stats(cb200_handle *h, int *host, int first, int count)
{
    memcpy(host, h->istat.data() + (long long)first * I_COUNT, sizeof(int) * I_COUNT * count);
    return 0;
}
extern "C" int cb200_get_profile(cb200_handle *, long long *, int) { return 0; }
extern "C" int cb200_get_array(cb200_handle *h, int which, double *host, int first, int count);
extern "C" int cb200_get_stats(cb200_handle *h, int *host, int first, int count);
extern "C" int cb200_get_array_async(cb200_handle *h, int which, double *host, int first, int count) { return cb200_get_array(h, which, host, first, count); }
extern "C" int cb200_get_stats_async(cb200_handle *h, int *host, int first, int count) { return cb200_get_stats(h, host, first, count); }
extern "C" int cb200_array_length(const cb200_handle *h, int which)
{
    if (which < 0 || which >= CB200_NUM_ARRAYS || h->len[which] == 0) return -1;
    return (int)h->len[which];
}
extern "C" void *cb200_device_ptr(cb200_handle *h, int which) { return h->arr[which].data(); }
extern "C" int cb200_values_changed(cb200_handle *) { return 0; }
extern "C" void *cb200_stream(cb200_handle *) { return nullptr; }
extern "C" int cb200_synchronize(cb200_handle *) { return 0; }
extern "C" int cb200_set_options(cb200_handle *h, const cb200_options *o)
{
    if (o->max_filter != h->opt.max_filter || o->gmres_restart != h->opt.gmres_restart)
        return fail("max_filter / gmres_restart are fixed at creation");
    memcpy(&h->opt, o, sizeof(Options));
    return 0;
}

static double g_red[34];
static void refresh_values(cb200_handle *h);
#define FOR_EACH_INSTANCE                       \
    Ctx ctx{0, 1, 0, g_red, h->scratch.data(), nullptr, nullptr}; \
    const DevProblem &P = h->P;                 \
    for (int b = 0; b < h->batch; b++) {        \
        Inst I = inst(h, b);
#define END_FOR }
static void refresh_values(cb200_handle *h)
{   // (the CUDA library does this only when cb200_set_array changed W or G values; here: always)
    if (h->generic) return;
    FOR_EACH_INSTANCE expand_values(ctx, P, I); END_FOR
}

extern "C" int cb200_cone(cb200_handle *h, int flags, int at_candidate)
{
    FOR_EACH_INSTANCE cone_eval(ctx, P, I, at_candidate ? I.cand : I.w, flags & 1, (flags >> 1) & 1, (flags >> 2) & 1); END_FOR
    return 0;
}
extern "C" int cb200_residual(cb200_handle *h)
{
    FOR_EACH_INSTANCE residual_eval(ctx, P, I); I.scal[S_THETA] = constraint_violation(ctx, P, I, I.w); END_FOR
    return 0;
}
extern "C" int cb200_search_direction(cb200_handle *h)
{
    refresh_values(h);
    FOR_EACH_INSTANCE I.istat[I_STATUS] = search_direction(ctx, P, I, h->opt); END_FOR
    return 0;
}
extern "C" int cb200_cone_search(cb200_handle *h)
{
    FOR_EACH_INSTANCE
        int st = cone_search(ctx, P, I, h->opt);
        double a = I.scal[S_STEP_SIZE];
        for (int i = 0; i < P.n + P.m; i++) I.cand[i] = I.w[i] - a * I.step[i];
        if (st != ST_OK) I.istat[I_STATUS] = st;
    END_FOR
    return 0;
}
extern "C" int cb200_apply_step(cb200_handle *h)
{
    FOR_EACH_INSTANCE
        const int m = P.m, p = P.p, N = P.N;
        const double a = I.scal[S_STEP_SIZE];
        for (int i = 0; i < N; i++) I.w[i] = I.w[i] - a * I.step[i];
        for (int i = 0; i < m + p; i++) I.w[N + i] = I.w[N + i] - a * I.step[N + i];
        for (int i = 0; i < p; i++) I.w[N + m + p + i] = I.cand[N + m + p + i];
        cone_eval(ctx, P, I, I.w, 0, 0, 1);
        I.scal[S_EQUALITY_VIOLATION] = scope_max(ctx, m, [&](int i) { return fabs(I.g[i]); });
        I.scal[S_CONE_PRODUCT_VIOLATION] = scope_max(ctx, p, [&](int i) { return fabs(I.prod[i]); });
    END_FOR
    return 0;
}
extern "C" int cb200_kkt_factor_solve(cb200_handle *h, int nsolves)
{
    refresh_values(h);
    FOR_EACH_INSTANCE
        kkt_entries(ctx, P, I);
        ldl_factor(ctx, P, I.panels, I.D, I.Dinv, KSrc{I.Wv, I.Gr, I.Cv, I.kx}, I.Lcsr, I.istat, nullptr);
        for (int k = 0; k < nsolves; k++) direction_symmetric(ctx, P, I, I.res, I.step);
    END_FOR
    return 0;
}
extern "C" int cb200_differentiate(cb200_handle *h, int nparam, const double *H, double *S)
{
    refresh_values(h);
    FOR_EACH_INSTANCE
        kkt_entries(ctx, P, I);
        ldl_factor(ctx, P, I.panels, I.D, I.Dinv, KSrc{I.Wv, I.Gr, I.Cv, I.kx}, I.Lcsr, I.istat, nullptr);
        for (int i = 0; i < nparam; i++) {
            const double *rhs = H + ((long long)b * nparam + i) * P.total;
            double *out = S + ((long long)b * nparam + i) * P.total;
            direction_symmetric(ctx, P, I, rhs, out);
            for (int k = 0; k < P.total; k++) out[k] = -1.0 * out[k];
        }
    END_FOR
    return 0;
}
static int scatter_slot(int which)
{
    return which == CB200_W_VALUES ? 0 : which == CB200_G_VALUES ? 1 : which == CB200_C_VALUES ? 2 : -1;
}
extern "C" int cb200_scatter_plan(cb200_handle *h, int which, int ncaches, const int *cache_len, const int *rows, const int *cols)
{
    if (h->generic) return fail("not available on a LinearSolver-seam handle");
    const int slot = scatter_slot(which);
    if (slot < 0) return fail("cb200_scatter_plan: which must be CB200_W_VALUES, CB200_G_VALUES or CB200_C_VALUES");
    if (slot > 0 && ncaches != 1) return fail("cb200_scatter_plan: the G and C values have one cache each");
    const HostProblem &H = h->hp;
    auto &sc = h->scatter[slot];
    std::string err = slot == 0 ? sc.plan.build(H.n, H.n, H.Wp.data(), H.Wi.data(), true, ncaches, cache_len, rows, cols)
                    : slot == 1 ? sc.plan.build(H.m, H.n, H.Gp.data(), H.Gi.data(), false, ncaches, cache_len, rows, cols)
                                : sc.plan.build(H.p, H.n, H.Cp.data(), H.Ci.data(), false, ncaches, cache_len, rows, cols);
    if (!err.empty()) { sc.plan = ScatterPlan(); sc.set = false; return fail(err); }
    sc.caches.assign((size_t)std::max<long long>(sc.plan.cache_total * h->batch, 1), 0.0);
    sc.set = true;
    return 0;
}
extern "C" void *cb200_scatter_buffer(cb200_handle *h, int which)
{
    const int slot = scatter_slot(which);
    return slot < 0 || !h->scatter[slot].set ? nullptr : (void *)h->scatter[slot].caches.data();
}
extern "C" int cb200_scatter(cb200_handle *h, int which, const double *caches_host, int first, int count)
{
    const int slot = scatter_slot(which);
    if (h->generic || slot < 0 || !h->scatter[slot].set) return fail("cb200_scatter: no scatter plan for this array");
    if (check(h, which, first, count)) return -1;
    auto &sc = h->scatter[slot];
    const long long L = sc.plan.cache_total;
    if (caches_host) memcpy(sc.caches.data() + first * L, caches_host, sizeof(double) * L * count);
    Ctx ctx{0, 1, 0, g_red, h->scratch.data(), nullptr, nullptr};
    for (int b = first; b < first + count; b++)
        scatter_caches(ctx, sc.plan.nnz, sc.plan.ncaches, sc.plan.idx.data(), sc.caches.data() + b * L,
                       h->arr[which].data() + (long long)b * h->len[which], 0, 1);
    return 0;
}
static int stage_slot(int which)
{
    return which == CB200_GRADIENT ? 0 : which == CB200_EQ_DUAL_GRAD ? 1 : which == CB200_CONE_DUAL_GRAD ? 2
         : which == CB200_EQUALITY ? 3 : which == CB200_CONE ? 4 : -1;
}
extern "C" int cb200_stage_plan(cb200_handle *h, int which, int accumulate, int count, const int *dst)
{
    if (h->generic) return fail("not available on a LinearSolver-seam handle");
    const int slot = stage_slot(which);
    if (slot < 0) return fail("cb200_stage_plan: which must be CB200_GRADIENT, CB200_EQ_DUAL_GRAD, CB200_CONE_DUAL_GRAD, CB200_EQUALITY or CB200_CONE");
    if (count > 0 && !dst) return fail("cb200_stage_plan: no index list");
    auto &sg = h->stage[slot];
    std::string err = sg.plan.build((int)h->len[which], accumulate != 0, count, dst);
    if (!err.empty()) { sg.plan = StagePlan(); sg.set = false; return fail(err); }
    sg.caches.assign((size_t)std::max<long long>(sg.plan.cache_total * h->batch, 1), 0.0);
    sg.set = true;
    return 0;
}
extern "C" void *cb200_stage_buffer(cb200_handle *h, int which)
{
    const int slot = stage_slot(which);
    return slot < 0 || !h->stage[slot].set ? nullptr : (void *)h->stage[slot].caches.data();
}
extern "C" int cb200_stage_scatter(cb200_handle *h, int which, const double *caches_host, int first, int count)
{
    const int slot = stage_slot(which);
    if (h->generic || slot < 0 || !h->stage[slot].set) return fail("cb200_stage_scatter: no stage plan for this array");
    if (check(h, which, first, count)) return -1;
    auto &sg = h->stage[slot];
    const long long L = sg.plan.cache_total;
    if (caches_host && L > 0) memcpy(sg.caches.data() + first * L, caches_host, sizeof(double) * L * count);
    Ctx ctx{0, 1, 0, g_red, h->scratch.data(), nullptr, nullptr};
    for (int b = first; b < first + count; b++)
        stage_gather(ctx, sg.plan.nout, sg.plan.accumulate, sg.plan.ptr.data(), sg.plan.src.data(), sg.caches.data() + b * L,
                     h->arr[which].data() + (long long)b * h->len[which], 0, 1);
    return 0;
}
extern "C" int cb200_jacobian_times(cb200_handle *h, const double *v, double *out)
{
    refresh_values(h);
    FOR_EACH_INSTANCE jacobian_times(ctx, P, I, v + (long long)b * P.total, out + (long long)b * P.total); END_FOR
    return 0;
}
extern "C" int cb200_lq_evaluate(cb200_handle *h, int flags, int at_candidate)
{
    refresh_values(h);
    FOR_EACH_INSTANCE lq_evaluate(ctx, P, I, at_candidate ? I.cand : I.w, flags); END_FOR
    return 0;
}
extern "C" int cb200_lq_begin(cb200_handle *h, int warmstart)
{
    refresh_values(h);
    FOR_EACH_INSTANCE solve_begin_lq(ctx, P, I, h->opt, warmstart); END_FOR
    return 0;
}
extern "C" int cb200_lq_step(cb200_handle *h, int iterations)
{
    refresh_values(h);
    for (int k = 0; k < iterations; k++) { FOR_EACH_INSTANCE solve_step_lq(ctx, P, I, h->opt); END_FOR }
    return 0;
}
extern "C" int cb200_filter_reset(cb200_handle *h)
{
    FOR_EACH_INSTANCE filter_reset(ctx, I, h->opt); END_FOR
    return 0;
}
extern "C" int cb200_filter_search(cb200_handle *h, int first, int count, const double *f_host, const double *g_host,
                                   const double *h_host, int *accepted_host)
{
    if (first < 0 || count <= 0 || !f_host || !accepted_host) return fail("cb200_filter_search: invalid arguments");
    FOR_EACH_INSTANCE
        accepted_host[b] = filter_search(ctx, P, I, h->opt, first, count, f_host + (long long)b * count,
                                         g_host + (long long)b * count * P.m, h_host + (long long)b * count * P.p);
    END_FOR
    return 0;
}
extern "C" int cb200_lq_set_order(cb200_handle *h, const int *order)
{   // a scheduling hint only: the emulation runs the instances one after the other in any case; the argument is validated
    if (!order) return 0;
    std::vector<char> seen((size_t)h->batch, 0);
    for (int i = 0; i < h->batch; i++) {
        if (order[i] < 0 || order[i] >= h->batch || seen[(size_t)order[i]]) { g_err = "cb200_lq_set_order: not a permutation of the instances"; return -1; }
        seen[(size_t)order[i]] = 1;
    }
    return 0;
}
extern "C" int cb200_allreduce_counts(cb200_handle *h, long long *counts)
{
    for (int k = 0; k < 4; k++) counts[k] = 0;
    for (int b = 0; b < h->batch; b++) counts[h->istat[(size_t)b * I_COUNT + I_CONVERGED] & 3]++;
    return 0;
}
extern "C" int cb200_lq_solve(cb200_handle *h, int max_steps, int check_every, long long *counts, int *steps_done)
{
    if (check_every <= 0) check_every = 1;
    int done = 0;
    long long c[4] = {h->batch, 0, 0, 0};
    while (done < max_steps) {
        int chunk = std::min(check_every, max_steps - done);
        cb200_lq_step(h, chunk);
        done += chunk;
        cb200_allreduce_counts(h, c);
        if (c[0] == 0) break;
    }
    if (counts) for (int k = 0; k < 4; k++) counts[k] = c[k];
    if (steps_done) *steps_done = done;
    return 0;
}
extern "C" int cb200_ldl_factorize(cb200_handle *h)
{
    Ctx ctx{0, 1, 0, g_red, h->scratch.data(), nullptr, nullptr};
    for (int b = 0; b < h->batch; b++) {
        double *pan = h->arr[CB200_PANELS].data() + (long long)b * h->P.panel_total;
        const double *Ax = h->arr[CB200_MATRIX_VALUES].data() + (long long)b * h->P.nnzA;
        ldl_factor(ctx, h->P, pan, h->arr[CB200_PIVOTS].data() + (long long)b * h->P.N,
                   h->arr[X_DINV].data() + (long long)b * h->P.N, KSrc{Ax, Ax, Ax, Ax},
                   h->arr[X_LCSR].data() + (long long)b * h->P.lcsr_total, h->istat.data() + (size_t)b * I_COUNT, nullptr);
    }
    return 0;
}
extern "C" int cb200_ldl_solve(cb200_handle *h)
{
    Ctx ctx{0, 1, 0, g_red, h->scratch.data(), nullptr, nullptr};
    for (int b = 0; b < h->batch; b++) {
        double *rhs = h->arr[CB200_RHS].data() + (long long)b * h->P.N;
        ldl_solve(ctx, h->P, h->arr[CB200_PANELS].data() + (long long)b * h->P.panel_total,
                  h->arr[CB200_PIVOTS].data() + (long long)b * h->P.N, h->arr[X_DINV].data() + (long long)b * h->P.N,
                  h->arr[X_KX].data() + (long long)b * h->P.kx_total,
                  h->arr[X_LCSR].data() + (long long)b * h->P.lcsr_total, rhs, rhs,
                  h->arr[X_XP].data() + (long long)b * h->P.N, h->istat.data() + (size_t)b * I_COUNT, nullptr);
    }
    return 0;
}
extern "C" int cb200_ldl_inertia(cb200_handle *h, int *out)
{
    for (int b = 0; b < h->batch; b++)
        for (int k = 0; k < 3; k++) out[3 * b + k] = h->istat[(size_t)b * I_COUNT + k];
    return 0;
}
extern "C" int cb200_ldl_linear_solve(cb200_handle *h, const double *Ax, const double *bvec, double *x, int factorize)
{
    if (factorize) { cb200_set_array(h, CB200_MATRIX_VALUES, Ax, 0, h->batch); cb200_ldl_factorize(h); }
    cb200_set_array(h, CB200_RHS, bvec, 0, h->batch);
    cb200_ldl_solve(h);
    return cb200_get_array(h, CB200_RHS, x, 0, h->batch);
}
extern "C" int cb200_nccl_unique_id(char *) { return fail("emulation library has no NCCL"); }
extern "C" int cb200_comm_init(cb200_handle *, int, int, const char *) { return fail("emulation library has no NCCL"); }
