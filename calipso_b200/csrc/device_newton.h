// device_newton.h -- the inner Newton iteration of solve! (src/solver/solve.jl:98-350) as one CTA-per-instance
// device routine, for problems whose callbacks are affine/quadratic (the LQ-conic family, calipso_b200/lqc.py):
// evaluate! (src/solver/evaluate.jl) then reduces to sparse mat-vecs with constant matrices and the whole loop --
// residual, KKT factorisation, refinement, cone search, filter line search, update -- stays on the device.
#pragma once

#include "device_core.h"

namespace cb200 {

enum {   // evaluate! flags (same bits as include/calipso_b200.h)
    EV_OBJECTIVE = 1, EV_GRADIENT = 2, EV_EQUALITY = 4, EV_CONE = 8, EV_EQUALITY_DUAL_GRAD = 16,
    EV_CONE_DUAL_GRAD = 32, EV_HESSIAN = 64, EV_EQUALITY_JAC = 128, EV_CONE_JAC = 256
};

// f = 1/2 x'Qx + q'x, grad = Qx + q, g = Gx + g0, h = Cx + h0, (g'y)_x = G'y, (h'z)_x = C'z at point w
CB_DEVN void lq_evaluate(const Ctx &ctx, const DevProblem &P, const Inst &I, const double *w, int flags)
{
    const int n = P.n, m = P.m, p = P.p;
    const double *x = w, *y = w + n + m + p, *z = w + n + 2 * m + p;
    // sparse rows by groups of four lanes (grouped_rows): short dependent-load chains, coalesced index reads
    if (flags & (EV_OBJECTIVE | EV_GRADIENT)) {
        const int *__restrict__ wp = P.Wfull.ptr, *__restrict__ wc = P.Wfull.col;
        double *red = I.tmp;       // per-row terms of the objective (summed below in a fixed order)
        grouped_rows<4>(
            ctx, n,
            [&](int i, int sub, int st) {
#if CB_ON_DEVICE
                if (st == 4) return sparse_dot4(I.Wf, nullptr, wc, x, wp[i], wp[i + 1], sub);     // same order, eight loads in flight
#endif
                double a = 0.0;
                for (int k = wp[i] + sub; k < wp[i + 1]; k += st) a += I.Wf[k] * x[wc[k]];
                return a;
            },
            [&](int i, double a) {
                if (flags & EV_GRADIENT) I.grad[i] = a + I.q[i];
                red[i] = x[i] * (0.5 * a + I.q[i]);
            });
        ctx.sync();
        if (flags & EV_OBJECTIVE) {
            double f = scope_sum(ctx, n, [&](int i) { return red[i]; });
            if (ctx.tid == 0) I.scal[S_OBJECTIVE] = f;
        }
    }
    if (flags & EV_EQUALITY) {
        const int *__restrict__ gp = P.Grow.ptr, *__restrict__ gc = P.Grow.col;
        grouped_rows<4>(
            ctx, m,
            [&](int i, int sub, int st) {
#if CB_ON_DEVICE
                if (st == 4) return sparse_dot4(I.Gr, nullptr, gc, x, gp[i], gp[i + 1], sub);
#endif
                double a = 0.0;
                for (int k = gp[i] + sub; k < gp[i + 1]; k += st) a += I.Gr[k] * x[gc[k]];
                return a;
            },
            [&](int i, double a) { I.g[i] = I.g0[i] + a; });
    }
    if (flags & EV_CONE)
        PAR_FOR(i, p) {
            double a = I.h0[i];
            for (int k = P.Crow.ptr[i]; k < P.Crow.ptr[i + 1]; k++) a += I.Cv[P.Crow.src[k]] * x[P.Crow.col[k]];
            I.h[i] = a;
        }
    if (flags & (EV_EQUALITY_DUAL_GRAD | EV_CONE_DUAL_GRAD)) {
        const int *__restrict__ gp = P.Gp, *__restrict__ gi = P.Gi, *__restrict__ cp = P.Cp, *__restrict__ ci = P.Ci;
        const double *__restrict__ Gv = I.Gv, *__restrict__ Cv = I.Cv;
        if (flags & EV_EQUALITY_DUAL_GRAD)
            grouped_rows<4>(
                ctx, n,
                [&](int j, int sub, int st) {
#if CB_ON_DEVICE
                    if (st == 4) return sparse_dot4(Gv, nullptr, gi, y, gp[j], gp[j + 1], sub);
#endif
                    double a = 0.0;
                    for (int k = gp[j] + sub; k < gp[j + 1]; k += st) a += Gv[k] * y[gi[k]];
                    return a;
                },
                [&](int j, double a) { I.gyx[j] = a; });
        if (flags & EV_CONE_DUAL_GRAD)
            PAR_FOR(j, n) {
                double a = 0.0;
                for (int k = cp[j]; k < cp[j + 1]; k++) a += Cv[k] * z[ci[k]];
                I.hzx[j] = a;
            }
    }
    ctx.sync();
}

// ------------------------------------------------------------------------------------------------ filter.jl
// pairs live in I.filter[0 .. 2F), cache in I.filter[2F .. 4F)
CB_DEV void filter_reset(const Ctx &ctx, const Inst &I, const Options &o)
{
    const int F = o.max_filter;
    PAR_FOR(i, 2 * F) { I.filter[i] = 1.0e8; I.filter[2 * F + i] = 1.0e8; }
    if (ctx.tid == 0) I.istat[I_FILTER_INDEX] = 0;
    ctx.sync();
}

CB_DEV int check_filter(const Ctx &ctx, const Inst &I, const Options &o, double cv, double merit)
{   // filter.jl:43-50 (all stored pairs, placeholders included)
    return !scope_any(ctx, o.max_filter, [&](int i) { return !(cv < I.filter[2 * i] || merit < I.filter[2 * i + 1]); });
}

CB_DEV void augment_filter_pair(const Ctx &ctx, const Inst &I, const Options &o, double cv, double merit)
{   // filter.jl:52-79
    const int F = o.max_filter;
    int idx = I.istat[I_FILTER_INDEX];
    ctx.sync();
    if (idx == 0) {
        if (ctx.tid == 0) { I.filter[0] = cv; I.filter[1] = merit; I.istat[I_FILTER_INDEX] = 1; }
    } else if (check_filter(ctx, I, o, cv, merit)) {
        if (ctx.tid == 0) {
            double *pairs = I.filter, *cache = I.filter + 2 * F;
            for (int i = 0; i < 2 * idx; i++) cache[i] = pairs[i];
            for (int i = 0; i < 2 * idx; i++) pairs[i] = 1.0e8;
            int k = 1;
            pairs[0] = cv; pairs[1] = merit;
            for (int i = 0; i < idx; i++)
                if (!(cache[2 * i] >= cv && cache[2 * i + 1] >= merit)) {
                    pairs[2 * k] = cache[2 * i];
                    pairs[2 * k + 1] = cache[2 * i + 1];
                    k++;
                }
            I.istat[I_FILTER_INDEX] = k;
        }
    }
    ctx.sync();
}

// line_search.jl
CB_DEV bool switching_condition(double step_size, double d, double merit_exponent, double violation,
                                double violation_exponent, double regularization)
{
    return d < 0.0 && step_size * pow(-d, merit_exponent) > regularization * pow(violation, violation_exponent);
}
CB_DEV bool sufficient_progress(double v, double vc, double M, double Mc, double vt, double mt, double mach)
{
    return (vc - 10.0 * mach * fabs(v) <= (1.0 - vt) * v) || (Mc - 10.0 * mach * fabs(M) <= M - mt * v);
}
CB_DEV bool armijo(double M, double Mc, double d, double step_size, double at, double mach)
{
    return Mc - M - 10.0 * mach * fabs(M) <= at * step_size * d;
}

// The residual (filter) line search of solve.jl:224-306 over candidates whose callback outputs were computed elsewhere:
// fc[k], gc[k][m], hc[k][p] = f, g, h at w - alpha_k step, alpha_k = alpha_cone 0.5^(first + k).  Returns the index
// (first + k) of the accepted candidate or -1.  Same tests, in the same order, as the on-device loop of
// newton_iteration_lq below; state that spans calls (merit, theta, slope, alpha_cone) lives in the scalar slots.
CB_DEVN int filter_search(const Ctx &ctx, const DevProblem &P, const Inst &I, const Options &o, int first, int count,
                          const double *fc, const double *gc, const double *hc)
{
    const int n = P.n, m = P.m, p = P.p, N = P.N;
    const double *w = I.w;
    double *c = I.cand;
    if (first == 0) {
        // current point: merit, merit gradient, slope, theta (solve.jl:141-150,170-172,231-233)
        const double M0 = merit_value(ctx, P, I, w);
        merit_gradient(ctx, P, I);
        const double th0 = constraint_violation(ctx, P, I, w);
        const double d0 = scope_sum(ctx, N, [&](int i) { return I.mgrad[i] * I.step[i]; });
        if (ctx.tid == 0) {
            I.scal[S_MERIT] = M0; I.scal[S_THETA] = th0; I.scal[S_MERIT_SLOPE] = d0; I.scal[S_STEP_SIZE_CONE] = I.scal[S_STEP_SIZE];
        }
        ctx.sync();
    }
    const double M = I.scal[S_MERIT], theta = I.scal[S_THETA], d = I.scal[S_MERIT_SLOPE], a0 = I.scal[S_STEP_SIZE_CONE];
    const double f_keep = I.scal[S_OBJECTIVE], phi_keep = I.scal[S_BARRIER];
    ctx.sync();
    int accepted = -1;
    double step_size = a0, Mh = 0.0, theta_h = 0.0;
    for (int k = 0; k < first; k++) step_size = o.scaling_line_search * step_size;
    for (int k = 0; k < count; k++) {
        PAR_FOR(i, N) c[i] = w[i] - step_size * I.step[i];                      // x, r, s of the candidate (solve.jl:224-229,292-294)
        PAR_FOR(i, m) I.g[i] = gc[(long long)k * m + i];
        PAR_FOR(i, p) I.h[i] = hc[(long long)k * p + i];
        if (ctx.tid == 0) I.scal[S_OBJECTIVE] = fc[k];
        ctx.sync();
        cone_eval(ctx, P, I, c, 1, 0, 0);
        Mh = merit_value(ctx, P, I, c);
        theta_h = constraint_violation(ctx, P, I, c);
        bool ok = false;
        if (check_filter(ctx, I, o, theta_h, Mh)) {
            if (theta <= o.slack_tolerance && switching_condition(step_size, d, o.merit_exponent, theta, o.violation_exponent, 1.0) &&
                armijo(M, Mh, d, step_size, o.armijo_tolerance, o.machine_tolerance))
                ok = true;
            else if (sufficient_progress(theta, theta_h, M, Mh, o.violation_tolerance, o.merit_tolerance, o.machine_tolerance))
                ok = true;
        }
        // the reference leaves the loop after max_residual_line_search halvings with the last candidate evaluated
        if (ok || first + k >= o.max_residual_line_search) { accepted = first + k; break; }
        step_size = o.scaling_line_search * step_size;
    }
    ctx.sync();
    if (accepted >= 0) {
        // augment_filter!(solver, ...), filter.jl:81-89 / solve.jl:298-306
        if (!switching_condition(step_size, d, o.merit_exponent, theta, o.violation_exponent, 1.0) ||
            !armijo(M, Mh, d, step_size, o.armijo_tolerance, o.machine_tolerance))
            augment_filter_pair(ctx, I, o, (1.0 - o.violation_tolerance) * theta, M - o.merit_tolerance * theta);
        if (ctx.tid == 0) {
            I.scal[S_STEP_SIZE] = step_size;
            I.scal[S_MERIT_CANDIDATE] = Mh;
            I.scal[S_THETA_CANDIDATE] = theta_h;
            I.istat[I_LINE_SEARCH] = accepted;
        }
    } else if (ctx.tid == 0) {          // nothing accepted yet: the current point's objective and barrier stay in their slots
        I.scal[S_OBJECTIVE] = f_keep;
        I.scal[S_BARRIER] = phi_keep;
    }
    ctx.sync();
    return accepted;
}

// ------------------------------------------------------------------------------------------------ solve! pieces
// solve.jl:8-95
CB_DEVN void solve_begin_lq(const Ctx &ctx, const DevProblem &P, const Inst &I, const Options &o, int warmstart)
{
    const int n = P.n, m = P.m, p = P.p;
    double *w = I.w;
    if (!warmstart) {
        lq_evaluate(ctx, P, I, w, EV_EQUALITY | EV_CONE);                 // initialize_slacks!, initialize.jl:15-29
        PAR_FOR(i, m) { w[n + i] = I.g[i]; w[n + m + p + i] = 0.0; }
        PAR_FOR(i, p) w[n + 2 * m + p + i] = 0.0;                          // initialize_duals!, :31-36
        PAR_FOR(i, P.q_nn) { w[n + m + i] = 1.0; w[n + 2 * m + 2 * p + i] = 1.0; }
        PAR_FOR(k, P.nsoc)
            for (int i = 0; i < P.soc_dims[k]; i++) {
                double v = i == 0 ? 1.0 : 0.1;
                w[n + m + P.soc_off[k] + i] = v;
                w[n + 2 * m + 2 * p + P.soc_off[k] + i] = v;
            }
    }
    if (ctx.tid == 0) {
        I.scal[S_KAPPA] = o.central_path_initial;
        I.scal[S_TAU] = fmax(0.99, 1.0 - o.central_path_initial);
        I.scal[S_RHO] = o.penalty_initial;
        I.istat[I_TOTAL_ITERATIONS] = 1;
        I.istat[I_OUTER] = 1;
        I.istat[I_INNER] = 1;
        I.istat[I_STATUS] = ST_OK;
        I.istat[I_CONVERGED] = 0;
    }
    PAR_FOR(i, m) I.lambda[i] = o.dual_initial;
    ctx.sync();
    lq_evaluate(ctx, P, I, w, EV_OBJECTIVE | EV_EQUALITY | EV_CONE);
    double ev = scope_max(ctx, m, [&](int i) { return fabs(I.g[i]); });
    double cv = scope_max(ctx, p, [&](int i) { return fabs(I.prod[i]); });  // stale product on purpose, solve.jl:85-91
    if (ctx.tid == 0) { I.scal[S_EQUALITY_VIOLATION] = ev; I.scal[S_CONE_PRODUCT_VIOLATION] = cv; }
    cone_eval(ctx, P, I, w, 0, 0, 1);
    filter_reset(ctx, I, o);
}

// solve.jl:356-368
CB_DEVN void outer_update(const Ctx &ctx, const DevProblem &P, const Inst &I, const Options &o)
{
    const double kappa = I.scal[S_KAPPA], rho = I.scal[S_RHO];
    ctx.sync();
    const double kn = fmax(o.residual_tolerance / 10.0, fmin(o.central_path_scaling * kappa, pow(kappa, o.central_path_exponent)));
    PAR_FOR(i, P.m) I.lambda[i] = I.lambda[i] + rho * I.w[P.n + i];
    if (ctx.tid == 0) {
        I.scal[S_KAPPA] = kn;
        I.scal[S_TAU] = fmax(0.99, 1.0 - kn);
        I.scal[S_RHO] = fmin(fmax(o.penalty_scaling * rho, 1.0 / kn), o.max_penalty);
        I.istat[I_OUTER]++;
        I.istat[I_INNER] = 1;
    }
    ctx.sync();
    filter_reset(ctx, I, o);
}

// One pass of the inner loop body, solve.jl:98-350.  Returns 0 continue, 1 outer-converged, 2 inner-converged
// (caller performs the outer update), negative = -status on error.
CB_DEVN int newton_iteration_lq(const Ctx &ctx, const DevProblem &P, const Inst &I, const Options &o)
{
    const int n = P.n, m = P.m, p = P.p, N = P.N;
    double *w = I.w;
    ProfTimer pt{I.prof, 0}, ptot{I.prof, 0};
    pt.start();
    ptot.start();
    lq_evaluate(ctx, P, I, w, EV_GRADIENT | EV_EQUALITY_DUAL_GRAD | EV_CONE_DUAL_GRAD);
    cone_eval(ctx, P, I, w, 1, 1, 0);
    const double M = merit_value(ctx, P, I, w);
    merit_gradient(ctx, P, I);
    residual_eval(ctx, P, I);
    const double kappa = I.scal[S_KAPPA];
    if (I.scal[S_RESIDUAL_VIOLATION] < o.residual_tolerance && I.scal[S_SLACK_VIOLATION] < o.slack_tolerance &&
        I.scal[S_EQUALITY_VIOLATION] <= o.equality_tolerance &&
        I.scal[S_CONE_PRODUCT_VIOLATION] <= o.complementarity_tolerance)
        return 1;
    else if (I.scal[S_OPTIMALITY_VIOLATION] <= fmax(o.central_path_update_tolerance * kappa, o.optimality_tolerance))
        return 2;
    const double theta = constraint_violation(ctx, P, I, w);
    pt.stop(PROF_CONE_RESIDUAL);
    // (second derivatives and Jacobians are constant for the LQ family: nothing to re-evaluate, solve.jl:175-185)
    int st = search_direction(ctx, P, I, o);
    if (st == ST_INERTIA_FAILURE) return -ST_INERTIA_FAILURE;
    // ST_REFINEMENT_FAILURE is not an error of solve!: the reference takes whatever `J \ R` returns after a failed
    // refinement (search_direction.jl:22) and carries on; here the step of the fallback is taken likewise and the event is
    // counted in I_UNREFINED_STEPS (cb200_get_stats) so that callers can see it.
    pt.start();
    st = cone_search(ctx, P, I, o);
    if (st != ST_OK) return -ST_CONE_SEARCH_FAILURE;
    double step_size = I.scal[S_STEP_SIZE];
    ctx.sync();
    double *c = I.cand;
    PAR_FOR(i, n + m) c[i] = w[i] - step_size * I.step[i];
    ctx.sync();
    lq_evaluate(ctx, P, I, c, EV_OBJECTIVE | EV_EQUALITY | EV_CONE);
    cone_eval(ctx, P, I, c, 1, 1, 0);
    double Mh = merit_value(ctx, P, I, c);
    double theta_h = constraint_violation(ctx, P, I, c);
    const double d = scope_sum(ctx, N, [&](int i) { return I.mgrad[i] * I.step[i]; });
    int residual_iteration = 0;
    while (residual_iteration < o.max_residual_line_search) {
        if (check_filter(ctx, I, o, theta_h, Mh)) {
            if (theta <= o.slack_tolerance &&
                switching_condition(step_size, d, o.merit_exponent, theta, o.violation_exponent, 1.0) &&
                armijo(M, Mh, d, step_size, o.armijo_tolerance, o.machine_tolerance))
                break;
            else if (sufficient_progress(theta, theta_h, M, Mh, o.violation_tolerance, o.merit_tolerance,
                                         o.machine_tolerance))
                break;
        }
        step_size = o.scaling_line_search * step_size;
        PAR_FOR(i, N) c[i] = w[i] - step_size * I.step[i];
        ctx.sync();
        lq_evaluate(ctx, P, I, c, EV_OBJECTIVE | EV_EQUALITY | EV_CONE);
        cone_eval(ctx, P, I, c, 1, 1, 0);
        Mh = merit_value(ctx, P, I, c);
        theta_h = constraint_violation(ctx, P, I, c);
        residual_iteration++;
    }
    // augment_filter!(solver, ...), filter.jl:81-89
    if (!switching_condition(step_size, d, o.merit_exponent, theta, o.violation_exponent, 1.0) ||
        !armijo(M, Mh, d, step_size, o.armijo_tolerance, o.machine_tolerance))
        augment_filter_pair(ctx, I, o, (1.0 - o.violation_tolerance) * theta, M - o.merit_tolerance * theta);
    // update, solve.jl:309-326
    PAR_FOR(i, N) w[i] = c[i];
    PAR_FOR(i, m + p) w[N + i] = w[N + i] - step_size * I.step[N + i];
    PAR_FOR(i, p) w[N + m + p + i] = c[N + m + p + i];
    ctx.sync();
    cone_eval(ctx, P, I, w, 0, 0, 1);
    double ev = scope_max(ctx, m, [&](int i) { return fabs(I.g[i]); });
    double cv = scope_max(ctx, p, [&](int i) { return fabs(I.prod[i]); });
    if (ctx.tid == 0) {
        I.scal[S_EQUALITY_VIOLATION] = ev;
        I.scal[S_CONE_PRODUCT_VIOLATION] = cv;
        I.scal[S_STEP_SIZE] = step_size;
        I.scal[S_MERIT] = M;
        I.scal[S_THETA] = theta;
        I.scal[S_MERIT_CANDIDATE] = Mh;
        I.scal[S_THETA_CANDIDATE] = theta_h;
        I.istat[I_LINE_SEARCH] = residual_iteration;
        I.istat[I_TOTAL_ITERATIONS]++;
        I.istat[I_INNER]++;
    }
    ctx.sync();
    pt.stop(PROF_EVAL_LINESEARCH);
    ptot.stop(PROF_TOTAL);
    return 0;
}

// One scheduling step of solve! for an instance: a Newton iteration, plus the outer update when the inner loop
// breaks (solve.jl:165-167,356-368) or exhausts max_residual_iterations.  Converged / failed instances are masked.
// I_CONVERGED: 0 running, 1 converged (solve! returns true), 2 gave up (returns false), 3 error.
CB_DEVN void solve_step_lq(const Ctx &ctx, const DevProblem &P, const Inst &I, const Options &o)
{
    int state = I.istat[I_CONVERGED];
    ctx.sync();
    if (state != 0) return;
    int rc = newton_iteration_lq(ctx, P, I, o);
    int inner = I.istat[I_INNER], outer = I.istat[I_OUTER];
    ctx.sync();
    if (rc == 1) {
        if (ctx.tid == 0) I.istat[I_CONVERGED] = 1;
    } else if (rc < 0) {
        if (ctx.tid == 0) { I.istat[I_CONVERGED] = 3; I.istat[I_STATUS] = -rc; }
    } else if (rc == 2 || inner > o.max_residual_iterations) {
        if (outer >= o.max_outer_iterations) {
            if (ctx.tid == 0) I.istat[I_CONVERGED] = 2;
        }
        outer_update(ctx, P, I, o);
    }
    ctx.sync();
}

}  // namespace cb200
