// symbolic.cpp -- see symbolic.h
#include "symbolic.h"

#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <string>

namespace cb200 {

// ---------------------------------------------------------------------------------------------------------------
// Quotient-graph minimum degree (stand-in for AMD.amd, reference call site src/solver/qdldl.jl:135).
// Variables keep a list of adjacent variables and adjacent elements; eliminating p creates element p whose boundary
// is the union of p's variable neighbours and the boundaries of p's elements (which are absorbed).  Degrees of the
// boundary variables are recomputed exactly with a stamped marker sweep.
void minimum_degree(int n, const int *Ap, const int *Ai, std::vector<int> &perm)
{
    std::vector<std::vector<int>> adjV(n), adjE(n), elem(n);
    for (int j = 0; j < n; j++)
        for (int q = Ap[j]; q < Ap[j + 1]; q++) {
            int i = Ai[q];
            if (i == j) continue;
            adjV[i].push_back(j);
            adjV[j].push_back(i);
        }
    std::vector<int> status(n, 0);  // 0 variable, 1 element, 2 absorbed element
    std::vector<int> degree(n), mark(n, -1);
    for (int i = 0; i < n; i++) {
        std::sort(adjV[i].begin(), adjV[i].end());
        adjV[i].erase(std::unique(adjV[i].begin(), adjV[i].end()), adjV[i].end());
        degree[i] = (int)adjV[i].size();
    }
    perm.assign(n, 0);
    int stamp = 0;
    std::vector<int> Lp;
    for (int k = 0; k < n; k++) {
        int p = -1;
        for (int i = 0; i < n; i++)
            if (status[i] == 0 && (p < 0 || degree[i] < degree[p])) p = i;
        perm[k] = p;
        // boundary of the new element
        stamp++;
        mark[p] = stamp;
        Lp.clear();
        for (int v : adjV[p])
            if (status[v] == 0 && mark[v] != stamp) { mark[v] = stamp; Lp.push_back(v); }
        for (int e : adjE[p]) {
            if (status[e] != 1) continue;
            for (int v : elem[e])
                if (status[v] == 0 && mark[v] != stamp) { mark[v] = stamp; Lp.push_back(v); }
            status[e] = 2;  // absorbed
            std::vector<int>().swap(elem[e]);
        }
        status[p] = 1;
        std::vector<int>().swap(adjV[p]);
        std::vector<int>().swap(adjE[p]);
        std::sort(Lp.begin(), Lp.end());
        elem[p] = Lp;
        // prune the boundary variables' lists: anything inside the new element is now represented by it
        for (int i : Lp) {
            auto &av = adjV[i];
            size_t w = 0;
            for (size_t r = 0; r < av.size(); r++)
                if (status[av[r]] == 0 && mark[av[r]] != stamp) av[w++] = av[r];
            av.resize(w);
            auto &ae = adjE[i];
            w = 0;
            for (size_t r = 0; r < ae.size(); r++)
                if (status[ae[r]] == 1 && ae[r] != p) ae[w++] = ae[r];
            ae.resize(w);
            ae.push_back(p);
        }
        // exact external degrees of the boundary variables
        for (int i : Lp) {
            stamp++;
            mark[i] = stamp;
            int cnt = 0;
            for (int v : adjV[i])
                if (mark[v] != stamp) { mark[v] = stamp; cnt++; }
            for (int e : adjE[i])
                for (int v : elem[e])
                    if (status[v] == 0 && mark[v] != stamp) { mark[v] = stamp; cnt++; }
            degree[i] = cnt;
        }
    }
}

// QDLDL_etree! (src/solver/qdldl.jl:358-395) on an upper-triangular CSC pattern; 0-based.
static long long etree_counts(int n, const std::vector<int> &Up, const std::vector<int> &Ui, std::vector<int> &etree,
                              std::vector<int> &Lnz)
{
    std::vector<int> work(n, -1);
    etree.assign(n, -1);
    Lnz.assign(n, 0);
    for (int j = 0; j < n; j++) {
        work[j] = j;
        for (int q = Up[j]; q < Up[j + 1]; q++) {
            int i = Ui[q];
            while (work[i] != j) {
                if (etree[i] < 0) etree[i] = j;
                Lnz[i]++;
                work[i] = j;
                i = etree[i];
            }
        }
    }
    long long s = 0;
    for (int i = 0; i < n; i++) s += Lnz[i];
    return s;
}

// upper CSC (columns hold rows <= col, unsorted) of P A P' for elimination order perm
static void permuted_upper(int n, const int *Ap, const int *Ai, const std::vector<int> &iperm, std::vector<int> &Up,
                           std::vector<int> &Ui)
{
    Up.assign(n + 1, 0);
    for (int j = 0; j < n; j++)
        for (int q = Ap[j]; q < Ap[j + 1]; q++) {
            int c = std::max(iperm[Ai[q]], iperm[j]);
            Up[c + 1]++;
        }
    for (int j = 0; j < n; j++) Up[j + 1] += Up[j];
    Ui.assign(Up[n], 0);
    std::vector<int> next(Up.begin(), Up.end() - 1);
    for (int j = 0; j < n; j++)
        for (int q = Ap[j]; q < Ap[j + 1]; q++) {
            int a = iperm[Ai[q]], b = iperm[j];
            Ui[next[std::max(a, b)]++] = std::min(a, b);
        }
}

const char *Symbolic::analyze_auto(int n, const int *Ap, const int *Ai, const int *user_perm, int big_task_threshold)
{
    // Plans, in order of preference: resident CTAs per SM, doubles of dynamic shared memory a CTA may plan with at that
    // residency ((233472 / c - 1 KB reserved - 4 KB static) / 8; the single-CTA figure capped by the 227 KB per-block limit),
    // supernode width cap, panel size above which the solves stream a chain panel in two column parts, threads per CTA.
    // Measured on B200 (profiles/r2_threads_per_cta_experiment.txt): the CTAs are bound by their own dependent chains, so
    // more, narrower CTAs per SM win as long as the whole numeric path stays on the shared-memory code.
    struct Plan { int ctas, budget, width, split, threads; bool leaves_first; };
    // (Plan 0 -- four CTAs of 192 threads -- is kept for experiments, CB200_PLAN=0: measured on B200 it delivers what three
    // CTAs of 256 threads deliver, because throughput follows the number of resident warps, 24 per SM either way at the
    // 80-register cap, not the number of resident instances; profiles/r2_threads_per_cta_experiment.txt.)
    // (three CTAs: 233472 / 3 = 77824 bytes per CTA - 1024 reserved - 3968 static = 72832 bytes = 9104 doubles; a plan that
    // asks for more silently runs at two CTAs per SM)
    static const Plan plans[] = {{4, 6656, 48, 2048, 192, true},  {3, 9100, 48, 2048, 256, false}, {3, 9100, 48, 2048, 256, true},
                                 {2, 13952, 48, 2048, 256, false}, {2, 13952, 48, 2048, 256, true}, {1, 28000, 48, 2048, 512, false},
                                 {1, 28000, 48, 2048, 512, true}};
    const int nplans = (int)(sizeof(plans) / sizeof(plans[0]));
    int first = 1;                                      // default: start at the three-CTA plan with x[N] resident
    if (const char *e = getenv("CB200_PLAN")) first = std::max(0, std::min(nplans - 1, atoi(e)));
    const char *msg = "";
    for (int k = first; k < nplans; k++) {
        width_cap = plans[k].width;
        part_split = plans[k].split;
        leaves_first = plans[k].leaves_first;
        msg = analyze(n, Ap, Ai, user_perm, big_task_threshold, plans[k].budget);
        if (msg[0]) return msg;
        ctas_per_sm = plans[k].ctas;
        smem_budget = plans[k].budget;
        threads = plans[k].threads;
        if (const char *e = getenv("CB200_PLAN_THREADS")) threads = atoi(e);      // tuning experiments (tools/)
        if (n_cta_tasks == 0 || (solve_smem && n_generic_cta_tasks == 0) || getenv("CB200_PLAN_ONLY")) break;
    }
    return msg;
}

const char *Symbolic::analyze(int n, const int *Ap, const int *Ai, const int *user_perm, int big_task_threshold,
                              int smem_budget_doubles)
{
    N = n;
    nnzA = Ap[n];
    for (int j = 0; j < n; j++) {
        bool diag = false;
        for (int q = Ap[j]; q < Ap[j + 1]; q++) {
            if (Ai[q] > j) return "matrix is not upper triangular";
            if (Ai[q] == j) diag = true;
        }
        if (!diag) return "upper triangle is missing a diagonal entry";
    }
    // 1. ordering
    std::vector<int> perm0;
    if (user_perm) {
        perm0.assign(user_perm, user_perm + n);
        std::vector<char> seen(n, 0);
        for (int v : perm0) {
            if (v < 0 || v >= n || seen[v]) return "perm is not a permutation";
            seen[v] = 1;
        }
    } else {
        // Two built-in orderings.  "amd" is the reference's own (amd(A), qdldl.jl:135; also available to callers through
        // cb200_amd_order + the perm argument); "mindeg" (exact external degrees) keeps the constraint multipliers of a
        // trajectory-optimisation KKT matrix as singleton leaves and gives one wide supernode per stage -- 42 levels instead
        // of 196 on BASELINE's cfg3 at 11% more fill -- which is the better shape for one-CTA-per-instance numerics and
        // therefore the default.  Results agree to rounding either way (any permutation gives the same solution).
        const char *ord = getenv("CB200_ORDERING");
        if (ord && std::string(ord) == "amd") amd_order(n, Ap, Ai, perm0);
        else minimum_degree(n, Ap, Ai, perm0);
    }
    std::vector<int> iperm0(n);
    for (int k = 0; k < n; k++) iperm0[perm0[k]] = k;
    // 2. elimination tree of the ordered matrix, postorder it (children ascending), compose
    std::vector<int> Up, Ui, et0, lnz0;
    permuted_upper(n, Ap, Ai, iperm0, Up, Ui);
    etree_counts(n, Up, Ui, et0, lnz0);
    {
        std::vector<int> head(n, -1), nextc(n, -1);
        for (int j = n - 1; j >= 0; j--)
            if (et0[j] >= 0) { nextc[j] = head[et0[j]]; head[et0[j]] = j; }
        std::vector<int> post;
        post.reserve(n);
        std::vector<int> stack;
        for (int r = 0; r < n; r++) {
            if (et0[r] >= 0) continue;
            stack.push_back(r);
            while (!stack.empty()) {
                int v = stack.back();
                int c = head[v];
                if (c >= 0) { head[v] = nextc[c]; stack.push_back(c); }
                else { post.push_back(v); stack.pop_back(); }
            }
        }
        perm.resize(n);
        for (int k = 0; k < n; k++) perm[k] = perm0[post[k]];
    }
    iperm.resize(n);
    for (int k = 0; k < n; k++) iperm[perm[k]] = k;
    permuted_upper(n, Ap, Ai, iperm, Up, Ui);
    nnzL = etree_counts(n, Up, Ui, etree, Lnz);
    nleaf = 0;
    if (leaves_first) {
        // a column is a singleton leaf iff it has no child in the elimination tree and does not start a multi-column
        // fundamental supernode (step 4 below); moving those columns to the front keeps children before parents
        std::vector<char> has_child(n, 0), leaf(n, 0);
        for (int j = 0; j < n; j++)
            if (etree[j] >= 0) has_child[etree[j]] = 1;
        for (int j = 0; j < n; j++)
            leaf[j] = !has_child[j] && !(j + 1 < n && etree[j] == j + 1 && Lnz[j] == Lnz[j + 1] + 1);
        std::vector<int> p2;
        p2.reserve(n);
        for (int j = 0; j < n; j++) if (leaf[j]) p2.push_back(perm[j]);
        nleaf = (int)p2.size();
        for (int j = 0; j < n; j++) if (!leaf[j]) p2.push_back(perm[j]);
        perm.swap(p2);
        for (int k = 0; k < n; k++) iperm[perm[k]] = k;
        permuted_upper(n, Ap, Ai, iperm, Up, Ui);
        nnzL = etree_counts(n, Up, Ui, etree, Lnz);
    }
    flops = 0;
    for (int j = 0; j < n; j++) flops += (long long)Lnz[j] * Lnz[j];
    // 3. column structures of L: struct(j) = lower(A)_j U (struct(children) \ {j})
    Lptr.assign(n + 1, 0);
    for (int j = 0; j < n; j++) Lptr[j + 1] = Lptr[j] + Lnz[j];
    Lrows.assign((size_t)Lptr[n], 0);
    {
        // lower entries by column = upper entries by row: transpose Up/Ui
        std::vector<int> cnt(n + 1, 0);
        for (int c = 0; c < n; c++)
            for (int q = Up[c]; q < Up[c + 1]; q++)
                if (Ui[q] != c) cnt[Ui[q] + 1]++;
        for (int j = 0; j < n; j++) cnt[j + 1] += cnt[j];
        std::vector<int> lowr(cnt[n]);
        std::vector<int> nx(cnt.begin(), cnt.end() - 1);
        for (int c = 0; c < n; c++)
            for (int q = Up[c]; q < Up[c + 1]; q++)
                if (Ui[q] != c) lowr[nx[Ui[q]]++] = c;
        std::vector<int> mark(n, -1), fill(n, 0);
        std::vector<std::vector<int>> children(n);
        for (int j = 0; j < n; j++)
            if (etree[j] >= 0) children[etree[j]].push_back(j);
        for (int j = 0; j < n; j++) {
            int *out = Lrows.data() + Lptr[j];
            int k = 0;
            mark[j] = j;
            for (int q = cnt[j]; q < cnt[j + 1]; q++) {
                int r = lowr[q];
                if (mark[r] != j) { mark[r] = j; out[k++] = r; }
            }
            for (int c : children[j])
                for (int q = Lptr[c]; q < Lptr[c + 1]; q++) {
                    int r = Lrows[q];
                    if (r != j && mark[r] != j) { mark[r] = j; out[k++] = r; }
                }
            if (k != Lnz[j]) return "internal error: column count mismatch";
            std::sort(out, out + k);
        }
    }
    // 4. fundamental supernodes (consecutive columns, parent = next column, nested structure) ...
    const int max_width = width_cap;    // keeps (rows + width) x width panels inside the shared-memory budget
    std::vector<int> fstart;
    for (int j = 0; j < n; j++) {
        bool merge = j > 0 && etree[j - 1] == j && Lnz[j - 1] == Lnz[j] + 1 && (j - fstart.back()) < max_width;
        if (!merge) fstart.push_back(j);
    }
    const int nf = (int)fstart.size();
    fstart.push_back(n);
    // ... then relaxed amalgamation: a supernode that has children of its own is merged into the supernode that
    // follows it when that one holds its parent and the padding (explicit zeros) stays small.  Childless supernodes
    // (the singleton leaves a minimum-degree ordering makes of the constraint rows) are never merged: they are
    // processed in bulk.  Walk right to left so that a group keeps the row structure of its rightmost member.
    std::vector<int> nchild(n, 0);
    for (int j = 0; j < n; j++)
        if (etree[j] >= 0) nchild[etree[j]]++;
    std::vector<int> gstart;   // group starts, collected right to left
    {
        int cur_start = fstart[nf - 1], cur_end = n, cur_R = Lnz[n - 1];
        long long cur_entries = 0, cur_zeros = 0;
        for (int j = cur_start; j < cur_end; j++) cur_entries += (cur_end - 1 - j) + cur_R;
        for (int f = nf - 2; f >= 0; f--) {
            int a = fstart[f], b = fstart[f + 1];
            bool can = etree[b - 1] >= cur_start && etree[b - 1] < cur_end && nchild[a] > 0 && (cur_end - a) <= max_width;
            long long add_entries = 0, add_zeros = 0;
            if (can) {
                for (int j = a; j < b; j++) {
                    long long padded = (cur_end - 1 - j) + cur_R;
                    add_entries += padded;
                    add_zeros += padded - Lnz[j];
                }
                long long z = cur_zeros + add_zeros, e = cur_entries + add_entries;
                can = z <= 16 || z * 10 <= e;   // at most 10% explicit zeros
            }
            if (can) {
                cur_start = a;
                cur_entries += add_entries;
                cur_zeros += add_zeros;
            } else {
                gstart.push_back(cur_start);
                cur_start = a; cur_end = b; cur_R = Lnz[b - 1];
                cur_entries = 0; cur_zeros = 0;
                for (int j = a; j < b; j++) cur_entries += (b - 1 - j) + cur_R;
            }
        }
        gstart.push_back(cur_start);
    }
    sn_start.assign(gstart.rbegin(), gstart.rend());
    ns = (int)sn_start.size();
    sn_start.push_back(n);
    sn_of.assign(n, 0);
    for (int s = 0; s < ns; s++)
        for (int j = sn_start[s]; j < sn_start[s + 1]; j++) sn_of[j] = s;
    rows_ptr.assign(ns + 1, 0);
    panel_off.assign(ns + 1, 0);
    max_w = max_nrow = 0;
    for (int s = 0; s < ns; s++) {
        int last = sn_start[s + 1] - 1;
        rows_ptr[s + 1] = rows_ptr[s] + Lnz[last];
    }
    rows.resize(rows_ptr[ns]);
    for (int s = 0; s < ns; s++) {
        int last = sn_start[s + 1] - 1, w = sn_start[s + 1] - sn_start[s];
        std::copy(Lrows.begin() + Lptr[last], Lrows.begin() + Lptr[last + 1], rows.begin() + rows_ptr[s]);
        int nrow = w + Lnz[last];
        long long sz = (long long)nrow * w;
        sz = (sz + 1) & ~1LL;  // 16-byte granularity
        panel_off[s + 1] = panel_off[s] + sz;
        max_w = std::max(max_w, w);
        max_nrow = std::max(max_nrow, nrow);
    }
    panel_total = panel_off[ns];
    // 5. pull-based update lists with relative indices
    std::vector<std::vector<UpdateEntry>> lists(ns);
    rel.clear();
    std::vector<long long> work(ns, 0);
    for (int d = 0; d < ns; d++) {
        const int *R = rows.data() + rows_ptr[d];
        int nR = rows_ptr[d + 1] - rows_ptr[d];
        int wd = sn_start[d + 1] - sn_start[d];
        int a = 0;
        while (a < nR) {
            int t = sn_of[R[a]];
            int b = a;
            while (b < nR && sn_of[R[b]] == t) b++;
            UpdateEntry u{d, a, b, (int)rel.size()};
            const int *Rt = rows.data() + rows_ptr[t];
            int nRt = rows_ptr[t + 1] - rows_ptr[t];
            int wt = sn_start[t + 1] - sn_start[t];
            for (int i = a; i < nR; i++) {
                int r = R[i];
                if (r < sn_start[t + 1]) rel.push_back(r - sn_start[t]);
                else {
                    const int *it = std::lower_bound(Rt, Rt + nRt, r);
                    if (it == Rt + nRt || *it != r) return "internal error: update row missing from target structure";
                    rel.push_back(wt + (int)(it - Rt));
                }
            }
            lists[t].push_back(u);
            work[t] += (long long)(nR - a) * (b - a) * wd;
            a = b;
        }
    }
    upd_ptr.assign(ns + 1, 0);
    upd.clear();
    for (int t = 0; t < ns; t++) {
        std::sort(lists[t].begin(), lists[t].end(), [](const UpdateEntry &x, const UpdateEntry &y) { return x.d < y.d; });
        for (auto &u : lists[t]) upd.push_back(u);
        upd_ptr[t + 1] = (int)upd.size();
    }
    // 6. levels and phases
    level.assign(ns, 0);
    nlevels = 0;
    for (int t = 0; t < ns; t++) {
        int l = 0;
        for (int q = upd_ptr[t]; q < upd_ptr[t + 1]; q++) l = std::max(l, level[upd[q].d] + 1);
        level[t] = l;
        nlevels = std::max(nlevels, l + 1);
        int w = sn_start[t + 1] - sn_start[t], nrow = w + rows_ptr[t + 1] - rows_ptr[t];
        work[t] += (long long)nrow * w * w / 2 + (long long)nrow * w;
    }
    order.resize(ns);
    std::iota(order.begin(), order.end(), 0);
    // class of a supernode inside its level: 0 singleton leaf (width 1, nothing to pull), 1 big (CTA), 2 small (warp)
    std::vector<char> cls(ns);
    for (int t = 0; t < ns; t++) {
        bool single_leaf = (sn_start[t + 1] - sn_start[t] == 1) && upd_ptr[t + 1] == upd_ptr[t];
        cls[t] = single_leaf ? 0 : (work[t] >= big_task_threshold ? 1 : 2);
    }
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        if (level[x] != level[y]) return level[x] < level[y];
        if (cls[x] != cls[y]) return cls[x] < cls[y];
        return x < y;
    });
    phases.clear();
    for (int i = 0; i < ns;) {
        int j = i;
        while (j < ns && level[order[j]] == level[order[i]] && cls[order[j]] == cls[order[i]]) j++;
        int mode = cls[order[i]] == 0 ? 2 : (cls[order[i]] == 1 ? 1 : 0);
        if (mode == 0 && j - i == 1) mode = 1;  // a lone small task: let the whole CTA help anyway
        phases.push_back(Phase{mode, i, j, 0, 0, -1});
        i = j;
    }
    // 6b. shared-memory staging plan for CTA-scope targets
    big_index.assign(ns, -1);
    big.clear(); ychunks.clear(); ystage_src.clear(); ystage_dst.clear(); ypiv.clear(); ymask.clear(); big_seq.clear();
    yb_row.clear(); yb_col.clear(); yb_ptr.assign(1, 0);
    dg_dst.clear(); dg_src.clear(); dg_piv.clear(); dg_ptr.assign(1, 0);
    max_sb_doubles = 0;
    kx_total = 0;
    scratch_doubles = 0;
    for (const Phase &ph : phases) {
        if (ph.mode != 1 || panel_total > 2000000000LL) continue;
        for (int q = ph.begin; q < ph.end; q++) {
            int t = order[q];
            int w = sn_start[t + 1] - sn_start[t], nrow = w + rows_ptr[t + 1] - rows_ptr[t];
            const int ldy = ((w + 7) & ~7) + 4;                           // Y holds the pivot rows only
            const int ldp = nrow + ((4 - nrow % 8) + 8) % 8;              // smallest value >= nrow that is 4 mod 8
            const int ntI = (w + 7) / 8;
            const int bottom_reserve = 64;                               // (see the fit test below)                              // scalars of below-pivot entries per chunk
            const long long fixed = (long long)ldp * w;
            // after the GEMM the Y area is reused by the panel factorisation: 8x8 pivot block, 8 reciprocals, w pivots,
            // 8 x (w rounded + 4) unscaled multipliers
            const long long aux = 64 + 8 + w + 8LL * (((w + 7) & ~7) + 4) + 8;
            int ktot = 0;
            bool bottoms_fit = true;
            for (int u = upd_ptr[t]; u < upd_ptr[t + 1]; u++) {
                const UpdateEntry &ue = upd[u];
                const int nRd = rows_ptr[ue.d + 1] - rows_ptr[ue.d];
                ktot += sn_start[ue.d + 1] - sn_start[ue.d];
                int nb = 0;
                for (int i = ue.a; i < nRd; i++) nb += rel[ue.rel + (i - ue.a)] >= w;
                if (nb > bottom_reserve) bottoms_fit = false;
            }
            if (!bottoms_fit) continue;                                  // generic path
            if (fixed + std::max<long long>(ktot > 0 ? (long long)(ldy + 1) * 8 + bottom_reserve : 0, aux) > smem_budget_doubles) continue;   // generic path
            int kc_max = 0;
            if (ktot > 0) {
                kc_max = (int)std::min<long long>((smem_budget_doubles - fixed - bottom_reserve) / (ldy + 1), 128);
                kc_max &= ~3;                                            // whole groups of 4 columns, <= 32 groups
            }
            BigTarget bt;
            bt.chunk_begin = (int)ychunks.size();
            bt.ldy = ldy;
            bt.ldp = ldp;
            bt.panel_doubles = (int)(panel_off[t + 1] - panel_off[t]);
            bt.asm_begin = bt.asm_end = 0;
            bt.h1 = w;
            if (bt.panel_doubles > part_split && w >= 4) bt.h1 = ((w + 1) / 2 + 1) & ~1;   // even => both parts 16-byte aligned
            int col = 0;
            YChunk ch{0, 0, (int)ystage_src.size(), 0, (int)ypiv.size(), 0, 0, (int)ymask.size()};
            ymask.resize(ymask.size() + ntI, 0u);
            int max_kc = 0;
            struct Bottom { int row, src, col; };
            std::vector<Bottom> bottoms;
            auto close_chunk = [&]() {
                const int kc4 = (col - ch.col_begin + 3) & ~3;
                for (int cc = 0; cc < col - ch.col_begin; cc++) {      // pivots of the Y columns: Dy[cc] = D[ypiv]
                    ystage_src.push_back(-1 - ypiv[ch.piv_begin + cc]);
                    ystage_dst.push_back(ldy * kc4 + cc);
                }
                std::stable_sort(bottoms.begin(), bottoms.end(), [](const Bottom &x, const Bottom &y) { return x.row < y.row; });
                ch.bg_begin = (int)yb_row.size();
                for (size_t i = 0; i < bottoms.size();) {
                    size_t j = i;
                    while (j < bottoms.size() && bottoms[j].row == bottoms[i].row) j++;
                    yb_row.push_back(bottoms[i].row);
                    for (size_t e = i; e < j; e++) {
                        ystage_src.push_back(bottoms[e].src);
                        ystage_dst.push_back((ldy + 1) * kc4 + (int)e);          // scalar e of the chunk, after Dy
                        yb_col.push_back(bottoms[e].col);
                    }
                    yb_ptr.push_back((int)yb_col.size());
                    i = j;
                }
                ch.bg_end = (int)yb_row.size();
                bottoms.clear();
                ch.col_end = col; ch.stage_end = (int)ystage_src.size();
                max_kc = std::max(max_kc, ch.col_end - ch.col_begin);
                ychunks.push_back(ch);
            };
            int nbottom = 0;
            std::vector<std::pair<int, std::pair<int, int>>> singles;   // (local row, (panel offset, pivot column))
            for (int u = upd_ptr[t]; u < upd_ptr[t + 1]; u++) {
                const UpdateEntry &ue = upd[u];
                int cd0 = sn_start[ue.d], wd = sn_start[ue.d + 1] - cd0;
                int nRd = rows_ptr[ue.d + 1] - rows_ptr[ue.d], nrowd = wd + nRd;
                if (nRd - ue.a == 1) {      // one row only: a diagonal update (or nothing, if the row is below the pivots)
                    const int lrow = rel[ue.rel];
                    if (lrow < w)
                        for (int k = 0; k < wd; k++)
                            singles.push_back({lrow, {(int)(panel_off[ue.d] + (wd + ue.a) + (long long)k * nrowd), cd0 + k}});
                    continue;
                }
                int nb_col = 0;
                for (int i = ue.a; i < nRd; i++) nb_col += rel[ue.rel + (i - ue.a)] >= w;
                for (int k = 0; k < wd; k++) {
                    if (col - ch.col_begin == kc_max || nbottom + nb_col > bottom_reserve) {   // chunk full
                        close_chunk();
                        ch = YChunk{col, col, (int)ystage_src.size(), 0, (int)ypiv.size(), 0, 0, (int)ymask.size()};
                        ymask.resize(ymask.size() + ntI, 0u);
                        nbottom = 0;
                    }
                    ypiv.push_back(cd0 + k);
                    for (int i = ue.a; i < nRd; i++) {
                        const int lrow = rel[ue.rel + (i - ue.a)], lcol = col - ch.col_begin;
                        const int src = (int)(panel_off[ue.d] + (wd + i) + (long long)k * nrowd);
                        if (lrow >= w) { bottoms.push_back(Bottom{lrow, src, lcol}); nbottom++; continue; }
                        ystage_src.push_back(src);
                        ystage_dst.push_back(lrow + lcol * ldy);
                        ymask[ch.mask_begin + lrow / 8] |= 1u << (lcol / 4);
                    }
                    col++;
                }
            }
            std::stable_sort(singles.begin(), singles.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
            bt.dg_begin = (int)dg_dst.size();
            for (size_t i = 0; i < singles.size();) {
                size_t j = i;
                while (j < singles.size() && singles[j].first == singles[i].first) j++;
                dg_dst.push_back(singles[i].first * (ldp + 1));
                for (size_t e = i; e < j; e++) { dg_src.push_back(singles[e].second.first); dg_piv.push_back(singles[e].second.second); }
                dg_ptr.push_back((int)dg_src.size());
                i = j;
            }
            bt.dg_end = (int)dg_dst.size();
            if (col > ch.col_begin) close_chunk();
            else ymask.resize(ch.mask_begin);
            bt.chunk_end = (int)ychunks.size();
            big_index[t] = (int)big.size();
            big.push_back(bt);
            big_seq.push_back(t);
            max_sb_doubles = std::max(max_sb_doubles, std::max(nrow * bt.h1, bt.panel_doubles - nrow * bt.h1));
            const long long ysize = (long long)((max_kc + 3) & ~3) * (ldy + 1) + bottom_reserve;
            scratch_doubles = (int)std::max<long long>(scratch_doubles, fixed + std::max<long long>(ysize, aux) + 8);
        }
    }
    bdesc.clear();
    for (size_t bi = 0; bi < big.size(); bi++) {
        const int t = big_seq[bi];
        bdesc.push_back(ChainDesc{t, sn_start[t], sn_start[t + 1] - sn_start[t], rows_ptr[t + 1] - rows_ptr[t], rows_ptr[t],
                                  (int)panel_off[t], big[bi].h1, 0});
    }
    for (Phase &ph : phases) ph.first_big = ph.mode == 1 ? big_index[order[ph.begin]] : -1;
    big_seq_bwd.assign(big_seq.rbegin(), big_seq.rend());     // the backward solve visits them in exactly the reverse order
    n_cta_tasks = n_generic_cta_tasks = 0;
    for (const Phase &ph : phases)
        if (ph.mode == 1)
            for (int q = ph.begin; q < ph.end; q++) { n_cta_tasks++; n_generic_cta_tasks += big_index[order[q]] < 0; }
    {   // shared-memory solve: the permuted vector, two part buffers (TMA double buffering), one pivot window
        max_sb_doubles = (max_sb_doubles + 1) & ~1;
        int max_nR_big = 0;
        for (int t = 0; t < ns; t++)
            if (big_index[t] >= 0) max_nR_big = std::max(max_nR_big, rows_ptr[t + 1] - rows_ptr[t]);
        if (leaves_first) {      // the leading columns must be exactly the singleton-leaf supernodes of the schedule
            int cnt = 0;
            bool ok = true;
            for (int t = 0; t < ns; t++)
                if (cls[t] == 0) { cnt++; ok = ok && sn_start[t] < nleaf; }
            if (!ok || cnt != nleaf) return "internal error: leaves-first ordering does not match the leaf classification";
        }
        long long need = (((long long)(N - nleaf) + 1) & ~1LL) + 2LL * max_sb_doubles + 64 + 16 + ((max_nR_big + 1) & ~1);
        solve_smem = (!big.empty() && need <= smem_budget_doubles) ? 1 : 0;
        if (solve_smem) scratch_doubles = (int)std::max<long long>(scratch_doubles, need);
        parts_fwd.clear(); parts_bwd.clear();
        auto parts_of = [&](int t, bool forward, std::vector<int> &out) {
            const BigTarget &bt = big[big_index[t]];
            const int w = sn_start[t + 1] - sn_start[t], nrow = w + rows_ptr[t + 1] - rows_ptr[t];
            const int off0 = (int)panel_off[t], len0 = bt.h1 == w ? bt.panel_doubles : nrow * bt.h1;
            const int off1 = off0 + len0, len1 = bt.panel_doubles - len0;
            if (bt.h1 == w) { out.push_back(off0); out.push_back(len0); }
            else if (forward) { out.insert(out.end(), {off0, len0, off1, len1}); }
            else { out.insert(out.end(), {off1, len1, off0, len0}); }
        };
        for (int t : big_seq) parts_of(t, true, parts_fwd);
        for (int t : big_seq_bwd) parts_of(t, false, parts_bwd);
    }
    // 6c. forward-solve row lists (row c of L, grouped by the supernode that stores it); singleton leaves apart,
    //     shared-memory (big) supernodes excluded: they push
    auto is_single_leaf = [&](int d) { return cls[d] == 0; };
    fwd_ptr.assign(n + 1, 0);
    lcsr_ptr.assign(n + 1, 0);
    for (int d = 0; d < ns; d++) {
        bool leaf = is_single_leaf(d);
        if (!leaf && big_index[d] >= 0) continue;   // shared-memory supernodes PUSH their forward updates
        for (int q = rows_ptr[d]; q < rows_ptr[d + 1]; q++) (leaf ? lcsr_ptr : fwd_ptr)[rows[q] + 1]++;
    }
    for (int c = 0; c < n; c++) { fwd_ptr[c + 1] += fwd_ptr[c]; lcsr_ptr[c + 1] += lcsr_ptr[c]; }
    fwd.assign(fwd_ptr[n], FwdEntry{0, 0, 0, 0});
    lcsr_col.assign(lcsr_ptr[n], 0);
    leaf_csr_pos.assign(rows.size(), -1);
    lcsr_total = lcsr_ptr[n];
    {
        std::vector<int> nx(fwd_ptr.begin(), fwd_ptr.end() - 1), nl(lcsr_ptr.begin(), lcsr_ptr.end() - 1);
        for (int d = 0; d < ns; d++) {
            int wd = sn_start[d + 1] - sn_start[d];
            int nrowd = wd + rows_ptr[d + 1] - rows_ptr[d];
            bool leaf = is_single_leaf(d);
            for (int q = rows_ptr[d]; q < rows_ptr[d + 1]; q++) {
                int c = rows[q];
                if (leaf) {
                    lcsr_col[nl[c]] = sn_start[d];
                    leaf_csr_pos[q] = nl[c];
                    nl[c]++;
                } else if (big_index[d] < 0) {
                    fwd[nx[c]] = FwdEntry{(int)panel_off[d] + wd + (q - rows_ptr[d]), sn_start[d], wd, nrowd};
                    nx[c]++;
                }
            }
        }
    }
    {   // per-phase bulk pulls into the pivot columns of the shared-memory supernodes
        std::vector<int> phase_of(ns, 0);
        for (size_t pi = 0; pi < phases.size(); pi++)
            for (int q = phases[pi].begin; q < phases[pi].end; q++) phase_of[order[q]] = (int)pi;
        struct Item { int phase, col; FwdEntry fe; };
        std::vector<Item> items;
        for (int d = 0; d < ns; d++) {
            if (is_single_leaf(d) || big_index[d] >= 0) continue;
            const int wd = sn_start[d + 1] - sn_start[d], nrowd = wd + rows_ptr[d + 1] - rows_ptr[d];
            for (int q = rows_ptr[d]; q < rows_ptr[d + 1]; q++) {
                const int c = rows[q];
                if (big_index[sn_of[c]] < 0) continue;
                items.push_back(Item{phase_of[d], c, FwdEntry{(int)panel_off[d] + wd + (q - rows_ptr[d]), sn_start[d], wd, nrowd}});
            }
        }
        std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) {
            return a.phase != b.phase ? a.phase < b.phase : a.col < b.col;
        });
        pfwd.clear(); prow.clear();
        pphase_ptr.assign(phases.size() + 1, 0);
        for (size_t i = 0; i < items.size();) {
            size_t j = i;
            while (j < items.size() && items[j].phase == items[i].phase && items[j].col == items[i].col) j++;
            prow.insert(prow.end(), {items[i].col, (int)pfwd.size(), (int)(pfwd.size() + (j - i)), 0});
            for (size_t k = i; k < j; k++) pfwd.push_back(items[k].fe);
            pphase_ptr[items[i].phase + 1]++;
            i = j;
        }
        for (size_t pi = 0; pi < phases.size(); pi++) pphase_ptr[pi + 1] += pphase_ptr[pi];
        max_big_nR = 0;
        for (int t = 0; t < ns; t++)
            if (big_index[t] >= 0) max_big_nR = std::max(max_big_nR, rows_ptr[t + 1] - rows_ptr[t]);
    }
    // chain runs: maximal sequences of consecutive CTA-scope phases whose tasks are all on the shared-memory path and
    // between which no bulk pull is scheduled.  The solves process a run without CTA-wide barriers (one warp runs the
    // triangular sweeps, the others the rectangular parts, device_core.h).  Stored in the otherwise unused fields of the
    // mode-1 phase records: ebegin = first phase of the run + 1, eend = last phase of the run + 1 (0: not in a run).
    {
        const int np = (int)phases.size();
        auto all_big = [&](int pi) {
            if (phases[pi].mode != 1) return false;
            for (int q = phases[pi].begin; q < phases[pi].end; q++)
                if (big_index[order[q]] < 0) return false;
            return true;
        };
        for (int pi = 0; pi < np;) {
            if (!all_big(pi)) { if (phases[pi].mode == 1) phases[pi].ebegin = phases[pi].eend = 0; pi++; continue; }
            int pj = pi;
            while (pj + 1 < np && all_big(pj + 1) && pphase_ptr[pj + 1] == pphase_ptr[pj]) pj++;
            for (int k = pi; k <= pj; k++) { phases[k].ebegin = pi + 1; phases[k].eend = pj + 1; }
            pi = pj + 1;
        }
    }
    lcsr_cols.clear();
    lcsr_rowinfo.clear();
    for (int c = 0; c < n; c++)
        if (lcsr_ptr[c + 1] > lcsr_ptr[c]) {
            lcsr_cols.push_back(c);
            lcsr_rowinfo.insert(lcsr_rowinfo.end(), {c, lcsr_ptr[c], lcsr_ptr[c + 1], 0});
        }
    leaf_info.assign(4 * (size_t)ns, 0);
    for (const Phase &ph : phases) {
        if (ph.mode != 2) continue;
        for (int q = ph.begin; q < ph.end; q++) {
            const int t = order[q];
            leaf_info[4 * (size_t)q + 0] = sn_start[t];
            leaf_info[4 * (size_t)q + 1] = rows_ptr[t + 1] - rows_ptr[t];
            leaf_info[4 * (size_t)q + 2] = (int)(panel_off[t] + 1);
            leaf_info[4 * (size_t)q + 3] = rows_ptr[t];
        }
    }
    // 6d. flat entry lists of the singleton-leaf phases (coalesced scaling pass of the factorisation)
    leaf_e_off.clear(); leaf_e_col.clear(); leaf_e_pos.clear();
    for (Phase &ph : phases) {
        if (ph.mode != 2) continue;
        ph.ebegin = (int)leaf_e_off.size();
        for (int q = ph.begin; q < ph.end; q++) {
            const int t = order[q];
            for (int r = rows_ptr[t]; r < rows_ptr[t + 1]; r++) {
                leaf_e_off.push_back((int)(panel_off[t] + 1 + (r - rows_ptr[t])));
                leaf_e_col.push_back(sn_start[t]);
                leaf_e_pos.push_back(leaf_csr_pos[r]);
            }
        }
        ph.eend = (int)leaf_e_off.size();
    }
    // 7. destination of every input entry in the panel storage
    dest.assign(nnzA, 0);
    for (int j = 0; j < n; j++)
        for (int q = Ap[j]; q < Ap[j + 1]; q++) {
            int a = iperm[Ai[q]], b = iperm[j];
            int c = std::min(a, b), r = std::max(a, b);
            int s = sn_of[c];
            int w = sn_start[s + 1] - sn_start[s];
            int nrow = w + rows_ptr[s + 1] - rows_ptr[s];
            int lr;
            if (r < sn_start[s + 1]) lr = r - sn_start[s];
            else {
                const int *Rs = rows.data() + rows_ptr[s];
                int nRs = rows_ptr[s + 1] - rows_ptr[s];
                const int *it = std::lower_bound(Rs, Rs + nRs, r);
                if (it == Rs + nRs || *it != r) return "internal error: entry missing from supernode structure";
                lr = w + (int)(it - Rs);
            }
            dest[q] = panel_off[s] + (long long)(c - sn_start[s]) * nrow + lr;
        }
    // 8. fused-assembly lists
    if (panel_total > 2000000000LL) return "factor storage exceeds 2^31 entries";
    {
        std::vector<int> entry_at((size_t)panel_total, -1);
        for (int q = 0; q < nnzA; q++) entry_at[(size_t)dest[q]] = q;
        leaf_e_src.assign(leaf_e_off.size(), -1);
        for (size_t e = 0; e < leaf_e_off.size(); e++) {
            leaf_e_src[e] = entry_at[(size_t)leaf_e_off[e]];
            if (leaf_e_src[e] < 0) return "internal error: leaf entry without an input entry";
        }
        leaf_piv_src.assign(ns, -1);
        basm_src.clear(); basm_dst.clear(); gasm_src.clear(); gasm_dst.clear(); gasm_zero.clear();
        std::vector<char> is_leaf(ns, 0);
        for (const Phase &ph : phases)
            if (ph.mode == 2)
                for (int q = ph.begin; q < ph.end; q++) {
                    is_leaf[order[q]] = 1;
                    leaf_piv_src[q] = entry_at[(size_t)panel_off[order[q]]];
                }
        for (int t = 0; t < ns; t++) {
            if (is_leaf[t]) continue;
            const int w = sn_start[t + 1] - sn_start[t], nrow = w + rows_ptr[t + 1] - rows_ptr[t];
            const int bi = big_index[t];
            if (bi >= 0) big[bi].asm_begin = (int)basm_src.size();
            for (int k = 0; k < w; k++)
                for (int i = 0; i < nrow; i++) {
                    const long long off = panel_off[t] + (long long)k * nrow + i;
                    const int q = entry_at[(size_t)off];
                    if (bi >= 0) {
                        if (q >= 0) { basm_src.push_back(q); basm_dst.push_back(k * big[bi].ldp + i); }
                    } else {
                        gasm_zero.push_back((int)off);
                        if (q >= 0) { gasm_src.push_back(q); gasm_dst.push_back((int)off); }
                    }
                }
            if (bi >= 0) big[bi].asm_end = (int)basm_src.size();
        }
    }
    pack_leaf_entries();
    return "";
}

}  // namespace cb200
