// symstats.cpp -- developer tool: print supernode / schedule statistics of the symbolic analysis for a KKT pattern.
// Build + run through tools/symstats.py.
#include <cstdio>
#include <map>
#include "../calipso_b200/csrc/host_setup.h"
using namespace cb200;
extern "C" int symstats(int n, int m, int p, int q_nn, int nsoc, const int *soc_dims, const int *Wp, const int *Wi,
                        const int *Gp, const int *Gi, const int *Cp, const int *Ci, int big_threshold, int verbose)
{
    HostProblem H;
    std::string msg = H.build(n, m, p, q_nn, nsoc, soc_dims, Wp, Wi, Gp, Gi, Cp, Ci, nullptr, big_threshold);
    if (!msg.empty()) { printf("error: %s\n", msg.c_str()); return -1; }
    const Symbolic &S = H.sym;
    printf("N %d nnzK %d nnzL %lld flops(sum Lnz^2) %lld ns %d levels %d phases %zu panel_total %lld kx_total %lld lcsr_total %lld scratch %d solve_smem %d nbig %zu\n",
           S.N, S.nnzA, S.nnzL, S.flops, S.ns, S.nlevels, S.phases.size(), S.panel_total, S.kx_total, S.lcsr_total,
           S.scratch_doubles, S.solve_smem, S.big.size());
    printf("plan: ctas/SM %d budget %d threads %d width cap %d; cta tasks %d generic %d; chunks %zu; max_sb %d\n", S.ctas_per_sm, S.smem_budget, S.threads,
           S.width_cap, S.n_cta_tasks, S.n_generic_cta_tasks, S.ychunks.size(), S.max_sb_doubles);
    long long leaf_cols = 0, leaf_rows = 0, big_cols = 0, small_cols = 0;
    std::map<int, int> leaf_hist;
    for (size_t pi = 0; pi < S.phases.size(); pi++) {
        const Phase &ph = S.phases[pi];
        long long cols = 0, rows = 0, maxw = 0, maxr = 0, ycols = 0;
        for (int q = ph.begin; q < ph.end; q++) {
            int t = S.order[q];
            int w = S.sn_start[t + 1] - S.sn_start[t], nR = S.rows_ptr[t + 1] - S.rows_ptr[t];
            cols += w; rows += nR; maxw = std::max<long long>(maxw, w); maxr = std::max<long long>(maxr, nR);
            for (int u = S.upd_ptr[t]; u < S.upd_ptr[t + 1]; u++) ycols += S.sn_start[S.upd[u].d + 1] - S.sn_start[S.upd[u].d];
            if (ph.mode == 2) { leaf_cols++; leaf_rows += nR; leaf_hist[nR]++; }
            else if (ph.mode == 1) big_cols += w; else small_cols += w;
        }
        if (verbose) printf("phase %3zu mode %d tasks %4d cols %5lld maxw %3lld maxR %3lld sumR %6lld ycols %5lld\n", pi, ph.mode, ph.end - ph.begin, cols, maxw, maxr, rows, ycols);
    }
    printf("leaf cols %lld (rows %lld) big cols %lld small cols %lld\n", leaf_cols, leaf_rows, big_cols, small_cols);
    printf("leaf nR histogram:");
    for (auto &kv : leaf_hist) printf(" %d:%d", kv.first, kv.second);
    printf("\n");
    if (verbose > 1)
        for (int t = 0; t < S.ns; t++) {
            int w = S.sn_start[t + 1] - S.sn_start[t], nR = S.rows_ptr[t + 1] - S.rows_ptr[t];
            if (w > 1 || S.upd_ptr[t + 1] > S.upd_ptr[t]) printf("sn %d c0 %d w %d nR %d level %d nupd %d big %d\n", t, S.sn_start[t], w, nR, S.level[t], S.upd_ptr[t + 1] - S.upd_ptr[t], S.big_index[t]);
        }
    return 0;
}
