// host_setup.h -- host-side construction of the shared (per-pattern) problem description: cone layout, row views of
// W/G/C, the pattern of the reduced KKT matrix K in the reference's (x, y, z) order
// (src/solver/residual_jacobian_variables.jl:110-167, upper triangle as kept by triu!, linear_solver.jl:23), its
// symbolic factorisation and the assembly destinations.  Pure C++; api.cu uploads the vectors to the device.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "device_core.h"
#include "symbolic.h"

namespace cb200 {

struct HostProblem {
    int n = 0, m = 0, p = 0, N = 0, total = 0, q_nn = 0, nsoc = 0, tri_total = 0, nnzW = 0, nnzG = 0, nnzC = 0;
    std::vector<int> soc_off, soc_d, soc_tri, Wp, Wi, Wdiag, Wfp, Wfc, Wfs, Gp, Gi, Cp, Ci, Grp, Gcj, Gsrc, Crp, Ccj,
        Csrc, Kp, Ki;
    Symbolic sym;

    static void csr_view(int nrows, int ncols, const int *cp, const int *ri, std::vector<int> &rp, std::vector<int> &cj,
                         std::vector<int> &src)
    {
        int nnz = cp[ncols];
        rp.assign(nrows + 1, 0);
        cj.assign(nnz, 0);
        src.assign(nnz, 0);
        for (int k = 0; k < nnz; k++) rp[ri[k] + 1]++;
        for (int i = 0; i < nrows; i++) rp[i + 1] += rp[i];
        std::vector<int> nx(rp.begin(), rp.end() - 1);
        for (int j = 0; j < ncols; j++)
            for (int k = cp[j]; k < cp[j + 1]; k++) {
                cj[nx[ri[k]]] = j;
                src[nx[ri[k]]] = k;
                nx[ri[k]]++;
            }
    }

    // returns "" or an error message
    std::string build(int n_, int m_, int p_, int q_nn_, int nsoc_, const int *soc_dims, const int *Wp_, const int *Wi_,
                      const int *Gp_, const int *Gi_, const int *Cp_, const int *Ci_, const int *perm, int big_threshold)
    {
        n = n_; m = m_; p = p_; N = n + m + p; total = n + 2 * m + 3 * p; q_nn = q_nn_; nsoc = nsoc_;
        if (n <= 0 || m < 0 || p < 0 || q_nn < 0 || nsoc < 0) return "invalid dimensions";
        nnzW = Wp_[n]; nnzG = Gp_[n]; nnzC = Cp_[n];
        Wp.assign(Wp_, Wp_ + n + 1); Wi.assign(Wi_, Wi_ + nnzW);
        Gp.assign(Gp_, Gp_ + n + 1); Gi.assign(Gi_, Gi_ + nnzG);
        Cp.assign(Cp_, Cp_ + n + 1); Ci.assign(Ci_, Ci_ + nnzC);
        soc_off.assign(nsoc, 0); soc_d.assign(soc_dims, soc_dims + nsoc); soc_tri.assign(nsoc, 0);
        int off = q_nn, tri = 0;
        for (int k = 0; k < nsoc; k++) {
            if (soc_d[k] < 0) return "negative cone dimension";
            soc_off[k] = off; soc_tri[k] = tri; off += soc_d[k]; tri += soc_d[k] * (soc_d[k] + 1) / 2;
        }
        if (off != p) return "cone dimensions do not sum to p";
        tri_total = tri;
        Wdiag.assign(n, -1);
        for (int j = 0; j < n; j++)
            for (int k = Wp[j]; k < Wp[j + 1]; k++) {
                if (Wi[k] < 0 || Wi[k] > j) return "W must be given as its upper triangle";
                if (Wi[k] == j) Wdiag[j] = k;
            }
        for (int j = 0; j < n; j++)
            if (Wdiag[j] < 0) return "W pattern must contain every diagonal entry";
        for (int k = 0; k < nnzG; k++) if (Gi[k] < 0 || Gi[k] >= m) return "G row index out of range";
        for (int k = 0; k < nnzC; k++) if (Ci[k] < 0 || Ci[k] >= p) return "C row index out of range";
        // symmetric W by rows
        Wfp.assign(n + 1, 0);
        for (int j = 0; j < n; j++)
            for (int k = Wp[j]; k < Wp[j + 1]; k++) { Wfp[Wi[k] + 1]++; if (Wi[k] != j) Wfp[j + 1]++; }
        for (int i = 0; i < n; i++) Wfp[i + 1] += Wfp[i];
        Wfc.assign(Wfp[n], 0); Wfs.assign(Wfp[n], 0);
        {
            std::vector<int> nx(Wfp.begin(), Wfp.end() - 1);
            for (int j = 0; j < n; j++)
                for (int k = Wp[j]; k < Wp[j + 1]; k++) {
                    int i = Wi[k];
                    Wfc[nx[i]] = j; Wfs[nx[i]] = k; nx[i]++;
                    if (i != j) { Wfc[nx[j]] = i; Wfs[nx[j]] = k; nx[j]++; }
                }
        }
        csr_view(m, n, Gp.data(), Gi.data(), Grp, Gcj, Gsrc);
        csr_view(p, n, Cp.data(), Ci.data(), Crp, Ccj, Csrc);
        // pattern of K: W columns, rows of G + diagonal, rows of C + cone block (rows above the diagonal) + diagonal
        std::vector<int> eW(nnzW), eG(nnzG), eC(nnzC), eY(m), eZnn(q_nn), eZsoc(tri);
        Kp.assign(N + 1, 0);
        Ki.clear();
        for (int j = 0; j < n; j++) {
            Kp[j] = (int)Ki.size();
            for (int k = Wp[j]; k < Wp[j + 1]; k++) { eW[k] = (int)Ki.size(); Ki.push_back(Wi[k]); }
        }
        for (int i = 0; i < m; i++) {
            Kp[n + i] = (int)Ki.size();
            for (int k = Grp[i]; k < Grp[i + 1]; k++) { eG[Gsrc[k]] = (int)Ki.size(); Ki.push_back(Gcj[k]); }
            eY[i] = (int)Ki.size(); Ki.push_back(n + i);
        }
        {
            int sk = 0, t = 0;
            for (int i = 0; i < p; i++) {
                Kp[n + m + i] = (int)Ki.size();
                for (int k = Crp[i]; k < Crp[i + 1]; k++) { eC[Csrc[k]] = (int)Ki.size(); Ki.push_back(Ccj[k]); }
                if (i < q_nn) { eZnn[i] = (int)Ki.size(); Ki.push_back(n + m + i); }
                else {
                    while (sk < nsoc && i >= soc_off[sk] + soc_d[sk]) sk++;
                    // entries of column i: rows soc_off..i; stored per cone column by column (= soc_tri order)
                    for (int r = soc_off[sk]; r <= i; r++) { eZsoc[t++] = (int)Ki.size(); Ki.push_back(n + m + r); }
                }
            }
        }
        Kp[N] = (int)Ki.size();
        const char *msg = sym.analyze_auto(N, Kp.data(), Ki.data(), perm, big_threshold);
        if (msg[0]) return msg;
        // value codes of the fused assembly: array id << 30 | offset with arrays 0 = W values, 1 = G values,
        // 2 = C values, 3 = computed entries kx = [W diagonal + eps_p (n) | y diagonal (m) | nonnegative z diagonal (q_nn) |
        // second-order z blocks (tri_total)]
        {
            std::vector<int> code(Ki.size(), -1);
            for (int k = 0; k < nnzW; k++) code[eW[k]] = k;
            for (int j = 0; j < n; j++) code[eW[Wdiag[j]]] = (int)((3u << 30) | (unsigned)j);
            {   // G values are read from the row-ordered copy Gr: position of CSC entry k in the row view
                std::vector<int> rowpos(nnzG);
                for (int r = 0; r < nnzG; r++) rowpos[Gsrc[r]] = r;
                for (int k = 0; k < nnzG; k++) code[eG[k]] = (1 << 30) | rowpos[k];
            }
            for (int k = 0; k < nnzC; k++) code[eC[k]] = (int)((2u << 30) | (unsigned)k);
            for (int i = 0; i < m; i++) code[eY[i]] = (int)((3u << 30) | (unsigned)(n + i));
            for (int i = 0; i < q_nn; i++) code[eZnn[i]] = (int)((3u << 30) | (unsigned)(n + m + i));
            for (int t = 0; t < tri; t++) code[eZsoc[t]] = (int)((3u << 30) | (unsigned)(n + m + q_nn + t));
            sym.map_sources([&](int e) { return code[e]; });
            sym.pack_leaf_entries();
            sym.kx_total = (long long)n + m + q_nn + tri;
        }
        return "";
    }
};

// L in QDLDL's CSC form (true structure, explicit padding zeros dropped) from the supernodal panels of one instance
inline void extract_factor(const Symbolic &S, const double *pan, int *Lp, int *Li, double *Lx)
{
    Lp[0] = 0;
    for (int s = 0; s < S.ns; s++) {
        const int c0 = S.sn_start[s], c1 = S.sn_start[s + 1], w = c1 - c0;
        const int nR = S.rows_ptr[s + 1] - S.rows_ptr[s], nrow = w + nR;
        const int *Rs = S.rows.data() + S.rows_ptr[s];
        const double *Ps = pan + S.panel_off[s];
        for (int c = c0; c < c1; c++) {
            int k = Lp[c];
            for (int q = S.Lptr[c]; q < S.Lptr[c + 1]; q++) {
                const int r = S.Lrows[q];
                const int lr = r < c1 ? r - c0 : w + (int)(std::lower_bound(Rs, Rs + nR, r) - Rs);
                Li[k] = r;
                Lx[k] = Ps[lr + (long long)(c - c0) * nrow];
                k++;
            }
            Lp[c + 1] = k;
        }
    }
}

// Fill the pointer fields of a DevProblem from any provider `up(vector) -> const T*` (device upload or host data()).
template <class Up> void fill_symbolic(DevProblem &P, const Symbolic &S, Up up)
{
    P.N = S.N; P.ns = S.ns; P.nphases = (int)S.phases.size(); P.max_w = S.max_w; P.max_nrow = S.max_nrow;
    P.panel_total = S.panel_total;
    P.perm = up(S.perm); P.sn_start = up(S.sn_start); P.rows_ptr = up(S.rows_ptr); P.rows = up(S.rows);
    P.panel_off = up(S.panel_off); P.upd_ptr = up(S.upd_ptr); P.upd = up(S.upd); P.rel = up(S.rel);
    P.order = up(S.order); P.phases = up(S.phases); P.fwd_ptr = up(S.fwd_ptr); P.fwd = up(S.fwd);
    P.big_index = up(S.big_index); P.big = up(S.big); P.bdesc = up(S.bdesc); P.ychunks = up(S.ychunks);
    P.ystage_src = up(S.ystage_src); P.ystage_dst = up(S.ystage_dst); P.ypiv = up(S.ypiv); P.ymask = up(S.ymask);
    P.yb_row = up(S.yb_row); P.yb_ptr = up(S.yb_ptr); P.yb_col = up(S.yb_col);
    P.dg_dst = up(S.dg_dst); P.dg_ptr = up(S.dg_ptr); P.dg_src = up(S.dg_src); P.dg_piv = up(S.dg_piv);
    P.kx_total = S.kx_total;
    P.big_seq = up(S.big_seq); P.big_seq_bwd = up(S.big_seq_bwd); P.nbig = (int)S.big_seq.size(); P.max_sb_doubles = S.max_sb_doubles;
    P.solve_smem = S.solve_smem; P.nleaf = S.leaves_first ? S.nleaf : 0;
    P.pfwd = up(S.pfwd); P.prow = up(S.prow); P.pphase_ptr = up(S.pphase_ptr); P.max_big_nR = S.max_big_nR;
    P.parts_fwd = up(S.parts_fwd); P.parts_bwd = up(S.parts_bwd);
    P.nparts_fwd = (int)S.parts_fwd.size() / 2; P.nparts_bwd = (int)S.parts_bwd.size() / 2;
    P.lcsr_ptr = up(S.lcsr_ptr); P.lcsr_col = up(S.lcsr_col); P.leaf_csr_pos = up(S.leaf_csr_pos);
    P.lcsr_total = S.lcsr_total;
    P.lcsr_cols = up(S.lcsr_cols); P.lcsr_ncols = (int)S.lcsr_cols.size();
    P.lcsr_rowinfo = up(S.lcsr_rowinfo); P.leaf_info = up(S.leaf_info);
    P.leaf_e_src = up(S.leaf_e_src); P.leaf_piv_src = up(S.leaf_piv_src); P.leaf_e4 = up(S.leaf_e4);
    P.basm_src = up(S.basm_src); P.basm_dst = up(S.basm_dst);
    P.gasm_src = up(S.gasm_src); P.gasm_dst = up(S.gasm_dst); P.gasm_zero = up(S.gasm_zero);
    P.n_gasm = (int)S.gasm_src.size(); P.n_gasm_zero = (int)S.gasm_zero.size();
    P.leaf_e_off = up(S.leaf_e_off); P.leaf_e_col = up(S.leaf_e_col); P.leaf_e_pos = up(S.leaf_e_pos);
}

// evaluate!'s scatter of the flat derivative caches (src/solver/evaluate.jl:37-42,55-60,73-78,95-100,109-114) turned
// into a gather.  The reference writes cache entry i to the dense matrix at sparsity[i] with `=`: where keys repeat
// (stage sparsities overlap: trajectory_optimization/methods.jl:26,41) the LAST entry wins; absent keys stay 0.0.  For
// `which` = W the caches are (objective, equality-dual, cone-dual) and the three matrices are then added in that order
// (residual_jacobian_variables.jl:11-13); G and C have one cache.  The plan holds, per cache and per entry of the
// pattern, the position (in the concatenated caches) of the last entry with that key, or -1.  Keys of W below the
// diagonal are dropped (the path reads the upper triangle only, linear_solver.jl:23); any other key outside the
// pattern is an error.
struct ScatterPlan {
    int ncaches = 0, nnz = 0;
    long long cache_total = 0;
    std::vector<int> idx;      // [ncaches][nnz]

    std::string build(int nrows, int ncols, const int *cp, const int *ri, bool upper_only, int ncaches_,
                      const int *cache_len, const int *rows, const int *cols)
    {
        ncaches = ncaches_; nnz = cp[ncols]; cache_total = 0;
        if (ncaches < 1 || ncaches > 3) return "scatter plan: 1 to 3 caches";
        idx.assign((size_t)ncaches * nnz, -1);
        long long off = 0;
        for (int c = 0; c < ncaches; c++) {
            if (cache_len[c] < 0) return "scatter plan: negative cache length";
            for (int e = 0; e < cache_len[c]; e++) {
                const int r = rows[off + e], col = cols[off + e];
                if (r < 0 || r >= nrows || col < 0 || col >= ncols) return "scatter plan: key out of range";
                if (upper_only && r > col) continue;
                const int *lo = ri + cp[col], *hi = ri + cp[col + 1];
                const int *it = std::lower_bound(lo, hi, r);
                if (it == hi || *it != r) return "scatter plan: key (" + std::to_string(r) + "," + std::to_string(col) + ") is not in the pattern";
                idx[(size_t)c * nnz + (it - ri)] = (int)(off + e);      // later entries overwrite: last write wins
            }
            off += cache_len[c];
        }
        if (off > 0x7fffffffLL) return "scatter plan: caches too long";
        cache_total = off;
        return "";
    }
};

// cb200_stage_plan: the index list of a stage-level scatter (dst[k] = output entry that concatenated cache entry k goes to,
// in the reference's program order) turned into per-output gather lists that keep that order.
struct StagePlan {
    int nout = 0, accumulate = 0;
    long long cache_total = 0;
    std::vector<int> ptr, src;     // [nout + 1], [cache_total]

    std::string build(int nout_, int accumulate_, int count, const int *dst)
    {
        if (count < 0) return "stage plan: negative length";
        nout = nout_; accumulate = accumulate_; cache_total = count;
        ptr.assign((size_t)nout + 1, 0);
        for (int k = 0; k < count; k++) {
            if (dst[k] < 0 || dst[k] >= nout) return "stage plan: index out of range";
            ptr[(size_t)dst[k] + 1]++;
        }
        for (int i = 0; i < nout; i++) ptr[(size_t)i + 1] += ptr[(size_t)i];
        src.assign((size_t)count, 0);
        std::vector<int> next(ptr.begin(), ptr.end() - 1);
        for (int k = 0; k < count; k++) src[(size_t)next[(size_t)dst[k]]++] = k;      // ascending k within an output: program order
        return "";
    }
};

template <class Up> void fill_problem(DevProblem &P, const HostProblem &H, Up up)
{
    P.n = H.n; P.m = H.m; P.p = H.p; P.total = H.total; P.q_nn = H.q_nn; P.nsoc = H.nsoc; P.tri_total = H.tri_total;
    P.nnzW = H.nnzW; P.nnzG = H.nnzG; P.nnzC = H.nnzC; P.nnzWf = (int)H.Wfc.size();
    fill_symbolic(P, H.sym, up);
    P.soc_off = up(H.soc_off); P.soc_dims = up(H.soc_d); P.soc_tri = up(H.soc_tri);
    P.Wp = up(H.Wp); P.Wi = up(H.Wi); P.Wdiag = up(H.Wdiag);
    P.Wfull = Csr{up(H.Wfp), up(H.Wfc), up(H.Wfs)};
    P.Gp = up(H.Gp); P.Gi = up(H.Gi); P.Cp = up(H.Cp); P.Ci = up(H.Ci);
    P.Grow = Csr{up(H.Grp), up(H.Gcj), up(H.Gsrc)};
    P.Crow = Csr{up(H.Crp), up(H.Ccj), up(H.Csrc)};
    P.nnzA = 0;
}

}  // namespace cb200
