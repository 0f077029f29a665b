// symbolic.h -- host-side symbolic analysis for the batched supernodal LDL^T (done once per sparsity pattern).
//
// Replaces, for the B200 path, what the reference does at Solver construction (src/solver/solver.jl:88-122 ->
// ldl_solver -> qdldl(A), src/solver/qdldl.jl:134-188): fill-reducing ordering (AMD.amd in the reference, :135),
// symmetric permutation (permute_symmetric, :642-742), elimination tree + column counts (QDLDL_etree!, :358-395).
// On top of that it builds what a GPU needs and QDLDL does not have: a postordered elimination tree, fundamental
// supernodes with dense column panels, pull-based ("left-looking") update lists with relative indices, and a level
// schedule, so that every numeric phase is race-free and deterministic without atomics.
#pragma once
#include <cstdint>
#include <vector>

namespace cb200 {

struct UpdateEntry {
    int d;       // descendant supernode whose rows [a,b) of its below-diagonal structure fall in the target's columns
    int a, b;    // row range within R_d
    int rel;     // offset into Symbolic::rel: local row position (in the target panel) of rows a..|R_d|-1 of d
};

struct Phase {
    int mode;        // 0: one warp per supernode (tasks run concurrently); 1: whole CTA per supernode (sequential);
                     // 2: singleton leaves (width 1, no incoming update), one thread per supernode
    int begin, end;  // range in Symbolic::order
    int ebegin, eend; // mode 2: range in the flat leaf-entry lists (leaf_e_off / leaf_e_col / leaf_e_pos);
                      // mode 1: chain run this phase belongs to (first phase + 1, last phase + 1; 0 = none), see analyze()
    int first_big;   // mode 1: index in big[] of the phase's first task (-1: not on the shared-memory path)
};

// What the numeric phases need to know about a shared-memory supernode, 32 bytes: kept in shared memory for the whole
// kernel so that the chain steps of the solves do not chase sn_start / rows_ptr / big_index / big through global memory.
struct ChainDesc {
    int s, c0, w, nR;
    int rows_off, panel_off, h1, pad;
};

// A CTA-scope target whose panel fits in shared memory pulls ALL its descendants' columns at once: they are staged
// as the columns of a dense (rows of the target) x (sum of descendant widths) matrix Y and applied as one GEMM
// S -= Y diag(D) Y_top'.  Chunks bound the shared-memory footprint.
struct YChunk {
    int col_begin, col_end;      // Y columns of this chunk
    int stage_begin, stage_end;  // range in ystage_src / ystage_dst
    int piv_begin;               // ypiv[piv_begin + c]: pivot (permuted column) of chunk column c
    int bg_begin, bg_end;        // bottom-row groups of this chunk (entries below the pivot rows, see yb_*)
    int mask_begin;              // ymask[mask_begin + ti]: bit g set <=> rows [8ti, 8ti+8) x columns [4g, 4g+4) of the chunk
                                 // hold a structural non-zero (the tensor-core GEMM skips the other 8x4 blocks)
};
struct BigTarget {
    int chunk_begin, chunk_end;
    int ldy;                     // leading dimension of Y (pivot rows only): w rounded up to a multiple of 8, plus 4 (=> 8x4 fragment
                                 // loads of the FP64 mma hit 16 distinct shared-memory banks per half-warp)
    int ldp;                     // leading dimension of the panel work area: smallest value >= rows that is 4 mod 8 (the
                                 // tensor-core tiles may read a few rows past the panel's last row: stores are predicated)
    int panel_doubles;           // size of the supernode's panel in the factor storage (even): one TMA bulk copy in the solves
    int asm_begin, asm_end;      // range in basm_src / basm_dst: the input entries of this supernode (fused assembly)
    int dg_begin, dg_end;        // range of diagonal-update groups (descendant columns with a single entry in this target)
    int h1;                      // the solves stream the panel in one part (h1 == w) or two: columns [0, h1) and [h1, w)
};
struct FwdEntry {                // one (descendant, row) pair of the forward-solve row lists, flattened
    int off;                     // panel offset of L_d[row, 0]
    int col0, width, stride;     // first pivot column of d, its width, its panel leading dimension
};

struct Symbolic {
    int N = 0;
    int nnzA = 0;                 // entries of the input upper triangle
    std::vector<int> perm, iperm; // perm[k] = natural index eliminated k-th; iperm = inverse
    std::vector<int> etree, Lnz;  // elimination tree / strictly-lower column counts of P A P' (QDLDL Appendix-B contract)
    long long nnzL = 0;           // sum(Lnz): true fill, without supernodal padding
    long long flops = 0;          // sum(Lnz^2)
    // supernodes
    int ns = 0;
    std::vector<int> sn_start;    // [ns+1] first column (permuted index) of each supernode
    std::vector<int> sn_of;       // [N] supernode of a permuted column
    std::vector<int> rows_ptr;    // [ns+1] into rows: below-diagonal row structure R_s (permuted indices, ascending)
    std::vector<int> rows;
    std::vector<long long> panel_off; // [ns+1] offset (in doubles) of the (w + |R_s|) x w column-major panel
    long long panel_total = 0;
    int max_w = 0, max_nrow = 0;
    // pull-based updates
    std::vector<int> upd_ptr;     // [ns+1]
    std::vector<UpdateEntry> upd;
    std::vector<int> rel;
    // schedule
    std::vector<int> level;       // [ns]
    std::vector<int> order;       // supernodes sorted by (level, big-first)
    std::vector<Phase> phases;
    int nlevels = 0;
    // true (unpadded) column structure of L, for extracting the factor in QDLDL's CSC form
    std::vector<int> Lptr, Lrows;
    // forward-solve row lists: for permuted column c, the (descendant supernode, local panel row) pairs holding row c.
    // Singleton-leaf descendants are kept apart: their contributions to every ancestor are known up front (x_leaf = b)
    // and are pulled in one bulk pass from a row-ordered (CSR) copy of the leaf columns written by the factorisation.
    std::vector<int> fwd_ptr;
    std::vector<FwdEntry> fwd;       // non-leaf descendants only
    // the same contributions for the pivot columns of the shared-memory supernodes only, grouped by the PHASE of the
    // contributing descendant: pulled in bulk right after that phase, so that the chain steps do no pulling at all
    std::vector<FwdEntry> pfwd;
    std::vector<int> prow;           // 4 ints per (phase, column): column, begin and end in pfwd, 0
    std::vector<int> pphase_ptr;     // [phases + 1] range of prow rows per phase
    int max_big_nR = 0;              // largest |R| of a shared-memory supernode (row window of the backward solve)
    std::vector<int> lcsr_ptr;       // [N+1] per permuted column: range in lcsr_col / the Lcsr value array
    std::vector<int> lcsr_col;       // pivot column of the leaf
    std::vector<int> leaf_csr_pos;   // [rows.size()] for entry q of rows[] of a singleton leaf: its position in Lcsr (-1 else)
    long long lcsr_total = 0;
    std::vector<int> lcsr_cols;      // columns c with lcsr_ptr[c + 1] > lcsr_ptr[c]
    std::vector<int> lcsr_rowinfo;   // 4 ints per such column: c, lcsr_ptr[c], lcsr_ptr[c + 1], 0 (one 16-byte load)
    std::vector<int> leaf_info;      // 4 ints per position of order[] that holds a singleton leaf: pivot column, |R|,
                                     // panel offset of its first below-diagonal entry, offset of R in rows[]
    // flat list of the below-diagonal entries of the singleton leaves, grouped by phase, leaf by leaf: panel offset of
    // the entry, pivot column of its leaf, position of its row-ordered copy in Lcsr
    std::vector<int> leaf_e_off, leaf_e_col, leaf_e_pos;
    std::vector<int> leaf_e4;        // packed {offset, column, position, value code} (filled by pack_leaf_entries)
    void pack_leaf_entries()
    {
        leaf_e4.resize(4 * leaf_e_off.size());
        for (size_t e = 0; e < leaf_e_off.size(); e++) {
            leaf_e4[4 * e] = leaf_e_off[e]; leaf_e4[4 * e + 1] = leaf_e_col[e];
            leaf_e4[4 * e + 2] = leaf_e_pos[e]; leaf_e4[4 * e + 3] = leaf_e_src[e];
        }
    }
    // shared-memory path for big targets
    std::vector<int> big_index;   // [ns] index into big, or -1
    std::vector<BigTarget> big;
    std::vector<ChainDesc> bdesc; // [big.size()]
    std::vector<YChunk> ychunks;
    std::vector<int> ystage_src, ystage_dst, ypiv;
    std::vector<unsigned> ymask;
    // entries of the staged columns BELOW the pivot rows do not enter Y (they are few: Y keeps w rows only); they are
    // staged as scalars after Dy and applied per target row r as  S[r, 0:w] -= sum_e l_e Dy[c_e] Y[0:w, c_e]
    std::vector<int> yb_row, yb_ptr;          // per group: local row r >= w; [groups + 1] range of entries
    std::vector<int> yb_col;                  // per entry: column of the chunk (its scalar sits at Dy + kc4 + entry - first entry of the chunk)
    // descendant columns whose only entry in a target is one row r: they never enter Y; S[r, r] -= sum l^2 d per
    // destination group (r < w), fixed order
    std::vector<int> dg_dst, dg_ptr;          // per group: offset in the work area S; [groups + 1] range in dg_src / dg_piv
    std::vector<int> dg_src, dg_piv;          // per entry: panel offset of l, pivot column of d
    std::vector<int> big_seq;     // shared-memory supernodes in forward schedule order (TMA prefetch chain)
    std::vector<int> big_seq_bwd; // ... and in backward order (the exact reverse)
    std::vector<int> parts_fwd, parts_bwd;   // TMA copies of the solves in issue order: (panel offset, doubles) pairs
    int max_sb_doubles = 0;       // largest part (see BigTarget::h1) of a shared-memory supernode's panel
    int solve_smem = 0;           // 1: x and two solve-block buffers fit in the CTA work area (ldl_solve fast path)
    long long kx_total = 0;     // (unused: the solves read the factor panels directly)
    int scratch_doubles = 0;      // shared-memory doubles a CTA needs
    // input entry k of the upper-triangular CSC -> offset in the panel storage
    std::vector<long long> dest;
    // Fused assembly: the factorisation reads the input values itself instead of a pre-assembled panel storage.
    // Sources are input-entry indices here; HostProblem::build rewrites them into value codes (array id << 30 | offset).
    std::vector<int> leaf_e_src;             // per flat leaf entry (see leaf_e_off)
    std::vector<int> leaf_piv_src;           // [ns] per position of order[]: pivot entry of a singleton leaf
    std::vector<int> basm_src, basm_dst;     // shared-memory supernodes: (input entry, offset in the work area S)
    std::vector<int> gasm_src, gasm_dst;     // all other supernodes: (input entry, panel offset), assembled in global memory
    std::vector<int> gasm_zero;              // ... after their panels were cleared (flat list of panel offsets)
    template <class F> void map_sources(F code)
    {
        for (int &v : leaf_e_src) v = code(v);
        for (int &v : leaf_piv_src) v = v >= 0 ? code(v) : v;
        for (int &v : basm_src) v = code(v);
        for (int &v : gasm_src) v = code(v);
    }

    // Build from an upper-triangular CSC pattern (sorted rows, every diagonal entry present).
    // user_perm (may be null): caller-specified elimination order, like qdldl(A; perm=p) (qdldl.jl:134-136); it is
    // still postordered (which does not change the fill) so that supernodes are contiguous.
    // Returns an empty string on success, else an error message.
    const char *analyze(int n, const int *Ap, const int *Ai, const int *user_perm, int big_task_threshold,
                        int smem_budget_doubles = 9100);   // 72.8 KB + 4 KB static + 1 KB reserved, three CTAs per SM
    // The same with the shared-memory budget chosen for the pattern: the analysis is run for three, two and one resident
    // CTA per SM (228 KB of shared memory per SM on sm_100) and the first plan that keeps the whole numeric path on the
    // shared-memory code (solve with x[N] resident, every CTA-scope supernode staged) is kept; if none does, the last one.
    const char *analyze_auto(int n, const int *Ap, const int *Ai, const int *user_perm, int big_task_threshold);
    int ctas_per_sm = 3;          // the occupancy the kept plan was sized for
    int smem_budget = 9100;       // ... and its budget in doubles
    int threads = 256;            // ... and the threads per CTA of the heavy kernels that go with it
    int width_cap = 48;           // widest supernode the analysis forms (panel (rows + width) x width must fit the budget)
    int part_split = 2048;        // chain panels larger than this many doubles are streamed in two column parts
    bool leaves_first = false;    // order every singleton leaf before every other column (same fill, still a topological order of
                                  // the elimination tree): the solves then keep only x[nleaf .. N) in shared memory -- the leaves
                                  // are final from the start (forward) and can be written straight to the result (backward)
    int nleaf = 0;                // number of leading singleton-leaf columns when leaves_first (else 0)
    int n_cta_tasks = 0, n_generic_cta_tasks = 0;   // CTA-scope supernodes / those that fell back to the global-memory code
};

// Default ordering: approximate minimum degree (amd.cpp) -- the published AMD algorithm with SuiteSparse's conventions and
// default controls, i.e. what the reference's `perm = amd(A)` (src/solver/qdldl.jl:135) computes.  Any triangle content;
// perm[k] = index eliminated k-th.
void amd_order(int n, const int *Ap, const int *Ai, std::vector<int> &perm, double dense_factor = 10.0, bool aggressive = true);
// Alternative (CB200_ORDERING=mindeg): quotient-graph minimum degree with element absorption and exact external degrees
// (ties -> lowest index); upper-triangular pattern.
void minimum_degree(int n, const int *Ap, const int *Ai, std::vector<int> &perm);

}  // namespace cb200
