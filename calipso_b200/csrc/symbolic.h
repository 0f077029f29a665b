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
    int mode;        // 0: one warp per supernode (tasks run concurrently); 1: whole CTA per supernode (sequential)
    int begin, end;  // range in Symbolic::order
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
    // forward-solve row lists: for permuted column c, the (descendant supernode, local panel row) pairs holding row c
    std::vector<int> fwd_ptr, fwd_d, fwd_row;
    // input entry k of the upper-triangular CSC -> offset in the panel storage
    std::vector<long long> dest;

    // Build from an upper-triangular CSC pattern (sorted rows, every diagonal entry present).
    // user_perm (may be null): caller-specified elimination order, like qdldl(A; perm=p) (qdldl.jl:134-136); it is
    // still postordered (which does not change the fill) so that supernodes are contiguous.
    // Returns an empty string on success, else an error message.
    const char *analyze(int n, const int *Ap, const int *Ai, const int *user_perm, int big_task_threshold);
};

// Approximate-minimum-degree stand-in: quotient-graph minimum degree with element absorption and exact external
// degrees (ties -> lowest index).  Not SuiteSparse AMD (absent here; any fill-reducing order gives the same solution
// to rounding).
void minimum_degree(int n, const int *Ap, const int *Ai, std::vector<int> &perm);

}  // namespace cb200
