// amd.cpp -- approximate minimum degree ordering for the symbolic analysis (see symbolic.h).
//
// The reference orders the reduced KKT matrix with `perm = amd(A)` (src/solver/qdldl.jl:135; AMD.jl -> SuiteSparse AMD with
// its default controls: dense-row threshold 10 sqrt(n), aggressive absorption on).  That library is not part of the
// reference tree; this is an implementation of the published algorithm (Amestoy, Davis, Duff, SIMAX 17(4), 1996; ACM TOMS
// Algorithm 837) with the same conventions -- quotient graph with elements and supervariables, approximate external
// degrees, element / aggressive absorption, mass elimination, hashed supervariable detection, LIFO degree buckets, and an
// assembly-tree postorder with the largest child last -- so that the elimination order, and with it etree, column counts
// and fill, are those of the reference's factorisation.  Checked against NVIDIA's independent implementation
// (cusolverSpXcsrsymamdHost) in tests/test_amd.py and against the oracle's separately written C version (oracle/amd.c).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "symbolic.h"

namespace cb200 {
namespace {

constexpr int NONE = -1;
inline int flip(int i) { return -i - 2; }

// Quotient graph in one growing index pool: node i owns pool[ptr[i] .. ptr[i] + len[i]); a variable's list holds its
// elements first (elen[i] of them), then its variable neighbours; an element's list holds its variables.
class QuotientGraph {
public:
    QuotientGraph(int n, const int *Ap, const int *Ai) : n_(n), ptr_(n), len_(n, 0)
    {
        // pattern of A + A' without the diagonal; the merge walks each column once, in step with the columns it meets, so
        // that every list comes out in ascending order (for symmetric patterns given as full matrix or as one triangle)
        std::vector<int> resume(n, 0);
        auto sweep = [&](auto &&emit) {
            for (int k = 0; k < n; k++) {
                int p = Ap[k];
                const int pend = Ap[k + 1];
                while (p < pend) {
                    const int j = Ai[p];
                    if (j > k) break;
                    p++;
                    if (j == k) break;
                    emit(j, k);
                    int q = resume[j];
                    const int qend = Ap[j + 1];
                    while (q < qend) {
                        const int i = Ai[q];
                        if (i > k) break;
                        q++;
                        if (i == k) break;
                        emit(i, j);
                    }
                    resume[j] = q;
                }
                resume[k] = p;
            }
            for (int j = 0; j < n; j++)
                for (int q = resume[j]; q < Ap[j + 1]; q++) emit(Ai[q], j);
        };
        sweep([&](int a, int b) { len_[a]++; len_[b]++; });
        long long total = 0;
        for (int j = 0; j < n; j++) { ptr_[j] = (int)total; total += len_[j]; }
        pool_.assign((size_t)(total + total / 5 + n + 64), 0);
        free_ = (int)total;
        std::vector<int> fill(ptr_);
        std::fill(resume.begin(), resume.end(), 0);
        sweep([&](int a, int b) { pool_[fill[a]++] = b; pool_[fill[b]++] = a; });
    }
    int n_;
    std::vector<int> ptr_, len_, pool_;
    int free_;
    void reserve(int upto)
    {
        if ((size_t)upto >= pool_.size()) pool_.resize(std::max<size_t>(pool_.size() * 2, (size_t)upto + 1024));
    }
};

struct Buckets {      // doubly linked LIFO lists by degree; during the hash phase the same heads double as hash buckets
    std::vector<int> head, next, last;
    explicit Buckets(int n) : head(n, NONE), next(n, NONE), last(n, NONE) {}
    void push(int deg, int i)
    {
        const int old = head[deg];
        if (old != NONE) last[old] = i;
        next[i] = old;
        last[i] = NONE;
        head[deg] = i;
    }
    void unlink(int i, int deg)
    {
        const int before = last[i], after = next[i];
        if (after != NONE) last[after] = before;
        if (before != NONE) next[before] = after;
        else head[deg] = after;
    }
};

}  // namespace

void amd_order(int n, const int *Ap, const int *Ai, std::vector<int> &perm, double dense_factor, bool aggressive)
{
    perm.assign(n, 0);
    if (n == 0) return;
    QuotientGraph g(n, Ap, Ai);
    std::vector<int> &ptr = g.ptr_, &len = g.len_;
    std::vector<int> nv(n, 1), elen(n, 0), degree(len), mark(n, 1);
    Buckets b(n);
    int dense = dense_factor < 0 ? n - 2 : (int)(dense_factor * std::sqrt((double)n));
    dense = std::min(n, std::max(16, dense));
    const int mark_limit = INT32_MAX - n;
    auto fresh_marks = [&](int tag) {
        if (tag < 2 || tag >= mark_limit) {
            for (int &m : mark)
                if (m != 0) m = 1;
            tag = 2;
        }
        return tag;
    };
    int tag = fresh_marks(0), eliminated = 0, mindeg = 0, largest_element = 0;
    for (int i = 0; i < n; i++) {
        const int d = degree[i];
        if (d == 0) { elen[i] = flip(1); eliminated++; ptr[i] = NONE; mark[i] = 0; }       // isolated: ordered at once
        else if (d > dense) { nv[i] = 0; elen[i] = NONE; eliminated++; ptr[i] = NONE; }    // dense: ordered last
        else b.push(d, i);
    }
    while (eliminated < n) {
        // ---- pivot of minimum approximate degree (ties: the most recently inserted)
        int pivot = NONE, d = mindeg;
        for (; d < n; d++)
            if ((pivot = b.head[d]) != NONE) break;
        mindeg = d;
        {
            const int after = b.next[pivot];
            if (after != NONE) b.last[after] = NONE;
            b.head[d] = after;
        }
        const int pivot_elements = elen[pivot];
        int pivot_size = nv[pivot];
        eliminated += pivot_size;
        nv[pivot] = -pivot_size;
        // ---- the new element: the pivot's variable neighbours and the variables of its elements (absorbed)
        int ext = 0, e_begin, e_end;
        auto take = [&](int i, int &out) {
            const int w = nv[i];
            if (w <= 0) return;
            ext += w;
            nv[i] = -w;
            g.pool_[out++] = i;
            b.unlink(i, degree[i]);
        };
        if (pivot_elements == 0) {
            e_begin = ptr[pivot];
            int out = e_begin;
            for (int p = e_begin, pe = e_begin + len[pivot]; p < pe; p++) take(g.pool_[p], out);
            e_end = out;
        } else {
            int p = ptr[pivot];
            e_begin = g.free_;
            const int own = len[pivot] - pivot_elements;
            for (int k = 0; k <= pivot_elements; k++) {
                int e, q, cnt;
                if (k == pivot_elements) { e = pivot; q = p; cnt = own; }
                else { e = g.pool_[p++]; q = ptr[e]; cnt = len[e]; }
                g.reserve(g.free_ + cnt);
                for (int c = 0; c < cnt; c++) take(g.pool_[q++], g.free_);
                if (e != pivot) { ptr[e] = flip(pivot); mark[e] = 0; }
            }
            e_end = g.free_;
        }
        degree[pivot] = ext;
        ptr[pivot] = e_begin;
        len[pivot] = e_end - e_begin;
        elen[pivot] = flip(pivot_size + ext);
        tag = fresh_marks(tag);
        // ---- scan 1: mark[e] - tag = |Le \ Lp| for every element e adjacent to a variable of the new element
        for (int s = e_begin; s < e_end; s++) {
            const int i = g.pool_[s], ne = elen[i];
            if (ne <= 0) continue;
            const int w = -nv[i], first = tag - w;
            for (int p = ptr[i], pe = ptr[i] + ne; p < pe; p++) {
                const int e = g.pool_[p];
                int m = mark[e];
                if (m >= tag) m -= w;
                else if (m != 0) m = degree[e] + first;
                mark[e] = m;
            }
        }
        // ---- scan 2: approximate degrees, absorption, mass elimination, hash keys
        for (int s = e_begin; s < e_end; s++) {
            const int i = g.pool_[s];
            const int p1 = ptr[i], p2 = p1 + elen[i];
            int out = p1, deg = 0;
            uint64_t key = 0;
            for (int p = p1; p < p2; p++) {
                const int e = g.pool_[p], m = mark[e];
                if (m == 0) continue;
                const int outside = m - tag;
                if (aggressive && outside <= 0) { ptr[e] = flip(pivot); mark[e] = 0; continue; }      // nothing outside the new element
                deg += outside;
                g.pool_[out++] = e;
                key += (uint64_t)e;
            }
            elen[i] = out - p1 + 1;
            const int vars_begin = out;
            for (int p = p2, pe = p1 + len[i]; p < pe; p++) {
                const int j = g.pool_[p], w = nv[j];
                if (w > 0) { deg += w; g.pool_[out++] = j; key += (uint64_t)j; }
            }
            if (elen[i] == 1 && vars_begin == out) {      // only the new element is left: eliminate with the pivot
                ptr[i] = flip(pivot);
                const int w = -nv[i];
                ext -= w;
                pivot_size += w;
                eliminated += w;
                nv[i] = 0;
                elen[i] = NONE;
                continue;
            }
            degree[i] = std::min(degree[i], deg);
            g.pool_[out] = g.pool_[vars_begin];
            g.pool_[vars_begin] = g.pool_[p1];
            g.pool_[p1] = pivot;
            len[i] = out - p1 + 1;
            const int h = (int)(key % (uint64_t)n);
            const int j = b.head[h];
            if (j <= NONE) { b.next[i] = flip(j); b.head[h] = flip(i); }       // bucket head kept flipped while no degree list lives there
            else { b.next[i] = b.last[j]; b.last[j] = i; }
            b.last[i] = h;
        }
        degree[pivot] = ext;
        largest_element = std::max(largest_element, ext);
        tag = fresh_marks(tag + largest_element);
        // ---- indistinguishable variables
        for (int s = e_begin; s < e_end; s++) {
            int i = g.pool_[s];
            if (nv[i] >= 0) continue;
            const int h = b.last[i];
            const int j0 = b.head[h];
            if (j0 == NONE) i = NONE;
            else if (j0 < NONE) { i = flip(j0); b.head[h] = NONE; }
            else { i = b.last[j0]; b.last[j0] = NONE; }
            while (i != NONE && b.next[i] != NONE) {
                const int ln = len[i], ne = elen[i];
                for (int p = ptr[i] + 1, pe = ptr[i] + ln; p < pe; p++) mark[g.pool_[p]] = tag;
                int prev = i, j = b.next[i];
                while (j != NONE) {
                    bool same = len[j] == ln && elen[j] == ne;
                    for (int p = ptr[j] + 1, pe = ptr[j] + ln; same && p < pe; p++) same = mark[g.pool_[p]] == tag;
                    if (same) {
                        ptr[j] = flip(i);
                        nv[i] += nv[j];
                        nv[j] = 0;
                        elen[j] = NONE;
                        j = b.next[j];
                        b.next[prev] = j;
                    } else {
                        prev = j;
                        j = b.next[j];
                    }
                }
                tag++;
                i = b.next[i];
            }
        }
        // ---- back into the degree lists; the element keeps its principal variables only
        int keep = e_begin;
        const int left = n - eliminated;
        for (int s = e_begin; s < e_end; s++) {
            const int i = g.pool_[s], w = -nv[i];
            if (w <= 0) continue;
            nv[i] = w;
            const int deg = std::min(degree[i] + ext - w, left - w);
            b.push(deg, i);
            mindeg = std::min(mindeg, deg);
            degree[i] = deg;
            g.pool_[keep++] = i;
        }
        nv[pivot] = pivot_size;
        len[pivot] = keep - e_begin;
        if (len[pivot] == 0) { ptr[pivot] = NONE; mark[pivot] = 0; }
        if (pivot_elements != 0) g.free_ = keep;
    }
    // ---- assembly tree: parent[] for elements, absorbed variables hang off their element
    std::vector<int> parent(n), size(n);
    for (int i = 0; i < n; i++) { parent[i] = flip(ptr[i]); size[i] = flip(elen[i]); }
    for (int i = 0; i < n; i++) {
        if (nv[i] != 0 || parent[i] == NONE) continue;
        int e = parent[i];
        while (nv[e] == 0) e = parent[e];
        for (int j = i; nv[j] == 0;) { const int up = parent[j]; parent[j] = e; j = up; }
    }
    // children lists (ascending index), the largest child moved to the end, depth-first numbering
    std::vector<int> child(n, NONE), sibling(n, NONE), rank(n, NONE), stack(n);
    for (int j = n - 1; j >= 0; j--)
        if (nv[j] > 0 && parent[j] != NONE) { sibling[j] = child[parent[j]]; child[parent[j]] = j; }
    for (int i = 0; i < n; i++) {
        if (nv[i] <= 0 || child[i] == NONE) continue;
        int prev = NONE, best = NONE, best_prev = NONE, best_size = NONE, tail = NONE;
        for (int f = child[i]; f != NONE; f = sibling[f]) {
            if (size[f] >= best_size) { best_size = size[f]; best_prev = prev; best = f; }
            prev = f;
            tail = f;
        }
        const int after = sibling[best];
        if (after != NONE) {
            if (best_prev == NONE) child[i] = after; else sibling[best_prev] = after;
            sibling[best] = NONE;
            sibling[tail] = best;
        }
    }
    int count = 0;
    for (int r = 0; r < n; r++) {
        if (parent[r] != NONE || nv[r] <= 0) continue;
        int top = 0;
        stack[0] = r;
        while (top >= 0) {
            const int i = stack[top];
            if (child[i] != NONE) {
                int cnt = 0;
                for (int f = child[i]; f != NONE; f = sibling[f]) cnt++;
                int h = top + cnt;
                for (int f = child[i]; f != NONE; f = sibling[f]) stack[h--] = f;
                top += cnt;
                child[i] = NONE;
            } else {
                top--;
                rank[i] = count++;
            }
        }
    }
    // ---- positions: elements in postorder, each preceded by the variables absorbed into it; dense variables last
    std::vector<int> by_rank(n, NONE), pos(n, NONE);
    for (int e = 0; e < n; e++)
        if (rank[e] != NONE) by_rank[rank[e]] = e;
    int at = 0;
    for (int k = 0; k < n && by_rank[k] != NONE; k++) { pos[by_rank[k]] = at; at += nv[by_rank[k]]; }
    for (int i = 0; i < n; i++) {
        if (nv[i] != 0) continue;
        const int e = parent[i];
        if (e != NONE) pos[i] = pos[e]++;
        else pos[i] = at++;
    }
    for (int i = 0; i < n; i++) perm[pos[i]] = i;
}

}  // namespace cb200
