/* amd.c -- ORACLE (test infrastructure only): restatement of the approximate-minimum-degree ordering the reference obtains
 * from a third-party dependency, `perm = amd(A)` at src/solver/qdldl.jl:135 (module AMD.jl, compat 0.4 in the reference's
 * Project.toml:19, a wrapper of SuiteSparse AMD's amd_l_order with the default controls dense = 10, aggressive = 1).
 *
 * The dependency is NOT in the reference tree and neither Julia nor SuiteSparse exist in this environment, so this file
 * restates the PUBLISHED algorithm (Amestoy, Davis, Duff, "An approximate minimum degree ordering algorithm", SIMAX 17(4)
 * 1996; "Algorithm 837: AMD", ACM TOMS 30(3) 2004) in the structure of its reference implementation:
 *   amd_order  -> pattern of A + A' without the diagonal (amd_aat / amd_1: every adjacency list ends up in ascending order
 *                 for a matrix with symmetric pattern, given as full matrix or as one triangle)
 *              -> amd_2: quotient-graph elimination with element absorption, aggressive absorption, mass elimination,
 *                 approximate external degrees (Scan 1 / Scan 2 bound), supervariable detection by hashing, degree lists
 *                 as LIFO stacks (ties: most recently inserted first), dense rows (degree > max(16, 10 sqrt(n))) last
 *              -> assembly-tree post-ordering (children by increasing size, largest last) -> permutation.
 * Parity statement: "bit-exact permutation" can only be claimed relative to this restatement; it is cross-checked on the
 * GPU box against NVIDIA's independent implementation of the same published algorithm (cusolverSpXcsrsymamdHost,
 * tests/test_amd.py) and, on CPU, against the product's separately written C++ version (calipso_b200/csrc/amd.cpp).
 *
 * The workspace grows instead of being garbage-collected: compaction preserves the order of every list, so the result
 * is the one the in-place algorithm gives. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EMPTY (-1)
#define FLIP(i) (-(i) - 2)

typedef struct {
    int *Iw;
    long long iwlen;
} amd_space;

static int grow(amd_space *s, long long need)
{
    if (need < s->iwlen) return 1;
    long long nl = s->iwlen * 2 > need + 1024 ? s->iwlen * 2 : need + 1024;
    int *p = (int *)realloc(s->Iw, (size_t)nl * sizeof(int));
    if (!p) return 0;
    s->Iw = p;
    s->iwlen = nl;
    return 1;
}

static int clear_flag(int wflg, int wbig, int *W, int n)
{
    if (wflg < 2 || wflg >= wbig) {
        for (int x = 0; x < n; x++)
            if (W[x] != 0) W[x] = 1;
        wflg = 2;
    }
    return wflg;
}

static int post_tree(int root, int k, int *Child, const int *Sibling, int *Order, int *Stack)
{
    int head = 0;
    Stack[0] = root;
    while (head >= 0) {
        int i = Stack[head];
        if (Child[i] != EMPTY) {
            /* push the children in reverse order: the first of the list is popped first, the biggest (last) one last */
            for (int f = Child[i]; f != EMPTY; f = Sibling[f]) head++;
            int h = head;
            for (int f = Child[i]; f != EMPTY; f = Sibling[f]) Stack[h--] = f;
            Child[i] = EMPTY;
        } else {
            head--;
            Order[i] = k++;
        }
    }
    return k;
}

static void postorder(int nn, const int *Parent, const int *Nv, const int *Fsize, int *Order, int *Child, int *Sibling,
                      int *Stack)
{
    for (int j = 0; j < nn; j++) { Child[j] = EMPTY; Sibling[j] = EMPTY; }
    for (int j = nn - 1; j >= 0; j--)
        if (Nv[j] > 0) {
            int parent = Parent[j];
            if (parent != EMPTY) { Sibling[j] = Child[parent]; Child[parent] = j; }
        }
    /* the largest child goes last in every child list */
    for (int i = 0; i < nn; i++)
        if (Nv[i] > 0 && Child[i] != EMPTY) {
            int fprev = EMPTY, maxfrsize = EMPTY, bigfprev = EMPTY, bigf = EMPTY;
            for (int f = Child[i]; f != EMPTY; f = Sibling[f]) {
                int frsize = Fsize[f];
                if (frsize >= maxfrsize) { maxfrsize = frsize; bigfprev = fprev; bigf = f; }
                fprev = f;
            }
            int fnext = Sibling[bigf];
            if (fnext != EMPTY) {
                if (bigfprev == EMPTY) Child[i] = fnext; else Sibling[bigfprev] = fnext;
                Sibling[bigf] = EMPTY;
                Sibling[fprev] = bigf;
            }
        }
    for (int i = 0; i < nn; i++) Order[i] = EMPTY;
    int k = 0;
    for (int i = 0; i < nn; i++)
        if (Parent[i] == EMPTY && Nv[i] > 0) k = post_tree(i, k, Child, Sibling, Order, Stack);
}

/* Ap, Ai: CSC pattern of a square matrix, sorted rows, no duplicates (any triangle content; the ordering is computed for
 * the pattern of A + A').  P receives the permutation: P[k] = index eliminated k-th.  dense < 0: no dense-row removal.
 * Returns 0 on success. */
int orc_amd_order(int n, const int *Ap, const int *Ai, int *P, double dense_ctl, int aggressive)
{
    if (n <= 0) return 0;
    int *Len = (int *)calloc((size_t)n, sizeof(int)), *Tp = (int *)malloc((size_t)n * sizeof(int));
    int *Pe = (int *)malloc((size_t)n * sizeof(int)), *Sp = (int *)malloc((size_t)n * sizeof(int));
    int *ws = (int *)malloc((size_t)7 * n * sizeof(int));
    if (!Len || !Tp || !Pe || !Sp || !ws) return -1;
    int *Nv = ws, *Next = ws + n, *Last = ws + 2 * n, *Head = ws + 3 * n, *Elen = ws + 4 * n, *Degree = ws + 5 * n, *W = ws + 6 * n;
    /* ---- amd_aat: column counts of A + A' without the diagonal */
    long long nzaat = 0;
    for (int k = 0; k < n; k++) {
        int p1 = Ap[k], p2 = Ap[k + 1], p;
        for (p = p1; p < p2;) {
            int j = Ai[p];
            if (j < k) { Len[j]++; Len[k]++; p++; }
            else if (j == k) { p++; break; }
            else break;
            int pj2 = Ap[j + 1], pj;
            for (pj = Tp[j]; pj < pj2;) {
                int i = Ai[pj];
                if (i < k) { Len[i]++; Len[j]++; pj++; }
                else if (i == k) { pj++; break; }
                else break;
            }
            Tp[j] = pj;
        }
        Tp[k] = p;
    }
    for (int j = 0; j < n; j++)
        for (int pj = Tp[j]; pj < Ap[j + 1]; pj++) { Len[Ai[pj]]++; Len[j]++; }
    for (int k = 0; k < n; k++) nzaat += Len[k];
    amd_space sp;
    sp.iwlen = nzaat + nzaat / 5 + n + 1024;
    sp.Iw = (int *)malloc((size_t)sp.iwlen * sizeof(int));
    if (!sp.Iw) return -1;
    /* ---- amd_1: the adjacency lists */
    long long pfree = 0;
    for (int j = 0; j < n; j++) { Pe[j] = (int)pfree; Sp[j] = (int)pfree; pfree += Len[j]; }
    {
        int *Iw = sp.Iw;
        for (int k = 0; k < n; k++) {
            int p1 = Ap[k], p2 = Ap[k + 1], p;
            for (p = p1; p < p2;) {
                int j = Ai[p];
                if (j < k) { Iw[Sp[j]++] = k; Iw[Sp[k]++] = j; p++; }
                else if (j == k) { p++; break; }
                else break;
                int pj2 = Ap[j + 1], pj;
                for (pj = Tp[j]; pj < pj2;) {
                    int i = Ai[pj];
                    if (i < k) { Iw[Sp[i]++] = j; Iw[Sp[j]++] = i; pj++; }
                    else if (i == k) { pj++; break; }
                    else break;
                }
                Tp[j] = pj;
            }
            Tp[k] = p;
        }
        for (int j = 0; j < n; j++)
            for (int pj = Tp[j]; pj < Ap[j + 1]; pj++) { int i = Ai[pj]; Iw[Sp[i]++] = j; Iw[Sp[j]++] = i; }
    }
    /* Tp was read before being written in the first pass of amd_aat only at Tp[j], j < k: already set (Tp[k] = p) */
    /* ---- amd_2 */
    int dense;
    if (dense_ctl < 0) dense = n - 2;
    else dense = (int)(dense_ctl * sqrt((double)n));
    if (dense < 16) dense = 16;
    if (dense > n) dense = n;
    const int wbig = 2147483647 - n;
    int nel = 0, mindeg = 0, lemax = 0, ndense = 0, me = EMPTY;
    for (int i = 0; i < n; i++) {
        Last[i] = EMPTY; Head[i] = EMPTY; Next[i] = EMPTY; Nv[i] = 1; W[i] = 1; Elen[i] = 0; Degree[i] = Len[i];
    }
    int wflg = clear_flag(0, wbig, W, n);
    for (int i = 0; i < n; i++) {
        int deg = Degree[i];
        if (deg == 0) { Elen[i] = FLIP(1); nel++; Pe[i] = EMPTY; W[i] = 0; }
        else if (deg > dense) { ndense++; Nv[i] = 0; Elen[i] = EMPTY; nel++; Pe[i] = EMPTY; }
        else {
            int inext = Head[deg];
            if (inext != EMPTY) Last[inext] = i;
            Next[i] = inext;
            Head[deg] = i;
        }
    }
    while (nel < n) {
        int deg;
        for (deg = mindeg; deg < n; deg++) { me = Head[deg]; if (me != EMPTY) break; }
        mindeg = deg;
        int inext = Next[me];
        if (inext != EMPTY) Last[inext] = EMPTY;
        Head[deg] = inext;
        const int elenme = Elen[me];
        int nvpiv = Nv[me];
        nel += nvpiv;
        /* ---- construct the new element */
        Nv[me] = -nvpiv;
        int degme = 0;
        long long pme1, pme2;
        int *Iw = sp.Iw;
        if (elenme == 0) {
            pme1 = Pe[me];
            pme2 = pme1 - 1;
            for (long long p = pme1; p <= pme1 + Len[me] - 1; p++) {
                int i = Iw[p], nvi = Nv[i];
                if (nvi > 0) {
                    degme += nvi;
                    Nv[i] = -nvi;
                    Iw[++pme2] = i;
                    int ilast = Last[i];
                    inext = Next[i];
                    if (inext != EMPTY) Last[inext] = ilast;
                    if (ilast != EMPTY) Next[ilast] = inext; else Head[Degree[i]] = inext;
                }
            }
        } else {
            long long p = Pe[me];
            pme1 = pfree;
            const int slenme = Len[me] - elenme;
            for (int knt1 = 1; knt1 <= elenme + 1; knt1++) {
                int e, ln;
                long long pj;
                if (knt1 > elenme) { e = me; pj = p; ln = slenme; }
                else { e = Iw[p++]; pj = Pe[e]; ln = Len[e]; }
                for (int knt2 = 1; knt2 <= ln; knt2++) {
                    int i = Iw[pj++], nvi = Nv[i];
                    if (nvi > 0) {
                        if (pfree >= sp.iwlen) { if (!grow(&sp, pfree + 1)) return -1; Iw = sp.Iw; }
                        degme += nvi;
                        Nv[i] = -nvi;
                        Iw[pfree++] = i;
                        int ilast = Last[i];
                        inext = Next[i];
                        if (inext != EMPTY) Last[inext] = ilast;
                        if (ilast != EMPTY) Next[ilast] = inext; else Head[Degree[i]] = inext;
                    }
                }
                if (e != me) { Pe[e] = FLIP(me); W[e] = 0; }
            }
            pme2 = pfree - 1;
        }
        Degree[me] = degme;
        Pe[me] = (int)pme1;
        Len[me] = (int)(pme2 - pme1 + 1);
        Elen[me] = FLIP(nvpiv + degme);
        wflg = clear_flag(wflg, wbig, W, n);
        /* ---- Scan 1: W[e] - wflg = |Le \ Lme| for all elements adjacent to the variables of Lme */
        for (long long pme = pme1; pme <= pme2; pme++) {
            int i = Iw[pme], eln = Elen[i];
            if (eln > 0) {
                int nvi = -Nv[i], wnvi = wflg - nvi;
                for (long long p = Pe[i]; p <= (long long)Pe[i] + eln - 1; p++) {
                    int e = Iw[p], we = W[e];
                    if (we >= wflg) we -= nvi;
                    else if (we != 0) we = Degree[e] + wnvi;
                    W[e] = we;
                }
            }
        }
        /* ---- Scan 2: degree update and element absorption */
        for (long long pme = pme1; pme <= pme2; pme++) {
            int i = Iw[pme];
            long long p1 = Pe[i], p2 = p1 + Elen[i] - 1, pn = p1;
            unsigned long long hash = 0;
            deg = 0;
            if (aggressive) {
                for (long long p = p1; p <= p2; p++) {
                    int e = Iw[p], we = W[e];
                    if (we != 0) {
                        int dext = we - wflg;
                        if (dext > 0) { deg += dext; Iw[pn++] = e; hash += (unsigned long long)e; }
                        else { Pe[e] = FLIP(me); W[e] = 0; }
                    }
                }
            } else {
                for (long long p = p1; p <= p2; p++) {
                    int e = Iw[p], we = W[e];
                    if (we != 0) { deg += we - wflg; Iw[pn++] = e; hash += (unsigned long long)e; }
                }
            }
            Elen[i] = (int)(pn - p1 + 1);
            long long p3 = pn, p4 = p1 + Len[i];
            for (long long p = p2 + 1; p < p4; p++) {
                int j = Iw[p], nvj = Nv[j];
                if (nvj > 0) { deg += nvj; Iw[pn++] = j; hash += (unsigned long long)j; }
            }
            if (Elen[i] == 1 && p3 == pn) {
                /* mass elimination: nothing left of i but the edge to me */
                Pe[i] = FLIP(me);
                int nvi = -Nv[i];
                degme -= nvi;
                nvpiv += nvi;
                nel += nvi;
                Nv[i] = 0;
                Elen[i] = EMPTY;
            } else {
                Degree[i] = Degree[i] < deg ? Degree[i] : deg;
                Iw[pn] = Iw[p3];
                Iw[p3] = Iw[p1];
                Iw[p1] = me;
                Len[i] = (int)(pn - p1 + 1);
                hash = hash % (unsigned long long)n;
                int j = Head[hash];
                if (j <= EMPTY) { Next[i] = FLIP(j); Head[hash] = FLIP(i); }
                else { Next[i] = Last[j]; Last[j] = i; }
                Last[i] = (int)hash;
            }
        }
        Degree[me] = degme;
        lemax = lemax > degme ? lemax : degme;
        wflg += lemax;
        wflg = clear_flag(wflg, wbig, W, n);
        /* ---- supervariable detection */
        for (long long pme = pme1; pme <= pme2; pme++) {
            int i = Iw[pme];
            if (Nv[i] < 0) {
                int hash = Last[i];
                int j = Head[hash];
                if (j == EMPTY) i = EMPTY;
                else if (j < EMPTY) { i = FLIP(j); Head[hash] = EMPTY; }
                else { i = Last[j]; Last[j] = EMPTY; }
                while (i != EMPTY && Next[i] != EMPTY) {
                    int ln = Len[i], eln = Elen[i];
                    for (long long p = (long long)Pe[i] + 1; p <= (long long)Pe[i] + ln - 1; p++) W[Iw[p]] = wflg;
                    int jlast = i;
                    j = Next[i];
                    while (j != EMPTY) {
                        int ok = (Len[j] == ln) && (Elen[j] == eln);
                        for (long long p = (long long)Pe[j] + 1; ok && p <= (long long)Pe[j] + ln - 1; p++)
                            if (W[Iw[p]] != wflg) ok = 0;
                        if (ok) {
                            Pe[j] = FLIP(i);
                            Nv[i] += Nv[j];
                            Nv[j] = 0;
                            Elen[j] = EMPTY;
                            j = Next[j];
                            Next[jlast] = j;
                        } else {
                            jlast = j;
                            j = Next[j];
                        }
                    }
                    wflg++;
                    i = Next[i];
                }
            }
        }
        /* ---- restore the degree lists, remove non-principal supervariables from the element */
        long long p = pme1;
        const int nleft = n - nel;
        for (long long pme = pme1; pme <= pme2; pme++) {
            int i = Iw[pme], nvi = -Nv[i];
            if (nvi > 0) {
                Nv[i] = nvi;
                deg = Degree[i] + degme - nvi;
                if (deg > nleft - nvi) deg = nleft - nvi;
                inext = Head[deg];
                if (inext != EMPTY) Last[inext] = i;
                Next[i] = inext;
                Last[i] = EMPTY;
                Head[deg] = i;
                if (deg < mindeg) mindeg = deg;
                Degree[i] = deg;
                Iw[p++] = i;
            }
        }
        /* ---- finalise the new element */
        Nv[me] = nvpiv;
        Len[me] = (int)(p - pme1);
        if (Len[me] == 0) { Pe[me] = EMPTY; W[me] = 0; }
        if (elenme != 0) pfree = p;
    }
    /* ---- post-ordering */
    for (int i = 0; i < n; i++) Pe[i] = FLIP(Pe[i]);
    for (int i = 0; i < n; i++) Elen[i] = FLIP(Elen[i]);
    for (int i = 0; i < n; i++)
        if (Nv[i] == 0) {
            int j = Pe[i];
            if (j == EMPTY) continue;            /* dense variable: no parent */
            while (Nv[j] == 0) j = Pe[j];
            int e = j;
            j = i;
            while (Nv[j] == 0) { int jnext = Pe[j]; Pe[j] = e; j = jnext; }
        }
    postorder(n, Pe, Nv, Elen, W, Head, Next, Last);
    for (int k = 0; k < n; k++) { Head[k] = EMPTY; Next[k] = EMPTY; }
    for (int e = 0; e < n; e++) { int k = W[e]; if (k != EMPTY) Head[k] = e; }
    nel = 0;
    for (int k = 0; k < n; k++) {
        int e = Head[k];
        if (e == EMPTY) break;
        Next[e] = nel;
        nel += Nv[e];
    }
    for (int i = 0; i < n; i++)
        if (Nv[i] == 0) {
            int e = Pe[i];
            if (e != EMPTY) { Next[i] = Next[e]; Next[e]++; }
            else Next[i] = nel++;
        }
    for (int i = 0; i < n; i++) Last[Next[i]] = i;
    memcpy(P, Last, (size_t)n * sizeof(int));
    (void)ndense;
    free(sp.Iw); free(ws); free(Sp); free(Pe); free(Tp); free(Len);
    return 0;
}
