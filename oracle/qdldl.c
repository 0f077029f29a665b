/*
 * qdldl.c -- CPU ORACLE (test infrastructure only): restatement of the reference's vendored QDLDL,
 * src/solver/qdldl.jl, with 0-based indices.  See oracle.h for scope and the parity statement.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

#define QDLDL_UNKNOWN (-1)

struct orc_qdldl {
    int n;
    int *perm, *iperm;      /* qdldl.jl:135,143 */
    int nnzA;
    int *Ap, *Ai;           /* triuA: upper triangle of P A P', rows unsorted within a column (qdldl.jl:669-742) */
    double *Ax;
    int *AtoPAPt;           /* qdldl.jl:731 */
    int *etree, *Lnz;       /* QDLDL_etree! qdldl.jl:358 */
    int sumLnz;
    int *Lp, *Li;
    double *Lx, *D, *Dinv;
    int *iwork;             /* 3n */
    unsigned char *bwork;   /* n */
    double *fwork;          /* n */
    int positive_inertia;
    long long factor_count;
};

/* ------------------------------------------------------------------ ordering stand-in (see oracle.h header) */
void orc_min_degree(int n, const int *Ap, const int *Ai, int *perm)
{
    /* Exact minimum (external) degree on the elimination graph of A + A', dense bitset adjacency.
     * NOT SuiteSparse AMD: the reference's ordering library is absent from this environment. */
    size_t words = ((size_t)n + 63) / 64;
    uint64_t *adj = (uint64_t *)calloc((size_t)n * words, sizeof(uint64_t));
    int *deg = (int *)malloc(sizeof(int) * (size_t)n);
    unsigned char *done = (unsigned char *)calloc((size_t)n, 1);
    int *nb = (int *)malloc(sizeof(int) * (size_t)n);
    for (int j = 0; j < n; j++)
        for (int q = Ap[j]; q < Ap[j + 1]; q++) {
            int i = Ai[q];
            if (i == j) continue;
            adj[(size_t)i * words + (size_t)(j >> 6)] |= 1ull << (j & 63);
            adj[(size_t)j * words + (size_t)(i >> 6)] |= 1ull << (i & 63);
        }
    for (int i = 0; i < n; i++) {
        int d = 0;
        for (size_t w = 0; w < words; w++) d += __builtin_popcountll(adj[(size_t)i * words + w]);
        deg[i] = d;
    }
    for (int k = 0; k < n; k++) {
        int best = -1;
        for (int i = 0; i < n; i++)
            if (!done[i] && (best < 0 || deg[i] < deg[best])) best = i;
        perm[k] = best;
        done[best] = 1;
        /* neighbours of the pivot */
        int cnt = 0;
        uint64_t *rb = adj + (size_t)best * words;
        for (size_t w = 0; w < words; w++) {
            uint64_t bits = rb[w];
            while (bits) {
                int b = __builtin_ctzll(bits);
                bits &= bits - 1;
                nb[cnt++] = (int)(w * 64) + b;
            }
        }
        /* clique the neighbours, remove the pivot */
        for (int a = 0; a < cnt; a++) {
            uint64_t *ra = adj + (size_t)nb[a] * words;
            for (size_t w = 0; w < words; w++) ra[w] |= rb[w];
            ra[(size_t)(nb[a] >> 6)] &= ~(1ull << (nb[a] & 63));
            ra[(size_t)(best >> 6)] &= ~(1ull << (best & 63));
            int d = 0;
            for (size_t w = 0; w < words; w++) d += __builtin_popcountll(ra[w]);
            deg[nb[a]] = d;
        }
        memset(rb, 0, words * sizeof(uint64_t));
    }
    free(adj);
    free(deg);
    free(done);
    free(nb);
}

/* ------------------------------------------------------------------ QDLDL_etree!  qdldl.jl:358-395 */
static int qdldl_etree(int n, const int *Ap, const int *Ai, int *work, int *Lnz, int *etree)
{
    for (int i = 0; i < n; i++) {
        work[i] = 0;
        Lnz[i] = 0;
        etree[i] = QDLDL_UNKNOWN;
        if (Ap[i] == Ap[i + 1]) return -1; /* empty column, qdldl.jl:368 */
    }
    for (int j = 0; j < n; j++) {
        work[j] = j + 1; /* marker j+1 so that 0 means "never" (reference marks with the 1-based j) */
        for (int p = Ap[j]; p < Ap[j + 1]; p++) {
            int i = Ai[p];
            if (i > j) return -1; /* not upper triangular, qdldl.jl:377 */
            while (work[i] != j + 1) {
                if (etree[i] == QDLDL_UNKNOWN) etree[i] = j;
                Lnz[i] += 1;
                work[i] = j + 1;
                i = etree[i];
            }
        }
    }
    int sum = 0;
    for (int i = 0; i < n; i++) sum += Lnz[i];
    return sum;
}

/* ------------------------------------------------------------------ QDLDL_factor!  qdldl.jl:400-589 */
static int qdldl_factor(int n, const int *Ap, const int *Ai, const double *Ax, int *Lp, int *Li, double *Lx,
                        double *D, double *Dinv, const int *Lnz, const int *etree, unsigned char *bwork,
                        int *iwork, double *fwork)
{
    int positiveValuesInD = 0;
    unsigned char *yMarkers = bwork;
    int *yIdx = iwork;
    int *elimBuffer = iwork + n;
    int *LNextSpaceInCol = iwork + 2 * n;
    double *yVals = fwork;

    Lp[0] = 0;
    for (int i = 0; i < n; i++) {
        Lp[i + 1] = Lp[i] + Lnz[i];
        yMarkers[i] = 0;
        yVals[i] = 0.0;
        D[i] = 0.0;
        LNextSpaceInCol[i] = Lp[i];
    }
    /* first pivot, qdldl.jl:449-458 (the reference takes Ax[1]: column 1 holds only its diagonal) */
    D[0] = Ax[0];
    if (D[0] == 0.0) return -1;
    if (D[0] > 0.0) positiveValuesInD++;
    Dinv[0] = 1.0 / D[0];

    for (int k = 1; k < n; k++) {
        int nnzY = 0;
        for (int i = Ap[k]; i < Ap[k + 1]; i++) {
            int bidx = Ai[i];
            if (bidx == k) {
                D[k] = Ax[i];
                continue;
            }
            yVals[bidx] = Ax[i];
            int nextIdx = bidx;
            if (!yMarkers[nextIdx]) {
                yMarkers[nextIdx] = 1;
                elimBuffer[0] = nextIdx;
                int nnzE = 1;
                nextIdx = etree[bidx];
                while (nextIdx != QDLDL_UNKNOWN && nextIdx < k) {
                    if (yMarkers[nextIdx]) break;
                    yMarkers[nextIdx] = 1;
                    elimBuffer[nnzE] = nextIdx;
                    nnzE++;
                    nextIdx = etree[nextIdx];
                }
                while (nnzE != 0) {
                    yIdx[nnzY] = elimBuffer[nnzE - 1];
                    nnzY++;
                    nnzE--;
                }
            }
        }
        for (int i = nnzY - 1; i >= 0; i--) {
            int cidx = yIdx[i];
            int tmpIdx = LNextSpaceInCol[cidx];
            double yVals_cidx = yVals[cidx];
            for (int j = Lp[cidx]; j < tmpIdx; j++) yVals[Li[j]] -= Lx[j] * yVals_cidx;
            Lx[tmpIdx] = yVals_cidx * Dinv[cidx];
            D[k] -= yVals_cidx * Lx[tmpIdx];
            Li[tmpIdx] = k;
            LNextSpaceInCol[cidx]++;
            yVals[cidx] = 0.0;
            yMarkers[cidx] = 0;
        }
        if (D[k] == 0.0) return -1; /* qdldl.jl:579: abort, later D entries stay 0.0 */
        if (D[k] > 0.0) positiveValuesInD++;
        Dinv[k] = 1.0 / D[k];
    }
    return positiveValuesInD;
}

/* ------------------------------------------------------------------ permute_symmetric  qdldl.jl:642-742 */
static void permute_symmetric(int n, const int *Ac, const int *Ar, const double *Av, const int *iperm, int *Pc,
                              int *Pr, double *Pv, int *AtoPAPt)
{
    int *num_entries = (int *)calloc((size_t)n, sizeof(int));
    for (int colA = 0; colA < n; colA++) {
        int colP = iperm[colA];
        for (int q = Ac[colA]; q < Ac[colA + 1]; q++) {
            int rowA = Ar[q];
            int rowP = iperm[rowA];
            if (rowA <= colA) {
                int col_idx = rowP > colP ? rowP : colP;
                num_entries[col_idx]++;
            }
        }
    }
    Pc[0] = 0;
    for (int k = 0; k < n; k++) {
        Pc[k + 1] = Pc[k] + num_entries[k];
        num_entries[k] = Pc[k];
    }
    int *row_starts = num_entries;
    for (int colA = 0; colA < n; colA++) {
        int colP = iperm[colA];
        for (int q = Ac[colA]; q < Ac[colA + 1]; q++) {
            int rowA = Ar[q];
            if (rowA <= colA) {
                int rowP = iperm[rowA];
                int col_idx = colP > rowP ? colP : rowP;
                int dst = row_starts[col_idx];
                Pr[dst] = colP < rowP ? colP : rowP;
                Pv[dst] = Av[q];
                AtoPAPt[q] = dst;
                row_starts[col_idx]++;
            }
        }
    }
    free(num_entries);
}

static void factor(orc_qdldl *F)
{
    F->positive_inertia = qdldl_factor(F->n, F->Ap, F->Ai, F->Ax, F->Lp, F->Li, F->Lx, F->D, F->Dinv, F->Lnz,
                                       F->etree, F->bwork, F->iwork, F->fwork);
    F->factor_count++;
}

orc_qdldl *orc_qdldl_new(int n, const int *Ap, const int *Ai, const double *Ax, const int *perm)
{
    orc_qdldl *F = (orc_qdldl *)calloc(1, sizeof(orc_qdldl));
    int nnz = Ap[n];
    F->n = n;
    F->nnzA = nnz;
    F->perm = (int *)malloc(sizeof(int) * (size_t)n);
    F->iperm = (int *)malloc(sizeof(int) * (size_t)n);
    if (perm)
        memcpy(F->perm, perm, sizeof(int) * (size_t)n);
    else
        if (orc_amd_order(n, Ap, Ai, F->perm, 10.0, 1) != 0) orc_min_degree(n, Ap, Ai, F->perm); /* amd(A), qdldl.jl:135 */
    for (int i = 0; i < n; i++) F->iperm[F->perm[i]] = i; /* invperm, qdldl.jl:143 */
    F->Ap = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    F->Ai = (int *)malloc(sizeof(int) * (size_t)nnz);
    F->Ax = (double *)malloc(sizeof(double) * (size_t)nnz);
    F->AtoPAPt = (int *)malloc(sizeof(int) * (size_t)nnz);
    permute_symmetric(n, Ap, Ai, Ax, F->iperm, F->Ap, F->Ai, F->Ax, F->AtoPAPt);
    F->etree = (int *)malloc(sizeof(int) * (size_t)n);
    F->Lnz = (int *)malloc(sizeof(int) * (size_t)n);
    F->iwork = (int *)malloc(sizeof(int) * 3 * (size_t)n);
    F->bwork = (unsigned char *)malloc((size_t)n);
    F->fwork = (double *)malloc(sizeof(double) * (size_t)n);
    F->sumLnz = qdldl_etree(n, F->Ap, F->Ai, F->iwork, F->Lnz, F->etree);
    if (F->sumLnz < 0) { /* "Input matrix is not upper triangular or has an empty column", qdldl.jl:72 */
        orc_qdldl_free(F);
        return NULL;
    }
    F->Lp = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    F->Li = (int *)malloc(sizeof(int) * (size_t)(F->sumLnz > 0 ? F->sumLnz : 1));
    F->Lx = (double *)malloc(sizeof(double) * (size_t)(F->sumLnz > 0 ? F->sumLnz : 1));
    F->D = (double *)malloc(sizeof(double) * (size_t)n);
    F->Dinv = (double *)malloc(sizeof(double) * (size_t)n);
    F->positive_inertia = -1;
    factor(F); /* qdldl.jl:173 */
    return F;
}

void orc_qdldl_free(orc_qdldl *F)
{
    if (!F) return;
    free(F->perm); free(F->iperm); free(F->Ap); free(F->Ai); free(F->Ax); free(F->AtoPAPt);
    free(F->etree); free(F->Lnz); free(F->Lp); free(F->Li); free(F->Lx); free(F->D); free(F->Dinv);
    free(F->iwork); free(F->bwork); free(F->fwork);
    free(F);
}

int orc_qdldl_refactor(orc_qdldl *F, const double *Ax)
{
    for (int i = 0; i < F->nnzA; i++) F->Ax[F->AtoPAPt[i]] = Ax[i]; /* update_values!, qdldl.jl:208-210 */
    factor(F);                                                      /* refactor!, qdldl.jl:269-278 */
    return F->positive_inertia;
}

void orc_qdldl_solve(orc_qdldl *F, double *b)
{
    int n = F->n;
    double *x = F->fwork;
    for (int j = 0; j < n; j++) x[j] = b[F->perm[j]]; /* permute!, qdldl.jl:628 */
    for (int i = 0; i < n; i++)                       /* QDLDL_Lsolve!, :592 */
        for (int j = F->Lp[i]; j < F->Lp[i + 1]; j++) x[F->Li[j]] -= F->Lx[j] * x[i];
    for (int i = 0; i < n; i++) x[i] *= F->Dinv[i];   /* :619 */
    for (int i = n - 1; i >= 0; i--)                  /* QDLDL_Ltsolve!, :604 */
        for (int j = F->Lp[i]; j < F->Lp[i + 1]; j++) x[i] -= F->Lx[j] * x[F->Li[j]];
    for (int j = 0; j < n; j++) b[F->perm[j]] = x[j]; /* ipermute!, :635 */
}

int orc_qdldl_n(const orc_qdldl *F) { return F->n; }
int orc_qdldl_nnzL(const orc_qdldl *F) { return F->sumLnz; }
int orc_qdldl_nnzA(const orc_qdldl *F) { return F->nnzA; }
const int *orc_qdldl_perm(const orc_qdldl *F) { return F->perm; }
const int *orc_qdldl_iperm(const orc_qdldl *F) { return F->iperm; }
const int *orc_qdldl_etree(const orc_qdldl *F) { return F->etree; }
const int *orc_qdldl_Lnz(const orc_qdldl *F) { return F->Lnz; }
const int *orc_qdldl_Lp(const orc_qdldl *F) { return F->Lp; }
const int *orc_qdldl_Li(const orc_qdldl *F) { return F->Li; }
const double *orc_qdldl_Lx(const orc_qdldl *F) { return F->Lx; }
const double *orc_qdldl_D(const orc_qdldl *F) { return F->D; }
const double *orc_qdldl_Dinv(const orc_qdldl *F) { return F->Dinv; }
const int *orc_qdldl_triuA_colptr(const orc_qdldl *F) { return F->Ap; }
const int *orc_qdldl_triuA_rowval(const orc_qdldl *F) { return F->Ai; }
const double *orc_qdldl_triuA_nzval(const orc_qdldl *F) { return F->Ax; }
const int *orc_qdldl_AtoPAPt(const orc_qdldl *F) { return F->AtoPAPt; }
int orc_qdldl_positive_inertia(const orc_qdldl *F) { return F->positive_inertia; }
long long orc_qdldl_factor_count(const orc_qdldl *F) { return F->factor_count; }
