/*
 * lis_precon_ilu.c -- ILU(k) (`-p ilu [-ilu_fill k]`) for CSR matrices: the level-of-fill
 * factorization on the host, once per solve, and the two triangular solves of every apply on the
 * GPU through the one-launch kernel (lis_sptrsv.c).
 *
 * Reference: src/precon/lis_precon_iluk.c -- lis_symbolic_fact_csr :262, lis_numerical_fact_csr
 * :637, lis_psolve_iluk_csr :880, lis_psolveh_iluk_csr :1086.  What is pinned to it, so that the
 * factors and every apply carry the same bits:
 *   - the pattern: level-of-fill IKJ elimination, pivots taken in ascending column order; a row of
 *     L ends up sorted by column, a row of U keeps A's storage order followed by its fill-ins in
 *     the order they were discovered (that order is the summation order of the U solve);
 *   - the numbers: l_ik = l_ik * (1/u_kk) with the reciprocal pivot stored in D, updates
 *     a_ij -= l_ik*u_kj unfused, one pivot after the other;
 *   - the OpenMP build factors each thread's diagonal block on its own (couplings that leave the
 *     block are dropped before the factorization); the block count here is the emulated thread
 *     count (`-omp_num_threads N`, default 1 = the serial reference), as for SSOR;
 *   - apply:  w = b - L w (unit diagonal),  x = D (w - U x);
 *     transposed apply (BiCG, BiCR):  w = D (b - U^T w),  x = w - L^T x  with the column-oriented
 *     update order of the reference (ascending pivot for U^T, descending for L^T).
 * The factorization is sequential by nature and stays on the host, as in the reference; the apply,
 * which runs once or twice per iteration, is the part on the path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lis_device.h"
#include "lis_host.h"

typedef struct {                  /* growable CSR rows */
    LIS_INT *ptr, *idx, *lev;
    LIS_SCALAR *val;
    size_t cap, nnz;
} ilu_rows;

typedef struct lisd_ilu {
    int n;
    ilu_rows L, U;                /* host factors (kept for the transposed solve) */
    lisd_tri *tL, *tU, *tLT, *tUT;
    double *d_w;
} lisd_ilu;

static void rows_free(ilu_rows *r)
{
    free(r->ptr); free(r->idx); free(r->lev); free(r->val);
    memset(r, 0, sizeof(*r));
}

static int rows_reserve(ilu_rows *r, size_t extra, int with_lev)
{
    if (r->nnz + extra <= r->cap) return 0;
    size_t cap = r->cap ? r->cap * 2 : 1024;
    while (cap < r->nnz + extra) cap *= 2;
    LIS_INT *idx = (LIS_INT *)realloc(r->idx, cap * sizeof(LIS_INT));
    if (!idx) return -1;
    r->idx = idx;
    if (with_lev) {
        LIS_INT *lev = (LIS_INT *)realloc(r->lev, cap * sizeof(LIS_INT));
        if (!lev) return -1;
        r->lev = lev;
    }
    r->cap = cap;
    return 0;
}

void lis_host_ilu_free(void *p)
{
    lisd_ilu *F = (lisd_ilu *)p;
    if (F == NULL) return;
    rows_free(&F->L); rows_free(&F->U);
    lisd_tri_free(F->tL); lisd_tri_free(F->tU); lisd_tri_free(F->tLT); lisd_tri_free(F->tUT);
    lisd_free(F->d_w);
    free(F);
}

/* pattern of L and U with fill levels <= levfill, rows [is,ie) of each block factored alone */
static LIS_INT ilu_symbolic(LIS_MATRIX A, LIS_INT levfill, int nb, lisd_ilu *F)
{
    const LIS_INT n = A->n;
    LIS_INT *lo_col = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n > 0 ? n : 1));   /* candidates left of the diagonal */
    LIS_INT *lo_lev = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n > 0 ? n : 1));
    LIS_INT *up_col = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n > 0 ? n : 1));   /* right of it, in discovery order */
    LIS_INT *up_lev = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n > 0 ? n : 1));
    LIS_INT *where = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n > 0 ? n : 1));    /* column -> slot in its list, -1 = absent */
    LIS_INT err = LIS_OUT_OF_MEMORY;
    F->L.ptr = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
    F->U.ptr = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
    if (!lo_col || !lo_lev || !up_col || !up_lev || !where || !F->L.ptr || !F->U.ptr) goto done;
    for (LIS_INT i = 0; i < n; i++) where[i] = -1;
    for (int b = 0; b < nb; b++) {
        LIS_INT is, ie;
        LIS_GET_ISIE(b, nb, n, is, ie);
        for (LIS_INT i = is; i < ie; i++) {
            LIS_INT nlo = 0, nup = 0;
            for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++) {
                const LIS_INT c = A->index[j];
                if (c < is || c >= ie || c == i) continue;          /* outside the block (or the halo), or the pivot itself */
                if (c < i) { lo_col[nlo] = c; lo_lev[nlo] = 0; where[c] = nlo++; }
                else { up_col[nup] = c; up_lev[nup] = 0; where[c] = nup++; }
            }
            for (LIS_INT p = 0; p < nlo; p++) {
                /* next pivot = smallest remaining column; bring it to slot p */
                LIS_INT m = p;
                for (LIS_INT q = p + 1; q < nlo; q++) if (lo_col[q] < lo_col[m]) m = q;
                if (m != p) {
                    const LIS_INT c = lo_col[p], l = lo_lev[p];
                    lo_col[p] = lo_col[m]; lo_lev[p] = lo_lev[m];
                    lo_col[m] = c; lo_lev[m] = l;
                    where[lo_col[p]] = p; where[c] = m;
                }
                const LIS_INT k = lo_col[p];
                for (LIS_INT j = F->U.ptr[k]; j < F->U.ptr[k + 1]; j++) {
                    const LIS_INT c = F->U.idx[j];
                    const LIS_INT lev = F->U.lev[j] + lo_lev[p] + 1;
                    if (lev > levfill) continue;
                    const LIS_INT at = where[c];
                    if (at == -1) {
                        if (c < i) { lo_col[nlo] = c; lo_lev[nlo] = lev; where[c] = nlo++; }
                        else if (c > i) { up_col[nup] = c; up_lev[nup] = lev; where[c] = nup++; }
                    } else if (c < i) { if (lev < lo_lev[at]) lo_lev[at] = lev; }
                    else { if (lev < up_lev[at]) up_lev[at] = lev; }
                }
            }
            if (rows_reserve(&F->L, (size_t)nlo, 0) || rows_reserve(&F->U, (size_t)nup, 1)) goto done;
            for (LIS_INT q = 0; q < nlo; q++) { F->L.idx[F->L.nnz++] = lo_col[q]; where[lo_col[q]] = -1; }
            for (LIS_INT q = 0; q < nup; q++) { F->U.idx[F->U.nnz] = up_col[q]; F->U.lev[F->U.nnz++] = up_lev[q]; where[up_col[q]] = -1; }
            if (F->L.nnz > 0x7fffffff || F->U.nnz > 0x7fffffff) { LIS_SETERR(LIS_ERR_OUT_OF_MEMORY, "ILU fill exceeds 2^31 entries\n"); err = LIS_ERR_OUT_OF_MEMORY; goto done; }
            F->L.ptr[i + 1] = (LIS_INT)F->L.nnz;
            F->U.ptr[i + 1] = (LIS_INT)F->U.nnz;
        }
    }
    err = LIS_SUCCESS;
done:
    free(lo_col); free(lo_lev); free(up_col); free(up_lev); free(where);
    if (err == LIS_OUT_OF_MEMORY) LIS_SETERR_MEM(n * sizeof(LIS_INT));
    return err;
}

/* values of L, U and the reciprocal pivots d[] on that pattern */
static LIS_INT ilu_numeric(LIS_MATRIX A, int nb, lisd_ilu *F, LIS_SCALAR *d)
{
    const LIS_INT n = A->n;
    LIS_INT *slot = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n > 0 ? n : 1));      /* column -> position in row i of L / U */
    F->L.val = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (F->L.nnz ? F->L.nnz : 1));
    F->U.val = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (F->U.nnz ? F->U.nnz : 1));
    if (!slot || !F->L.val || !F->U.val) { free(slot); LIS_SETERR_MEM(n * sizeof(LIS_INT)); return LIS_OUT_OF_MEMORY; }
    const LIS_INT *lp = F->L.ptr, *li = F->L.idx, *up = F->U.ptr, *ui = F->U.idx;
    LIS_SCALAR *lv = F->L.val, *uv = F->U.val;
    for (LIS_INT i = 0; i < n; i++) slot[i] = -1;
    for (int b = 0; b < nb; b++) {
        LIS_INT is, ie;
        LIS_GET_ISIE(b, nb, n, is, ie);
        for (LIS_INT i = is; i < ie; i++) {
            for (LIS_INT j = lp[i]; j < lp[i + 1]; j++) { slot[li[j]] = j; lv[j] = 0.0; }
            for (LIS_INT j = up[i]; j < up[i + 1]; j++) { slot[ui[j]] = j; uv[j] = 0.0; }
            slot[i] = 0;                                            /* any value but -1: the pivot is present */
            d[i] = 0.0;
            for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++) {
                const LIS_INT c = A->index[j];
                if (c < is || c >= ie) continue;
                if (c < i) lv[slot[c]] = A->value[j];
                else if (c == i) d[i] = A->value[j];
                else uv[slot[c]] = A->value[j];
            }
            for (LIS_INT j = lp[i]; j < lp[i + 1]; j++) {
                const LIS_INT k = li[j];
                lv[j] *= d[k];
                const LIS_SCALAR m = lv[j];
                for (LIS_INT q = up[k]; q < up[k + 1]; q++) {
                    const LIS_INT c = ui[q];
                    if (slot[c] == -1) continue;
                    if (c < i) lv[slot[c]] -= m * uv[q];
                    else if (c == i) d[i] -= m * uv[q];
                    else uv[slot[c]] -= m * uv[q];
                }
            }
            for (LIS_INT j = lp[i]; j < lp[i + 1]; j++) slot[li[j]] = -1;
            for (LIS_INT j = up[i]; j < up[i + 1]; j++) slot[ui[j]] = -1;
            slot[i] = -1;
            d[i] = 1.0 / d[i];
        }
    }
    free(slot);
    return LIS_SUCCESS;
}

/* ILUT (`-p ilut [-iluc_drop tol] [-iluc_rate m]`), the serial factorization of
 * src/precon/lis_precon_ilut.c:364-620 into the same L / U / D containers as ILU(k), so that the apply
 * (lis_precon_ilut.c:623-830 and :832-1030: the loops of the ILU(k) apply) runs on the same sweeps.
 * Row i: the entries left of the diagonal are eliminated in ascending column order (fill-ins join the
 * queue), l = w * d[k], every update lxu = -l * u_kj is dropped when it would create a new entry and
 * |lxu| < tol * mean|a_i*|; d[i] = 1/w_ii.  Then each part keeps at most lfil = int(nnz/(2n) * m) entries:
 * the reference sorts |w| ASCENDING and keeps the first lfil, i.e. the smallest (lis_sort_di, :556-560) --
 * followed here, with its quicksort scheme, which decides among equal magnitudes at the cut.  What is kept stays in position order: L ascending by column, U in A's order then fill-ins in
 * discovery order (the summation order of the U solve). */
typedef struct { LIS_SCALAR mag; LIS_INT pos; } ilut_key;
/* Which of several equal magnitudes survive the cut is decided by the reference's sort, so this is that
 * sort's scheme (src/system/lis_sort.c:431-470, lis_sort_di): quicksort on the magnitudes, pivot = the
 * middle element parked at the right end, both scans with strict comparisons, halves [lo, j] and [i, hi]. */
static void ilut_sort(ilut_key *k, LIS_INT lo, LIS_INT hi)
{
    while (lo < hi) {
        const LIS_INT mid = (lo + hi) / 2;
        const LIS_SCALAR pivot = k[mid].mag;
        ilut_key t = k[mid]; k[mid] = k[hi]; k[hi] = t;
        LIS_INT i = lo, j = hi;
        while (i <= j) {
            while (k[i].mag < pivot) i++;
            while (k[j].mag > pivot) j--;
            if (i <= j) { t = k[i]; k[i] = k[j]; k[j] = t; i++; j--; }
        }
        ilut_sort(k, lo, j);
        lo = i;                                            /* right half in place of the second recursive call */
    }
}
static int ilut_pos_cmp(const void *a, const void *b) { return (*(const LIS_INT *)a > *(const LIS_INT *)b) - (*(const LIS_INT *)a < *(const LIS_INT *)b); }

static LIS_INT ilut_keep(ilu_rows *R, LIS_INT count, LIS_INT lfil, const LIS_INT *cols, const LIS_SCALAR *vals, ilut_key *keys, LIS_INT *sel)
{
    const LIS_INT len = lfil < count ? lfil : count;
    if (rows_reserve(R, (size_t)(len > 0 ? len : 0), 0)) return LIS_OUT_OF_MEMORY;
    {
        LIS_SCALAR *nv = (LIS_SCALAR *)realloc(R->val, sizeof(LIS_SCALAR) * (R->cap ? R->cap : 1));
        if (!nv) return LIS_OUT_OF_MEMORY;
        R->val = nv;
    }
    if (len == count) {
        for (LIS_INT j = 0; j < count; j++) sel[j] = j;
    } else {
        for (LIS_INT j = 0; j < count; j++) { keys[j].mag = fabs(vals[j]); keys[j].pos = j; }
        ilut_sort(keys, 0, count - 1);
        for (LIS_INT j = 0; j < len; j++) sel[j] = keys[j].pos;
        qsort(sel, (size_t)len, sizeof(LIS_INT), ilut_pos_cmp);
    }
    for (LIS_INT j = 0; j < len; j++) { R->idx[R->nnz] = cols[sel[j]]; R->val[R->nnz] = vals[sel[j]]; R->nnz++; }
    return LIS_SUCCESS;
}

/* nb > 1: the OpenMP build's variant (lis_precon_ilut.c:105-362) -- every thread factors its own diagonal block
 * (couplings that leave it are skipped, the row norm is taken over the block's entries only) and the fill cap
 * is int(nnz/(2n)) * rate, truncated BEFORE the multiplication there. */
static LIS_INT ilut_factor(LIS_MATRIX A, LIS_SCALAR tol, LIS_SCALAR rate, int nb, lisd_ilu *F, LIS_SCALAR *d)
{
    const LIS_INT n = A->n;
    const LIS_INT lfil = n <= 0 ? 0 : nb > 1 ? (LIS_INT)((double)(LIS_INT)((double)A->ptr[n] / (2.0 * n)) * rate)
                                             : (LIS_INT)(((double)A->ptr[n] / (2.0 * n)) * rate);
    LIS_INT err = LIS_OUT_OF_MEMORY;
    LIS_INT *where = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n + 1));       /* column -> slot (lower: 0.., diagonal: -2, upper: n + k), -1 absent */
    LIS_INT *lcol = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n + 1)), *ucol = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n + 1));
    LIS_INT *sel = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n + 1));
    LIS_SCALAR *lval = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(n + 1)), *uval = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(n + 1));
    ilut_key *keys = (ilut_key *)malloc(sizeof(ilut_key) * (size_t)(n + 1));
    char *done = (char *)calloc((size_t)(n + 1), 1);
    F->L.ptr = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
    F->U.ptr = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
    if (!where || !lcol || !ucol || !sel || !lval || !uval || !keys || !done || !F->L.ptr || !F->U.ptr) { LIS_SETERR_MEM(n); goto out; }
    for (LIS_INT i = 0; i < n; i++) where[i] = -1;
    for (int blk = 0; blk < nb; blk++) {
    LIS_INT is, ie;
    LIS_GET_ISIE(blk, nb, n, is, ie);
    for (LIS_INT i = is; i < ie; i++) {
        LIS_REAL tnorm = 0;
        LIS_INT nl = 0, nu = 0, cnt = 0;
        LIS_SCALAR wd = 0;
        for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++) {
            if (A->index[j] < is || A->index[j] >= ie) continue;
            tnorm += fabs(A->value[j]);
            cnt++;
        }
        tnorm = tnorm / (double)(nb > 1 ? cnt : A->ptr[i + 1] - A->ptr[i]);
        const LIS_REAL tolnorm = tol * tnorm;
        where[i] = -2;
        for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++) {
            const LIS_INT c = A->index[j];
            if (c < is || c >= ie) continue;
            /* a column stored twice overwrites the slot map like the reference's iw[] does: both copies stay in the list */
            if (c < i) { lcol[nl] = c; lval[nl] = A->value[j]; where[c] = nl; done[nl] = 0; nl++; }
            else if (c == i) wd = A->value[j];
            else { ucol[nu] = c; uval[nu] = A->value[j]; where[c] = n + nu; nu++; }
        }
        for (LIS_INT step = 0; step < nl; step++) {
            LIS_INT q = -1;                                       /* next pivot: the smallest column not yet eliminated */
            for (LIS_INT k = 0; k < nl; k++) if (!done[k] && (q < 0 || lcol[k] < lcol[q])) q = k;
            done[q] = 1;
            const LIS_INT k = lcol[q];
            const LIS_SCALAR fact = lval[q] * d[k];
            lval[q] = fact;
            where[k] = -1;
            for (LIS_INT p = F->U.ptr[k]; p < F->U.ptr[k + 1]; p++) {
                const LIS_INT c = F->U.idx[p];
                const LIS_INT slot = where[c];
                const LIS_SCALAR lxu = -fact * F->U.val[p];
                if (fabs(lxu) < tolnorm && slot == -1) continue;
                if (c >= i) {
                    if (slot == -1) { ucol[nu] = c; uval[nu] = lxu; where[c] = n + nu; nu++; }
                    else if (slot == -2) wd += lxu;
                    else uval[slot - n] += lxu;
                } else {
                    if (slot == -1) { lcol[nl] = c; lval[nl] = lxu; where[c] = nl; done[nl] = 0; nl++; }
                    else lval[slot] += lxu;
                }
            }
        }
        where[i] = -1;
        for (LIS_INT k = 0; k < nu; k++) where[ucol[k]] = -1;
        for (LIS_INT k = 0; k < nl; k++) where[lcol[k]] = -1;
        d[i] = 1.0 / wd;
        /* L in ascending column order = the order the reference's selection loop leaves behind */
        for (LIS_INT a = 1; a < nl; a++) {
            const LIS_INT c = lcol[a]; const LIS_SCALAR v = lval[a];
            LIS_INT b = a - 1;
            while (b >= 0 && lcol[b] > c) { lcol[b + 1] = lcol[b]; lval[b + 1] = lval[b]; b--; }
            lcol[b + 1] = c; lval[b + 1] = v;
        }
        if (ilut_keep(&F->L, nl, lfil, lcol, lval, keys, sel) || ilut_keep(&F->U, nu, lfil, ucol, uval, keys, sel)) { LIS_SETERR_MEM(nl + nu); goto out; }
        F->L.ptr[i + 1] = (LIS_INT)F->L.nnz;
        F->U.ptr[i + 1] = (LIS_INT)F->U.nnz;
    }
    }
    if (F->L.idx == NULL) { F->L.idx = (LIS_INT *)calloc(1, sizeof(LIS_INT)); F->L.val = (LIS_SCALAR *)calloc(1, sizeof(LIS_SCALAR)); }
    if (F->U.idx == NULL) { F->U.idx = (LIS_INT *)calloc(1, sizeof(LIS_INT)); F->U.val = (LIS_SCALAR *)calloc(1, sizeof(LIS_SCALAR)); }
    err = LIS_SUCCESS;
out:
    free(where); free(lcol); free(ucol); free(sel); free(lval); free(uval); free(keys); free(done);
    return err;
}

/* T = R^T as CSR; a row of T lists its entries by ascending (or descending) source row: the
 * order in which the reference's column-oriented loops subtract them */
static LIS_INT rows_transpose(LIS_INT n, const ilu_rows *R, int descending, ilu_rows *T)
{
    memset(T, 0, sizeof(*T));
    T->ptr = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
    T->idx = (LIS_INT *)malloc(sizeof(LIS_INT) * (R->nnz ? R->nnz : 1));
    T->val = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (R->nnz ? R->nnz : 1));
    LIS_INT *cur = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n > 0 ? n : 1));
    if (!T->ptr || !T->idx || !T->val || !cur) { free(cur); rows_free(T); LIS_SETERR_MEM(R->nnz * 12); return LIS_OUT_OF_MEMORY; }
    for (size_t j = 0; j < R->nnz; j++) T->ptr[R->idx[j] + 1]++;
    for (LIS_INT i = 0; i < n; i++) T->ptr[i + 1] += T->ptr[i];
    for (LIS_INT i = 0; i < n; i++) cur[i] = T->ptr[i];
    for (LIS_INT s = 0; s < n; s++) {
        const LIS_INT i = descending ? n - 1 - s : s;
        for (LIS_INT j = R->ptr[i]; j < R->ptr[i + 1]; j++) {
            const LIS_INT at = cur[R->idx[j]]++;
            T->idx[at] = i; T->val[at] = R->val[j];
        }
    }
    T->nnz = R->nnz;
    free(cur);
    return LIS_SUCCESS;
}

LIS_INT lis_host_ilu_create(LIS_SOLVER solver, LIS_PRECON precon)
{
    LIS_MATRIX A = solver->A, B = NULL;
    LIS_INT err;
    if (A->matrix_type == LIS_MATRIX_BSR || solver->options[LIS_OPTIONS_STORAGE] == LIS_MATRIX_BSR ||
        solver->options[LIS_OPTIONS_STORAGE] == LIS_MATRIX_VBR) {
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "block ILU (BSR/VBR storage) is not available; use CSR\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    if (A->matrix_type != LIS_MATRIX_CSR) {            /* factor a CSR copy (lis_precon_iluk.c:117-133) */
        err = lis_matrix_duplicate(A, &B);
        if (err) return err;
        lis_matrix_set_type(B, LIS_MATRIX_CSR);
        err = lis_matrix_convert(A, B);
        if (err) { lis_matrix_destroy(B); return err; }
        A = B;
    }
    int nb = lis_host_num_threads();
    if (nb < 1) nb = 1;
    if (nb > A->n && A->n > 0) nb = A->n;
    lisd_ilu *F = (lisd_ilu *)calloc(1, sizeof(lisd_ilu));
    if (!F) { if (B) lis_matrix_destroy(B); LIS_SETERR_MEM(sizeof(lisd_ilu)); return LIS_OUT_OF_MEMORY; }
    F->n = A->n;
    precon->b200_ilu = F;
    err = lis_vector_duplicate(solver->A, &precon->D);
    LIS_SCALAR *d = NULL;
    if (!err) { d = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(A->n > 0 ? A->n : 1)); if (!d) { LIS_SETERR_MEM(A->n * 8); err = LIS_OUT_OF_MEMORY; } }
    if (!err && solver->options[LIS_OPTIONS_PRECON] == LIS_PRECON_TYPE_ILUT) {
        err = ilut_factor(A, solver->params[LIS_PARAMS_DROP - LIS_OPTIONS_LEN], solver->params[LIS_PARAMS_RATE - LIS_OPTIONS_LEN], nb, F, d);
    } else {
        if (!err) err = ilu_symbolic(A, solver->options[LIS_OPTIONS_FILL], nb, F);
        if (!err) err = ilu_numeric(A, nb, F, d);
    }
    if (!err && A->n > 0) err = lis_vector_set_values2(LIS_INS_VALUE, precon->D->is + precon->D->origin, A->n, d, precon->D);
    free(d);
    if (B) lis_matrix_destroy(B);
    if (err) return err;
    if (!lisd_available()) return LIS_SUCCESS;          /* host-only use: the apply reports the missing device */
    err = lisd_tri_build(F->n, F->L.ptr, F->L.idx, F->L.val, &F->tL);
    if (!err) err = lisd_tri_build(F->n, F->U.ptr, F->U.idx, F->U.val, &F->tU);
    if (!err) err = lisd_malloc((void **)&F->d_w, sizeof(double) * (size_t)(F->n > 0 ? F->n : 1));
    return err;
}

/* x = U^-1 L^-1 b: src/precon/lis_precon_iluk.c:1028-1049 */
LIS_INT lis_psolve_iluk(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    LIS_PRECON precon = solver->precon;
    lisd_ilu *F = (lisd_ilu *)precon->b200_ilu;
    LIS_INT err = lisd_require("lis_psolve_iluk");
    if (err) return err;
    if (F == NULL || F->tL == NULL) { LIS_SETERR(LIS_ERR_ILL_ARG, "ILU factors are not set up\n"); return LIS_ERR_ILL_ARG; }
    err = lisd_vec_device(b);
    if (!err) err = lisd_vec_device(x);
    if (!err) err = lisd_vec_device(precon->D);
    if (err) return err;
    err = lisd_tri_solve(F->tL, 1, NULL, b->value, F->d_w, "ILU forward solve");
    if (err) return err;
    return lisd_tri_solve(F->tU, 0, precon->D->value, F->d_w, x->value, "ILU backward solve");
}

/* x = L^-T U^-T b: src/precon/lis_precon_iluk.c:1224-1240 */
LIS_INT lis_psolveh_iluk(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    LIS_PRECON precon = solver->precon;
    lisd_ilu *F = (lisd_ilu *)precon->b200_ilu;
    LIS_INT err = lisd_require("lis_psolveh_iluk");
    if (err) return err;
    if (F == NULL || F->tL == NULL) { LIS_SETERR(LIS_ERR_ILL_ARG, "ILU factors are not set up\n"); return LIS_ERR_ILL_ARG; }
    if (F->tUT == NULL) {
        ilu_rows T;
        err = rows_transpose(F->n, &F->U, 0, &T);
        if (!err) { err = lisd_tri_build(F->n, T.ptr, T.idx, T.val, &F->tUT); rows_free(&T); }
        if (!err) err = rows_transpose(F->n, &F->L, 1, &T);
        if (!err) { err = lisd_tri_build(F->n, T.ptr, T.idx, T.val, &F->tLT); rows_free(&T); }
        if (err) return err;
    }
    err = lisd_vec_device(b);
    if (!err) err = lisd_vec_device(x);
    if (!err) err = lisd_vec_device(precon->D);
    if (err) return err;
    err = lisd_tri_solve(F->tUT, 0, precon->D->value, b->value, F->d_w, "ILU transposed forward solve");
    if (err) return err;
    return lisd_tri_solve(F->tLT, 1, NULL, F->d_w, x->value, "ILU transposed backward solve");
}
