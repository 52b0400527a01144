/*
 * lis_oracle.c -- CPU restatement of the Lis 2.1.11 SpMV/Krylov hot path (see lis_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg as the checker.  Never linked into the product.
 * Parity status: PINNED against the compiled reference (oracle/_ref) and tests/golden/.
 *
 * Build with -ffp-contract=off (oracle/Makefile): the reference's default flags
 * (-O3 -fomit-frame-pointer, no -march; configure.ac:505) never emit FMA, so every
 * multiply and add below must round separately.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lis_oracle.h"

/* include/lis.h:1067-1078 */
void orc_get_isie(int id, int nprocs, int n, int *is, int *ie)
{
    int len;
    if (id < n % nprocs) { len = n / nprocs + 1; *is = len * id; }
    else                 { len = n / nprocs;     *is = len * id + n % nprocs; }
    *ie = *is + len;
}

/* ============================== SpMV =================================================== */

/* src/matvec/lis_matvec_csr.c:90-110 (unsplit branch) */
void orc_spmv_csr(int n, const int *ptr, const int *idx, const double *val, const double *x, double *y)
{
    for (int i = 0; i < n; i++) {
        double t = 0.0;
        for (int j = ptr[i]; j < ptr[i + 1]; j++) t += val[j] * x[idx[j]];
        y[i] = t;
    }
}

/* src/matvec/lis_matvec_csr.c:64-87 (is_splited branch): D, then L, then U */
void orc_spmv_csr_split(int n, const double *diag, const int *lptr, const int *lidx, const double *lval,
                        const int *uptr, const int *uidx, const double *uval, const double *x, double *y)
{
    for (int i = 0; i < n; i++) {
        double t = diag[i] * x[i];
        for (int j = lptr[i]; j < lptr[i + 1]; j++) t += lval[j] * x[lidx[j]];
        for (int j = uptr[i]; j < uptr[i + 1]; j++) t += uval[j] * x[uidx[j]];
        y[i] = t;
    }
}

/* src/matvec/lis_matvec_ell.c:92-130: y zeroed, then one sweep per slot (slot-major) */
void orc_spmv_ell(int n, int maxnzr, const int *idx, const double *val, const double *x, double *y)
{
    for (int i = 0; i < n; i++) y[i] = 0.0;
    for (int j = 0; j < maxnzr; j++) {
        const size_t jj = (size_t)j * n;
        for (int i = 0; i < n; i++) y[i] += val[jj + i] * x[idx[jj + i]];
    }
}

/* src/matvec/lis_matvec_dia.c:126-174: per "thread" block [is,ie), value is thread-blocked:
 * value[is*nnd + j*(ie-is) + (i-is)] */
void orc_spmv_dia(int n, int nnd, const int *off, const double *val, const double *x, double *y, int nthreads)
{
    for (int t = 0; t < nthreads; t++) {
        int is, ie;
        orc_get_isie(t, nthreads, n, &is, &ie);
        for (int i = is; i < ie; i++) y[i] = 0.0;
        for (int j = 0; j < nnd; j++) {
            const int jj = off[j];
            const int js = is > -jj ? is : -jj;
            const int je = ie < n - jj ? ie : n - jj;
            const size_t k = (size_t)is * nnd + (size_t)j * (ie - is);
            for (int i = js; i < je; i++) y[i] += val[k + (i - is)] * x[jj + i];
        }
    }
}

/* src/matvec/lis_matvec_jad.c:144-198: per-thread jagged diagonals, ptr has
 * nthreads*(maxnzr+1) entries; w accumulates, then y[perm[i]] = w[i] */
void orc_spmv_jad(int n, int maxnzr, const int *jptr, const int *perm, const int *idx, const double *val,
                  const double *x, double *y, int nthreads)
{
    double *w = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int t = 0; t < nthreads; t++) {
        int is, ie;
        orc_get_isie(t, nthreads, n, &is, &ie);
        for (int i = is; i < ie; i++) w[i] = 0.0;
        for (int j = 0; j < maxnzr; j++) {
            int k = is;
            const int js = jptr[t * (maxnzr + 1) + j], je = jptr[t * (maxnzr + 1) + j + 1];
            for (int i = js; i < je; i++) { w[k] += val[i] * x[idx[i]]; k++; }
        }
        for (int i = is; i < ie; i++) y[perm[i]] = w[i];
    }
    free(w);
}

/* src/matvec/lis_matvec_bsr.c:125-148 (generic) == the unrolled RxC kernels :152-858:
 * per block row, blocks in storage order, inside a block column by column */
void orc_spmv_bsr(int n, int nr, int bnr, int bnc, const int *bptr, const int *bidx, const double *val,
                  const double *x, double *y)
{
    const int bs = bnr * bnc;
    double *t = (double *)malloc(sizeof(double) * (size_t)bnr);
    for (int bi = 0; bi < nr; bi++) {
        for (int i = 0; i < bnr; i++) t[i] = 0.0;
        for (int bc = bptr[bi]; bc < bptr[bi + 1]; bc++) {
            const int bj = bidx[bc] * bnc;
            size_t k = (size_t)bc * bs;
            for (int j = 0; j < bnc; j++)
                for (int i = 0; i < bnr; i++) { t[i] += val[k] * x[bj + j]; k++; }
        }
        for (int i = 0; i < bnr; i++)
            if (bi * bnr + i < n) y[bi * bnr + i] = t[i];
    }
    free(t);
}

/* src/matvec/lis_matvec_csc.c:128-144 (serial branch): column scatter */
void orc_spmv_csc(int n, const int *ptr, const int *idx, const double *val, const double *x, double *y)
{
    for (int i = 0; i < n; i++) y[i] = 0.0;
    for (int i = 0; i < n; i++) {
        const double t = x[i];
        for (int j = ptr[i]; j < ptr[i + 1]; j++) y[idx[j]] += val[j] * t;
    }
}

/* ============================== format builders ========================================= */

/* lis_matrix_sort_csr, src/matrix/lis_matrix_csr.c:1486-1521: every row ascending by column.
 * (Stable insertion sort here; identical to the reference's quicksort when a row has no
 * duplicate columns, which is all the reference itself guarantees.) */
void orc_sort_csr_rows(int n, const int *ptr, int *idx, double *val)
{
    for (int i = 0; i < n; i++) {
        for (int j = ptr[i] + 1; j < ptr[i + 1]; j++) {
            const int c = idx[j]; const double v = val[j];
            int k = j - 1;
            while (k >= ptr[i] && idx[k] > c) { idx[k + 1] = idx[k]; val[k + 1] = val[k]; k--; }
            idx[k + 1] = c; val[k + 1] = v;
        }
    }
}

/* src/matrix/lis_matrix_ell.c:1000-1006 */
int orc_csr2ell_maxnzr(int n, const int *ptr)
{
    int m = 0;
    for (int i = 0; i < n; i++) if (ptr[i + 1] - ptr[i] > m) m = ptr[i + 1] - ptr[i];
    return m;
}

/* src/matrix/lis_matrix_ell.c:1029-1052: pad = (0.0, column i), slot k = k-th stored entry */
void orc_csr2ell(int n, const int *ptr, const int *idx, const double *val, int maxnzr, int *eidx, double *eval)
{
    for (int j = 0; j < maxnzr; j++)
        for (int i = 0; i < n; i++) { eval[(size_t)j * n + i] = 0.0; eidx[(size_t)j * n + i] = i; }
    for (int i = 0; i < n; i++) {
        int k = 0;
        for (int j = ptr[i]; j < ptr[i + 1]; j++, k++) {
            eval[(size_t)k * n + i] = val[j];
            eidx[(size_t)k * n + i] = idx[j];
        }
    }
}

static int cmp_int(const void *a, const void *b)
{
    const int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* distinct offsets (column - row), ascending: src/matrix/lis_matrix_dia.c:1219-1237 */
static int dia_offsets(int n, const int *ptr, const int *idx, int *out /* may be NULL */)
{
    const int nnz = ptr[n];
    if (nnz == 0) return 0;
    int *iw = (int *)malloc(sizeof(int) * (size_t)nnz);
    for (int i = 0; i < n; i++)
        for (int j = ptr[i]; j < ptr[i + 1]; j++) iw[j] = idx[j] - i;
    qsort(iw, (size_t)nnz, sizeof(int), cmp_int);
    int nnd = 0;
    for (int i = 0; i < nnz; i++)
        if (i == 0 || iw[i] != iw[i - 1]) { if (out) out[nnd] = iw[i]; nnd++; }
    free(iw);
    return nnd;
}

int orc_csr2dia_nnd(int n, const int *ptr, const int *idx) { return dia_offsets(n, ptr, idx, NULL); }

/* src/matrix/lis_matrix_dia.c:1246-1284: rows must already be sorted (:1217); layout is
 * thread-blocked value[is*nnd + k*(ie-is) + (i-is)], explicit zeros elsewhere */
void orc_csr2dia(int n, const int *ptr, const int *idx, const double *val, int nnd, int *off, double *dval, int nthreads)
{
    dia_offsets(n, ptr, idx, off);
    memset(dval, 0, sizeof(double) * (size_t)n * (size_t)nnd);
    for (int t = 0; t < nthreads; t++) {
        int is, ie;
        orc_get_isie(t, nthreads, n, &is, &ie);
        for (int i = is; i < ie; i++) {
            int k = 0;
            for (int j = ptr[i]; j < ptr[i + 1]; j++) {
                const int jj = idx[j] - i;
                while (jj != off[k]) k++;
                dval[(size_t)is * nnd + (size_t)k * (ie - is) + (i - is)] = val[j];
            }
        }
    }
}

int orc_csr2jad_maxnzr(int n, const int *ptr) { return orc_csr2ell_maxnzr(n, ptr); }

/* src/matrix/lis_matrix_jad.c:1682-1751: per thread block, rows ordered by descending length
 * (order among equal lengths is an artefact of the reference's quicksort and does not affect
 * y; a stable order is used here), j-th jagged diagonal = j-th stored entry of each row */
void orc_csr2jad(int n, const int *ptr, const int *idx, const double *val, int maxnzr,
                 int *perm, int *jptr, int *jidx, double *jval, int nthreads)
{
    for (int t = 0; t < nthreads; t++) {
        int is, ie;
        orc_get_isie(t, nthreads, n, &is, &ie);
        int *jp = jptr + t * (maxnzr + 1);
        memset(jp, 0, sizeof(int) * (size_t)(maxnzr + 1));
        /* counting sort by descending length, stable in row index */
        int *cnt = (int *)calloc((size_t)maxnzr + 2, sizeof(int));
        for (int i = is; i < ie; i++) {
            const int len = ptr[i + 1] - ptr[i];
            cnt[maxnzr - len + 1]++;
            for (int j = 0; j < len; j++) jp[j + 1]++;
        }
        for (int l = 0; l <= maxnzr; l++) cnt[l + 1] += cnt[l];
        for (int i = is; i < ie; i++) {
            const int len = ptr[i + 1] - ptr[i];
            perm[is + cnt[maxnzr - len]++] = i;
        }
        free(cnt);
        jp[0] = ptr[is];
        for (int j = 0; j < maxnzr; j++) jp[j + 1] += jp[j];
        for (int i = is; i < ie; i++) {
            const int js = ptr[perm[i]], je = ptr[perm[i] + 1];
            for (int j = js; j < je; j++) {
                const int l = jp[j - js] + i - is;
                jval[l] = val[j];
                jidx[l] = idx[j];
            }
        }
    }
}

/* src/matrix/lis_matrix_bsr.c:398-446: blocks per block row, first-seen order */
int orc_csr2bsr_bnnz(int n, const int *ptr, const int *idx, int bnr, int bnc, int *bptr)
{
    const int nr = 1 + (n - 1) / bnr, nc = 1 + (n - 1) / bnc;
    char *seen = (char *)calloc((size_t)nc, 1);
    int *list = (int *)malloc(sizeof(int) * (size_t)nc);
    bptr[0] = 0;
    for (int bi = 0; bi < nr; bi++) {
        int cnt = 0;
        for (int ii = 0; ii < bnr && bi * bnr + ii < n; ii++)
            for (int j = ptr[bi * bnr + ii]; j < ptr[bi * bnr + ii + 1]; j++) {
                const int bj = idx[j] / bnc;
                if (!seen[bj]) { seen[bj] = 1; list[cnt++] = bj; }
            }
        for (int k = 0; k < cnt; k++) seen[list[k]] = 0;
        bptr[bi + 1] = bptr[bi] + cnt;
    }
    free(seen); free(list);
    return bptr[nr];
}

/* src/matrix/lis_matrix_bsr.c:469-525: block (bi,bj) column-major, ij = j*bnr + ii */
void orc_csr2bsr(int n, const int *ptr, const int *idx, const double *val, int bnr, int bnc,
                 const int *bptr, int *bidx, double *bval)
{
    const int nr = 1 + (n - 1) / bnr, nc = 1 + (n - 1) / bnc, bs = bnr * bnc;
    int *pos = (int *)calloc((size_t)nc, sizeof(int));      /* 1 + value offset of block, 0 = none */
    for (int bi = 0; bi < nr; bi++) {
        int kk = bptr[bi];
        for (int ii = 0; ii < bnr && bi * bnr + ii < n; ii++)
            for (int k = ptr[bi * bnr + ii]; k < ptr[bi * bnr + ii + 1]; k++) {
                const int bj = idx[k] / bnc, j = idx[k] % bnc;
                if (pos[bj] == 0) {
                    const int kv = kk * bs;
                    pos[bj] = kv + 1;
                    bidx[kk] = bj;
                    for (int q = 0; q < bs; q++) bval[kv + q] = 0.0;
                    bval[kv + j * bnr + ii] = val[k];
                    kk++;
                } else {
                    bval[pos[bj] - 1 + j * bnr + ii] = val[k];
                }
            }
        for (int j = bptr[bi]; j < bptr[bi + 1]; j++) pos[bidx[j]] = 0;
    }
    free(pos);
}

/* src/matrix/lis_matrix_csc.c:1036-1065: counting transpose, rows ascending inside a column */
void orc_csr2csc(int n, const int *ptr, const int *idx, const double *val, int *cptr, int *cidx, double *cval)
{
    int *iw = (int *)calloc((size_t)n + 1, sizeof(int));
    for (int i = 0; i < n; i++)
        for (int j = ptr[i]; j < ptr[i + 1]; j++) iw[idx[j]]++;
    cptr[0] = 0;
    for (int i = 0; i < n; i++) { cptr[i + 1] = cptr[i] + iw[i]; iw[i] = cptr[i]; }
    for (int i = 0; i < n; i++)
        for (int j = ptr[i]; j < ptr[i + 1]; j++) {
            const int l = iw[idx[j]]++;
            cval[l] = val[j];
            cidx[l] = i;
        }
    free(iw);
}

/* ============================== BLAS-1 ================================================== */

/* OpenMP `omp for` with the default static schedule: thread k gets one contiguous chunk;
 * ceil-sized chunks first (libgomp: q = n/nt, r = n%nt, the first r threads get q+1). */
static void omp_static_chunk(int k, int nt, int n, int *s, int *e)
{
    const int q = n / nt, r = n % nt;
    if (k < r) { *s = k * (q + 1); *e = *s + q + 1; }
    else       { *s = k * q + r;   *e = *s + q; }
}

/* src/vector/lis_vector_ops.c:89-117 */
double orc_dot(int n, const double *x, const double *y, int nthreads)
{
    double dot = 0.0;
    for (int k = 0; k < nthreads; k++) {
        int s, e; omp_static_chunk(k, nthreads, n, &s, &e);
        double tmp = 0.0;
        for (int i = s; i < e; i++) tmp += x[i] * y[i];
        if (nthreads == 1) return tmp;     /* serial build: one running sum (:110-117) */
        dot += tmp;
    }
    return dot;
}

/* src/vector/lis_vector_ops.c:236-266 */
double orc_nrm2(int n, const double *x, int nthreads) { return sqrt(orc_dot(n, x, x, nthreads)); }

/* src/vector/lis_vector_ops.c:278-340 */
double orc_nrm1(int n, const double *x, int nthreads)
{
    double sum = 0.0;
    for (int k = 0; k < nthreads; k++) {
        int s, e; omp_static_chunk(k, nthreads, n, &s, &e);
        double tmp = 0.0;
        for (int i = s; i < e; i++) tmp += fabs(x[i]);
        if (nthreads == 1) return tmp;
        sum += tmp;
    }
    return sum;
}

/* src/vector/lis_vector_ops.c:344-414 */
double orc_nrmi(int n, const double *x)
{
    double m = 0.0;
    for (int i = 0; i < n; i++) if (fabs(x[i]) > m) m = fabs(x[i]);
    return m;
}

/* src/vector/lis_vector_ops.c:418-480 */
double orc_sum(int n, const double *x, int nthreads)
{
    double sum = 0.0;
    for (int k = 0; k < nthreads; k++) {
        int s, e; omp_static_chunk(k, nthreads, n, &s, &e);
        double tmp = 0.0;
        for (int i = s; i < e; i++) tmp += x[i];
        if (nthreads == 1) return tmp;
        sum += tmp;
    }
    return sum;
}

/* src/vector/lis_vector_opv.c:176, :216, :256, :287, :328, :368, :460, :522, :430 */
void orc_axpy(int n, double a, const double *x, double *y)  { for (int i = 0; i < n; i++) y[i] += a * x[i]; }
void orc_xpay(int n, const double *x, double a, double *y)  { for (int i = 0; i < n; i++) y[i] = x[i] + a * y[i]; }
void orc_axpyz(int n, double a, const double *x, const double *y, double *z) { for (int i = 0; i < n; i++) z[i] = a * x[i] + y[i]; }
void orc_scale(int n, double a, double *x)                  { for (int i = 0; i < n; i++) x[i] = a * x[i]; }
void orc_pmul(int n, const double *x, const double *y, double *z) { for (int i = 0; i < n; i++) z[i] = x[i] * y[i]; }
void orc_pdiv(int n, const double *x, const double *y, double *z) { for (int i = 0; i < n; i++) z[i] = x[i] / y[i]; }
void orc_reciprocal(int n, double *x)                       { for (int i = 0; i < n; i++) x[i] = 1.0 / x[i]; }
void orc_shift(int n, double s, double *x)                  { for (int i = 0; i < n; i++) x[i] = x[i] - s; }
void orc_abs(int n, double *x)                              { for (int i = 0; i < n; i++) x[i] = fabs(x[i]); }

/* ============================== preconditioner pieces =================================== */

/* src/matrix/lis_matrix_csr.c:540-553 */
void orc_csr_get_diagonal(int n, const int *ptr, const int *idx, const double *val, double *d)
{
    for (int i = 0; i < n; i++) {
        d[i] = 0.0;
        for (int j = ptr[i]; j < ptr[i + 1]; j++)
            if (idx[j] == i) { d[i] = val[j]; break; }
    }
}

/* src/matrix/lis_matrix_csr.c:829-846 */
void orc_csr_split_count(int n, const int *ptr, const int *idx, int *nnzl, int *nnzu)
{
    int l = 0, u = 0;
    for (int i = 0; i < n; i++)
        for (int j = ptr[i]; j < ptr[i + 1]; j++) {
            if (idx[j] < i) l++; else if (idx[j] > i) u++;
        }
    *nnzl = l; *nnzu = u;
}

/* src/matrix/lis_matrix_csr.c:903-932: storage order kept inside L and U; the LAST diagonal
 * entry of a row wins.  Rows without a stored diagonal: the reference leaves D[i] as whatever
 * lis_malloc returned (lis_matrix_diag_duplicateM never initialises it) -- undefined there,
 * defined as 0 here and in lis_b200. */
void orc_csr_split(int n, const int *ptr, const int *idx, const double *val,
                   int *lptr, int *lidx, double *lval, int *uptr, int *uidx, double *uval, double *diag)
{
    int l = 0, u = 0;
    lptr[0] = 0; uptr[0] = 0;
    for (int i = 0; i < n; i++) {
        diag[i] = 0.0;
        for (int j = ptr[i]; j < ptr[i + 1]; j++) {
            if (idx[j] < i)      { lidx[l] = idx[j]; lval[l] = val[j]; l++; }
            else if (idx[j] > i) { uidx[u] = idx[j]; uval[u] = val[j]; u++; }
            else diag[i] = val[j];
        }
        lptr[i + 1] = l; uptr[i + 1] = u;
    }
}

/* src/matrix/lis_matrix_csr.c:1578-1628.  nthreads>1: independent block per thread, couplings
 * leaving the block dropped (:1590, :1601); nthreads==1: the global sweep (:1607-1628) */
void orc_ssor_sweep(int n, const int *lptr, const int *lidx, const double *lval,
                    const int *uptr, const int *uidx, const double *uval, const double *wd,
                    const double *b, double *x, int nthreads)
{
    for (int k = 0; k < nthreads; k++) {
        int is, ie;
        orc_get_isie(k, nthreads, n, &is, &ie);
        for (int i = is; i < ie; i++) {
            double t = b[i];
            for (int j = lptr[i]; j < lptr[i + 1]; j++) {
                const int jj = lidx[j];
                if (jj < is) continue;
                t -= lval[j] * x[jj];
            }
            x[i] = t * wd[i];
        }
        for (int i = ie - 1; i >= is; i--) {
            double t = 0.0;
            for (int j = uptr[i]; j < uptr[i + 1]; j++) {
                const int jj = uidx[j];
                if (jj < is || jj >= ie) continue;
                t += uval[j] * x[jj];
            }
            x[i] -= t * wd[i];
        }
    }
}

/* ============================== Krylov drivers ========================================== */

typedef struct {
    int n; const int *ptr, *idx; const double *val;
    int precon, nthreads, ssor_blocks;
    /* jacobi */ double *dinv;
    /* ssor   */ int *lptr, *lidx, *uptr, *uidx; double *lval, *uval, *diag, *wd;
} orc_sys_t;

/* lis_precon_create_{jacobi,ssor}: src/precon/lis_precon_jacobi.c:60-86, lis_precon_ssor.c:57-95
 * WD = 1/(omega*D): src/matrix/lis_matrix_diag.c:663-672 (scale) and :775-783 (inverse) */
static void sys_setup(orc_sys_t *S, int n, const int *ptr, const int *idx, const double *val, const orc_solver_t *s)
{
    memset(S, 0, sizeof(*S));
    S->n = n; S->ptr = ptr; S->idx = idx; S->val = val;
    S->precon = s->precon; S->nthreads = s->nthreads > 0 ? s->nthreads : 1;
    S->ssor_blocks = s->ssor_blocks > 0 ? s->ssor_blocks : S->nthreads;
    if (s->precon == 1) {
        S->dinv = (double *)malloc(sizeof(double) * (size_t)n);
        orc_csr_get_diagonal(n, ptr, idx, val, S->dinv);
        orc_reciprocal(n, S->dinv);
    } else if (s->precon == 3) {
        int nl, nu;
        orc_csr_split_count(n, ptr, idx, &nl, &nu);
        S->lptr = (int *)malloc(sizeof(int) * ((size_t)n + 1)); S->uptr = (int *)malloc(sizeof(int) * ((size_t)n + 1));
        S->lidx = (int *)malloc(sizeof(int) * (size_t)(nl + 1)); S->uidx = (int *)malloc(sizeof(int) * (size_t)(nu + 1));
        S->lval = (double *)malloc(sizeof(double) * (size_t)(nl + 1)); S->uval = (double *)malloc(sizeof(double) * (size_t)(nu + 1));
        S->diag = (double *)malloc(sizeof(double) * (size_t)n); S->wd = (double *)malloc(sizeof(double) * (size_t)n);
        orc_csr_split(n, ptr, idx, val, S->lptr, S->lidx, S->lval, S->uptr, S->uidx, S->uval, S->diag);
        for (int i = 0; i < n; i++) S->wd[i] = 1.0 / (s->ssor_omega * S->diag[i]);
    }
}

static void sys_free(orc_sys_t *S)
{
    free(S->dinv); free(S->lptr); free(S->lidx); free(S->uptr); free(S->uidx);
    free(S->lval); free(S->uval); free(S->diag); free(S->wd);
}

/* lis_matvec on the solver's matrix: split order once SSOR setup has split A */
static void sys_matvec(const orc_sys_t *S, const double *x, double *y)
{
    if (S->precon == 3)
        orc_spmv_csr_split(S->n, S->diag, S->lptr, S->lidx, S->lval, S->uptr, S->uidx, S->uval, x, y);
    else
        orc_spmv_csr(S->n, S->ptr, S->idx, S->val, x, y);
}

/* lis_psolve: none = copy (src/precon/lis_precon.c lis_psolve_none), jacobi, ssor */
static void sys_psolve(const orc_sys_t *S, const double *b, double *x)
{
    if (S->precon == 1) orc_pmul(S->n, b, S->dinv, x);
    else if (S->precon == 3)
        orc_ssor_sweep(S->n, S->lptr, S->lidx, S->lval, S->uptr, S->uidx, S->uval, S->wd, b, x, S->ssor_blocks);
    else memcpy(x, b, sizeof(double) * (size_t)S->n);
}

/* lis_solver_get_initial_residual with the default zero initial guess and nrm2_r criterion,
 * src/solver/lis_solver.c:957-1090.  Returns nonzero when already converged. */
static int initial_residual(const orc_sys_t *S, const double *b, const double *x, int x_is_zero,
                            double *r, double *bnrm2, orc_solver_t *s, double *rhistory)
{
    const int n = S->n;
    if (!x_is_zero) { sys_matvec(S, x, r); orc_xpay(n, b, -1.0, r); }
    else memcpy(r, b, sizeof(double) * (size_t)n);
    double nrm2 = orc_nrm2(n, r, S->nthreads);
    *bnrm2 = nrm2;
    if (*bnrm2 == 0.0) *bnrm2 = 1.0; else *bnrm2 = 1.0 / *bnrm2;
    nrm2 = nrm2 * *bnrm2;
    (void)rhistory;
    if (nrm2 <= fabs(s->tol)) { s->retcode = 0; s->iter = 1; s->resid = nrm2; return 1; }
    return 0;
}

static int is_zero_vec(int n, const double *x)
{
    for (int i = 0; i < n; i++) if (x[i] != 0.0) return 0;
    return 1;
}

/* src/solver/lis_solver_cg.c:129-235 */
int orc_cg(int n, const int *ptr, const int *idx, const double *val, const double *b, double *x,
           orc_solver_t *s, double *rhistory)
{
    orc_sys_t S; sys_setup(&S, n, ptr, idx, val, s);
    const int nt = S.nthreads;
    double *z = (double *)calloc((size_t)n, sizeof(double)), *q = (double *)calloc((size_t)n, sizeof(double));
    double *r = (double *)calloc((size_t)n, sizeof(double)), *p = (double *)calloc((size_t)n, sizeof(double));
    double rho_old = 1.0, beta, rho, alpha, dot_pq, bnrm2, nrm2 = 0.0;
    int iter, ret = 4;
    rhistory[0] = 1.0;
    if (initial_residual(&S, b, x, is_zero_vec(n, x), r, &bnrm2, s, rhistory)) { ret = 0; goto done; }
    for (iter = 1; iter <= s->maxiter; iter++) {
        sys_psolve(&S, r, z);
        rho = orc_dot(n, r, z, nt);
        beta = rho / rho_old;
        orc_xpay(n, z, beta, p);
        sys_matvec(&S, p, q);
        dot_pq = orc_dot(n, p, q, nt);
        if (dot_pq == 0.0) { s->retcode = 2; s->iter = iter; s->resid = nrm2; ret = 2; goto done; }
        alpha = rho / dot_pq;
        orc_axpy(n, alpha, p, x);
        orc_axpy(n, -alpha, q, r);
        nrm2 = orc_nrm2(n, r, nt) * bnrm2;
        rhistory[iter] = nrm2;
        if (s->tol >= nrm2) { s->retcode = 0; s->iter = iter; s->resid = nrm2; ret = 0; goto done; }
        rho_old = rho;
    }
    s->retcode = 4; s->iter = iter; s->resid = nrm2;
done:
    free(z); free(q); free(r); free(p); sys_free(&S);
    return ret;
}

/* src/solver/lis_solver_bicgstab.c:137-315; s aliases r (:160-161); rtld = r0 (lis_solver.c:1862) */
int orc_bicgstab(int n, const int *ptr, const int *idx, const double *val, const double *b, double *x,
                 orc_solver_t *s, double *rhistory)
{
    orc_sys_t S; sys_setup(&S, n, ptr, idx, val, s);
    const int nt = S.nthreads;
    double *rtld = (double *)calloc((size_t)n, sizeof(double)), *r = (double *)calloc((size_t)n, sizeof(double));
    double *t = (double *)calloc((size_t)n, sizeof(double)), *p = (double *)calloc((size_t)n, sizeof(double));
    double *v = (double *)calloc((size_t)n, sizeof(double)), *phat = (double *)calloc((size_t)n, sizeof(double));
    double *shat = (double *)calloc((size_t)n, sizeof(double));
    double alpha = 1.0, omega = 1.0, rho_old = 1.0, beta, rho, d1, d2, bnrm2, nrm2 = 0.0;
    int iter, ret = 4;
    rhistory[0] = 1.0;
    if (initial_residual(&S, b, x, is_zero_vec(n, x), r, &bnrm2, s, rhistory)) { ret = 0; goto done; }
    memcpy(rtld, r, sizeof(double) * (size_t)n);
    for (iter = 1; iter <= s->maxiter; iter++) {
        rho = orc_dot(n, rtld, r, nt);
        if (rho == 0.0) { s->retcode = 2; s->iter = iter; s->resid = nrm2; ret = 2; goto done; }
        if (iter == 1) memcpy(p, r, sizeof(double) * (size_t)n);
        else {
            beta = (rho / rho_old) * (alpha / omega);
            orc_axpy(n, -omega, v, p);
            orc_xpay(n, r, beta, p);
        }
        sys_psolve(&S, p, phat);
        sys_matvec(&S, phat, v);
        d1 = orc_dot(n, rtld, v, nt);
        alpha = rho / d1;
        orc_axpy(n, -alpha, v, r);
        nrm2 = orc_nrm2(n, r, nt) * bnrm2;
        if (nrm2 <= s->tol) {
            rhistory[iter] = nrm2;
            orc_axpy(n, alpha, phat, x);
            s->retcode = 0; s->iter = iter; s->resid = nrm2; ret = 0; goto done;
        }
        sys_psolve(&S, r, shat);
        sys_matvec(&S, shat, t);
        d1 = orc_dot(n, t, r, nt);
        d2 = orc_dot(n, t, t, nt);
        omega = d1 / d2;
        orc_axpy(n, alpha, phat, x);
        orc_axpy(n, omega, shat, x);
        orc_axpy(n, -omega, t, r);
        nrm2 = orc_nrm2(n, r, nt) * bnrm2;
        rhistory[iter] = nrm2;
        if (s->tol >= nrm2) { s->retcode = 0; s->iter = iter; s->resid = nrm2; ret = 0; goto done; }
        if (omega == 0.0) { s->retcode = 2; s->iter = iter; s->resid = nrm2; ret = 2; goto done; }
        rho_old = rho;
    }
    s->retcode = 4; s->iter = iter; s->resid = nrm2;
done:
    free(rtld); free(r); free(t); free(p); free(v); free(phat); free(shat); sys_free(&S);
    return ret;
}

/* src/solver/lis_solver_gmres.c:135-343.  Quirks kept: v[0] is first M^-1(b-Ax) (:179-181) and
 * then overwritten with the unpreconditioned residual (:184); the restart residual is rebuilt
 * from the basis (:321-333); MAXITER returns iter+1 (:337). */
int orc_gmres(int n, const int *ptr, const int *idx, const double *val, const double *b, double *x,
              orc_solver_t *s, double *rhistory)
{
    orc_sys_t S; sys_setup(&S, n, ptr, idx, val, s);
    const int nt = S.nthreads;
    const int m = s->restart, h_dim = m + 1;
    const int cs = (m + 1) * h_dim, sn = (m + 2) * h_dim;
    double *h = (double *)calloc((size_t)(h_dim + 1) * (size_t)(h_dim + 2), sizeof(double));
    double *sv = (double *)calloc((size_t)m + 2, sizeof(double));
    double *r = (double *)calloc((size_t)n, sizeof(double)), *z = (double *)calloc((size_t)n, sizeof(double));
    double **v = (double **)malloc(sizeof(double *) * (size_t)(m + 2));
    for (int k = 0; k < m + 2; k++) v[k] = (double *)calloc((size_t)n, sizeof(double));
    double bnrm2, nrm2 = 0.0, rnorm, t, aa, bb, rr, a2, b2;
    int iter = 0, ret = 4, i, ii = 0, i1 = 0, iih, k, j, jj;
    rhistory[0] = 1.0;

    sys_matvec(&S, x, z);
    orc_xpay(n, b, -1.0, z);
    sys_psolve(&S, z, v[0]);
    if (initial_residual(&S, b, x, is_zero_vec(n, x), v[0], &bnrm2, s, rhistory)) { ret = 0; goto done; }

    while (iter < s->maxiter) {
        rnorm = orc_nrm2(n, v[0], nt);
        orc_scale(n, 1.0 / rnorm, v[0]);
        for (k = 0; k < m + 1; k++) sv[k] = 0.0;
        sv[0] = rnorm;
        i = 0;
        do {
            iter++; i++;
            ii = i - 1; i1 = i; iih = (i - 1) * h_dim;
            sys_psolve(&S, v[ii], z);
            sys_matvec(&S, z, v[i1]);
            for (k = 0; k < i; k++) {
                t = orc_dot(n, v[i1], v[k], nt);
                h[k + iih] = t;
                orc_axpy(n, -t, v[k], v[i1]);
            }
            t = orc_nrm2(n, v[i1], nt);
            h[i1 + iih] = t;
            orc_scale(n, 1.0 / t, v[i1]);
            for (k = 1; k <= ii; k++) {
                jj = k - 1;
                t = h[jj + iih];
                aa = h[jj + cs] * t;
                aa += h[jj + sn] * h[k + iih];
                bb = -h[jj + sn] * t;
                bb += h[jj + cs] * h[k + iih];
                h[jj + iih] = aa;
                h[k + iih] = bb;
            }
            aa = h[ii + iih]; bb = h[i1 + iih];
            a2 = aa * aa; b2 = bb * bb;
            rr = sqrt(a2 + b2);
            if (rr == 0.0) rr = 1.0e-17;
            h[ii + cs] = aa / rr;
            h[ii + sn] = bb / rr;
            sv[i1] = -h[ii + sn] * sv[ii];
            sv[ii] = h[ii + cs] * sv[ii];
            aa = h[ii + cs] * h[ii + iih];
            aa += h[ii + sn] * h[i1 + iih];
            h[ii + iih] = aa;
            nrm2 = fabs(sv[i1]) * bnrm2;
            rhistory[iter] = nrm2;
            if (s->tol >= nrm2) break;
        } while (i < m && iter < s->maxiter);

        sv[ii] = sv[ii] / h[ii + ii * h_dim];
        for (k = 1; k <= ii; k++) {
            jj = ii - k;
            t = sv[jj];
            for (j = jj + 1; j <= ii; j++) t -= h[jj + j * h_dim] * sv[j];
            sv[jj] = t / h[jj + jj * h_dim];
        }
        for (k = 0; k < n; k++) z[k] = sv[0] * v[0][k];
        for (j = 1; j <= ii; j++) orc_axpy(n, sv[j], v[j], z);
        sys_psolve(&S, z, r);
        orc_axpy(n, 1.0, r, x);
        if (s->tol >= nrm2) { s->retcode = 0; s->iter = iter; s->resid = nrm2; ret = 0; goto done; }
        for (j = 1; j <= i; j++) {
            jj = i1 - j + 1;
            sv[jj - 1] = -h[jj - 1 + sn] * sv[jj];
            sv[jj] = h[jj - 1 + cs] * sv[jj];
        }
        for (j = 0; j <= i1; j++) {
            t = sv[j];
            if (j == 0) t = t - 1.0;
            orc_axpy(n, t, v[j], v[0]);
        }
    }
    s->retcode = 4; s->iter = iter + 1; s->resid = nrm2;
done:
    for (k = 0; k < m + 2; k++) free(v[k]);
    free(v); free(h); free(sv); free(r); free(z); sys_free(&S);
    return ret;
}
