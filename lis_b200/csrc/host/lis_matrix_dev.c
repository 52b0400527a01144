/*
 * lis_matrix_dev.c -- device mirrors of LIS_MATRIX storage and the lis_matvec dispatcher.
 *
 * The public arrays of a matrix (A->ptr/index/value, ...) stay where the caller put them
 * (host memory, same ownership rules as the reference, src/matrix/lis_matrix_csr.c:98-103).
 * The first kernel that needs the matrix uploads a private copy into HBM; the mirror is
 * dropped whenever the library itself changes the host arrays (sort, split, merge, convert,
 * destroy).  A caller who edits A->value behind the library's back after the first SpMV
 * must call lis_matrix_b200_invalidate(A).
 *
 * lis_matvec follows src/matvec/lis_matvec.c:55-187: argument check, halo exchange for
 * row-partitioned matrices (LIS_MATVEC_SENDRECV, include/lis_matvec.h:31-44), then the
 * per-format kernel.  Public entry points are host-synchronous because the reference drivers
 * time them with lis_wtime() (test/spmvtest1.c:219-221).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lis_device.h"
#include "lis_host.h"
#include "lis_b200_kernels.h"

/* ------------------------------------------------------------------ upload helpers */
static LIS_INT up(void **dst, const void *src, size_t bytes, size_t pad_bytes)
{
    LIS_INT err = lisd_malloc(dst, bytes + pad_bytes);
    if (err) return err;
    if (pad_bytes) { err = lisd_memset((char *)*dst + bytes, 0, pad_bytes); if (err) return err; }
    return lisd_upload(*dst, src, bytes);
}

static void csr_free(lisd_csr *c)
{
    lisd_free(c->ptr); lisd_free(c->idx); lisd_free(c->val);
    memset(c, 0, sizeof(*c));
}

/* idx/val are padded so that 128-bit loads starting at any multiple of 4 entries stay in
 * bounds (the CSR kernel rounds its window start down to a multiple of 4) */
static LIS_INT csr_upload(lisd_csr *c, int n, const LIS_INT *ptr, const LIS_INT *idx, const LIS_SCALAR *val)
{
    const size_t nnz = (size_t)ptr[n];
    LIS_INT err;
    memset(c, 0, sizeof(*c));
    c->n = n; c->nnz = (int)nnz;
    err = up((void **)&c->ptr, ptr, ((size_t)n + 1) * sizeof(int), 16);
    if (!err) err = up((void **)&c->idx, idx, nnz * sizeof(int), 8 * sizeof(int));
    if (!err) err = up((void **)&c->val, val, nnz * sizeof(double), 8 * sizeof(double));
    if (err) { csr_free(c); return err; }
    /* kernel choice, once per matrix: short-row matrices take the TMA-staged row-block kernel,
     * long / very ragged rows the product-tile kernel.  LIS_B200_CSR_KERNEL=tile|tma overrides
     * (tma only where the plan exists). */
    {
        const char *force = getenv("LIS_B200_CSR_KERNEL");
        int rows = 0, tile = 0, stages = 0;
        if (!(force && strcmp(force, "tile") == 0) && lisb200_spmv_csr_tma_plan(n, ptr, &rows, &tile, &stages) == 0) {
            c->tma_rows = rows; c->tma_tile = tile; c->tma_stages = stages;
        }
    }
    return LIS_SUCCESS;
}

static void mirror_free(lisd_matrix *M)
{
    if (M == NULL) return;
    csr_free(&M->csr); csr_free(&M->L); csr_free(&M->U);
    csr_free(&M->csrT); csr_free(&M->LT); csr_free(&M->UT);
    if (!(M->shared & LISD_SH_IDX)) lisd_free(M->idx);
    if (!(M->shared & LISD_SH_PERM)) lisd_free(M->perm);
    if (!(M->shared & LISD_SH_BIDX)) lisd_free(M->bidx);
    if (!(M->shared & LISD_SH_VAL)) lisd_free(M->val);
    lisd_free(M->off); lisd_free(M->jptr); lisd_free(M->bptr);
    lisd_free(M->diag); lisd_free(M->wd);
    free(M->pipe_row); free(M->pipe_need);
    if (M->sweep) lisd_sweep_free(M->sweep);
    if (M->sweep_global) lisd_sweep_free(M->sweep_global);
    free(M);
}

void lisd_mirror_free(lisd_matrix *M) { mirror_free(M); }

void lisd_matrix_drop(LIS_MATRIX A)
{
    if (A->b200_dev) { mirror_free((lisd_matrix *)A->b200_dev); A->b200_dev = NULL; }
}

LIS_INT lis_matrix_b200_invalidate(LIS_MATRIX A)
{
    if (!lis_is_malloc(A)) return LIS_ERR_ILL_ARG;
    lisd_matrix_drop(A);
    return LIS_SUCCESS;
}

/* lis_matrix_shift_diagonal changed the host arrays: make the same edit in HBM where the mirror is an
 * unsplit CSR or a split matrix (one small kernel instead of re-uploading the matrix -- Rayleigh
 * quotient iteration shifts twice per step), drop the mirror otherwise.  Transposed mirrors and
 * sweep schedules are rebuilt on demand: their "first diagonal entry" may be another stored one. */
LIS_INT lisd_matrix_shift_diagonal(LIS_MATRIX A, LIS_SCALAR sigma)
{
    lisd_matrix *M = (lisd_matrix *)A->b200_dev;
    if (M == NULL) return LIS_SUCCESS;
    const int usable = M->type == A->matrix_type && M->splited == A->is_splited && M->n == A->n &&
                       (M->splited || M->type == LIS_MATRIX_CSR) && M->sweep == NULL && M->sweep_global == NULL;
    if (!usable) { lisd_matrix_drop(A); return LIS_SUCCESS; }
    if (M->has_t) { csr_free(&M->csrT); csr_free(&M->LT); csr_free(&M->UT); M->has_t = 0; }
    lisd_mark_busy();
    if (M->splited) return lisd_check(lisb200_shift(A->n, sigma, M->diag, lisd_stream()), "lis_matrix_shift_diagonal");
    return lisd_check(lisb200_csr_shift_diagonal(A->n, M->csr.ptr, M->csr.idx, M->csr.val, sigma, lisd_stream()), "lis_matrix_shift_diagonal");
}

LIS_INT lisd_matrix_refresh_wd(LIS_MATRIX A)
{
    lisd_matrix *M = (lisd_matrix *)A->b200_dev;
    if (M == NULL || A->WD == NULL) return LIS_SUCCESS;
    if (M->wd == NULL) {
        LIS_INT err = lisd_malloc((void **)&M->wd, (size_t)(A->n > 0 ? A->n : 1) * sizeof(double));
        if (err) return err;
    }
    return lisd_upload(M->wd, A->WD->value, (size_t)A->n * sizeof(double));
}

LIS_INT lisd_matrix_get(LIS_MATRIX A, lisd_matrix **out)
{
    LIS_INT err = lisd_require("matrix upload");
    if (err) return err;
    lisd_matrix *M = (lisd_matrix *)A->b200_dev;
    if (M && M->type == A->matrix_type && M->splited == A->is_splited && M->n == A->n) { *out = M; return LIS_SUCCESS; }
    if (M) lisd_matrix_drop(A);
    M = (lisd_matrix *)calloc(1, sizeof(lisd_matrix));
    if (M == NULL) { LIS_SETERR_MEM(sizeof(lisd_matrix)); return LIS_OUT_OF_MEMORY; }
    const int n = A->n;
    M->type = A->matrix_type; M->n = n; M->np = A->np; M->splited = A->is_splited;
    err = LIS_SUCCESS;
    if (A->is_splited) {
        if (A->matrix_type != LIS_MATRIX_CSR) { free(M); LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED; }
        err = csr_upload(&M->L, n, A->L->ptr, A->L->index, A->L->value);
        if (!err) err = csr_upload(&M->U, n, A->U->ptr, A->U->index, A->U->value);
        if (!err) err = up((void **)&M->diag, A->D->value, (size_t)n * sizeof(double), 16);
        if (!err && A->WD) err = up((void **)&M->wd, A->WD->value, (size_t)n * sizeof(double), 16);
    } else {
        switch (A->matrix_type) {
        case LIS_MATRIX_CSR:
            err = csr_upload(&M->csr, n, A->ptr, A->index, A->value);
            break;
        case LIS_MATRIX_CSC: {
            /* y[idx[j]] += val[j]*x[i] over columns i ascending == row sums in ascending column
             * order (src/matvec/lis_matvec_csc.c:128-144): the row-major transpose of the CSC
             * arrays, run through the CSR kernel, adds the same products in the same order */
            LIS_INT *tp, *ti;
            LIS_SCALAR *tv;
            err = lis_host_transpose(A->np, n, A->ptr, A->index, A->value, &tp, &ti, &tv);     /* np columns (halo included) -> n rows */
            if (!err) { err = csr_upload(&M->csr, n, tp, ti, tv); lis_free2(3, tp, ti, tv); }
            break;
        }
        case LIS_MATRIX_ELL: {
            const size_t cnt = (size_t)n * (size_t)A->maxnzr;
            M->maxnzr = A->maxnzr; M->ld = n;
            err = up((void **)&M->idx, A->index, cnt * sizeof(int), 16);
            if (!err) err = up((void **)&M->val, A->value, cnt * sizeof(double), 16);
            break;
        }
        case LIS_MATRIX_DIA: {
            const size_t cnt = (size_t)n * (size_t)A->nnd;
            M->nnd = A->nnd; M->ld = n;
            err = up((void **)&M->off, A->index, (size_t)A->nnd * sizeof(int), 16);
            if (!err) err = up((void **)&M->val, A->value, cnt * sizeof(double), 16);
            break;
        }
        case LIS_MATRIX_JAD: {
            const size_t nnz = (size_t)A->ptr[A->maxnzr];
            M->maxnzr = A->maxnzr;
            err = up((void **)&M->jptr, A->ptr, ((size_t)A->maxnzr + 1) * sizeof(int), 16);
            if (!err) err = up((void **)&M->perm, A->row, (size_t)n * sizeof(int), 16);
            if (!err) err = up((void **)&M->idx, A->index, nnz * sizeof(int), 16);
            if (!err) err = up((void **)&M->val, A->value, nnz * sizeof(double), 16);
            break;
        }
        case LIS_MATRIX_BSR: {
            const size_t cnt = (size_t)A->bnnz * (size_t)A->bnr * (size_t)A->bnc;
            M->nr = A->nr; M->bnr = A->bnr; M->bnc = A->bnc; M->bnnz = A->bnnz;
            err = up((void **)&M->bptr, A->bptr, ((size_t)A->nr + 1) * sizeof(int), 16);
            if (!err) err = up((void **)&M->bidx, A->bindex, (size_t)A->bnnz * sizeof(int), 16);
            if (!err) err = up((void **)&M->val, A->value, cnt * sizeof(double), 16);
            break;
        }
        case LIS_MATRIX_MSR: {
            /* t = value[i]*x[i]; t += off-diagonals (src/matvec/lis_matvec_msr.c:91-102) is the split-order
             * product with D = value[0..n), L = the off-diagonal rows, U = nothing */
            LIS_INT *rp = (LIS_INT *)malloc(((size_t)n + 1) * sizeof(LIS_INT));
            LIS_INT *zp = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
            LIS_INT zi = 0; LIS_SCALAR zv = 0.0;
            if (!rp || !zp) { free(rp); free(zp); err = LIS_OUT_OF_MEMORY; LIS_SETERR_MEM(n); break; }
            for (int i = 0; i <= n; i++) rp[i] = A->index[i] - (n + 1);
            err = csr_upload(&M->L, n, rp, A->index + n + 1, A->value + n + 1);
            if (!err) err = csr_upload(&M->U, n, zp, &zi, &zv);
            if (!err) err = up((void **)&M->diag, A->value, (size_t)n * sizeof(double), 16);
            free(rp); free(zp);
            break;
        }
        case LIS_MATRIX_COO: case LIS_MATRIX_BSC: case LIS_MATRIX_VBR: case LIS_MATRIX_DNS: {
            /* rows rebuilt in the order the reference's serial product adds into y[i] (explicit zeros of
             * the dense blocks included), run through the CSR kernels: host/lis_formats_ext.c */
            LIS_INT nnz, *tp, *ti;
            LIS_SCALAR *tv;
            err = lis_host_ordered_rows(A, 1, &nnz, &tp, &ti, &tv);
            if (!err) { err = csr_upload(&M->csr, n, tp, ti, tv); lis_free2(3, tp, ti, tv); }
            break;
        }
        default:
            LIS_SETERR1(LIS_ERR_NOT_IMPLEMENTED, "storage format %D has no B200 kernel\n", A->matrix_type);
            err = LIS_ERR_NOT_IMPLEMENTED;
        }
    }
    if (err) { mirror_free(M); return err; }
    A->b200_dev = M;
    *out = M;
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ y = A x */
static LIS_INT matvec_launch(LIS_MATRIX A, lisd_matrix *M, const double *x, double *y)
{
    void *st = lisd_stream();
    const int n = A->n;
    int rc;
    if (M->splited || M->type == LIS_MATRIX_MSR)
        rc = lisb200_spmv_csr_split(n, M->diag, M->L.ptr, M->L.idx, M->L.val, M->U.ptr, M->U.idx, M->U.val, x, y, st);
    else switch (M->type) {
    case LIS_MATRIX_CSR:
    case LIS_MATRIX_CSC:
    case LIS_MATRIX_COO: case LIS_MATRIX_BSC: case LIS_MATRIX_VBR: case LIS_MATRIX_DNS:      /* ordered-row mirrors */
        if (M->csr.tma_rows) rc = lisb200_spmv_csr_tma(n, M->csr.tma_rows, M->csr.tma_tile, M->csr.tma_stages, M->csr.ptr, M->csr.idx, M->csr.val, x, y, st);
        else rc = lisb200_spmv_csr(n, M->csr.ptr, M->csr.idx, M->csr.val, x, y, st);
        break;
    case LIS_MATRIX_ELL: rc = lisb200_spmv_ell(n, M->maxnzr, M->ld, M->idx, M->val, x, y, st); break;
    case LIS_MATRIX_DIA: rc = lisb200_spmv_dia(n, M->np, M->nnd, M->ld, M->off, M->val, x, y, st); break;
    case LIS_MATRIX_JAD: rc = lisb200_spmv_jad(n, M->maxnzr, M->jptr, M->perm, M->idx, M->val, x, y, st); break;
    case LIS_MATRIX_BSR: rc = lisb200_spmv_bsr_cols(n, M->np > n ? M->np : n, M->nr, M->bnr, M->bnc, M->bptr, M->bidx, M->val, x, y, st); break;
    default: LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED;
    }
    lisd_mark_busy();
    return lisd_check(rc, "lis_matvec");
}

/* rows [r0, r1) of an unsplit CSR mirror on stream st (r0 a multiple of 1024: row-pointer slices stay
 * aligned and the row blocks are those of the whole-matrix TMA plan) */
static LIS_INT csr_rows_launch(lisd_matrix *M, int r0, int r1, const double *x, double *y, void *st, const char *what)
{
    int rc;
    if (r1 <= r0) return LIS_SUCCESS;
    if (M->csr.tma_rows) rc = lisb200_spmv_csr_tma(r1 - r0, M->csr.tma_rows, M->csr.tma_tile, M->csr.tma_stages, M->csr.ptr + r0, M->csr.idx, M->csr.val, x, y + r0, st);
    else rc = lisb200_spmv_csr(r1 - r0, M->csr.ptr + r0, M->csr.idx, M->csr.val, x, y + r0, st);
    lisd_mark_busy();
    return lisd_check(rc, what);
}

/* Row-partitioned CSR: the longest run of 1024-row blocks none of whose rows reads a halo entry
 * (column >= n).  For a slab of a stencil grid that is everything but the first and last planes.
 * Those rows do not need the exchange: they run on a second stream while it is in flight (the
 * reference's LIS_MATVEC_SENDRECV, include/lis_matvec.h:31-44, finishes the exchange first). */
static int g_overlap = -1;             /* -1: take LIS_B200_OVERLAP from the environment on first use; 0 off; 1 products and the
                                        * fused CG step; 2 products only */
static int overlap_enabled(void)
{
    if (g_overlap < 0) {
        const char *e = getenv("LIS_B200_OVERLAP");
        g_overlap = (e && e[0] == '0') ? 0 : (e && strcmp(e, "spmv") == 0) ? 2 : 1;
    }
    return g_overlap != 0;
}
static int overlap_dot_enabled(void) { return overlap_enabled() && g_overlap == 1; }
LIS_INT lis_b200_set_overlap(LIS_INT on) { overlap_enabled(); const int old = g_overlap; g_overlap = on == 2 ? 2 : on ? 1 : 0; return old; }

static void overlap_plan(LIS_MATRIX A, lisd_matrix *M)
{
    const int n = A->n, blk = 1024;
    const char *e = getenv("LIS_B200_OVERLAP");
    int best_lo = 0, best_hi = 0, run_lo = -1;
    M->ov_built = -1;
    if (n < 4 * blk) return;
    const int nb = n / blk;                              /* whole blocks only; the tail belongs to the boundary part */
    for (int b = 0; b <= nb; b++) {
        int touches = 1;
        if (b < nb) {
            touches = 0;
            for (LIS_INT j = A->ptr[b * blk]; j < A->ptr[(b + 1) * blk]; j++) if (A->index[j] >= n) { touches = 1; break; }
        }
        if (!touches) { if (run_lo < 0) run_lo = b; }
        else if (run_lo >= 0) {
            if (b - run_lo > best_hi - best_lo) { best_lo = run_lo; best_hi = b; }
            run_lo = -1;
        }
    }
    if (best_hi <= best_lo) return;
    if ((long long)(best_hi - best_lo) * blk * 2 < n && !(e && strcmp(e, "force") == 0)) return;   /* under half the rows: not worth a second launch */
    M->ov_lo = best_lo * blk; M->ov_hi = best_hi * blk;
    M->ov_built = 1;
}

/* grow x to np entries so the halo has somewhere to land (LIS_MATVEC_SENDRECV) */
static LIS_INT vec_reserve(LIS_VECTOR x, size_t count)
{
    if (x->b200_capacity >= count) return LIS_SUCCESS;
    LIS_SCALAR *nv;
    LIS_INT managed, err;
    err = lisd_alloc_vector(count, &nv, &managed);
    if (err) return err;
    if (!managed) { free(nv); LIS_SETERR(LIS_ERR_DEVICE, "no device\n"); return LIS_ERR_DEVICE; }
    err = lisd_vec_device(x);
    if (err) return err;
    err = lisd_memset(nv, 0, count * sizeof(LIS_SCALAR));
    if (!err) err = lisd_check(lisb200_copy(x->n, x->value, nv, lisd_stream()), "vector grow");
    if (err) return err;
    lisd_sync();
    lisd_free_vector(x->value, x->b200_managed);
    x->value = nv; x->b200_managed = 1; x->b200_capacity = count; x->b200_resident = 1;
    return LIS_SUCCESS;
}

LIS_INT lisd_matvec(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y)
{
    LIS_INT err = lisd_require("lis_matvec");
    if (err) return err;
    if (x == y || x->value == y->value) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "lis_matvec: x and y must not alias\n");
        return LIS_ERR_ILL_ARG;
    }
    lisd_matrix *M;
    err = lisd_matrix_get(A, &M);
    if (err) return err;
    if (A->np > A->n) {
        err = vec_reserve(x, (size_t)A->np + (size_t)A->pad_comm);
        if (err) return err;
    }
    err = lisd_vec_device(x);
    if (!err) err = lisd_vec_device(y);
    if (err) return err;
    if (A->nprocs > 1 && A->commtable) {
        if (M->type == LIS_MATRIX_CSR && !M->splited) {
            /* halo exchange inside the kernel over peer memory, where every rank can map its neighbours and every
             * rank's slab takes the TMA row-block kernel (agreed on collectively at the first product) */
            unsigned long long epoch = 0;
            const struct lisb200_p2p *tb = lisd_p2p_begin(A, M->csr.tma_rows != 0, &epoch);
            if (tb) {
                if (M->ov_built == 0) overlap_plan(A, M);
                const int lo = M->ov_built == 1 ? M->ov_lo : 0, hi = M->ov_built == 1 ? M->ov_hi : 0;
                lisd_mark_busy();
                return lisd_check(lisb200_spmv_csr_tma_p2p(A->n, M->csr.tma_rows, M->csr.tma_tile, M->csr.tma_stages, M->csr.ptr, M->csr.idx,
                                                           M->csr.val, x->value, y->value, 0, NULL, NULL, NULL, tb, epoch, lo, hi,
                                                           lisd_stream()), "lis_matvec (halo exchange in the kernel)");
            }
        }
        if (M->type == LIS_MATRIX_CSR && !M->splited && M->ov_built == 0 && overlap_enabled()) overlap_plan(A, M);
        if (M->type == LIS_MATRIX_CSR && !M->splited && M->ov_built == 1 && overlap_enabled()) {
            /* exchange on the main stream (enqueued first, so its few CTAs are resident first), interior
             * rows on the second stream meanwhile, then the rows that read halo entries */
            void *aux;
            err = lisd_aux_fork(&aux);
            if (!err) err = lisd_halo_exchange(A, x);
            if (!err) err = csr_rows_launch(M, M->ov_lo, M->ov_hi, x->value, y->value, aux, "lis_matvec (interior rows)");
            if (!err) err = csr_rows_launch(M, 0, M->ov_lo, x->value, y->value, lisd_stream(), "lis_matvec (boundary rows)");
            if (!err) err = csr_rows_launch(M, M->ov_hi, A->n, x->value, y->value, lisd_stream(), "lis_matvec (boundary rows)");
            { LIS_INT e2 = lisd_aux_join(); if (!err) err = e2; }
            return err;
        }
        err = lisd_halo_exchange(A, x);
        if (err) return err;
    }
    return matvec_launch(A, M, x->value, y->value);
}

/* the fused SpMV+dot on rows [r0, r1): its share of <x,y> goes to mapped scalar `slot`; `lane` picks
 * the partial-sum scratch and ticket counter, one per stream that may be running such a launch */
static LIS_INT csr_rows_dot_launch(lisd_matrix *M, int r0, int r1, const double *x, double *y, double *partial, int lane, int slot, void *st,
                                   const char *what)
{
    const int rc = lisb200_spmv_csr_tma_dot_rows(r1 - r0, M->csr.tma_rows, M->csr.tma_tile, M->csr.tma_stages, M->csr.ptr + r0, M->csr.idx,
                                                 M->csr.val, x, y + r0, x + r0, partial + (size_t)lane * (size_t)lisb200_reduce_slots(),
                                                 lisd_counter() + lane, lisd_scalar_dev(slot), st);
    lisd_mark_busy();
    return lisd_check(rc, what);
}

/* row-partitioned CG step q = A p, <p,q>: interior rows (fused with their share of the dot) on the
 * second stream while the halo exchange is in flight, the rows that read halo entries behind it.
 * The dot is the sum of the (up to three) range shares in row order: a fixed order, the same on every
 * run, but not the single tree of the one-launch path -- like any change of the rank count it moves
 * the last bits of <p,q>, not the result class (tests/test_multi_rank.py checks CG against the oracle). */
static LIS_INT matvec_dot_overlapped(LIS_MATRIX A, lisd_matrix *M, LIS_VECTOR x, LIS_VECTOR y, LIS_SCALAR *dot_xy)
{
    double vals[3] = {0.0, 0.0, 0.0};
    void *aux;
    double *partial = lisd_partial(2 * (size_t)lisb200_reduce_slots());
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    LIS_INT err = lisd_aux_fork(&aux);
    if (!err) err = lisd_halo_exchange(A, x);
    /* every rank fills all three slots (an empty range stores 0): the cross-rank sum is slot by slot */
    if (!err) err = csr_rows_dot_launch(M, 0, M->ov_lo, x->value, y->value, partial, 0, 0, lisd_stream(), "lis_matvec+dot (boundary rows)");
    if (!err) err = csr_rows_dot_launch(M, M->ov_lo, M->ov_hi, x->value, y->value, partial, 1, 1, aux, "lis_matvec+dot (interior rows)");
    if (!err) err = csr_rows_dot_launch(M, M->ov_hi, A->n, x->value, y->value, partial, 0, 2, lisd_stream(), "lis_matvec+dot (boundary rows)");
    { LIS_INT e2 = lisd_aux_join(); if (!err) err = e2; }
    if (err) return err;
    err = lisd_reduce_finish(vals, 3, 0);
    if (err) return err;
    *dot_xy = (vals[0] + vals[1]) + vals[2];
    return LIS_SUCCESS;
}

/* y = A x and <x,y> in one pass where a fused kernel exists (unsplit CSR / CSC mirror) */
LIS_INT lisd_matvec_dot(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y, LIS_SCALAR *dot_xy)
{
    lisd_matrix *M;
    LIS_INT err = lisd_require("lis_matvec");
    if (err) return err;
    err = lisd_matrix_get(A, &M);
    if (err) return err;
    const int fusable = !M->splited && (M->type == LIS_MATRIX_CSR || M->type == LIS_MATRIX_CSC || M->type == LIS_MATRIX_COO ||
                                        M->type == LIS_MATRIX_BSC || M->type == LIS_MATRIX_VBR || M->type == LIS_MATRIX_DNS);
    if (!fusable) {
        err = lisd_matvec(A, x, y);
        if (err) return err;
        return lis_vector_dot(x, y, dot_xy);
    }
    if (x == y || x->value == y->value) { LIS_SETERR(LIS_ERR_ILL_ARG, "lis_matvec: x and y must not alias\n"); return LIS_ERR_ILL_ARG; }
    if (A->np > A->n) { err = vec_reserve(x, (size_t)A->np + (size_t)A->pad_comm); if (err) return err; }
    err = lisd_vec_device(x);
    if (!err) err = lisd_vec_device(y);
    if (err) return err;
    if (A->nprocs > 1 && A->commtable) {
        if (M->type == LIS_MATRIX_CSR) {
            unsigned long long epoch = 0;
            const struct lisb200_p2p *tb = lisd_p2p_begin(A, M->csr.tma_rows != 0, &epoch);
            if (tb) {
                if (M->ov_built == 0) overlap_plan(A, M);
                const int lo = M->ov_built == 1 ? M->ov_lo : 0, hi = M->ov_built == 1 ? M->ov_hi : 0;
                double *partial = lisd_partial(0);
                if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
                lisd_mark_busy();
                err = lisd_check(lisb200_spmv_csr_tma_p2p(A->n, M->csr.tma_rows, M->csr.tma_tile, M->csr.tma_stages, M->csr.ptr, M->csr.idx,
                                                          M->csr.val, x->value, y->value, 1, partial, lisd_counter(), lisd_scalar_dev(0), tb,
                                                          epoch, lo, hi, lisd_stream()), "lis_matvec+dot (halo exchange in the kernel)");
                if (err) return err;
                return lisd_reduce_finish(dot_xy, 1, 0);
            }
        }
        if (M->type == LIS_MATRIX_CSR && M->csr.tma_rows && M->ov_built == 0 && overlap_dot_enabled()) overlap_plan(A, M);
        if (M->type == LIS_MATRIX_CSR && M->csr.tma_rows && M->ov_built == 1 && overlap_dot_enabled())
            return matvec_dot_overlapped(A, M, x, y, dot_xy);
        err = lisd_halo_exchange(A, x);
        if (err) return err;
    }
    double *partial = lisd_partial(M->csr.tma_rows ? 0 : (size_t)lisb200_spmv_csr_dot_slots(A->n));
    if (partial == NULL) { LIS_SETERR_MEM(0); return LIS_ERR_OUT_OF_MEMORY; }
    lisd_mark_busy();
    if (M->csr.tma_rows)
        err = lisd_check(lisb200_spmv_csr_tma_dot(A->n, M->csr.tma_rows, M->csr.tma_tile, M->csr.tma_stages, M->csr.ptr, M->csr.idx, M->csr.val,
                                                  x->value, y->value, partial, lisd_counter(), lisd_scalar_dev(0),
                                                  lisd_stream()), "lis_matvec+dot");
    else
        err = lisd_check(lisb200_spmv_csr_dot(A->n, M->csr.ptr, M->csr.idx, M->csr.val, x->value, y->value, partial,
                                              lisd_counter(), lisd_scalar_dev(0), lisd_stream()), "lis_matvec+dot");
    if (err) return err;
    return lisd_reduce_finish(dot_xy, 1, 0);
}

LIS_INT lis_matvec(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y)
{
    LIS_INT err = lis_host_matrix_check_input(A);
    if (err) return err;
    if (A->n != x->n || A->n != y->n) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "lis_matvec: sizes of A, x and y do not match\n");
        return LIS_ERR_ILL_ARG;
    }
    err = lisd_matvec(A, x, y);
    if (err) return err;
    return lisd_sync();
}

/* lis_matvec without the closing stream synchronisation: the product is enqueued on the library stream
 * (lis_b200_stream) and the call returns; lis_b200_sync() -- or any host-synchronous lis.h call -- waits for
 * it.  For callers that keep the queue full (bench.py's device-timed leg). */
LIS_INT lis_b200_matvec_async(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y)
{
    LIS_INT err = lis_host_matrix_check_input(A);
    if (err) return err;
    if (A->n != x->n || A->n != y->n) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "lis_matvec: sizes of A, x and y do not match\n");
        return LIS_ERR_ILL_ARG;
    }
    return lisd_matvec(A, x, y);
}
LIS_INT lis_b200_sync(void) { return lisd_sync(); }

/* ------------------------------------------------------------------ y = A x with HOST x and y
 * What an application that keeps its vectors in host arrays pays per product is two PCIe
 * transfers around a 2 ms kernel; done one after the other (lis_vector_scatter, lis_matvec,
 * lis_vector_gather) the link is idle in one direction at any time.  Here the rows are cut into
 * chunks, and chunk c's product starts as soon as the x entries its rows read have landed, its y
 * slice leaving while later x chunks are still arriving: both directions of the link and the SMs
 * overlap.  For a banded matrix (stencils: the band is one grid plane) a chunk needs x up to the
 * next chunk only; a matrix whose rows read all of x degenerates to copy-in, then products
 * overlapped with copy-out.  Every row is still summed by the same kernel in the same order:
 * x, y, host_y hold the same bits as after the three separate calls.
 * Row-partitioned matrices: host_x / host_y are this rank's slices; the chunks that read halo
 * entries run last, behind the halo exchange, which itself needs all of the local x.
 * Unsplit CSR; anything else takes the three calls. */
#define LISD_PIPE_MIN_ROWS  (1 << 18)       /* 2 MiB of x per chunk at least */
#define LISD_PIPE_MAX_CHUNKS 32

static LIS_INT pipe_plan(LIS_MATRIX A, lisd_matrix *M)
{
    const int n = A->n;
    int nch = n / LISD_PIPE_MIN_ROWS;
    const char *force = getenv("LIS_B200_PIPE_CHUNKS");     /* tests: chunking at small n */
    if (nch > LISD_PIPE_MAX_CHUNKS) nch = LISD_PIPE_MAX_CHUNKS;
    if (force && atoi(force) > 0) nch = atoi(force);
    if (nch > n / 4) nch = n / 4;
    if (nch < 2) { M->pipe_n = -1; return LIS_SUCCESS; }       /* too small to be worth cutting */
    /* chunk boundaries on multiples of 1024 rows: row-pointer slices stay 16-byte aligned for the
     * bulk copies and a chunk's row blocks coincide with the whole-matrix ones the TMA plan was
     * made for (forced small chunks in tests: multiples of the row-block size) */
    const int align = (force && atoi(force) > 0) ? (M->csr.tma_rows ? M->csr.tma_rows : 4) : 1024;
    int rows = ((n + nch - 1) / nch + align - 1) / align * align;
    nch = (n + rows - 1) / rows;
    if (nch < 2) { M->pipe_n = -1; return LIS_SUCCESS; }
    M->pipe_row = (int *)malloc(sizeof(int) * (size_t)(nch + 1));
    M->pipe_need = (int *)malloc(sizeof(int) * (size_t)nch);
    if (!M->pipe_row || !M->pipe_need) { LIS_SETERR_MEM(nch * sizeof(int)); return LIS_OUT_OF_MEMORY; }
    for (int c = 0; c <= nch; c++) M->pipe_row[c] = (long long)c * rows < n ? c * rows : n;
    for (int c = 0; c < nch; c++) {
        LIS_INT mx = 0;
        const LIS_INT *idx = A->index;
        for (LIS_INT j = A->ptr[M->pipe_row[c]]; j < A->ptr[M->pipe_row[c + 1]]; j++) if (idx[j] > mx) mx = idx[j];
        /* a column beyond the owned range is a halo entry (row-partitioned matrix): such a chunk
         * runs after the halo exchange, i.e. after all of x has landed */
        M->pipe_need[c] = mx >= n ? nch : (mx / rows < nch ? mx / rows : nch - 1);
    }
    M->pipe_n = nch;
    return LIS_SUCCESS;
}

LIS_INT lis_b200_matvec_host(LIS_MATRIX A, LIS_SCALAR host_x[], LIS_VECTOR x, LIS_VECTOR y, LIS_SCALAR host_y[])
{
    LIS_INT err = lis_host_matrix_check_input(A);
    if (err) return err;
    if (A->n != x->n || A->n != y->n) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "lis_b200_matvec_host: sizes of A, x and y do not match\n");
        return LIS_ERR_ILL_ARG;
    }
    if (x == y || x->value == y->value) { LIS_SETERR(LIS_ERR_ILL_ARG, "lis_b200_matvec_host: x and y must not alias\n"); return LIS_ERR_ILL_ARG; }
    err = lisd_require("lis_b200_matvec_host");
    if (err) return err;
    lisd_matrix *M;
    err = lisd_matrix_get(A, &M);
    if (err) return err;
    int pipelined = A->matrix_type == LIS_MATRIX_CSR && !M->splited && x->b200_managed && y->b200_managed;
    if (pipelined && M->pipe_n == 0) { err = pipe_plan(A, M); if (err) return err; }
    if (!pipelined || M->pipe_n < 2) {
        err = lis_vector_set_values2(LIS_INS_VALUE, x->is + x->origin, x->n, host_x, x);
        if (!err) err = lis_matvec(A, x, y);
        if (!err) err = lis_vector_get_values(y, y->is + y->origin, y->n, host_y);
        return err;
    }
    /* the copies run between the host arrays and staging vectors in plain device memory (asynchronous
     * by specification; vector storage is managed memory, for which the runtime promises less); x and y
     * themselves are filled from the staging vectors on the device at the end (2 x 16 B per row of HBM
     * traffic, hidden under the PCIe time) */
    double *xs, *ys;
    err = lisd_vec_device(x);
    if (!err) err = lisd_vec_device(y);
    if (!err) err = lisd_pipe_staging((size_t)A->np + (size_t)A->pad_comm, (size_t)A->n, &xs, &ys);
    if (!err) err = lisd_pipe_begin(M->pipe_n);
    if (err) return err;
    const int *row = M->pipe_row;
    const int nchunks = M->pipe_n;
    for (int c = 0; c < M->pipe_n && !err; c++)
        err = lisd_pipe_h2d(c, xs + row[c], host_x + row[c], (size_t)(row[c + 1] - row[c]) * sizeof(LIS_SCALAR));
    int landed = -1;                                  /* inputs the main stream already waits for */
    void *st = lisd_stream();
    /* pass 0: chunks whose rows read owned entries only, as their inputs land; pass 1 (row-partitioned
     * matrices): the halo exchange once all of x is there, then the chunks that read halo entries */
    for (int pass = 0; pass < 2 && !err; pass++) {
        if (pass == 1) {
            if (!(A->nprocs > 1 && A->commtable)) break;
            if (landed < nchunks - 1) { landed = nchunks - 1; err = lisd_pipe_wait_in(landed); if (err) break; }
            err = lisd_halo_exchange_raw(A, xs);
            if (err) break;
        }
        for (int c = 0; c < nchunks && !err; c++) {
            const int nr = row[c + 1] - row[c];
            int rc;
            if ((M->pipe_need[c] >= nchunks) != (pass == 1)) continue;
            if (pass == 0 && M->pipe_need[c] > landed) { landed = M->pipe_need[c]; err = lisd_pipe_wait_in(landed); if (err) break; }
            if (M->csr.tma_rows) rc = lisb200_spmv_csr_tma(nr, M->csr.tma_rows, M->csr.tma_tile, M->csr.tma_stages, M->csr.ptr + row[c], M->csr.idx, M->csr.val, xs, ys + row[c], st);
            else rc = lisb200_spmv_csr(nr, M->csr.ptr + row[c], M->csr.idx, M->csr.val, xs, ys + row[c], st);
            lisd_mark_busy();
            err = lisd_check(rc, "lis_b200_matvec_host");
            if (!err) err = lisd_pipe_d2h(c, host_y + row[c], ys + row[c], (size_t)nr * sizeof(LIS_SCALAR));
        }
    }
    /* x chunks nobody waited for (beyond every row's reach) still have to land before x is used again */
    if (!err && landed < nchunks - 1) err = lisd_pipe_wait_in(nchunks - 1);
    if (!err) err = lisd_d2d(x->value, xs, (size_t)A->n * sizeof(LIS_SCALAR));
    if (!err) err = lisd_d2d(y->value, ys, (size_t)A->n * sizeof(LIS_SCALAR));
    {
        LIS_INT e2 = lisd_pipe_end();
        if (!err) err = e2;
    }
    {
        LIS_INT e2 = lisd_sync();
        if (!err) err = e2;
    }
    return err;
}

/* the chunking lis_b200_matvec_host uses for A (tests): returns the chunk count (0: not chunked),
 * fills up to cap+1 row bounds and cap `need` entries */
LIS_INT lis_b200_matvec_host_plan(LIS_MATRIX A, LIS_INT cap, LIS_INT *rows, LIS_INT *need)
{
    lisd_matrix *M = (lisd_matrix *)A->b200_dev;
    if (M == NULL || M->pipe_n < 2) return 0;
    for (int c = 0; c < M->pipe_n && c < cap; c++) { rows[c] = M->pipe_row[c]; rows[c + 1] = M->pipe_row[c + 1]; need[c] = M->pipe_need[c]; }
    return M->pipe_n;
}

/* ------------------------------------------------------------------ y = A^H x
 * The reference's serial lis_matvech_csr (src/matvec/lis_matvec_csr.c:113-260) is a scatter,
 * y[idx[j]] += val[j]*x[i] over rows i ascending: for every output entry the products arrive in
 * ascending row order.  A counting transpose keeps exactly that order inside each row of A^T,
 * so A^T in CSR through the ordinary (gather) CSR kernels adds the same products in the same
 * order.  Split matrices: y = D x, then all of L's contributions, then all of U's (:170-199) ==
 * the split kernel on (D, L^T, U^T).  CSC storage already is the CSR of A^T. */
/* n rows, ncols columns (ncols = np on a row-partitioned matrix: the halo columns become rows n..np of the
 * transpose, whose sums go back to their owners afterwards) */
static LIS_INT transposed_upload_cols(lisd_csr *dst, LIS_INT n, LIS_INT ncols, const LIS_INT *ptr, const LIS_INT *idx, const LIS_SCALAR *val)
{
    LIS_INT *tp, *ti;
    LIS_SCALAR *tv;
    LIS_INT err = lis_host_transpose(n, ncols, ptr, idx, val, &tp, &ti, &tv);
    if (err) return err;
    err = csr_upload(dst, (int)ncols, tp, ti, tv);
    lis_free2(3, tp, ti, tv);
    return err;
}

static LIS_INT transposed_upload(lisd_csr *dst, LIS_INT n, const LIS_INT *ptr, const LIS_INT *idx, const LIS_SCALAR *val)
{
    LIS_INT *tp, *ti;
    LIS_SCALAR *tv;
    LIS_INT err = lis_host_transpose(n, n, ptr, idx, val, &tp, &ti, &tv);
    if (err) return err;
    err = csr_upload(dst, (int)n, tp, ti, tv);
    lis_free2(3, tp, ti, tv);
    return err;
}

LIS_INT lisd_matvech(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y)
{
    LIS_INT err = lisd_require("lis_matvech");
    if (err) return err;
    if (x == y || x->value == y->value) { LIS_SETERR(LIS_ERR_ILL_ARG, "lis_matvech: x and y must not alias\n"); return LIS_ERR_ILL_ARG; }
    if (A->nprocs > 1 && A->commtable) {
        /* row-partitioned: y[0..np) = A_loc^T x with the halo columns as extra rows, then lis_reduce sends those
         * sums to their owners (src/matvec/lis_matvec.c:199-205 + lis_matrix_mpi.c:958).  CSR, unsplit. */
        if (A->matrix_type != LIS_MATRIX_CSR || A->is_splited) {
            LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "lis_matvech on a row-partitioned matrix: unsplit CSR only\n");
            return LIS_ERR_NOT_IMPLEMENTED;
        }
        lisd_matrix *Mp;
        err = lisd_matrix_get(A, &Mp);
        if (err) return err;
        if (!Mp->has_t) {
            err = transposed_upload_cols(&Mp->csrT, A->n, A->np, A->ptr, A->index, A->value);
            if (err) return err;
            Mp->has_t = 1;
        }
        err = vec_reserve(y, (size_t)A->np + (size_t)A->pad_comm);
        if (!err) err = lisd_vec_device(x);
        if (!err) err = lisd_vec_device(y);
        if (err) return err;
        int rc;
        if (Mp->csrT.tma_rows)
            rc = lisb200_spmv_csr_tma(A->np, Mp->csrT.tma_rows, Mp->csrT.tma_tile, Mp->csrT.tma_stages, Mp->csrT.ptr, Mp->csrT.idx, Mp->csrT.val, x->value, y->value, lisd_stream());
        else
            rc = lisb200_spmv_csr(A->np, Mp->csrT.ptr, Mp->csrT.idx, Mp->csrT.val, x->value, y->value, lisd_stream());
        lisd_mark_busy();
        err = lisd_check(rc, "lis_matvech");
        if (err) return err;
        return lisd_halo_reduce_raw(A, y->value);
    }
    if (A->is_splited && A->matrix_type != LIS_MATRIX_CSR) { LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED; }
    lisd_matrix *M;
    err = lisd_matrix_get(A, &M);
    if (err) return err;
    const int n = A->n;
    if (!M->has_t) {
        if (M->splited) {
            err = transposed_upload(&M->LT, n, A->L->ptr, A->L->index, A->L->value);
            if (!err) err = transposed_upload(&M->UT, n, A->U->ptr, A->U->index, A->U->value);
        } else if (A->matrix_type == LIS_MATRIX_CSR) {
            err = transposed_upload(&M->csrT, n, A->ptr, A->index, A->value);
        } else if (A->matrix_type == LIS_MATRIX_CSC) {
            err = csr_upload(&M->csrT, n, A->ptr, A->index, A->value);       /* CSC arrays == CSR of A^T */
        } else {
            /* every other format: A^T with each row in the order the reference's serial lis_matvech_<fmt>
             * scatters into it (host/lis_formats_ext.c); MSR keeps its diagonal apart */
            LIS_INT *tp, *ti;
            LIS_SCALAR *tv;
            err = lis_host_transposed_rows(A, &tp, &ti, &tv);
            if (!err) {
                err = csr_upload(A->matrix_type == LIS_MATRIX_MSR ? &M->LT : &M->csrT, n, tp, ti, tv);
                lis_free2(3, tp, ti, tv);
            }
        }
        if (err) return err;
        M->has_t = 1;
    }
    const int msr = !M->splited && M->type == LIS_MATRIX_MSR;
    err = lisd_vec_device(x);
    if (!err) err = lisd_vec_device(y);
    if (err) return err;
    void *st = lisd_stream();
    int rc;
    if (M->splited)
        rc = lisb200_spmv_csr_split(n, M->diag, M->LT.ptr, M->LT.idx, M->LT.val, M->UT.ptr, M->UT.idx, M->UT.val, x->value, y->value, st);
    else if (msr)                   /* y = d*x first, then the scattered off-diagonals; U (the forward mirror's empty part) adds nothing */
        rc = lisb200_spmv_csr_split(n, M->diag, M->LT.ptr, M->LT.idx, M->LT.val, M->U.ptr, M->U.idx, M->U.val, x->value, y->value, st);
    else if (M->csrT.tma_rows)
        rc = lisb200_spmv_csr_tma(n, M->csrT.tma_rows, M->csrT.tma_tile, M->csrT.tma_stages, M->csrT.ptr, M->csrT.idx, M->csrT.val, x->value, y->value, st);
    else
        rc = lisb200_spmv_csr(n, M->csrT.ptr, M->csrT.idx, M->csrT.val, x->value, y->value, st);
    lisd_mark_busy();
    return lisd_check(rc, "lis_matvech");
}

LIS_INT lis_matvech(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y)
{
    LIS_INT err = lis_host_matrix_check_input(A);
    if (err) return err;
    if (A->n != x->n || A->n != y->n) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "lis_matvech: sizes of A, x and y do not match\n");
        return LIS_ERR_ILL_ARG;
    }
    err = lisd_matvech(A, x, y);
    if (err) return err;
    return lisd_sync();
}

/* ---- per-format seam with raw pointers (include/lis_matvec.h:76-205 of the reference).
 * x and y must be device-accessible: vector storage handed out by this library is. */
static void raw_matvec(LIS_MATRIX A, LIS_INT type, LIS_SCALAR x[], LIS_SCALAR y[])
{
    lisd_matrix *M;
    if (A->matrix_type != type) { LIS_SETERR(LIS_ERR_ILL_ARG, "matrix storage format does not match the kernel\n"); return; }
    if (lisd_matrix_get(A, &M)) return;
    if (matvec_launch(A, M, x, y)) return;
    lisd_sync();
}
void lis_matvec_csr(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_CSR, x, y); }
void lis_matvec_csc(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_CSC, x, y); }
void lis_matvec_ell(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_ELL, x, y); }
void lis_matvec_dia(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_DIA, x, y); }
void lis_matvec_jad(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_JAD, x, y); }
void lis_matvec_bsr(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_BSR, x, y); }
void lis_matvec_msr(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_MSR, x, y); }
void lis_matvec_coo(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_COO, x, y); }
void lis_matvec_bsc(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_BSC, x, y); }
void lis_matvec_vbr(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_VBR, x, y); }
void lis_matvec_dns(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]) { raw_matvec(A, LIS_MATRIX_DNS, x, y); }
