/*
 * lis_formats_ext.c -- the five storage formats outside the named hot path: MSR, COO, BSC, VBR, DNS.
 *
 * The reference's spmvtest drivers walk formats 1..10 when no format is given (test/spmvtest1.c:188-204),
 * so a drop-in has to answer for these too.  What is here:
 *   - lis_matrix_malloc_<fmt> / lis_matrix_set_<fmt>      (src/matrix/lis_matrix_<fmt>.c, adopt the caller's arrays)
 *   - CSR -> <fmt> with the array layouts the reference's serial build produces, and <fmt> -> CSR
 *   - lis_host_ordered_rows(): the matrix as CSR arrays whose entries sit, row by row, in the order in
 *     which the reference's serial lis_matvec_<fmt> adds their products into y[i] (explicit zeros of
 *     dense blocks included).  That is the device mirror (host/lis_matrix_dev.c): the CSR kernels then
 *     add the same products in the same order, so y carries the reference's bits --
 *       COO  y=0; y[row[k]] += v[k]*x[col[k]], k ascending        src/matvec/lis_matvec_coo.c:78-87
 *       BSC  y=0; block columns ascending, blocks in storage order, j then i   lis_matvec_bsc.c:121-146
 *       VBR  y=0; block rows, blocks in storage order, j then i    lis_matvec_vbr.c:113-135
 *       DNS  y=0; y[i] += value[j*n+i]*x[j], j ascending           lis_matvec_dns.c:66-78
 *       MSR  t = value[i]*x[i]; t += off-diagonals                 lis_matvec_msr.c:91-102
 *     (MSR runs on the split-order kernel: diagonal first, then one off-diagonal CSR).
 * These formats execute at CSR speed on their mirror; they are not tuned beyond that and are
 * single-process only.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lis_device.h"
#include "lis_host.h"

static void *tracked(size_t count, size_t size, char *tag) { return lis_malloc((count > 0 ? count : 1) * size, tag); }

static LIS_INT single_process(LIS_MATRIX A, const char *fmt)
{
    if (A->np != A->n || A->nprocs > 1) {
        LIS_SETERR1(LIS_ERR_NOT_IMPLEMENTED, "%s storage is not available for row-partitioned (multi-GPU) matrices\n", fmt);
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ malloc / set */
LIS_INT lis_matrix_malloc_msr(LIS_INT n, LIS_INT nnz, LIS_INT ndz, LIS_INT **index, LIS_SCALAR **value)
{
    (void)n;
    *index = (LIS_INT *)tracked((size_t)nnz + ndz + 1, sizeof(LIS_INT), "lis_matrix_malloc_msr::index");
    *value = (LIS_SCALAR *)tracked((size_t)nnz + ndz + 1, sizeof(LIS_SCALAR), "lis_matrix_malloc_msr::value");
    if (!*index || !*value) { LIS_SETERR_MEM((size_t)nnz * sizeof(LIS_SCALAR)); lis_free2(2, *index, *value); *index = NULL; *value = NULL; return LIS_OUT_OF_MEMORY; }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_msr(LIS_INT nnz, LIS_INT ndz, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = lis_host_matrix_check_set(A);
    if (err) return err;
    A->index = index; A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_MSR;
    A->nnz = nnz; A->ndz = ndz;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_malloc_coo(LIS_INT nnz, LIS_INT **row, LIS_INT **col, LIS_SCALAR **value)
{
    *row = (LIS_INT *)tracked((size_t)nnz, sizeof(LIS_INT), "lis_matrix_malloc_coo::row");
    *col = (LIS_INT *)tracked((size_t)nnz, sizeof(LIS_INT), "lis_matrix_malloc_coo::col");
    *value = (LIS_SCALAR *)tracked((size_t)nnz, sizeof(LIS_SCALAR), "lis_matrix_malloc_coo::value");
    if (!*row || !*col || !*value) { LIS_SETERR_MEM((size_t)nnz * sizeof(LIS_SCALAR)); lis_free2(3, *row, *col, *value); *row = *col = NULL; *value = NULL; return LIS_OUT_OF_MEMORY; }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_coo(LIS_INT nnz, LIS_INT *row, LIS_INT *col, LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = lis_host_matrix_check_set(A);
    if (err) return err;
    A->row = row; A->col = col; A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_COO;
    A->nnz = nnz;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_malloc_bsc(LIS_INT n, LIS_INT bnr, LIS_INT bnc, LIS_INT bnnz, LIS_INT **bptr, LIS_INT **bindex, LIS_SCALAR **value)
{
    const LIS_INT nc = n > 0 ? 1 + (n - 1) / bnc : 0;
    *bptr = (LIS_INT *)tracked((size_t)nc + 1, sizeof(LIS_INT), "lis_matrix_malloc_bsc::bptr");
    *bindex = (LIS_INT *)tracked((size_t)bnnz, sizeof(LIS_INT), "lis_matrix_malloc_bsc::bindex");
    *value = (LIS_SCALAR *)tracked((size_t)bnnz * bnr * bnc, sizeof(LIS_SCALAR), "lis_matrix_malloc_bsc::value");
    if (!*bptr || !*bindex || !*value) { LIS_SETERR_MEM((size_t)bnnz * sizeof(LIS_SCALAR)); lis_free2(3, *bptr, *bindex, *value); *bptr = *bindex = NULL; *value = NULL; return LIS_OUT_OF_MEMORY; }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_bsc(LIS_INT bnr, LIS_INT bnc, LIS_INT bnnz, LIS_INT *bptr, LIS_INT *bindex, LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = lis_host_matrix_check_set(A);
    if (err) return err;
    if (bnr <= 0 || bnc <= 0) { LIS_SETERR2(LIS_ERR_ILL_ARG, "bnr=%D <= 0 or bnc=%D <= 0\n", bnr, bnc); return LIS_ERR_ILL_ARG; }
    A->bptr = bptr; A->bindex = bindex; A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_BSC;
    A->is_block = LIS_TRUE;
    A->bnnz = bnnz;
    A->nr = A->n > 0 ? 1 + (A->n - 1) / bnr : 0;
    A->nc = A->gn > 0 ? 1 + (A->gn - 1) / bnc : 0;
    A->bnr = bnr; A->bnc = bnc;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_malloc_vbr(LIS_INT n, LIS_INT nnz, LIS_INT nr, LIS_INT nc, LIS_INT bnnz, LIS_INT **row, LIS_INT **col, LIS_INT **ptr,
                              LIS_INT **bptr, LIS_INT **bindex, LIS_SCALAR **value)
{
    (void)n;
    *row = (LIS_INT *)tracked((size_t)nr + 1, sizeof(LIS_INT), "lis_matrix_malloc_vbr::row");
    *col = (LIS_INT *)tracked((size_t)nc + 1, sizeof(LIS_INT), "lis_matrix_malloc_vbr::col");
    *ptr = (LIS_INT *)tracked((size_t)bnnz + 1, sizeof(LIS_INT), "lis_matrix_malloc_vbr::ptr");
    *bptr = (LIS_INT *)tracked((size_t)nr + 1, sizeof(LIS_INT), "lis_matrix_malloc_vbr::bptr");
    *bindex = (LIS_INT *)tracked((size_t)bnnz, sizeof(LIS_INT), "lis_matrix_malloc_vbr::bindex");
    *value = (LIS_SCALAR *)tracked((size_t)nnz, sizeof(LIS_SCALAR), "lis_matrix_malloc_vbr::value");
    if (!*row || !*col || !*ptr || !*bptr || !*bindex || !*value) {
        LIS_SETERR_MEM((size_t)nnz * sizeof(LIS_SCALAR));
        lis_free2(6, *row, *col, *ptr, *bptr, *bindex, *value);
        *row = *col = *ptr = *bptr = *bindex = NULL; *value = NULL;
        return LIS_OUT_OF_MEMORY;
    }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_vbr(LIS_INT nnz, LIS_INT nr, LIS_INT nc, LIS_INT bnnz, LIS_INT *row, LIS_INT *col, LIS_INT *ptr, LIS_INT *bptr,
                           LIS_INT *bindex, LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = lis_host_matrix_check_set(A);
    if (err) return err;
    A->row = row; A->col = col; A->ptr = ptr; A->bptr = bptr; A->bindex = bindex; A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_VBR;
    A->is_block = LIS_TRUE;
    A->nnz = nnz; A->bnnz = bnnz; A->nr = nr; A->nc = nc;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_malloc_dns(LIS_INT n, LIS_INT np, LIS_SCALAR **value)
{
    *value = (LIS_SCALAR *)tracked((size_t)n * (size_t)np, sizeof(LIS_SCALAR), "lis_matrix_malloc_dns::value");
    if (!*value) { LIS_SETERR_MEM((size_t)n * np * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_dns(LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = lis_host_matrix_check_set(A);
    if (err) return err;
    A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_DNS;                       /* nnz is left alone, as in the reference */
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ CSR -> X */
static LIS_INT finish(LIS_MATRIX Aout, LIS_INT err)
{
    if (err) return err;
    err = lis_matrix_assemble(Aout);
    if (err) lis_matrix_storage_destroy(Aout);
    return err;
}

/* index[0..n] row starts (first one n+1), value[0..n) the diagonal (0 where a row stores none, counted in
 * ndz), off-diagonals behind in CSR order: src/matrix/lis_matrix_msr.c:982-1100 */
static LIS_INT csr2msr(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, nnz = Ain->ptr[n];
    LIS_INT *index, err, ndz = 0;
    LIS_SCALAR *value;
    for (LIS_INT i = 0; i < n; i++) {
        int has = 0;
        for (LIS_INT j = Ain->ptr[i]; j < Ain->ptr[i + 1]; j++) if (Ain->index[j] == i) has = 1;
        ndz += !has;
    }
    err = lis_matrix_malloc_msr(n, nnz, ndz, &index, &value);
    if (err) return err;
    LIS_INT k = n + 1;
    for (LIS_INT i = 0; i < n; i++) {
        index[i] = k;
        value[i] = 0.0;
        for (LIS_INT j = Ain->ptr[i]; j < Ain->ptr[i + 1]; j++) {
            if (Ain->index[j] == i) value[i] = Ain->value[j];            /* a repeated diagonal entry: the last one stays, as there */
            else { value[k] = Ain->value[j]; index[k] = Ain->index[j]; k++; }
        }
    }
    index[n] = k;
    value[n] = 0.0;
    err = lis_matrix_set_msr(nnz, ndz, index, value, Aout);
    if (err) { lis_free2(2, index, value); return err; }
    return finish(Aout, LIS_SUCCESS);
}

/* entries in CSR order with their row numbers: src/matrix/lis_matrix_coo.c:729-782 */
static LIS_INT csr2coo(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, nnz = Ain->ptr[n];
    LIS_INT *row, *col, err;
    LIS_SCALAR *value;
    err = lis_matrix_malloc_coo(nnz, &row, &col, &value);
    if (err) return err;
    for (LIS_INT i = 0; i < n; i++)
        for (LIS_INT j = Ain->ptr[i]; j < Ain->ptr[i + 1]; j++) { row[j] = i; col[j] = Ain->index[j]; value[j] = Ain->value[j]; }
    err = lis_matrix_set_coo(nnz, row, col, value, Aout);
    if (err) { lis_free2(3, row, col, value); return err; }
    return finish(Aout, LIS_SUCCESS);
}

/* column-major dense, value[j*n+i]: src/matrix/lis_matrix_dns.c:745-815 */
static LIS_INT csr2dns(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n, np = Ain->np;
    LIS_SCALAR *value;
    LIS_INT err;
    if ((double)n * (double)np * sizeof(LIS_SCALAR) > 64e9) { LIS_SETERR_MEM((size_t)n * np * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    err = lis_matrix_malloc_dns(n, np, &value);
    if (err) return err;
    memset(value, 0, (size_t)n * (size_t)np * sizeof(LIS_SCALAR));
    for (LIS_INT i = 0; i < n; i++)
        for (LIS_INT j = Ain->ptr[i]; j < Ain->ptr[i + 1]; j++) value[(size_t)Ain->index[j] * n + i] = Ain->value[j];
    err = lis_matrix_set_dns(value, Aout);
    if (err) { lis_free(value); return err; }
    return finish(Aout, LIS_SUCCESS);
}

/* The reference goes CSR -> CSC -> BSC (src/matrix/lis_matrix_ops.c:231-248, lis_matrix_bsc.c:352-560):
 * block columns of bnc columns; inside one, blocks in the order their block row is first met walking
 * the columns left to right, each column top to bottom; a block is column-major and zero-filled. */
static LIS_INT csr2bsc(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n;
    const LIS_INT bnr = Aout->conv_bnr, bnc = Aout->conv_bnc, bs = bnr * bnc;
    LIS_INT *cp = NULL, *ci = NULL, *bptr = NULL, *bindex = NULL, err;
    LIS_SCALAR *cv = NULL, *value = NULL;
    if (bnr != bnc) {
        /* the reference builder files entry (i,j) of a block at i + j*bnc while its product reads j*bnr + i
         * (lis_matrix_bsc.c:499 vs lis_matvec_bsc.c:131-141): only square blocks are consistent there */
        LIS_SETERR2(LIS_ERR_NOT_IMPLEMENTED, "BSC with %D x %D blocks: square blocks only\n", bnr, bnc);
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    err = lis_host_transpose(n, n, Ain->ptr, Ain->index, Ain->value, &cp, &ci, &cv);       /* CSC: rows ascending inside a column */
    if (err) return err;
    const LIS_INT nr = 1 + (n - 1) / bnr, nc = 1 + (n - 1) / bnc;
    LIS_INT *pos = (LIS_INT *)calloc((size_t)nr, sizeof(LIS_INT));
    LIS_INT *list = (LIS_INT *)malloc((size_t)nr * sizeof(LIS_INT));
    LIS_INT *cnt = (LIS_INT *)malloc(((size_t)nc + 1) * sizeof(LIS_INT));
    if (!pos || !list || !cnt) { err = LIS_OUT_OF_MEMORY; LIS_SETERR_MEM(nr); goto out; }
    cnt[0] = 0;
    for (LIS_INT bj = 0; bj < nc; bj++) {
        LIS_INT c = 0;
        for (LIS_INT jj = 0; jj < bnc && bj * bnc + jj < n; jj++)
            for (LIS_INT k = cp[bj * bnc + jj]; k < cp[bj * bnc + jj + 1]; k++) {
                const LIS_INT bi = ci[k] / bnr;
                if (!pos[bi]) { pos[bi] = 1; list[c++] = bi; }
            }
        for (LIS_INT k = 0; k < c; k++) pos[list[k]] = 0;
        cnt[bj + 1] = cnt[bj] + c;
    }
    const LIS_INT bnnz = cnt[nc];
    err = lis_matrix_malloc_bsc(n, bnr, bnc, bnnz, &bptr, &bindex, &value);
    if (err) goto out;
    memcpy(bptr, cnt, ((size_t)nc + 1) * sizeof(LIS_INT));
    for (LIS_INT bj = 0; bj < nc; bj++) {
        LIS_INT kk = bptr[bj];
        for (LIS_INT jj = 0; jj < bnc && bj * bnc + jj < n; jj++)
            for (LIS_INT k = cp[bj * bnc + jj]; k < cp[bj * bnc + jj + 1]; k++) {
                const LIS_INT bi = ci[k] / bnr, ii = ci[k] % bnr;
                if (pos[bi] == 0) {
                    pos[bi] = kk + 1;
                    bindex[kk] = bi;
                    for (LIS_INT q = 0; q < bs; q++) value[(size_t)kk * bs + q] = 0.0;
                    kk++;
                }
                value[(size_t)(pos[bi] - 1) * bs + (size_t)jj * bnr + ii] = cv[k];
            }
        for (LIS_INT k = bptr[bj]; k < bptr[bj + 1]; k++) pos[bindex[k]] = 0;
    }
    err = lis_matrix_set_bsc(bnr, bnc, bnnz, bptr, bindex, value, Aout);
    if (err) { lis_free2(3, bptr, bindex, value); goto out; }
    err = finish(Aout, LIS_SUCCESS);                   /* nnz stays 0, as lis_matrix_set_bsc leaves it */
out:
    free(pos); free(list); free(cnt);
    lis_free2(3, cp, ci, cv);
    return err;
}

/* the partition the reference derives when the caller gave none (lis_matrix_vbr.c:262-337): a boundary
 * wherever a run of consecutive column numbers starts or ends in any (sorted) row; the same list cuts
 * rows and columns */
static LIS_INT vbr_partition(LIS_MATRIX Ain, LIS_INT *nblk, LIS_INT **row, LIS_INT **col);
LIS_INT lis_host_vbr_partition(LIS_MATRIX Ain, LIS_INT *nblk, LIS_INT **row, LIS_INT **col) { return vbr_partition(Ain, nblk, row, col); }
static LIS_INT vbr_partition(LIS_MATRIX Ain, LIS_INT *nblk, LIS_INT **row, LIS_INT **col)
{
    const LIS_INT n = Ain->n;
    char *cut = (char *)calloc((size_t)n + 2, 1);
    if (cut == NULL) { LIS_SETERR_MEM(n); return LIS_OUT_OF_MEMORY; }
    for (LIS_INT i = 0; i < n; i++) {
        const LIS_INT s = Ain->ptr[i], e = Ain->ptr[i + 1];
        if (s >= e) continue;
        cut[Ain->index[s]] = 1;
        for (LIS_INT j = s + 1; j < e; j++)
            if (Ain->index[j - 1] != Ain->index[j] - 1) { cut[Ain->index[j]] = 1; cut[Ain->index[j - 1] + 1] = 1; }
        cut[Ain->index[e - 1] + 1] = 1;
    }
    cut[n] = 1;                    /* (the reference leaves rows behind the last cut out when column n-1 is empty) */
    LIS_INT k = 0;
    for (LIS_INT i = 1; i <= n; i++) k += cut[i] != 0;
    *row = (LIS_INT *)tracked((size_t)k + 1, sizeof(LIS_INT), "lis_matrix_get_vbr_rowcol::row");
    *col = (LIS_INT *)tracked((size_t)k + 1, sizeof(LIS_INT), "lis_matrix_get_vbr_rowcol::col");
    if (!*row || !*col) { free(cut); lis_free2(2, *row, *col); LIS_SETERR_MEM(k); return LIS_OUT_OF_MEMORY; }
    (*row)[0] = (*col)[0] = 0;
    k = 0;
    for (LIS_INT i = 1; i <= n; i++) if (cut[i]) { k++; (*row)[k] = (*col)[k] = i; }
    *nblk = k;
    free(cut);
    return LIS_SUCCESS;
}

/* src/matrix/lis_matrix_vbr.c:501-744: sorts Ain's rows, then per block row the blocks in first-met
 * order, each column-major (bnr rows) and zero-filled; ptr[] = start of every block in value[] */
static LIS_INT csr2vbr(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_INT n = Ain->n;
    LIS_INT nr = 0, *row = NULL, *col = NULL, *ptr = NULL, *bptr = NULL, *bindex = NULL, err;
    LIS_SCALAR *value = NULL;
    lis_matrix_sort_csr(Ain);
    err = vbr_partition(Ain, &nr, &row, &col);
    if (err) return err;
    const LIS_INT nc = nr;
    LIS_INT *blk_of = (LIS_INT *)malloc((size_t)(n > 0 ? n : 1) * sizeof(LIS_INT));
    LIS_INT *pos = (LIS_INT *)calloc((size_t)(nc > 0 ? nc : 1), sizeof(LIS_INT));
    LIS_INT *list = (LIS_INT *)malloc((size_t)(nc > 0 ? nc : 1) * sizeof(LIS_INT));
    LIS_INT *bcnt = (LIS_INT *)malloc(((size_t)nr + 1) * sizeof(LIS_INT));
    LIS_INT *vcnt = (LIS_INT *)malloc(((size_t)nr + 1) * sizeof(LIS_INT));
    if (!blk_of || !pos || !list || !bcnt || !vcnt) { err = LIS_OUT_OF_MEMORY; LIS_SETERR_MEM(n); goto fail; }
    for (LIS_INT b = 0; b < nc; b++) for (LIS_INT j = col[b]; j < col[b + 1]; j++) blk_of[j] = b;
    bcnt[0] = vcnt[0] = 0;
    for (LIS_INT bi = 0; bi < nr; bi++) {
        const LIS_INT h = row[bi + 1] - row[bi];
        LIS_INT c = 0, words = 0;
        for (LIS_INT i = row[bi]; i < row[bi + 1]; i++)
            for (LIS_INT j = Ain->ptr[i]; j < Ain->ptr[i + 1]; j++) {
                const LIS_INT bj = blk_of[Ain->index[j]];
                if (!pos[bj]) { pos[bj] = 1; list[c++] = bj; words += h * (col[bj + 1] - col[bj]); }
            }
        for (LIS_INT k = 0; k < c; k++) pos[list[k]] = 0;
        bcnt[bi + 1] = bcnt[bi] + c;
        vcnt[bi + 1] = vcnt[bi] + words;
    }
    const LIS_INT bnnz = bcnt[nr], nnz = vcnt[nr];
    ptr = (LIS_INT *)tracked((size_t)bnnz + 1, sizeof(LIS_INT), "lis_matrix_convert_csr2vbr::ptr");
    bptr = (LIS_INT *)tracked((size_t)nr + 1, sizeof(LIS_INT), "lis_matrix_convert_csr2vbr::bptr");
    bindex = (LIS_INT *)tracked((size_t)bnnz, sizeof(LIS_INT), "lis_matrix_convert_csr2vbr::bindex");
    value = (LIS_SCALAR *)tracked((size_t)nnz, sizeof(LIS_SCALAR), "lis_matrix_convert_csr2vbr::value");
    if (!ptr || !bptr || !bindex || !value) { err = LIS_OUT_OF_MEMORY; LIS_SETERR_MEM(nnz); goto fail; }
    memcpy(bptr, bcnt, ((size_t)nr + 1) * sizeof(LIS_INT));
    ptr[0] = 0;
    for (LIS_INT bi = 0; bi < nr; bi++) {
        const LIS_INT h = row[bi + 1] - row[bi];
        LIS_INT kk = bptr[bi], kv = vcnt[bi];
        ptr[kk] = kv;
        for (LIS_INT i = row[bi]; i < row[bi + 1]; i++)
            for (LIS_INT k = Ain->ptr[i]; k < Ain->ptr[i + 1]; k++) {
                const LIS_INT bj = blk_of[Ain->index[k]], j = Ain->index[k] - col[bj];
                if (pos[bj] == 0) {
                    const LIS_INT words = h * (col[bj + 1] - col[bj]);
                    memset(value + kv, 0, (size_t)words * sizeof(LIS_SCALAR));
                    bindex[kk] = bj;
                    pos[bj] = kv + 1;
                    kv += words;
                    ptr[kk + 1] = kv;
                    kk++;
                }
                value[(size_t)(pos[bj] - 1) + (size_t)j * h + (i - row[bi])] = Ain->value[k];
            }
        for (LIS_INT k = bptr[bi]; k < bptr[bi + 1]; k++) pos[bindex[k]] = 0;
    }
    ptr[bnnz] = nnz;
    err = lis_matrix_set_vbr(nnz, nr, nc, bnnz, row, col, ptr, bptr, bindex, value, Aout);
    if (err) goto fail;
    free(blk_of); free(pos); free(list); free(bcnt); free(vcnt);
    return finish(Aout, LIS_SUCCESS);
fail:
    free(blk_of); free(pos); free(list); free(bcnt); free(vcnt);
    lis_free2(6, row, col, ptr, bptr, bindex, value);
    return err;
}

LIS_INT lis_host_ext_from_csr(LIS_MATRIX Acsr, LIS_MATRIX Aout, int *handled)
{
    static const char *names[] = {"MSR", "BSC", "VBR", "COO", "DNS"};
    LIS_INT err;
    *handled = 1;
    switch (Aout->matrix_type) {
    case LIS_MATRIX_MSR: err = single_process(Acsr, names[0]); return err ? err : csr2msr(Acsr, Aout);
    case LIS_MATRIX_BSC: err = single_process(Acsr, names[1]); return err ? err : csr2bsc(Acsr, Aout);
    case LIS_MATRIX_VBR: err = single_process(Acsr, names[2]); return err ? err : csr2vbr(Acsr, Aout);
    case LIS_MATRIX_COO: err = single_process(Acsr, names[3]); return err ? err : csr2coo(Acsr, Aout);
    case LIS_MATRIX_DNS: err = single_process(Acsr, names[4]); return err ? err : csr2dns(Acsr, Aout);
    default: *handled = 0; return LIS_SUCCESS;
    }
}

/* ------------------------------------------------------------------ X -> CSR-shaped arrays
 * keep_zeros = 1: every stored slot, in the order lis_matvec_<fmt> accumulates (the device mirror);
 * keep_zeros = 0: the back-conversion to CSR, which drops the zero fill of dense blocks like the
 * reference (lis_matrix_bsc.c:563-740, lis_matrix_vbr.c:746-873, lis_matrix_dns.c:819-900,
 * lis_matrix_msr.c:1102-1190: diagonal first when it is not 0).
 * COO rows keep the order of k (a stable counting sort); the reference's coo2csr runs an unstable
 * quicksort over row[] (lis_matrix_coo.c:803, lis_sort_iid), which leaves the order inside a row to the
 * pivot sequence -- here the CSR product of the result adds in the COO product's order instead. */
typedef struct { LIS_INT r, c; LIS_SCALAR v; } slot_t;

/* walks every stored slot of A in accumulation order; visit() gets (row, column, value) */
typedef void (*visit_fn)(void *ctx, LIS_INT r, LIS_INT c, LIS_SCALAR v);

/* msr_offdiag_only: leave MSR's diagonal slots out (the transposed mirror keeps them apart) */
static void walk_ex(LIS_MATRIX A, visit_fn visit, void *ctx, int msr_offdiag_only)
{
    const LIS_INT n = A->n;
    switch (A->matrix_type) {
    /* the four hot-path formats below are walked only for the transposed mirror (lis_matvech): storage
     * order, padding slots included -- the reference's serial lis_matvech_<fmt> scatters in exactly this
     * order (lis_matvec_ell.c:230-245, lis_matvec_dia.c:320-345, lis_matvec_jad.c:330-350, lis_matvec_bsr.c:985-1005) */
    case LIS_MATRIX_ELL:
        for (LIS_INT j = 0; j < A->maxnzr; j++)
            for (LIS_INT i = 0; i < n; i++) visit(ctx, i, A->index[(size_t)j * n + i], A->value[(size_t)j * n + i]);
        break;
    case LIS_MATRIX_DIA:
        for (LIS_INT j = 0; j < A->nnd; j++) {
            const LIS_INT off = A->index[j], js = off < 0 ? -off : 0, je = n - off < n ? n - off : n;
            for (LIS_INT i = js; i < je; i++) visit(ctx, i, i + off, A->value[(size_t)j * n + i]);
        }
        break;
    case LIS_MATRIX_JAD:
        for (LIS_INT j = 0; j < A->maxnzr; j++)
            for (LIS_INT p = A->ptr[j], k = 0; p < A->ptr[j + 1]; p++, k++) visit(ctx, A->row[k], A->index[p], A->value[p]);
        break;
    case LIS_MATRIX_BSR: {
        const LIS_INT bnr = A->bnr, bnc = A->bnc, bs = bnr * bnc;
        for (LIS_INT bi = 0; bi < A->nr; bi++)
            for (LIS_INT bc = A->bptr[bi]; bc < A->bptr[bi + 1]; bc++)
                for (LIS_INT j = 0; j < bnc; j++)
                    for (LIS_INT i = 0; i < bnr; i++) {
                        const LIS_INT r = bi * bnr + i, c = A->bindex[bc] * bnc + j;
                        if (r < n && c < n) visit(ctx, r, c, A->value[(size_t)bc * bs + (size_t)j * bnr + i]);
                    }
        break;
    }
    case LIS_MATRIX_MSR:
        for (LIS_INT i = 0; i < n; i++) {
            if (!msr_offdiag_only) visit(ctx, i, i, A->value[i]);
            for (LIS_INT j = A->index[i]; j < A->index[i + 1]; j++) visit(ctx, i, A->index[j], A->value[j]);
        }
        break;
    case LIS_MATRIX_COO:
        for (LIS_INT k = 0; k < A->nnz; k++) visit(ctx, A->row[k], A->col[k], A->value[k]);
        break;
    case LIS_MATRIX_BSC: {
        const LIS_INT bnr = A->bnr, bnc = A->bnc, bs = bnr * bnc;
        for (LIS_INT bj = 0; bj < A->nc; bj++)
            for (LIS_INT bc = A->bptr[bj]; bc < A->bptr[bj + 1]; bc++)
                for (LIS_INT j = 0; j < bnc; j++)
                    for (LIS_INT i = 0; i < bnr; i++) {
                        const LIS_INT r = A->bindex[bc] * bnr + i, c = bj * bnc + j;
                        if (r < n && c < n) visit(ctx, r, c, A->value[(size_t)bc * bs + (size_t)j * bnr + i]);
                    }
        break;
    }
    case LIS_MATRIX_VBR:
        for (LIS_INT bi = 0; bi < A->nr; bi++)
            for (LIS_INT bc = A->bptr[bi]; bc < A->bptr[bi + 1]; bc++) {
                const LIS_INT bj = A->bindex[bc];
                LIS_INT k = A->ptr[bc];
                for (LIS_INT j = A->col[bj]; j < A->col[bj + 1]; j++)
                    for (LIS_INT i = A->row[bi]; i < A->row[bi + 1]; i++) visit(ctx, i, j, A->value[k++]);
            }
        break;
    case LIS_MATRIX_DNS:
        for (LIS_INT j = 0; j < A->np; j++)
            for (LIS_INT i = 0; i < n; i++) visit(ctx, i, j, A->value[(size_t)j * n + i]);
        break;
    default: break;
    }
}

static void walk(LIS_MATRIX A, visit_fn visit, void *ctx) { walk_ex(A, visit, ctx, 0); }

typedef struct { LIS_INT *ptr, *fill, *index; LIS_SCALAR *value; int keep_zeros, msr; } fill_ctx;

static int kept(const fill_ctx *f, LIS_INT r, LIS_INT c, LIS_SCALAR v)
{
    (void)r; (void)c;
    return f->keep_zeros || v != 0.0;
}
static void count_visit(void *ctx, LIS_INT r, LIS_INT c, LIS_SCALAR v) { fill_ctx *f = (fill_ctx *)ctx; if (kept(f, r, c, v)) f->ptr[r + 1]++; }
static void fill_visit(void *ctx, LIS_INT r, LIS_INT c, LIS_SCALAR v)
{
    fill_ctx *f = (fill_ctx *)ctx;
    if (!kept(f, r, c, v)) return;
    f->index[f->fill[r]] = c; f->value[f->fill[r]] = v; f->fill[r]++;
}

/* A (MSR/COO/BSC/VBR/DNS) as tracked CSR arrays, rows in accumulation order.  MSR with keep_zeros keeps
 * every stored off-diagonal and the diagonal slot; without, only the diagonal is tested against 0
 * (the reference's msr2csr). */
LIS_INT lis_host_ordered_rows(LIS_MATRIX A, int keep_zeros, LIS_INT *nnz_out, LIS_INT **ptr_out, LIS_INT **index_out, LIS_SCALAR **value_out)
{
    const LIS_INT n = A->n;
    fill_ctx f;
    memset(&f, 0, sizeof(f));
    f.keep_zeros = keep_zeros || A->matrix_type == LIS_MATRIX_COO;
    f.ptr = (LIS_INT *)tracked((size_t)n + 1, sizeof(LIS_INT), "lis_host_ordered_rows::ptr");
    if (f.ptr == NULL) { LIS_SETERR_MEM(n); return LIS_OUT_OF_MEMORY; }
    memset(f.ptr, 0, ((size_t)n + 1) * sizeof(LIS_INT));
    if (A->matrix_type == LIS_MATRIX_MSR && !keep_zeros) {
        /* off-diagonals stay even when 0; the diagonal slot only when it is not */
        for (LIS_INT i = 0; i < n; i++) f.ptr[i + 1] = A->index[i + 1] - A->index[i] + (A->value[i] != 0.0);
    } else {
        walk(A, count_visit, &f);
    }
    for (LIS_INT i = 0; i < n; i++) f.ptr[i + 1] += f.ptr[i];
    const LIS_INT nnz = f.ptr[n];
    f.index = (LIS_INT *)tracked((size_t)nnz, sizeof(LIS_INT), "lis_host_ordered_rows::index");
    f.value = (LIS_SCALAR *)tracked((size_t)nnz, sizeof(LIS_SCALAR), "lis_host_ordered_rows::value");
    f.fill = (LIS_INT *)malloc(((size_t)n + 1) * sizeof(LIS_INT));
    if (!f.index || !f.value || !f.fill) { lis_free2(3, f.ptr, f.index, f.value); free(f.fill); LIS_SETERR_MEM(nnz); return LIS_OUT_OF_MEMORY; }
    memcpy(f.fill, f.ptr, ((size_t)n + 1) * sizeof(LIS_INT));
    if (A->matrix_type == LIS_MATRIX_MSR && !keep_zeros) {
        for (LIS_INT i = 0; i < n; i++) {
            LIS_INT k = f.ptr[i];
            if (A->value[i] != 0.0) { f.index[k] = i; f.value[k] = A->value[i]; k++; }
            for (LIS_INT j = A->index[i]; j < A->index[i + 1]; j++) { f.index[k] = A->index[j]; f.value[k] = A->value[j]; k++; }
        }
    } else {
        walk(A, fill_visit, &f);
    }
    free(f.fill);
    *nnz_out = nnz; *ptr_out = f.ptr; *index_out = f.index; *value_out = f.value;
    return LIS_SUCCESS;
}

/* The mirror lis_matvech runs on: A^T as CSR arrays, each row of A^T (= column c of A) listing its entries in
 * the order in which the reference's serial lis_matvech_<fmt> adds them into y[c] -- every one of those
 * routines zeroes y and scatters y[col] += v * x[row] while walking the storage once, front to back, so
 * the order is "storage order, filed by column" (a stable counting sort).  MSR sets y[i] = d[i]*x[i] first
 * (lis_matvec_msr.c:215-232): its mirror holds the off-diagonals only and runs on the split-order kernel
 * behind the diagonal.  Explicit zeros and padding slots stay in (they are added there too). */
typedef struct { LIS_INT *ptr, *fill, *index; LIS_SCALAR *value; } tr_ctx;
static void tr_count(void *ctx, LIS_INT r, LIS_INT c, LIS_SCALAR v) { (void)r; (void)v; ((tr_ctx *)ctx)->ptr[c + 1]++; }
static void tr_fill(void *ctx, LIS_INT r, LIS_INT c, LIS_SCALAR v)
{
    tr_ctx *t = (tr_ctx *)ctx;
    t->index[t->fill[c]] = r; t->value[t->fill[c]] = v; t->fill[c]++;
}

LIS_INT lis_host_transposed_rows(LIS_MATRIX A, LIS_INT **ptr_out, LIS_INT **index_out, LIS_SCALAR **value_out)
{
    const LIS_INT n = A->n;
    tr_ctx t;
    memset(&t, 0, sizeof(t));
    switch (A->matrix_type) {
    case LIS_MATRIX_ELL: case LIS_MATRIX_DIA: case LIS_MATRIX_JAD: case LIS_MATRIX_BSR:
    case LIS_MATRIX_MSR: case LIS_MATRIX_COO: case LIS_MATRIX_BSC: case LIS_MATRIX_VBR: case LIS_MATRIX_DNS: break;
    default: LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED;
    }
    t.ptr = (LIS_INT *)tracked((size_t)n + 1, sizeof(LIS_INT), "lis_host_transposed_rows::ptr");
    if (t.ptr == NULL) { LIS_SETERR_MEM(n); return LIS_OUT_OF_MEMORY; }
    memset(t.ptr, 0, ((size_t)n + 1) * sizeof(LIS_INT));
    walk_ex(A, tr_count, &t, 1);
    for (LIS_INT i = 0; i < n; i++) t.ptr[i + 1] += t.ptr[i];
    const LIS_INT nnz = t.ptr[n];
    t.index = (LIS_INT *)tracked((size_t)nnz, sizeof(LIS_INT), "lis_host_transposed_rows::index");
    t.value = (LIS_SCALAR *)tracked((size_t)nnz, sizeof(LIS_SCALAR), "lis_host_transposed_rows::value");
    t.fill = (LIS_INT *)malloc(((size_t)n + 1) * sizeof(LIS_INT));
    if (!t.index || !t.value || !t.fill) { lis_free2(3, t.ptr, t.index, t.value); free(t.fill); LIS_SETERR_MEM(nnz); return LIS_OUT_OF_MEMORY; }
    memcpy(t.fill, t.ptr, ((size_t)n + 1) * sizeof(LIS_INT));
    walk_ex(A, tr_fill, &t, 1);
    free(t.fill);
    *ptr_out = t.ptr; *index_out = t.index; *value_out = t.value;
    return LIS_SUCCESS;
}

LIS_INT lis_host_ext_to_csr(LIS_MATRIX Ain, LIS_MATRIX Aout, int *handled)
{
    LIS_INT nnz, *ptr, *index, err;
    LIS_SCALAR *value;
    *handled = 1;
    switch (Ain->matrix_type) {
    case LIS_MATRIX_MSR: case LIS_MATRIX_BSC: case LIS_MATRIX_VBR: case LIS_MATRIX_COO: case LIS_MATRIX_DNS: break;
    default: *handled = 0; return LIS_SUCCESS;
    }
    err = lis_host_ordered_rows(Ain, 0, &nnz, &ptr, &index, &value);
    if (err) return err;
    err = lis_matrix_set_csr(nnz, ptr, index, value, Aout);
    if (err) { lis_free2(3, ptr, index, value); return err; }
    return finish(Aout, LIS_SUCCESS);
}

/* d[i] = the stored diagonal entry (src/matrix/lis_matrix_<fmt>.c get_diagonal): MSR value[i]; COO the
 * first k with row == col == i; BSC / VBR / DNS the diagonal slot of the dense storage (first block found) */
typedef struct { LIS_SCALAR *d; char *seen; } diag_ctx;
static void diag_visit(void *ctx, LIS_INT r, LIS_INT c, LIS_SCALAR v)
{
    diag_ctx *q = (diag_ctx *)ctx;
    if (r == c && !q->seen[r]) { q->d[r] = v; q->seen[r] = 1; }
}

LIS_INT lis_host_ext_get_diagonal(LIS_MATRIX A, LIS_SCALAR *d, int *handled)
{
    *handled = 1;
    switch (A->matrix_type) {
    case LIS_MATRIX_MSR: case LIS_MATRIX_BSC: case LIS_MATRIX_VBR: case LIS_MATRIX_COO: case LIS_MATRIX_DNS: break;
    default: *handled = 0; return LIS_SUCCESS;
    }
    diag_ctx q;
    q.d = d;
    q.seen = (char *)calloc((size_t)(A->n > 0 ? A->n : 1), 1);
    if (q.seen == NULL) { LIS_SETERR_MEM(A->n); return LIS_OUT_OF_MEMORY; }
    for (LIS_INT i = 0; i < A->n; i++) d[i] = 0.0;
    walk(A, diag_visit, &q);
    free(q.seen);
    return LIS_SUCCESS;
}

/* A <- A - sigma*I on the stored diagonal slots (src/matrix/lis_matrix_<fmt>.c, lis_matrix_shift_diagonal_<fmt>:
 * MSR value[i]; COO every k with row == col; JAD / BSR / BSC / VBR / DNS the slot of (i, i)) */
LIS_INT lis_host_ext_shift_diagonal(LIS_MATRIX A, LIS_SCALAR sigma, int *handled)
{
    const LIS_INT n = A->n;
    *handled = 1;
    switch (A->matrix_type) {
    case LIS_MATRIX_MSR:
        for (LIS_INT i = 0; i < n; i++) A->value[i] -= sigma;
        break;
    case LIS_MATRIX_COO:
        for (LIS_INT k = 0; k < A->nnz; k++) if (A->row[k] == A->col[k]) A->value[k] -= sigma;
        break;
    case LIS_MATRIX_DNS:
        for (LIS_INT i = 0; i < n; i++) A->value[(size_t)i * n + i] -= sigma;
        break;
    case LIS_MATRIX_JAD:
        for (LIS_INT j = 0; j < A->maxnzr; j++)
            for (LIS_INT p = A->ptr[j], k = 0; p < A->ptr[j + 1]; p++, k++) if (A->row[k] == A->index[p]) A->value[p] -= sigma;
        break;
    case LIS_MATRIX_BSR: case LIS_MATRIX_BSC: {
        const LIS_INT bnr = A->bnr, bnc = A->bnc, bs = bnr * bnc, outer = A->matrix_type == LIS_MATRIX_BSR ? A->nr : A->nc;
        for (LIS_INT bo = 0; bo < outer; bo++)
            for (LIS_INT bc = A->bptr[bo]; bc < A->bptr[bo + 1]; bc++)
                for (LIS_INT j = 0; j < bnc; j++)
                    for (LIS_INT i = 0; i < bnr; i++) {
                        const LIS_INT r = (A->matrix_type == LIS_MATRIX_BSR ? bo : A->bindex[bc]) * bnr + i;
                        const LIS_INT c = (A->matrix_type == LIS_MATRIX_BSR ? A->bindex[bc] : bo) * bnc + j;
                        if (r == c && r < n) A->value[(size_t)bc * bs + (size_t)j * bnr + i] -= sigma;
                    }
        break;
    }
    case LIS_MATRIX_VBR:
        for (LIS_INT bi = 0; bi < A->nr; bi++)
            for (LIS_INT bc = A->bptr[bi]; bc < A->bptr[bi + 1]; bc++) {
                const LIS_INT bj = A->bindex[bc], h = A->row[bi + 1] - A->row[bi];
                for (LIS_INT j = A->col[bj]; j < A->col[bj + 1]; j++)
                    if (j >= A->row[bi] && j < A->row[bi + 1]) A->value[(size_t)A->ptr[bc] + (size_t)(j - A->col[bj]) * h + (j - A->row[bi])] -= sigma;
            }
        break;
    default: *handled = 0; break;
    }
    return LIS_SUCCESS;
}
