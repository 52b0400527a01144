/*
 * lis_matrix.c -- LIS_MATRIX objects: creation, caller-array adoption (set_csr & friends),
 * row-wise assembly (set_value), assemble, duplicate, destroy, diagonal extraction, and the
 * D/L/U split that the SSOR preconditioner and the split-order SpMV need.
 *
 * Host C.  Follows the object model of the reference (src/matrix/lis_matrix.c:71-1021,
 * lis_matrix_csr.c:764-949 split, :1257-1310 merge, lis_matrix_ops.c:727-780 get_diagonal);
 * the matrix arrays stay on the host exactly as the caller handed them over, and a private
 * device mirror (lis_matrix_dev.c) is built the first time a kernel needs the matrix.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lis_device.h"
#include "lis_host.h"
#include "lis_b200_kernels.h"

#define LIS_MATRIX_W_ANNZ 10

/* ------------------------------------------------------------------ helpers */
static void matrix_init(LIS_MATRIX A)
{
    memset(A, 0, sizeof(struct LIS_MATRIX_STRUCT));
    A->label = LIS_LABEL_MATRIX;
    A->matrix_type = LIS_MATRIX_CSR;
    A->status = LIS_MATRIX_DECIDING_SIZE;
    A->w_annz = LIS_MATRIX_W_ANNZ;
    A->conv_bnr = 2;
    A->conv_bnc = 2;
    A->is_destroy = LIS_TRUE;
}

/* the reference's lis_matrix_check levels (src/matrix/lis_matrix.c:93-196), condensed */
enum { CHECK_ALL, CHECK_SIZE, CHECK_NULL, CHECK_NOT_ASSEMBLED, CHECK_SET };

static LIS_INT matrix_check(LIS_MATRIX A, int level)
{
    if (!lis_is_malloc(A)) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "matrix A is undefined\n");
        return LIS_ERR_ILL_ARG;
    }
    switch (level) {
    case CHECK_SIZE:
        if (A->status == LIS_MATRIX_DECIDING_SIZE) {
            LIS_SETERR(LIS_ERR_ILL_ARG, "matrix size is undefined\n");
            return LIS_ERR_ILL_ARG;
        }
        break;
    case CHECK_NULL:
        break;
    case CHECK_NOT_ASSEMBLED:
        if (A->status != LIS_MATRIX_DECIDING_SIZE && A->status != LIS_MATRIX_NULL && A->status != LIS_MATRIX_ASSEMBLING) {
            LIS_SETERR(LIS_ERR_ILL_ARG, "matrix A has already been assembled\n");
            return LIS_ERR_ILL_ARG;
        }
        break;
    case CHECK_SET:
        if (A->status == LIS_MATRIX_DECIDING_SIZE) {
            LIS_SETERR(LIS_ERR_ILL_ARG, "matrix size is undefined\n");
            return LIS_ERR_ILL_ARG;
        }
        if (A->status != LIS_MATRIX_NULL) {
            LIS_SETERR(LIS_ERR_ILL_ARG, "matrix A has already been assigned\n");
            return LIS_ERR_ILL_ARG;
        }
        break;
    default:   /* CHECK_ALL: usable as an input operand */
        if (A->status == LIS_MATRIX_DECIDING_SIZE) {
            LIS_SETERR(LIS_ERR_ILL_ARG, "matrix size is undefined\n");
            return LIS_ERR_ILL_ARG;
        }
        if (A->status <= LIS_MATRIX_ASSEMBLING) {
            LIS_SETERR(LIS_ERR_ILL_ARG, "matrix A is not assembled\n");
            return LIS_ERR_ILL_ARG;
        }
        break;
    }
    return LIS_SUCCESS;
}

LIS_INT lis_host_matrix_check_input(LIS_MATRIX A) { return matrix_check(A, CHECK_ALL); }
LIS_INT lis_host_matrix_check_set(LIS_MATRIX A) { return matrix_check(A, CHECK_SET); }

/* ------------------------------------------------------------------ lifetime */
LIS_INT lis_matrix_create(LIS_Comm comm, LIS_MATRIX *Amat)
{
    *Amat = (LIS_MATRIX)lis_malloc(sizeof(struct LIS_MATRIX_STRUCT), "lis_matrix_create::Amat");
    if (*Amat == NULL) { LIS_SETERR_MEM(sizeof(struct LIS_MATRIX_STRUCT)); return LIS_OUT_OF_MEMORY; }
    matrix_init(*Amat);
    (*Amat)->comm = comm;
    (*Amat)->nprocs = lisd_nranks();
    (*Amat)->my_rank = lisd_rank();
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_size(LIS_MATRIX A, LIS_INT local_n, LIS_INT global_n)
{
    LIS_INT nprocs, my_rank, is, ie, err;
    LIS_INT *ranges;

    err = matrix_check(A, CHECK_NULL);
    if (err) return err;
    if (global_n > 0 && local_n > global_n) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "local n(=%D) is larger than global n(=%D)\n", local_n, global_n);
        return LIS_ERR_ILL_ARG;
    }
    if (local_n < 0 || global_n < 0) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "local n(=%D) or global n(=%D) are less than 0\n", local_n, global_n);
        return LIS_ERR_ILL_ARG;
    }
    if (lisd_nranks() == 1 && local_n == 0 && global_n == 0) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "local n(=%D) and global n(=%D) are 0\n", local_n, global_n);
        return LIS_ERR_ILL_ARG;
    }
    if (global_n > 0 && global_n < lisd_nranks()) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "global n(=%D) is smaller than nprocs(=%D)\n", global_n, lisd_nranks());
        return LIS_ERR_ILL_ARG;
    }
    err = lis_ranges_create(A->comm, &local_n, &global_n, &ranges, &is, &ie, &nprocs, &my_rank);
    if (err) return err;
    A->status = LIS_MATRIX_NULL;
    A->ranges = ranges;
    A->n = local_n; A->gn = global_n; A->np = local_n;
    A->my_rank = my_rank; A->nprocs = nprocs;
    A->is = is; A->ie = ie;
    return LIS_SUCCESS;
}

static void core_destroy(LIS_MATRIX_CORE C)
{
    if (C == NULL) return;
    lis_free(C->ptr); lis_free(C->row); lis_free(C->col); lis_free(C->index);
    lis_free(C->bptr); lis_free(C->bindex); lis_free(C->value); lis_free(C->work);
    lis_free(C);
}

LIS_INT lis_matrix_diag_destroy(LIS_MATRIX_DIAG D)
{
    if (D == NULL) return LIS_SUCCESS;
    lis_free(D->value); lis_free(D->work); lis_free(D->bns); lis_free(D->ptr);
    lis_free(D);
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_DLU_destroy(LIS_MATRIX A)
{
    if (A->D) lis_matrix_diag_destroy(A->D);
    if (A->L) core_destroy(A->L);
    if (A->U) core_destroy(A->U);
    A->D = NULL; A->L = NULL; A->U = NULL;
    A->is_splited = LIS_FALSE;
    lisd_matrix_drop(A);
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_storage_destroy(LIS_MATRIX A)
{
    if (A->is_destroy) {
        lis_free(A->ptr); lis_free(A->row); lis_free(A->col); lis_free(A->index);
        lis_free(A->bptr); lis_free(A->bindex); lis_free(A->value); lis_free(A->work);
        lis_free(A->conv_row); lis_free(A->conv_col);
        if (A->w_index || A->w_value) {
            for (LIS_INT i = 0; i < A->n; i++) {
                if (A->w_index) lis_free(A->w_index[i]);
                if (A->w_value) lis_free(A->w_value[i]);
            }
        }
        lis_free(A->w_nnz); lis_free(A->w_row); lis_free(A->w_index); lis_free(A->w_value);
    }
    A->ptr = A->row = A->col = A->index = A->bptr = A->bindex = NULL;
    A->value = A->work = NULL;
    A->conv_row = A->conv_col = NULL;
    A->w_nnz = A->w_row = NULL; A->w_index = NULL; A->w_value = NULL;
    lisd_matrix_drop(A);
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_destroy(LIS_MATRIX A)
{
    if (lis_is_malloc(A)) {
        lis_matrix_storage_destroy(A);
        lis_matrix_DLU_destroy(A);
        lis_matrix_diag_destroy(A->WD);
        if (A->l2g_map) lis_free(A->l2g_map);
        if (A->commtable) lisd_commtable_destroy(A->commtable);
        if (A->ranges) lis_free(A->ranges);
        lis_free(A);
    }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_duplicate(LIS_MATRIX Ain, LIS_MATRIX *Aout)
{
    LIS_INT err = matrix_check(Ain, CHECK_ALL);
    if (err) return err;
    *Aout = (LIS_MATRIX)lis_malloc(sizeof(struct LIS_MATRIX_STRUCT), "lis_matrix_duplicate::Aout");
    if (*Aout == NULL) { LIS_SETERR_MEM(sizeof(struct LIS_MATRIX_STRUCT)); return LIS_OUT_OF_MEMORY; }
    LIS_MATRIX B = *Aout;
    matrix_init(B);
    if (Ain->nprocs > 1) {
        const LIS_INT nh = Ain->np - Ain->n;
        if (nh > 0 && Ain->l2g_map) {
            B->l2g_map = (LIS_INT *)lis_malloc((size_t)nh * sizeof(LIS_INT), "lis_matrix_duplicate::l2g_map");
            if (B->l2g_map == NULL) { lis_matrix_destroy(B); *Aout = NULL; return LIS_OUT_OF_MEMORY; }
            memcpy(B->l2g_map, Ain->l2g_map, (size_t)nh * sizeof(LIS_INT));
        }
        if (Ain->ranges) {
            B->ranges = (LIS_INT *)lis_malloc((size_t)(Ain->nprocs + 1) * sizeof(LIS_INT), "lis_matrix_duplicate::ranges");
            if (B->ranges == NULL) { lis_matrix_destroy(B); *Aout = NULL; return LIS_OUT_OF_MEMORY; }
            memcpy(B->ranges, Ain->ranges, (size_t)(Ain->nprocs + 1) * sizeof(LIS_INT));
        }
    }
    B->status = LIS_MATRIX_NULL;
    B->is_block = Ain->is_block;
    B->n = Ain->n; B->gn = Ain->gn; B->np = Ain->np;
    B->comm = Ain->comm; B->my_rank = Ain->my_rank; B->nprocs = Ain->nprocs;
    B->is = Ain->is; B->ie = Ain->ie; B->origin = Ain->origin;
    B->is_destroy = Ain->is_destroy;
    if (Ain->nprocs > 1 && Ain->commtable) {
        err = lisd_commtable_duplicate(Ain, B);
        if (err) { lis_matrix_destroy(B); *Aout = NULL; return err; }
        B->is_comm = LIS_TRUE;
    }
    return LIS_SUCCESS;
}

/* take over every field of `src` (a temporary) into `dst`; frees the src shell */
void lis_host_matrix_adopt(LIS_MATRIX dst, LIS_MATRIX src)
{
    lis_matrix_storage_destroy(dst);
    lis_matrix_DLU_destroy(dst);
    lis_matrix_diag_destroy(dst->WD);
    if (dst->l2g_map) lis_free(dst->l2g_map);
    if (dst->commtable) lisd_commtable_destroy(dst->commtable);
    if (dst->ranges) lis_free(dst->ranges);
    memcpy(dst, src, sizeof(struct LIS_MATRIX_STRUCT));
    lis_free(src);
}

/* ------------------------------------------------------------------ queries */
LIS_INT lis_matrix_get_range(LIS_MATRIX A, LIS_INT *is, LIS_INT *ie)
{
    LIS_INT err = matrix_check(A, CHECK_SIZE);
    if (err) return err;
    *is = A->is; *ie = A->ie;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_get_size(LIS_MATRIX A, LIS_INT *local_n, LIS_INT *global_n)
{
    LIS_INT err = matrix_check(A, CHECK_SIZE);
    if (err) return err;
    *local_n = A->n; *global_n = A->gn;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_get_nnz(LIS_MATRIX A, LIS_INT *nnz)
{
    LIS_INT err = matrix_check(A, CHECK_SIZE);
    if (err) return err;
    *nnz = A->nnz;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_type(LIS_MATRIX A, LIS_INT matrix_type)
{
    LIS_INT err = matrix_check(A, CHECK_NOT_ASSEMBLED);
    if (err) return err;
    if (matrix_type < LIS_MATRIX_CSR || matrix_type > LIS_MATRIX_DNS) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "matrix_type is %D (Set between 1 to %D)\n", matrix_type, LIS_MATRIX_DNS);
        return LIS_ERR_ILL_ARG;
    }
    A->matrix_type = matrix_type;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_get_type(LIS_MATRIX A, LIS_INT *matrix_type)
{
    LIS_INT err = matrix_check(A, CHECK_NULL);
    if (err) return err;
    *matrix_type = A->matrix_type;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_is_assembled(LIS_MATRIX A)
{
    return A->status != LIS_MATRIX_NULL ? !LIS_SUCCESS : LIS_SUCCESS;
}

LIS_INT lis_matrix_set_blocksize(LIS_MATRIX A, LIS_INT bnr, LIS_INT bnc, LIS_INT row[], LIS_INT col[])
{
    LIS_INT err = matrix_check(A, CHECK_NULL);
    if (err) return err;
    if (bnr <= 0 || bnc <= 0) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "bnr=%D <= 0 or bnc=%D <= 0\n", bnr, bnc);
        return LIS_ERR_ILL_ARG;
    }
    if ((row == NULL) != (col == NULL)) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "either row[] or col[] is NULL\n");
        return LIS_ERR_ILL_ARG;
    }
    if (row != NULL) {                       /* variable block partition: VBR only */
        LIS_SETERR_IMP;
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    A->conv_bnr = bnr;
    A->conv_bnc = bnc;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_unset(LIS_MATRIX A)
{
    LIS_INT err = matrix_check(A, CHECK_SIZE);
    if (err) return err;
    if (A->is_copy) lis_matrix_storage_destroy(A);
    A->row = A->col = A->ptr = A->index = A->bptr = A->bindex = NULL;
    A->value = NULL;
    A->is_copy = LIS_FALSE;
    A->status = LIS_MATRIX_NULL;
    lisd_matrix_drop(A);
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ row-wise assembly */
LIS_INT lis_matrix_malloc(LIS_MATRIX A, LIS_INT nnz_row, LIS_INT nnz[])
{
    LIS_INT err = matrix_check(A, CHECK_NOT_ASSEMBLED);
    if (err) return err;
    const LIS_INT n = A->n;
    if (A->w_nnz == NULL) {
        A->w_nnz = (LIS_INT *)lis_malloc((size_t)(n > 0 ? n : 1) * sizeof(LIS_INT), "lis_matrix_malloc::A->w_nnz");
        if (A->w_nnz == NULL) { LIS_SETERR_MEM(n * sizeof(LIS_INT)); return LIS_OUT_OF_MEMORY; }
    }
    if (nnz == NULL) {
        A->w_annz = nnz_row;
        for (LIS_INT k = 0; k < n; k++) A->w_nnz[k] = nnz_row;
    } else {
        for (LIS_INT k = 0; k < n; k++) A->w_nnz[k] = nnz[k];
    }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_value(LIS_INT flag, LIS_INT i, LIS_INT j, LIS_SCALAR value, LIS_MATRIX A)
{
    LIS_INT err = matrix_check(A, CHECK_NOT_ASSEMBLED);
    if (err) return err;
    const LIS_INT n = A->n, gn = A->gn, is = A->is;
    if (A->origin) { i--; j--; }
    if (i < 0 || j < 0) {
        LIS_SETERR3(LIS_ERR_ILL_ARG, "i(=%D) or j(=%D) are less than %D\n", i + A->origin, j + A->origin, A->origin);
        return LIS_ERR_ILL_ARG;
    }
    if (i >= gn || j >= gn) {
        LIS_SETERR3(LIS_ERR_ILL_ARG, "i(=%D) or j(=%D) are larger than global n=(%D)\n", i + A->origin, j + A->origin, gn);
        return LIS_ERR_ILL_ARG;
    }
    if (i < is || i >= A->ie) {
        LIS_SETERR3(LIS_ERR_ILL_ARG, "row i(=%D) is outside the local range [%D,%D)\n", i + A->origin, is, A->ie);
        return LIS_ERR_ILL_ARG;
    }
    if (A->status == LIS_MATRIX_NULL) {
        const size_t nn = (size_t)(n > 0 ? n : 1);
        if (A->w_nnz == NULL) {
            A->w_nnz = (LIS_INT *)lis_malloc(nn * sizeof(LIS_INT), "lis_matrix_set_value::A->w_nnz");
            if (A->w_nnz == NULL) { LIS_SETERR_MEM(nn * sizeof(LIS_INT)); return LIS_OUT_OF_MEMORY; }
            for (LIS_INT k = 0; k < n; k++) A->w_nnz[k] = A->w_annz;
        }
        A->w_row = (LIS_INT *)lis_calloc(nn * sizeof(LIS_INT), "lis_matrix_set_value::w_row");
        A->w_index = (LIS_INT **)lis_calloc(nn * sizeof(LIS_INT *), "lis_matrix_set_value::w_index");
        A->w_value = (LIS_SCALAR **)lis_calloc(nn * sizeof(LIS_SCALAR *), "lis_matrix_set_value::w_value");
        if (!A->w_row || !A->w_index || !A->w_value) { LIS_SETERR_MEM(nn * sizeof(void *)); return LIS_OUT_OF_MEMORY; }
        for (LIS_INT k = 0; k < n; k++) {
            const size_t cap = (size_t)(A->w_nnz[k] > 0 ? A->w_nnz[k] : 1);
            A->w_index[k] = (LIS_INT *)lis_malloc(cap * sizeof(LIS_INT), "lis_matrix_set_value::w_index[k]");
            A->w_value[k] = (LIS_SCALAR *)lis_malloc(cap * sizeof(LIS_SCALAR), "lis_matrix_set_value::w_value[k]");
            if (!A->w_index[k] || !A->w_value[k]) { LIS_SETERR_MEM(cap * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
            if (A->w_nnz[k] < 1) A->w_nnz[k] = 1;
        }
        A->status = LIS_MATRIX_ASSEMBLING;
        A->is_copy = LIS_TRUE;
    }
    const LIS_INT r = i - is;
    LIS_INT k;
    for (k = 0; k < A->w_row[r]; k++)
        if (A->w_index[r][k] == j) break;
    if (k < A->w_row[r]) {
        if (flag == LIS_INS_VALUE) A->w_value[r][k] = value;
        else A->w_value[r][k] += value;
        return LIS_SUCCESS;
    }
    if (A->w_nnz[r] == A->w_row[r]) {
        const LIS_INT grow = A->w_annz > 0 ? A->w_annz : LIS_MATRIX_W_ANNZ;
        A->w_nnz[r] += grow;
        LIS_INT *ni = (LIS_INT *)lis_realloc(A->w_index[r], (size_t)A->w_nnz[r] * sizeof(LIS_INT));
        LIS_SCALAR *nv = (LIS_SCALAR *)lis_realloc(A->w_value[r], (size_t)A->w_nnz[r] * sizeof(LIS_SCALAR));
        if (ni) A->w_index[r] = ni;
        if (nv) A->w_value[r] = nv;
        if (!ni || !nv) { LIS_SETERR_MEM((size_t)A->w_nnz[r] * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    }
    k = A->w_row[r]++;
    A->w_index[r][k] = j;
    A->w_value[r][k] = value;
    return LIS_SUCCESS;
}

/* rows in insertion order, like lis_matrix_convert_rco2csr (src/matrix/lis_matrix_rco.c) */
static LIS_INT assemble_from_rows(LIS_MATRIX A, LIS_INT target_type)
{
    const LIS_INT n = A->n;
    LIS_INT nnz = 0, err;
    LIS_INT *ptr, *index;
    LIS_SCALAR *value;
    for (LIS_INT i = 0; i < n; i++) nnz += A->w_row[i];
    err = lis_matrix_malloc_csr(n, nnz, &ptr, &index, &value);
    if (err) return err;
    ptr[0] = 0;
    for (LIS_INT i = 0; i < n; i++) {
        LIS_INT k = ptr[i];
        for (LIS_INT j = 0; j < A->w_row[i]; j++, k++) {
            index[k] = A->w_index[i][j];
            value[k] = A->w_value[i][j];
        }
        ptr[i + 1] = k;
    }
    /* release the row buffers, install the CSR arrays */
    for (LIS_INT i = 0; i < n; i++) { lis_free(A->w_index[i]); lis_free(A->w_value[i]); }
    lis_free(A->w_nnz); lis_free(A->w_row); lis_free(A->w_index); lis_free(A->w_value);
    A->w_nnz = A->w_row = NULL; A->w_index = NULL; A->w_value = NULL;
    A->ptr = ptr; A->index = index; A->value = value;
    A->nnz = nnz;
    A->is_copy = LIS_TRUE;
    A->matrix_type = LIS_MATRIX_CSR;
    A->status = LIS_MATRIX_CSR;
    if (A->nprocs > 1) {
        err = lisd_matrix_g2l(A);
        if (err) return err;
        err = lisd_commtable_create(A);
        if (err) return err;
        A->is_comm = LIS_TRUE;
    }
    if (target_type != LIS_MATRIX_CSR) {
        LIS_MATRIX B;
        err = lis_matrix_duplicate(A, &B);
        if (err) return err;
        lis_matrix_set_type(B, target_type);
        err = lis_matrix_convert(A, B);
        if (err) { lis_matrix_destroy(B); return err; }
        lis_host_matrix_adopt(A, B);
    }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_assemble(LIS_MATRIX A)
{
    LIS_INT err = matrix_check(A, CHECK_SIZE);
    if (err) return err;
    if (A->status == LIS_MATRIX_ASSEMBLING) return assemble_from_rows(A, A->matrix_type);
    if (A->status < 0 && A->status > LIS_MATRIX_DECIDING_SIZE && A->n >= 0) {
        A->status = -A->status;
        A->matrix_type = A->status;
    } else if (A->status == LIS_MATRIX_NULL) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "matrix A has no entries\n");
        return LIS_ERR_ILL_ARG;
    }
    if (A->nprocs > 1 && !A->is_pmat) {
        if (A->l2g_map == NULL && !A->is_comm) {
            err = lisd_matrix_g2l(A);
            if (err) return err;
        }
        if (A->commtable == NULL) {
            err = lisd_commtable_create(A);
            if (err) return err;
            A->is_comm = LIS_TRUE;
        }
    }
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ array allocation + adoption
 * lis_matrix_malloc_* hand out tracked blocks; lis_matrix_set_* adopt the caller's arrays
 * (tracked or plain malloc) without copying -- src/matrix/lis_matrix_csr.c:60-110 etc. */
static void *tracked(size_t count, size_t size, char *tag)
{
    return lis_malloc((count > 0 ? count : 1) * size, tag);
}

LIS_INT lis_matrix_malloc_csr(LIS_INT n, LIS_INT nnz, LIS_INT **ptr, LIS_INT **index, LIS_SCALAR **value)
{
    *ptr = (LIS_INT *)tracked((size_t)n + 1, sizeof(LIS_INT), "lis_matrix_malloc_csr::ptr");
    *index = (LIS_INT *)tracked((size_t)nnz, sizeof(LIS_INT), "lis_matrix_malloc_csr::index");
    *value = (LIS_SCALAR *)tracked((size_t)nnz, sizeof(LIS_SCALAR), "lis_matrix_malloc_csr::value");
    if (!*ptr || !*index || !*value) {
        LIS_SETERR_MEM((size_t)nnz * sizeof(LIS_SCALAR));
        lis_free2(3, *ptr, *index, *value);
        *ptr = NULL; *index = NULL; *value = NULL;
        return LIS_OUT_OF_MEMORY;
    }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_malloc_csc(LIS_INT n, LIS_INT nnz, LIS_INT **ptr, LIS_INT **index, LIS_SCALAR **value)
{
    return lis_matrix_malloc_csr(n, nnz, ptr, index, value);
}

LIS_INT lis_matrix_malloc_ell(LIS_INT n, LIS_INT maxnzr, LIS_INT **index, LIS_SCALAR **value)
{
    const size_t cnt = (size_t)n * (size_t)maxnzr;
    *index = (LIS_INT *)tracked(cnt, sizeof(LIS_INT), "lis_matrix_malloc_ell::index");
    *value = (LIS_SCALAR *)tracked(cnt, sizeof(LIS_SCALAR), "lis_matrix_malloc_ell::value");
    if (!*index || !*value) {
        LIS_SETERR_MEM(cnt * sizeof(LIS_SCALAR));
        lis_free2(2, *index, *value);
        *index = NULL; *value = NULL;
        return LIS_OUT_OF_MEMORY;
    }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_malloc_dia(LIS_INT n, LIS_INT nnd, LIS_INT **index, LIS_SCALAR **value)
{
    const size_t cnt = (size_t)n * (size_t)nnd;
    *index = (LIS_INT *)tracked((size_t)nnd, sizeof(LIS_INT), "lis_matrix_malloc_dia::index");
    *value = (LIS_SCALAR *)tracked(cnt, sizeof(LIS_SCALAR), "lis_matrix_malloc_dia::value");
    if (!*index || !*value) {
        LIS_SETERR_MEM(cnt * sizeof(LIS_SCALAR));
        lis_free2(2, *index, *value);
        *index = NULL; *value = NULL;
        return LIS_OUT_OF_MEMORY;
    }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_malloc_jad(LIS_INT n, LIS_INT nnz, LIS_INT maxnzr, LIS_INT **perm, LIS_INT **ptr, LIS_INT **index, LIS_SCALAR **value)
{
    *perm = (LIS_INT *)tracked((size_t)n, sizeof(LIS_INT), "lis_matrix_malloc_jad::perm");
    *ptr = (LIS_INT *)tracked((size_t)maxnzr + 1, sizeof(LIS_INT), "lis_matrix_malloc_jad::ptr");
    *index = (LIS_INT *)tracked((size_t)nnz, sizeof(LIS_INT), "lis_matrix_malloc_jad::index");
    *value = (LIS_SCALAR *)tracked((size_t)nnz, sizeof(LIS_SCALAR), "lis_matrix_malloc_jad::value");
    if (!*perm || !*ptr || !*index || !*value) {
        LIS_SETERR_MEM((size_t)nnz * sizeof(LIS_SCALAR));
        lis_free2(4, *perm, *ptr, *index, *value);
        *perm = NULL; *ptr = NULL; *index = NULL; *value = NULL;
        return LIS_OUT_OF_MEMORY;
    }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_malloc_bsr(LIS_INT n, LIS_INT bnr, LIS_INT bnc, LIS_INT bnnz, LIS_INT **bptr, LIS_INT **bindex, LIS_SCALAR **value)
{
    const LIS_INT nr = n > 0 ? 1 + (n - 1) / bnr : 0;
    const size_t cnt = (size_t)bnnz * (size_t)bnr * (size_t)bnc;
    *bptr = (LIS_INT *)tracked((size_t)nr + 1, sizeof(LIS_INT), "lis_matrix_malloc_bsr::bptr");
    *bindex = (LIS_INT *)tracked((size_t)bnnz, sizeof(LIS_INT), "lis_matrix_malloc_bsr::bindex");
    *value = (LIS_SCALAR *)tracked(cnt, sizeof(LIS_SCALAR), "lis_matrix_malloc_bsr::value");
    if (!*bptr || !*bindex || !*value) {
        LIS_SETERR_MEM(cnt * sizeof(LIS_SCALAR));
        lis_free2(3, *bptr, *bindex, *value);
        *bptr = NULL; *bindex = NULL; *value = NULL;
        return LIS_OUT_OF_MEMORY;
    }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_csr(LIS_INT nnz, LIS_INT *ptr, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = matrix_check(A, CHECK_SET);
    if (err) return err;
    A->ptr = ptr; A->index = index; A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_CSR;
    A->nnz = nnz;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_csc(LIS_INT nnz, LIS_INT *ptr, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = matrix_check(A, CHECK_SET);
    if (err) return err;
    A->ptr = ptr; A->index = index; A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_CSC;
    A->nnz = nnz;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_ell(LIS_INT maxnzr, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = matrix_check(A, CHECK_SET);
    if (err) return err;
    A->index = index; A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_ELL;
    A->maxnzr = maxnzr;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_dia(LIS_INT nnd, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = matrix_check(A, CHECK_SET);
    if (err) return err;
    A->index = index; A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_DIA;
    A->nnd = nnd;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_jad(LIS_INT nnz, LIS_INT maxnzr, LIS_INT *perm, LIS_INT *ptr, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = matrix_check(A, CHECK_SET);
    if (err) return err;
    A->row = perm; A->ptr = ptr; A->index = index; A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_JAD;
    A->nnz = nnz;
    A->maxnzr = maxnzr;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_set_bsr(LIS_INT bnr, LIS_INT bnc, LIS_INT bnnz, LIS_INT *bptr, LIS_INT *bindex, LIS_SCALAR *value, LIS_MATRIX A)
{
    LIS_INT err = matrix_check(A, CHECK_SET);
    if (err) return err;
    if (bnr <= 0 || bnc <= 0) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "bnr=%D <= 0 or bnc=%D <= 0\n", bnr, bnc);
        return LIS_ERR_ILL_ARG;
    }
    A->bptr = bptr; A->bindex = bindex; A->value = value;
    A->is_copy = LIS_FALSE;
    A->status = -LIS_MATRIX_BSR;
    A->is_block = LIS_TRUE;
    A->bnnz = bnnz;
    A->nr = A->n > 0 ? 1 + (A->n - 1) / bnr : 0;
    A->nc = A->gn > 0 ? 1 + (A->gn - 1) / bnc : 0;
    if (A->np > A->n) A->nc = 1 + (A->np - 1) / bnc;
    A->bnr = bnr; A->bnc = bnc;
    return LIS_SUCCESS;
}

/* In-place update of one stored entry of an ASSEMBLED matrix (the reference's "psd" calls, for time-stepping codes
 * that keep the pattern and rewrite the numbers: src/matrix/lis_matrix.c:806-860, lis_matrix_csr.c:205-268): CSR
 * only, the first stored (i, j) of the row; an (i, j) that is not stored is silently left alone, as there.  The
 * device mirror is dropped and re-uploaded by the next product. */
LIS_INT lis_matrix_psd_set_value_csr(LIS_INT flag, LIS_INT i, LIS_INT j, LIS_SCALAR value, LIS_MATRIX A)
{
    const LIS_INT n = A->n, gn = A->gn, is = A->is, ie = A->ie;
    if (A->origin) { i--; j--; }
    if (i < 0 || j < 0) {
        LIS_SETERR3(LIS_ERR_ILL_ARG, "i(=%D) or j(=%D) are less than %D\n", i + A->origin, j + A->origin, A->origin);
        return LIS_ERR_ILL_ARG;
    }
    if (i >= gn || j >= gn) {
        LIS_SETERR3(LIS_ERR_ILL_ARG, "i(=%D) or j(=%D) are larger than global n=(%D)\n", i + A->origin, j + A->origin, gn);
        return LIS_ERR_ILL_ARG;
    }
    if (i < is || i >= ie) {
        LIS_SETERR3(LIS_ERR_ILL_ARG, "row i(=%D) is outside the local range [%D,%D)\n", i + A->origin, is, ie);
        return LIS_ERR_ILL_ARG;
    }
    for (LIS_INT k = A->ptr[i - is]; k < A->ptr[i - is + 1]; k++) {
        const LIS_INT c = A->index[k];
        const LIS_INT jg = c < n ? c + is : (A->l2g_map ? A->l2g_map[c - n] : c);
        if (jg == j) {
            if (flag == LIS_INS_VALUE) A->value[k] = value; else A->value[k] += value;
            lisd_matrix_drop(A);
            break;
        }
    }
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_psd_set_value(LIS_INT flag, LIS_INT i, LIS_INT j, LIS_SCALAR value, LIS_MATRIX A)
{
    LIS_INT err = matrix_check(A, CHECK_SIZE);
    if (err) return err;
    if (A->status == LIS_MATRIX_CSR && !A->is_splited) return lis_matrix_psd_set_value_csr(flag, i, j, value, A);
    if (A->status > 0) { LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED; }             /* assembled, another format */
    return lis_matrix_set_value(flag, i, j, value, A);                                   /* not assembled yet: the ordinary path */
}

LIS_INT lis_matrix_psd_reset_scale(LIS_MATRIX A) { A->is_scaled = LIS_FALSE; return LIS_SUCCESS; }

/* the block partition lis_matrix_convert derives for VBR when the caller gave none */
LIS_INT lis_matrix_get_vbr_rowcol(LIS_MATRIX Ain, LIS_INT *nr, LIS_INT *nc, LIS_INT **row, LIS_INT **col)
{
    LIS_INT err = lis_host_matrix_check_input(Ain);
    if (err) return err;
    if (Ain->matrix_type != LIS_MATRIX_CSR) { LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED; }
    err = lis_host_vbr_partition(Ain, nr, row, col);
    if (!err) *nc = *nr;
    return err;
}

/* ------------------------------------------------------------------ CSR utilities */
/* a dense n x n block given row by row, entry by entry through lis_matrix_set_value
 * (src/matrix/lis_matrix.c:859-875) */
LIS_INT lis_matrix_set_values(LIS_INT flag, LIS_INT n, LIS_SCALAR value[], LIS_MATRIX A)
{
    for (LIS_INT i = 0; i < n; i++)
        for (LIS_INT j = 0; j < n; j++) lis_matrix_set_value(flag, i, j, value[(size_t)i * n + j], A);
    return LIS_SUCCESS;
}

/* the factors lis_matrix_scale left in Dv (1/d or 1/sqrt|d|), applied to another CSR matrix with the same rows -- the
 * private split copy the sweeps of SSOR/GS/SOR run on under -storage <scalar format> (host/lis_precon.c create_ssor) */
LIS_INT lis_host_matrix_scale_like(LIS_MATRIX C, LIS_VECTOR Dv, LIS_INT action)
{
    const LIS_INT n = C->n;
    if (C->matrix_type != LIS_MATRIX_CSR || C->nprocs > 1) { LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED; }
    LIS_SCALAR *d = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(n > 0 ? n : 1));
    if (d == NULL) { LIS_SETERR_MEM(n * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    LIS_INT err = n > 0 ? lis_vector_get_values(Dv, Dv->is + Dv->origin, n, d) : LIS_SUCCESS;
    if (err) { free(d); return err; }
    const int symm = action == LIS_SCALE_SYMM_DIAG;
    if (C->is_splited) {
        for (LIS_INT i = 0; i < n; i++) {
            C->D->value[i] = 1.0;
            for (LIS_INT j = C->L->ptr[i]; j < C->L->ptr[i + 1]; j++) { if (symm) C->L->value[j] = C->L->value[j] * d[i] * d[C->L->index[j]]; else C->L->value[j] *= d[i]; }
            for (LIS_INT j = C->U->ptr[i]; j < C->U->ptr[i + 1]; j++) { if (symm) C->U->value[j] = C->U->value[j] * d[i] * d[C->U->index[j]]; else C->U->value[j] *= d[i]; }
        }
    } else {
        for (LIS_INT i = 0; i < n; i++)
            for (LIS_INT j = C->ptr[i]; j < C->ptr[i + 1]; j++) { if (symm) C->value[j] = C->value[j] * d[i] * d[C->index[j]]; else C->value[j] *= d[i]; }
    }
    free(d);
    lisd_matrix_drop(C);                            /* mirror and sweep schedule are rebuilt from the scaled arrays */
    C->is_scaled = LIS_TRUE;
    return LIS_SUCCESS;
}

/* -scale: A <- D^-1 A, b <- D^-1 b (LIS_SCALE_JACOBI) or A <- D^-1/2 A D^-1/2, b <- D^-1/2 b
 * (LIS_SCALE_SYMM_DIAG), D = diag(A); the scaling vector stays in Dv (src/matrix/lis_matrix_ops.c:579-712,
 * CSR loops src/matrix/lis_matrix_csr.c:607-693).  Once per solve, on the host arrays the caller sees
 * (they stay scaled, A->is_scaled, like in the reference); the device mirror is dropped.  One process; CSR (also split),
 * and unsplit CSC / ELL / DIA / JAD / BSR. */
LIS_INT lis_matrix_scale(LIS_MATRIX A, LIS_VECTOR B, LIS_VECTOR Dv, LIS_INT action)
{
    const LIS_INT n = A->n;
    LIS_INT err = lis_host_matrix_check_input(A);
    if (err) return err;
    const LIS_INT mt = A->matrix_type;
    const int other_fmt = mt == LIS_MATRIX_ELL || mt == LIS_MATRIX_DIA || mt == LIS_MATRIX_JAD || mt == LIS_MATRIX_BSR || mt == LIS_MATRIX_CSC;
    if ((mt != LIS_MATRIX_CSR && !(other_fmt && !A->is_splited)) || (A->nprocs > 1 && mt != LIS_MATRIX_CSR)) {
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "-scale needs a CSR, CSC, ELL, DIA, JAD or BSR matrix (row-partitioned: CSR)\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    if (action != LIS_SCALE_JACOBI && action != LIS_SCALE_SYMM_DIAG) { LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED; }
    err = lis_matrix_get_diagonal(A, Dv);
    if (err) return err;
    const LIS_INT np = A->np > n ? A->np : n;
    LIS_SCALAR *d = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(np > 0 ? np : 1));
    if (d == NULL) { LIS_SETERR_MEM(np * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    err = n > 0 ? lis_vector_get_values(Dv, Dv->is + Dv->origin, n, d) : LIS_SUCCESS;
    if (err) { free(d); return err; }
    if (A->nprocs > 1 && action == LIS_SCALE_SYMM_DIAG) {
        /* row-partitioned: the columns of the halo need the diagonal entries their owners hold -- one halo exchange
         * of the diagonal, like lis_send_recv(A->commtable, d) at src/matrix/lis_matrix_ops.c:604 (collective) */
        double *d_tmp = NULL;
        err = lisd_malloc((void **)&d_tmp, sizeof(double) * (size_t)(np > 0 ? np : 1));
        if (!err && n > 0) err = lisd_upload(d_tmp, d, sizeof(double) * (size_t)n);
        if (!err) err = lisd_halo_exchange_raw(A, d_tmp);
        if (!err) err = lisd_sync();
        if (!err && np > n) err = lisd_download(d + n, d_tmp + n, sizeof(double) * (size_t)(np - n));
        if (d_tmp) lisd_free(d_tmp);
        if (err) { free(d); return err; }
        for (LIS_INT i = n; i < np; i++) d[i] = 1.0 / sqrt(fabs(d[i]));
    }
    if (other_fmt) {
        /* the other formats, each with the expression of its reference loop (the order of the multiplications differs
         * between them): lis_matrix_scale[_symm]_{ell,dia,jad,bsr,csc}, src/matrix/lis_matrix_<fmt>.c */
        const int symm = action == LIS_SCALE_SYMM_DIAG;
        for (LIS_INT i = 0; i < n; i++) d[i] = symm ? 1.0 / sqrt(fabs(d[i])) : 1.0 / d[i];
        if (mt == LIS_MATRIX_ELL) {
            for (LIS_INT j = 0; j < A->maxnzr; j++)
                for (LIS_INT i = 0; i < n; i++) {
                    const size_t o = (size_t)j * n + i;
                    if (symm) A->value[o] *= d[i] * d[A->index[o]]; else A->value[o] *= d[i];
                }
        } else if (mt == LIS_MATRIX_DIA) {
            for (LIS_INT j = 0; j < A->nnd; j++) {
                const LIS_INT jj = A->index[j], is = jj < 0 ? -jj : 0, ie = jj > 0 ? n - jj : n;
                for (LIS_INT i = is; i < ie; i++) {
                    const size_t o = (size_t)j * n + i;
                    if (symm) A->value[o] *= d[i] * d[i + jj]; else A->value[o] *= d[i];
                }
            }
        } else if (mt == LIS_MATRIX_JAD) {
            for (LIS_INT j = 0; j < A->maxnzr; j++) {
                LIS_INT k = 0;
                for (LIS_INT i = A->ptr[j]; i < A->ptr[j + 1]; i++, k++) {
                    if (symm) A->value[i] *= d[A->row[k]] * d[A->index[i]]; else A->value[i] *= d[A->row[k]];
                }
            }
        } else if (mt == LIS_MATRIX_BSR) {
            const LIS_INT bnr = A->bnr, bnc = A->bnc, bs = bnr * bnc;
            for (LIS_INT bi = 0; bi < A->nr; bi++)
                for (LIS_INT bj = A->bptr[bi]; bj < A->bptr[bi + 1]; bj++) {
                    const LIS_INT bjj = A->bindex[bj];
                    for (LIS_INT j = 0; j < bnc; j++)
                        for (LIS_INT i = 0; i < bnr; i++) {
                            const LIS_INT r = bi * bnr + i, c = bjj * bnc + j;
                            if (r >= n || c >= n) continue;         /* padding of the last block row / column: structural zeros */
                            const size_t o = (size_t)bj * bs + (size_t)j * bnr + i;
                            if (symm) A->value[o] *= d[r] * d[c]; else A->value[o] *= d[r];
                        }
                }
        } else {                                                    /* CSC */
            for (LIS_INT i = 0; i < n; i++)
                for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++) {
                    if (symm) A->value[j] = A->value[j] * d[i] * d[A->index[j]]; else A->value[j] *= d[A->index[j]];
                }
        }
    } else if (action == LIS_SCALE_SYMM_DIAG) {
        for (LIS_INT i = 0; i < n; i++) d[i] = 1.0 / sqrt(fabs(d[i]));
        if (A->is_splited) {
            for (LIS_INT i = 0; i < n; i++) {
                A->D->value[i] = 1.0;
                for (LIS_INT j = A->L->ptr[i]; j < A->L->ptr[i + 1]; j++) A->L->value[j] = A->L->value[j] * d[i] * d[A->L->index[j]];
                for (LIS_INT j = A->U->ptr[i]; j < A->U->ptr[i + 1]; j++) A->U->value[j] = A->U->value[j] * d[i] * d[A->U->index[j]];
            }
        } else {
            for (LIS_INT i = 0; i < n; i++)
                for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++) A->value[j] = A->value[j] * d[i] * d[A->index[j]];
        }
    } else {
        for (LIS_INT i = 0; i < n; i++) d[i] = 1.0 / d[i];
        if (A->is_splited) {
            for (LIS_INT i = 0; i < n; i++) {
                A->D->value[i] = 1.0;
                for (LIS_INT j = A->L->ptr[i]; j < A->L->ptr[i + 1]; j++) A->L->value[j] *= d[i];
                for (LIS_INT j = A->U->ptr[i]; j < A->U->ptr[i + 1]; j++) A->U->value[j] *= d[i];
            }
        } else {
            for (LIS_INT i = 0; i < n; i++)
                for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++) A->value[j] *= d[i];
        }
    }
    lisd_matrix_drop(A);
    if (n > 0) err = lis_vector_set_values2(LIS_INS_VALUE, Dv->is + Dv->origin, n, d, Dv);
    free(d);
    if (!err) err = lis_vector_pmul(B, Dv, B);                 /* b[i] = b[i]*d[i] */
    if (err) return err;
    A->is_scaled = LIS_TRUE;
    B->is_scaled = LIS_TRUE;
    return LIS_SUCCESS;
}

/* A <- A - sigma*I on the host arrays (src/matrix/lis_matrix_ops.c:780-830; per format
 * lis_matrix_csr.c:565-603, lis_matrix_csc.c, lis_matrix_ell.c, lis_matrix_dia.c): the first stored
 * diagonal entry of each row; a row without one is left alone.  A CSR / split device mirror gets the
 * same edit by a kernel, other mirrors are dropped. */
LIS_INT lis_matrix_shift_diagonal(LIS_MATRIX A, LIS_SCALAR sigma)
{
    const LIS_INT n = A->n;
    LIS_INT err = lis_host_matrix_check_input(A);
    if (err) return err;
    if (A->is_splited) {
        if (A->matrix_type != LIS_MATRIX_CSR) { LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED; }
        for (LIS_INT i = 0; i < n; i++) A->D->value[i] -= sigma;
    } else switch (A->matrix_type) {
    case LIS_MATRIX_CSR:
    case LIS_MATRIX_CSC:
        for (LIS_INT i = 0; i < n; i++)
            for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++)
                if (A->index[j] == i) { A->value[j] -= sigma; break; }
        break;
    case LIS_MATRIX_ELL:
        for (LIS_INT i = 0; i < n; i++)
            for (LIS_INT j = 0; j < A->maxnzr; j++)
                if (A->index[(size_t)j * n + i] == i) { A->value[(size_t)j * n + i] -= sigma; break; }
        break;
    case LIS_MATRIX_DIA:
        for (LIS_INT j = 0; j < A->nnd; j++)
            if (A->index[j] == 0) { for (LIS_INT i = 0; i < n; i++) A->value[(size_t)j * n + i] -= sigma; break; }
        break;
    default: {
        int handled = 0;                                   /* MSR, JAD, BSR, BSC, VBR, COO, DNS: lis_formats_ext.c */
        lis_host_ext_shift_diagonal(A, sigma, &handled);
        if (!handled) {
            LIS_SETERR1(LIS_ERR_NOT_IMPLEMENTED, "lis_matrix_shift_diagonal: storage format %D\n", A->matrix_type);
            return LIS_ERR_NOT_IMPLEMENTED;
        }
    }
    }
    return lisd_matrix_shift_diagonal(A, sigma);
}

/* every row ascending by column: lis_matrix_sort_csr, src/matrix/lis_matrix_csr.c:1486-1521 */
LIS_INT lis_matrix_sort_csr(LIS_MATRIX A)
{
    if (!A->is_sorted) {
        for (LIS_INT i = 0; i < A->n; i++) lis_sort_id(A->ptr[i], A->ptr[i + 1] - 1, A->index, A->value);
        A->is_sorted = LIS_TRUE;
        lisd_matrix_drop(A);
    }
    return LIS_SUCCESS;
}

/* D (scalar diagonal) shell with n zero entries: lis_matrix_diag_duplicateM */
static LIS_INT diag_create(LIS_MATRIX A, LIS_MATRIX_DIAG *Dout)
{
    LIS_MATRIX_DIAG D = (LIS_MATRIX_DIAG)lis_calloc(sizeof(struct LIS_MATRIX_DIAG_STRUCT), "lis_matrix_diag::D");
    if (D == NULL) { LIS_SETERR_MEM(sizeof(struct LIS_MATRIX_DIAG_STRUCT)); return LIS_OUT_OF_MEMORY; }
    D->value = (LIS_SCALAR *)lis_calloc((size_t)(A->n > 0 ? A->n : 1) * sizeof(LIS_SCALAR), "lis_matrix_diag::value");
    if (D->value == NULL) { lis_free(D); LIS_SETERR_MEM(A->n * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    D->label = LIS_LABEL_MATRIX;
    D->n = A->n; D->gn = A->gn; D->np = A->np; D->nr = A->n; D->bn = 1;
    D->is = A->is; D->ie = A->ie; D->my_rank = A->my_rank; D->nprocs = A->nprocs; D->comm = A->comm;
    *Dout = D;
    return LIS_SUCCESS;
}

LIS_INT lis_host_diag_create(LIS_MATRIX A, LIS_MATRIX_DIAG *Dout) { return diag_create(A, Dout); }

/* D, L, U of a CSR matrix (src/matrix/lis_matrix_csr.c:764-949, serial branch): storage order kept inside L and U,
 * the last diagonal entry of a row wins, halo columns (>= n) land in U.  Two passes over the rows on the host worker
 * threads -- count, (serial prefix sums), fill. */
typedef struct { LIS_MATRIX A; LIS_INT *lp, *up; LIS_MATRIX_CORE L, U; LIS_MATRIX_DIAG D; } split_ctx;

static void split_count_rows(size_t r0, size_t r1, void *ctx)
{
    split_ctx *c = (split_ctx *)ctx;
    const LIS_MATRIX A = c->A;
    for (size_t r = r0; r < r1; r++) {
        const LIS_INT i = (LIS_INT)r;
        LIS_INT nl = 0, nu = 0;
        for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++) {
            if (A->index[j] < i) nl++;
            else if (A->index[j] > i) nu++;
        }
        c->lp[i + 1] = nl; c->up[i + 1] = nu;
    }
}

static void split_fill_rows(size_t r0, size_t r1, void *ctx)
{
    split_ctx *c = (split_ctx *)ctx;
    const LIS_MATRIX A = c->A;
    for (size_t r = r0; r < r1; r++) {
        const LIS_INT i = (LIS_INT)r;
        LIS_INT kl = c->L->ptr[i], ku = c->U->ptr[i];
        for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++) {
            const LIS_INT col = A->index[j];
            if (col < i) { c->L->index[kl] = col; c->L->value[kl] = A->value[j]; kl++; }
            else if (col > i) { c->U->index[ku] = col; c->U->value[ku] = A->value[j]; ku++; }
            else c->D->value[i] = A->value[j];
        }
    }
}

static LIS_INT split_csr(LIS_MATRIX A)
{
    const LIS_INT n = A->n;
    LIS_INT err;
    split_ctx c = { A, NULL, NULL, NULL, NULL, NULL };
    c.lp = (LIS_INT *)malloc(sizeof(LIS_INT) * ((size_t)n + 1));
    c.up = (LIS_INT *)malloc(sizeof(LIS_INT) * ((size_t)n + 1));
    if (!c.lp || !c.up) { free(c.lp); free(c.up); LIS_SETERR_MEM(n * sizeof(LIS_INT)); return LIS_OUT_OF_MEMORY; }
    c.lp[0] = 0; c.up[0] = 0;
    lis_host_parallel_for((size_t)n, 65536, split_count_rows, &c);
    for (LIS_INT i = 0; i < n; i++) { c.lp[i + 1] += c.lp[i]; c.up[i + 1] += c.up[i]; }
    const LIS_INT nnzl = c.lp[n], nnzu = c.up[n];
    LIS_MATRIX_CORE L = (LIS_MATRIX_CORE)lis_calloc(sizeof(struct LIS_MATRIX_CORE_STRUCT), "lis_matrix_split::L");
    LIS_MATRIX_CORE U = (LIS_MATRIX_CORE)lis_calloc(sizeof(struct LIS_MATRIX_CORE_STRUCT), "lis_matrix_split::U");
    LIS_MATRIX_DIAG D = NULL;
    if (!L || !U) { lis_free2(2, L, U); free(c.lp); free(c.up); LIS_SETERR_MEM(sizeof(struct LIS_MATRIX_CORE_STRUCT)); return LIS_OUT_OF_MEMORY; }
    err = lis_matrix_malloc_csr(n, nnzl, &L->ptr, &L->index, &L->value);
    if (!err) err = lis_matrix_malloc_csr(n, nnzu, &U->ptr, &U->index, &U->value);
    if (!err) err = diag_create(A, &D);
    if (err) { core_destroy(L); core_destroy(U); free(c.lp); free(c.up); return err; }
    memcpy(L->ptr, c.lp, sizeof(LIS_INT) * ((size_t)n + 1));
    memcpy(U->ptr, c.up, sizeof(LIS_INT) * ((size_t)n + 1));
    free(c.lp); free(c.up);
    c.L = L; c.U = U; c.D = D;
    lis_host_parallel_for((size_t)n, 65536, split_fill_rows, &c);
    L->nnz = nnzl; U->nnz = nnzu;
    A->L = L; A->U = U; A->D = D;
    return LIS_SUCCESS;
}

LIS_INT lis_matrix_split(LIS_MATRIX A)
{
    if (A->is_splited) return LIS_SUCCESS;
    if (A->matrix_type != LIS_MATRIX_CSR) {
        /* the reference splits every format; the B200 hot path keeps the triangular sweep in CSR */
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "lis_matrix_split: only CSR is supported (use -storage csr with SSOR)\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    LIS_INT err = split_csr(A);
    if (err) return err;
    A->is_splited = LIS_TRUE;
    lisd_matrix_drop(A);
    return LIS_SUCCESS;
}

/* src/matrix/lis_matrix_csr.c:1257-1310: rows become [L entries, diagonal, U entries] */
LIS_INT lis_matrix_merge(LIS_MATRIX A)
{
    if (!A->is_splited || (A->is_save && A->is_splited)) return LIS_SUCCESS;
    if (A->matrix_type != LIS_MATRIX_CSR) { LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED; }
    const LIS_INT n = A->n;
    LIS_INT nnz = A->L->nnz + A->U->nnz + n, err;
    LIS_INT *ptr, *index;
    LIS_SCALAR *value;
    err = lis_matrix_malloc_csr(n, nnz, &ptr, &index, &value);
    if (err) return err;
    nnz = 0; ptr[0] = 0;
    for (LIS_INT i = 0; i < n; i++) {
        for (LIS_INT j = A->L->ptr[i]; j < A->L->ptr[i + 1]; j++) { index[nnz] = A->L->index[j]; value[nnz] = A->L->value[j]; nnz++; }
        index[nnz] = i; value[nnz] = A->D->value[i]; nnz++;
        for (LIS_INT j = A->U->ptr[i]; j < A->U->ptr[i + 1]; j++) { index[nnz] = A->U->index[j]; value[nnz] = A->U->value[j]; nnz++; }
        ptr[i + 1] = nnz;
    }
    if (A->is_destroy) lis_free2(3, A->ptr, A->index, A->value);
    A->nnz = nnz; A->ptr = ptr; A->index = index; A->value = value;
    A->is_sorted = LIS_FALSE;
    lis_matrix_DLU_destroy(A);          /* also drops the device mirror */
    if (A->WD) { lis_matrix_diag_destroy(A->WD); A->WD = NULL; A->use_wd = 0; }
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ diagonal
 * d[i] = A(i,i): first stored (i,i) entry, 0 when absent; D->value when split
 * (src/matrix/lis_matrix_ops.c:727-780 and the per-format get_diagonal loops). */
static LIS_INT host_get_diagonal(LIS_MATRIX A, LIS_SCALAR *v);
LIS_INT lis_matrix_get_diagonal(LIS_MATRIX A, LIS_VECTOR d)
{
    LIS_INT err = matrix_check(A, CHECK_ALL);
    if (err) return err;
    if (A->n != d->n) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "length of matrix A and vector d is not equal\n");
        return LIS_ERR_ILL_ARG;
    }
    const LIS_INT n = A->n;
    if (A->matrix_type == LIS_MATRIX_CSR && !A->is_splited && lisd_available()) {
        lisd_matrix *M;
        err = lisd_matrix_get(A, &M);
        if (err) return err;
        err = lisd_vec_device(d);
        if (err) return err;
        lisd_mark_busy();
        err = lisd_check(lisb200_csr_get_diagonal(n, M->csr.ptr, M->csr.idx, M->csr.val, d->value, lisd_stream()),
                         "lis_matrix_get_diagonal");
        if (err) return err;
        return lisd_sync();
    }
    /* other formats: a one-off host pass over the caller-visible arrays */
    LIS_SCALAR *v = lisd_vec_host_view(d, 0);
    if (v == NULL) return LIS_ERR_OUT_OF_MEMORY;
    err = host_get_diagonal(A, v);
    const LIS_INT err2 = lisd_vec_host_done(d, v, err == LIS_SUCCESS);
    return err ? err : err2;
}

static LIS_INT host_get_diagonal(LIS_MATRIX A, LIS_SCALAR *v)
{
    const LIS_INT n = A->n;
    if (A->is_splited) {
        for (LIS_INT i = 0; i < n; i++) v[i] = A->D->value[i];
        return LIS_SUCCESS;
    }
    for (LIS_INT i = 0; i < n; i++) v[i] = 0.0;
    switch (A->matrix_type) {
    case LIS_MATRIX_CSR:
        for (LIS_INT i = 0; i < n; i++)
            for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++)
                if (A->index[j] == i) { v[i] = A->value[j]; break; }
        break;
    case LIS_MATRIX_CSC:
        for (LIS_INT i = 0; i < n; i++)
            for (LIS_INT j = A->ptr[i]; j < A->ptr[i + 1]; j++)
                if (A->index[j] == i) { v[i] = A->value[j]; break; }
        break;
    case LIS_MATRIX_ELL:
        for (LIS_INT i = 0; i < n; i++)
            for (LIS_INT j = 0; j < A->maxnzr; j++)
                if (A->index[(size_t)j * n + i] == i) { v[i] = A->value[(size_t)j * n + i]; break; }
        break;
    case LIS_MATRIX_DIA:
        for (LIS_INT j = 0; j < A->nnd; j++)
            if (A->index[j] == 0) { memcpy(v, A->value + (size_t)j * n, (size_t)n * sizeof(LIS_SCALAR)); break; }
        break;
    case LIS_MATRIX_JAD:
        for (LIS_INT i = 0; i < n; i++) {
            const LIS_INT r = A->row[i];
            for (LIS_INT j = 0; j < A->maxnzr; j++) {
                if (i >= A->ptr[j + 1] - A->ptr[j]) break;
                const LIS_INT k = A->ptr[j] + i;
                if (A->index[k] == r) { v[r] = A->value[k]; break; }
            }
        }
        break;
    case LIS_MATRIX_BSR: {
        const LIS_INT bnr = A->bnr, bnc = A->bnc, bs = bnr * bnc;
        for (LIS_INT bi = 0; bi < A->nr; bi++)
            for (LIS_INT ii = 0; ii < bnr && bi * bnr + ii < n; ii++) {
                const LIS_INT r = bi * bnr + ii;
                for (LIS_INT bc = A->bptr[bi]; bc < A->bptr[bi + 1]; bc++) {
                    const LIS_INT c0 = A->bindex[bc] * bnc;
                    if (r >= c0 && r < c0 + bnc) { v[r] = A->value[(size_t)bc * bs + (size_t)(r - c0) * bnr + ii]; break; }
                }
            }
        break;
    }
    default: {
        int handled = 0;                                   /* MSR, COO, BSC, VBR, DNS */
        LIS_INT err = lis_host_ext_get_diagonal(A, v, &handled);
        if (handled) return err;
        LIS_SETERR_IMP;
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    }
    return LIS_SUCCESS;
}
