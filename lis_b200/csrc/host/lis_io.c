/*
 * lis_io.c -- the on-disk formats either side of the hot path that the reference drivers
 * touch: Matrix Market input (with Lis' extension that appends b and x to the file,
 * src/system/lis_input_mm.c:61-1069, used by test/test1.c), vector input, and the
 * solution / matrix writers (src/system/lis_output.c:200-440).  Host C, ASCII only.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include "lis_device.h"
#include "lis_host.h"

#define LINE_MAX_LEN 1024

static void lower(char *s) { for (; *s; s++) *s = (char)tolower((unsigned char)*s); }

/* banner: %%MatrixMarket <matrix|vector> <coordinate|array> <real|integer|pattern> <general|symmetric> */
typedef struct { int is_vector, is_array, is_pattern, is_symm; } mm_banner_t;

static LIS_INT read_banner(FILE *f, mm_banner_t *bn)
{
    char buf[LINE_MAX_LEN], w0[64] = "", w1[64] = "", w2[64] = "", w3[64] = "", w4[64] = "";
    memset(bn, 0, sizeof(*bn));
    if (fgets(buf, sizeof(buf), f) == NULL) { LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }
    sscanf(buf, "%63s %63s %63s %63s %63s", w0, w1, w2, w3, w4);
    lower(w0); lower(w1); lower(w2); lower(w3); lower(w4);
    if (strcmp(w0, "%%matrixmarket") != 0) { LIS_SETERR(LIS_ERR_FILE_IO, "Not Matrix Market banner\n"); return LIS_ERR_FILE_IO; }
    bn->is_vector = strcmp(w1, "vector") == 0;
    if (!bn->is_vector && strcmp(w1, "matrix") != 0) { LIS_SETERR(LIS_ERR_FILE_IO, "Not Matrix Market format (matrix)\n"); return LIS_ERR_FILE_IO; }
    bn->is_array = strcmp(w2, "array") == 0;
    if (!bn->is_array && strcmp(w2, "coordinate") != 0) { LIS_SETERR(LIS_ERR_FILE_IO, "Not Coodinate or Array format\n"); return LIS_ERR_FILE_IO; }
    if (strcmp(w3, "complex") == 0) { LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "complex Matrix Market files are not supported\n"); return LIS_ERR_NOT_IMPLEMENTED; }
    bn->is_pattern = strcmp(w3, "pattern") == 0;
    bn->is_symm = strcmp(w4, "symmetric") == 0;
    if (!bn->is_symm && strcmp(w4, "general") != 0 && w4[0]) { LIS_SETERR(LIS_ERR_FILE_IO, "Not general or symmetric\n"); return LIS_ERR_FILE_IO; }
    return LIS_SUCCESS;
}

static LIS_INT next_data_line(FILE *f, char *buf, size_t cap)
{
    do {
        if (fgets(buf, (int)cap, f) == NULL) { LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }
    } while (buf[0] == '%' || buf[0] == '\n' || buf[0] == '\r');
    return LIS_SUCCESS;
}

/* n lines "i value" (1-based) into the locally owned slice of v */
static LIS_INT read_mm_vec_body(FILE *f, LIS_VECTOR v, LIS_INT gn)
{
    char buf[LINE_MAX_LEN];
    lisd_vec_host(v);
    for (LIS_INT i = 0; i < gn; i++) {
        int idx; double val;
        if (fgets(buf, sizeof(buf), f) == NULL || sscanf(buf, "%d %lg", &idx, &val) != 2) { LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }
        idx--;
        if (idx >= v->is && idx < v->ie) v->value[idx - v->is] = val;
    }
    return LIS_SUCCESS;
}

static LIS_INT input_mm(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, FILE *f)
{
    mm_banner_t bn;
    char buf[LINE_MAX_LEN];
    const LIS_INT want_type = A->matrix_type;
    LIS_INT err = read_banner(f, &bn);
    if (err) return err;
    if (bn.is_vector || bn.is_array) { LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "only coordinate matrices are supported\n"); return LIS_ERR_NOT_IMPLEMENTED; }
    err = next_data_line(f, buf, sizeof(buf));
    if (err) return err;
    int nr = 0, nc = 0, nnz = 0, isb = 0, isx = 0;
    const int got = sscanf(buf, "%d %d %d %d %d", &nr, &nc, &nnz, &isb, &isx);
    if (got != 3 && got != 5) { LIS_SETERR(LIS_ERR_FILE_IO, "matrix size line is not correct\n"); return LIS_ERR_FILE_IO; }
    if (nr != nc) { LIS_SETERR(LIS_ERR_FILE_IO, "matrix is not square\n"); return LIS_ERR_FILE_IO; }
    err = lis_matrix_set_size(A, 0, nr);
    if (err) return err;
    if (A->my_rank == 0) printf("matrix size = %d x %d (%d nonzero entries)\n\n", nr, nc, nnz);
    const LIS_INT n = A->n, is = A->is, ie = A->ie;

    int *ri = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    int *ci = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
    double *va = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    LIS_INT *ptr = NULL, *index = NULL, *fill = NULL;
    LIS_SCALAR *value = NULL;
    if (!ri || !ci || !va) { err = LIS_OUT_OF_MEMORY; LIS_SETERR_MEM(nnz); goto done; }
    for (int k = 0; k < nnz; k++) {
        double v = 1.0;
        if (fgets(buf, sizeof(buf), f) == NULL) { LIS_SETERR_FIO; err = LIS_ERR_FILE_IO; goto done; }
        const int c = bn.is_pattern ? sscanf(buf, "%d %d", &ri[k], &ci[k]) + 1 : sscanf(buf, "%d %d %lg", &ri[k], &ci[k], &v);
        if (c != 3) { LIS_SETERR_FIO; err = LIS_ERR_FILE_IO; goto done; }
        ri[k]--; ci[k]--; va[k] = v;
        if (ri[k] < 0 || ri[k] >= nr || ci[k] < 0 || ci[k] >= nr) { LIS_SETERR(LIS_ERR_FILE_IO, "index out of range\n"); err = LIS_ERR_FILE_IO; goto done; }
    }
    /* rows in file order; a symmetric file contributes the mirrored entry too */
    ptr = (LIS_INT *)lis_calloc(((size_t)n + 1) * sizeof(LIS_INT), "lis_input_mm::ptr");
    fill = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
    if (!ptr || !fill) { err = LIS_OUT_OF_MEMORY; LIS_SETERR_MEM(n); goto done; }
    for (int k = 0; k < nnz; k++) {
        if (ri[k] >= is && ri[k] < ie) ptr[ri[k] - is + 1]++;
        if (bn.is_symm && ri[k] != ci[k] && ci[k] >= is && ci[k] < ie) ptr[ci[k] - is + 1]++;
    }
    for (LIS_INT i = 0; i < n; i++) { ptr[i + 1] += ptr[i]; fill[i] = ptr[i]; }
    index = (LIS_INT *)lis_malloc((size_t)(ptr[n] > 0 ? ptr[n] : 1) * sizeof(LIS_INT), "lis_input_mm::index");
    value = (LIS_SCALAR *)lis_malloc((size_t)(ptr[n] > 0 ? ptr[n] : 1) * sizeof(LIS_SCALAR), "lis_input_mm::value");
    if (!index || !value) { err = LIS_OUT_OF_MEMORY; LIS_SETERR_MEM(ptr[n]); goto done; }
    for (int k = 0; k < nnz; k++) {
        if (ri[k] >= is && ri[k] < ie) { const LIS_INT l = fill[ri[k] - is]++; index[l] = ci[k]; value[l] = va[k]; }
        if (bn.is_symm && ri[k] != ci[k] && ci[k] >= is && ci[k] < ie) { const LIS_INT l = fill[ci[k] - is]++; index[l] = ri[k]; value[l] = va[k]; }
    }
    err = lis_matrix_set_csr(ptr[n], ptr, index, value, A);
    if (err) goto done;
    ptr = NULL; index = NULL; value = NULL;              /* adopted */
    err = lis_matrix_assemble(A);
    if (err) goto done;
    /* optional right-hand side and initial guess appended to the file */
    if (isb && b) {
        if (lis_vector_is_null(b)) { err = lis_vector_set_size(b, A->n, 0); if (err) goto done; }
        err = read_mm_vec_body(f, b, nr);
        if (err) goto done;
    } else if (isb) {
        for (int k = 0; k < nr; k++) if (fgets(buf, sizeof(buf), f) == NULL) break;
    }
    if (isx && x) {
        if (lis_vector_is_null(x)) { err = lis_vector_set_size(x, A->n, 0); if (err) goto done; }
        err = read_mm_vec_body(f, x, nr);
        if (err) goto done;
    }
    if (want_type != LIS_MATRIX_CSR) {
        LIS_MATRIX B;
        err = lis_matrix_duplicate(A, &B);
        if (err) goto done;
        lis_matrix_set_type(B, want_type);
        err = lis_matrix_convert(A, B);
        if (err) { lis_matrix_destroy(B); goto done; }
        lis_host_matrix_adopt(A, B);
    }
done:
    free(ri); free(ci); free(va); free(fill);
    if (err) lis_free2(3, ptr, index, value);
    return err;
}

LIS_INT lis_input(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, char *filename)
{
    if (!lis_is_malloc(A)) { LIS_SETERR(LIS_ERR_ILL_ARG, "matrix A is undefined\n"); return LIS_ERR_ILL_ARG; }
    if (A->status != LIS_MATRIX_DECIDING_SIZE && A->status != LIS_MATRIX_NULL) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "matrix A has already been assigned\n");
        return LIS_ERR_ILL_ARG;
    }
    FILE *f = fopen(filename, "r");
    if (f == NULL) { LIS_SETERR1(LIS_ERR_FILE_IO, "cannot open file %s\n", filename); return LIS_ERR_FILE_IO; }
    char buf[LINE_MAX_LEN];
    if (fgets(buf, sizeof(buf), f) == NULL) { fclose(f); LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }
    rewind(f);
    LIS_INT err;
    if (strncmp(buf, "%%MatrixMarket", 14) == 0) err = input_mm(A, b, x, f);
    else { LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "only Matrix Market files are supported (Harwell-Boeing is outside the hot path)\n"); err = LIS_ERR_NOT_IMPLEMENTED; }
    fclose(f);
    return err;
}

LIS_INT lis_input_matrix(LIS_MATRIX A, char *filename) { return lis_input(A, NULL, NULL, filename); }

/* vectors: Matrix Market "vector coordinate" (size line, then "i value") or PLAIN (one value
 * per line) -- src/system/lis_input.c:216-330 */
LIS_INT lis_input_vector(LIS_VECTOR v, char *filename)
{
    FILE *f = fopen(filename, "r");
    if (f == NULL) { LIS_SETERR1(LIS_ERR_FILE_IO, "cannot open file %s\n", filename); return LIS_ERR_FILE_IO; }
    char buf[LINE_MAX_LEN];
    LIS_INT err = LIS_SUCCESS;
    if (fgets(buf, sizeof(buf), f) == NULL) { fclose(f); LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }
    if (strncmp(buf, "%%MatrixMarket", 14) == 0) {
        int gn = 0;
        err = next_data_line(f, buf, sizeof(buf));
        if (!err && sscanf(buf, "%d", &gn) != 1) { LIS_SETERR_FIO; err = LIS_ERR_FILE_IO; }
        if (!err && lis_vector_is_null(v)) err = lis_vector_set_size(v, 0, gn);
        if (!err && v->gn != gn) { LIS_SETERR(LIS_ERR_FILE_IO, "vector size does not match\n"); err = LIS_ERR_FILE_IO; }
        if (!err) err = read_mm_vec_body(f, v, gn);
    } else {
        /* PLAIN: count the lines first when the vector has no size yet */
        rewind(f);
        if (lis_vector_is_null(v)) {
            int cnt = 0; double d;
            while (fgets(buf, sizeof(buf), f)) if (sscanf(buf, "%lg", &d) == 1) cnt++;
            rewind(f);
            err = lis_vector_set_size(v, 0, cnt);
        }
        if (!err) {
            lisd_vec_host(v);
            for (LIS_INT i = 0; i < v->gn; i++) {
                double d;
                if (fgets(buf, sizeof(buf), f) == NULL || sscanf(buf, "%lg", &d) != 1) { LIS_SETERR_FIO; err = LIS_ERR_FILE_IO; break; }
                if (i >= v->is && i < v->ie) v->value[i - v->is] = d;
            }
        }
    }
    fclose(f);
    return err;
}

LIS_INT lis_output_vector(LIS_VECTOR v, LIS_INT format, char *filename)
{
    if (lis_vector_is_null(v)) { LIS_SETERR(LIS_ERR_ILL_ARG, "vector v is undefined\n"); return LIS_ERR_ILL_ARG; }
    LIS_SCALAR *all = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(v->gn > 0 ? v->gn : 1));
    if (all == NULL) { LIS_SETERR_MEM(v->gn); return LIS_OUT_OF_MEMORY; }
    LIS_INT err = lis_vector_gather(v, all);
    if (err) { free(all); return err; }
    if (lisd_rank() != 0) { free(all); return LIS_SUCCESS; }
    FILE *f = fopen(filename, "w");
    if (f == NULL) { free(all); LIS_SETERR1(LIS_ERR_FILE_IO, "cannot open file %s\n", filename); return LIS_ERR_FILE_IO; }
    switch (format) {
    case LIS_FMT_PLAIN:
        for (LIS_INT i = 0; i < v->gn; i++) fprintf(f, "%28.20e\n", (double)all[i]);
        break;
    case LIS_FMT_MM:
        fprintf(f, "%%%%MatrixMarket vector coordinate real general\n");
        fprintf(f, "%d\n", (int)v->gn);
        for (LIS_INT i = 0; i < v->gn; i++) fprintf(f, "%d %28.20e\n", (int)(i + 1), (double)all[i]);
        break;
    case LIS_FMT_LIS:
        fprintf(f, "#LIS A vec\n1\n# 0 %d\n", (int)v->gn);
        for (LIS_INT i = 0; i < v->gn; i++) { fprintf(f, "%28.20e ", (double)all[i]); if ((i + 1) % 3 == 0) fprintf(f, "\n"); }
        if (v->gn % 3 != 0) fprintf(f, "\n");
        break;
    default:
        fclose(f); free(all);
        LIS_SETERR(LIS_ERR_ILL_ARG, "ill format option\n");
        return LIS_ERR_ILL_ARG;
    }
    fclose(f);
    free(all);
    return LIS_SUCCESS;
}

/* Matrix Market coordinate writer (single process) */
LIS_INT lis_output_matrix(LIS_MATRIX A, LIS_INT format, char *path)
{
    LIS_INT err = lis_host_matrix_check_input(A);
    if (err) return err;
    if (format != LIS_FMT_MM) { LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "only LIS_FMT_MM is supported\n"); return LIS_ERR_NOT_IMPLEMENTED; }
    if (A->nprocs > 1) { LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "matrix output is single-process only\n"); return LIS_ERR_NOT_IMPLEMENTED; }
    LIS_MATRIX C = A, T = NULL;
    if (A->matrix_type != LIS_MATRIX_CSR || A->is_splited) {
        err = lis_matrix_duplicate(A, &T);
        if (err) return err;
        lis_matrix_set_type(T, LIS_MATRIX_CSR);
        err = lis_matrix_convert(A, T);
        if (err) { lis_matrix_destroy(T); return err; }
        C = T;
    }
    FILE *f = fopen(path, "w");
    if (f == NULL) { if (T) lis_matrix_destroy(T); LIS_SETERR1(LIS_ERR_FILE_IO, "cannot open file %s\n", path); return LIS_ERR_FILE_IO; }
    fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n");
    fprintf(f, "%d %d %d 0 0\n", (int)C->gn, (int)C->gn, (int)C->ptr[C->n]);
    for (LIS_INT i = 0; i < C->n; i++)
        for (LIS_INT j = C->ptr[i]; j < C->ptr[i + 1]; j++)
            fprintf(f, "%d %d %28.20e\n", (int)(i + 1), (int)(C->index[j] + 1), (double)C->value[j]);
    fclose(f);
    if (T) lis_matrix_destroy(T);
    return LIS_SUCCESS;
}
