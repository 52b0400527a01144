/*
 * lis_io.c -- the on-disk formats either side of the hot path that the reference drivers
 * touch: Matrix Market input (with Lis' extension that appends b and x to the file,
 * src/system/lis_input_mm.c:61-1069, used by test/test1.c), vector input, and the
 * Harwell-Boeing (RUA) input (src/system/lis_input_hb.c), and the solution / matrix writers
 * (src/system/lis_output.c:200-440).  Host C, ASCII only.
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

/* binary records of Lis' LIS_FMT_MMB extension (include/lis_io.h:104-115 of the reference): the
 * size line carries a sixth field, 1 + (1 if the writer was little-endian) */
typedef struct { LIS_INT i; LIS_SCALAR value; } mmb_vec_t;
typedef struct { LIS_INT i; LIS_INT j; LIS_SCALAR value; } mmb_mat_t;

static void bswap_bytes(void *p, size_t n)
{
    unsigned char *b = (unsigned char *)p;
    for (size_t k = 0; k < n / 2; k++) { const unsigned char t = b[k]; b[k] = b[n - 1 - k]; b[n - 1 - k] = t; }
}
static int host_little_endian(void) { const int one = 1; return *(const char *)&one; }

/* n lines "i value" (1-based), or n binary records, into the locally owned slice of v */
static LIS_INT read_mm_vec_body(FILE *f, LIS_VECTOR v, LIS_INT gn, int isbin)
{
    char buf[LINE_MAX_LEN];
    const int swap = isbin && host_little_endian() != isbin - 1;
    LIS_SCALAR *hv = lisd_vec_host_view(v, 1);
    if (hv == NULL) return LIS_ERR_OUT_OF_MEMORY;
    for (LIS_INT i = 0; i < gn; i++) {
        int idx; double val;
        if (isbin) {
            mmb_vec_t r;
            if (fread(&r, sizeof(r), 1, f) != 1) { lisd_vec_host_done(v, hv, 0); LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }
            if (swap) { bswap_bytes(&r.i, sizeof(r.i)); bswap_bytes(&r.value, sizeof(r.value)); }
            idx = (int)r.i; val = (double)r.value;
        } else if (fgets(buf, sizeof(buf), f) == NULL || sscanf(buf, "%d %lg", &idx, &val) != 2) { lisd_vec_host_done(v, hv, 0); LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }
        idx--;
        if (idx >= v->is && idx < v->ie) hv[idx - v->is] = val;
    }
    return lisd_vec_host_done(v, hv, 1);
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
    int nr = 0, nc = 0, nnz = 0, isb = 0, isx = 0, isbin = 0;
    const int got = sscanf(buf, "%d %d %d %d %d %d", &nr, &nc, &nnz, &isb, &isx, &isbin);
    if (got != 3 && got != 5 && got != 6) { LIS_SETERR(LIS_ERR_FILE_IO, "matrix size line is not correct\n"); return LIS_ERR_FILE_IO; }
    if (got < 6) isbin = 0;
    if (got < 5) { isb = 0; isx = 0; }
    const int swap = isbin && host_little_endian() != isbin - 1;
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
        if (isbin) {
            mmb_mat_t r;
            if (fread(&r, sizeof(r), 1, f) != 1) { LIS_SETERR_FIO; err = LIS_ERR_FILE_IO; goto done; }
            if (swap) { bswap_bytes(&r.i, sizeof(r.i)); bswap_bytes(&r.j, sizeof(r.j)); bswap_bytes(&r.value, sizeof(r.value)); }
            ri[k] = (int)r.i; ci[k] = (int)r.j; v = (double)r.value;
        } else {
            if (fgets(buf, sizeof(buf), f) == NULL) { LIS_SETERR_FIO; err = LIS_ERR_FILE_IO; goto done; }
            const int c = bn.is_pattern ? sscanf(buf, "%d %d", &ri[k], &ci[k]) + 1 : sscanf(buf, "%d %d %lg", &ri[k], &ci[k], &v);
            if (c != 3) { LIS_SETERR_FIO; err = LIS_ERR_FILE_IO; goto done; }
        }
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
        err = read_mm_vec_body(f, b, nr, isbin);
        if (err) goto done;
    } else if (isb) {
        if (isbin) fseek(f, (long)sizeof(mmb_vec_t) * nr, SEEK_CUR);
        else for (int k = 0; k < nr; k++) if (fgets(buf, sizeof(buf), f) == NULL) break;
    }
    if (isx && x) {
        if (lis_vector_is_null(x)) { err = lis_vector_set_size(x, A->n, 0); if (err) goto done; }
        err = read_mm_vec_body(f, x, nr, isbin);
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

/* ------------------------------------------------------------------ Harwell-Boeing (RUA)
 * What src/system/lis_input_hb.c:129-466 accepts: real, unsymmetric, assembled, square; the
 * right-hand sides a file may carry are skipped (b and x are left to the caller, test/test1.c
 * then builds b = A*1).  Fixed-width Fortran fields: "(10I8)" -> 10 per line, 8 wide; the count and
 * width are taken the way the reference takes them (text before / after the edit letter through
 * atoi), and every field goes through atoi / atof, so "1.5D+00" reads as 1.5 there and here.
 * The file holds compressed columns: loaded as CSC, then converted to the requested storage. */
static void hb_fmt(const char *field, int size, int *count, int *width)
{
    char tmp[64];
    char *p, *s, *t;
    if (size > 63) size = 63;
    strncpy(tmp, field, (size_t)size); tmp[size] = '\0';
    lower(tmp);
    *count = 0; *width = 0;
    p = strchr(tmp, '(');
    if (p == NULL) return;
    s = p + 1;
    p = strchr(s, ')');
    if (p) *p = '\0';
    p = strchr(s, 'i');
    if (p == NULL) {
        p = strchr(s, 'e');
        if (p == NULL) p = strchr(s, 'd');
        if (p == NULL) return;
        t = strchr(s, '.');
        if (t) *t = '\0';
    }
    *p = '\0';
    *count = atoi(s);
    *width = atoi(p + 1);
}

static LIS_INT input_hb(LIS_MATRIX A, FILE *f)
{
    char buf[LINE_MAX_LEN], mtx[64] = "", dat[128];
    int totcrd = 0, ptrcrd = 0, indcrd = 0, valcrd = 0, rhscrd = 0;
    int nrow = 0, ncol = 0, nnzero = 0, neltvl = 0;
    int iptr, iind, ival, irhs, wptr, wind, wval, wrhs;
    const LIS_INT want_type = A->matrix_type;
    LIS_INT *ptr = NULL, *index = NULL, err;
    LIS_SCALAR *value = NULL;
    if (A->nprocs > 1) { LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "Harwell-Boeing input is single-process only\n"); return LIS_ERR_NOT_IMPLEMENTED; }
    if (fgets(buf, sizeof(buf), f) == NULL) { LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }                 /* line 1: title, key */
    if (fgets(buf, sizeof(buf), f) == NULL ||
        sscanf(buf, "%14d%14d%14d%14d%14d", &totcrd, &ptrcrd, &indcrd, &valcrd, &rhscrd) < 4) { LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }
    if (fgets(buf, sizeof(buf), f) == NULL ||
        sscanf(buf, "%63s %d %d %d %d", mtx, &nrow, &ncol, &nnzero, &neltvl) != 5) { LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }
    lower(mtx);
    if (mtx[0] != 'r') { LIS_SETERR(LIS_ERR_FILE_IO, "Not real\n"); return LIS_ERR_FILE_IO; }
    if (mtx[1] != 'u') { LIS_SETERR(LIS_ERR_FILE_IO, "Not unsymmetric\n"); return LIS_ERR_FILE_IO; }
    if (mtx[2] != 'a') { LIS_SETERR(LIS_ERR_FILE_IO, "Not assembled\n"); return LIS_ERR_FILE_IO; }
    if (nrow != ncol) { LIS_SETERR(LIS_ERR_FILE_IO, "matrix is not square\n"); return LIS_ERR_FILE_IO; }
    printf("matrix size = %d x %d (%d nonzero entries)\n\n", nrow, ncol, nnzero);
    memset(buf, 0, sizeof(buf));
    if (fgets(buf, sizeof(buf), f) == NULL) { LIS_SETERR_FIO; return LIS_ERR_FILE_IO; }                 /* line 4: formats */
    hb_fmt(buf, 16, &iptr, &wptr);
    hb_fmt(buf + 16, 16, &iind, &wind);
    hb_fmt(buf + 32, 20, &ival, &wval);
    hb_fmt(buf + 52, 20, &irhs, &wrhs);
    if (rhscrd != 0 && fgets(buf, sizeof(buf), f) == NULL) { LIS_SETERR_FIO; return LIS_ERR_FILE_IO; } /* line 5 */
    if (wptr <= 0 || wind <= 0 || wval <= 0 || wptr > 127 || wind > 127 || wval > 127 || nnzero < 0) {
        LIS_SETERR(LIS_ERR_FILE_IO, "Harwell-Boeing format line is not understood\n");
        return LIS_ERR_FILE_IO;
    }
    err = lis_matrix_set_size(A, 0, nrow);
    if (err) return err;
    const LIS_INT n = A->n;
    err = lis_matrix_malloc_csr(n, nnzero, &ptr, &index, &value);
    if (err) return err;
#define HB_READ(cards, per, width, limit, STORE)                                                  \
    do {                                                                                           \
        LIS_INT k = 0;                                                                             \
        for (int c = 0; c < (cards); c++) {                                                        \
            if (fgets(buf, sizeof(buf), f) == NULL) { LIS_SETERR_FIO; err = LIS_ERR_FILE_IO; goto fail; } \
            const size_t len = strlen(buf);                                                        \
            const char *p = buf;                                                                   \
            for (int j = 0; j < (per) && k < (limit); j++) {                                       \
                if ((size_t)(p - buf) >= len) { dat[0] = '\0'; }                                   \
                else { strncpy(dat, p, (size_t)(width)); dat[(width)] = '\0'; }                    \
                STORE;                                                                             \
                p += (width);                                                                      \
                k++;                                                                               \
            }                                                                                      \
        }                                                                                          \
    } while (0)
    HB_READ(ptrcrd, iptr, wptr, n + 1, ptr[k] = atoi(dat) - 1);
    HB_READ(indcrd, iind, wind, nnzero, index[k] = atoi(dat) - 1);
    HB_READ(valcrd, ival, wval, nnzero, value[k] = atof(dat));
#undef HB_READ
    if (ptr[0] != 0 || ptr[n] != nnzero) { LIS_SETERR(LIS_ERR_FILE_IO, "Harwell-Boeing column pointers do not match the entry count\n"); err = LIS_ERR_FILE_IO; goto fail; }
    for (LIS_INT k = 0; k < nnzero; k++)
        if (index[k] < 0 || index[k] >= n) { LIS_SETERR(LIS_ERR_FILE_IO, "index out of range\n"); err = LIS_ERR_FILE_IO; goto fail; }
    lis_matrix_set_type(A, LIS_MATRIX_CSC);
    err = lis_matrix_set_csc(nnzero, ptr, index, value, A);
    if (err) goto fail;
    ptr = NULL; index = NULL; value = NULL;
    err = lis_matrix_assemble(A);
    if (err) return err;
    if (want_type != LIS_MATRIX_CSC) {
        /* through CSR, like lis_input_hb_csr + lis_input_hb (:391-407, :72-99) */
        LIS_MATRIX B;
        err = lis_matrix_duplicate(A, &B);
        if (err) return err;
        lis_matrix_set_type(B, LIS_MATRIX_CSR);
        err = lis_matrix_convert(A, B);
        if (err) { lis_matrix_destroy(B); return err; }
        lis_host_matrix_adopt(A, B);
        if (want_type != LIS_MATRIX_CSR) {
            err = lis_matrix_duplicate(A, &B);
            if (err) return err;
            lis_matrix_set_type(B, want_type);
            err = lis_matrix_convert(A, B);
            if (err) { lis_matrix_destroy(B); return err; }
            lis_host_matrix_adopt(A, B);
        }
    }
    return LIS_SUCCESS;
fail:
    lis_free2(3, ptr, index, value);
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
    else err = input_hb(A, f);                     /* anything else is taken for Harwell-Boeing, src/system/lis_input.c:100-118 */
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
        if (!err) err = read_mm_vec_body(f, v, gn, 0);
    } else {
        /* PLAIN: count the lines first when the vector has no size yet */
        rewind(f);
        if (lis_vector_is_null(v)) {
            int cnt = 0; double d;
            while (fgets(buf, sizeof(buf), f)) if (sscanf(buf, "%lg", &d) == 1) cnt++;
            rewind(f);
            err = lis_vector_set_size(v, 0, cnt);
        }
        LIS_SCALAR *hv = err ? NULL : lisd_vec_host_view(v, 1);
        if (!err && hv == NULL) err = LIS_ERR_OUT_OF_MEMORY;
        if (!err) {
            for (LIS_INT i = 0; i < v->gn; i++) {
                double d;
                if (fgets(buf, sizeof(buf), f) == NULL || sscanf(buf, "%lg", &d) != 1) { LIS_SETERR_FIO; err = LIS_ERR_FILE_IO; break; }
                if (i >= v->is && i < v->ie) hv[i - v->is] = d;
            }
            const LIS_INT err2 = lisd_vec_host_done(v, hv, err == LIS_SUCCESS);
            if (!err) err = err2;
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

/* Matrix Market coordinate writer (single process): the matrix, then b and x when they are set
 * -- Lis' extension of the size line "n n nnz isb isx", which lis_input reads back
 * (src/system/lis_output.c:62-143, lis_output_mm.c:404-466 header, :58-150 vectors, :560-720 CSR).
 * LIS_FMT_MMB: same header with a sixth field (1 + little-endian), records in binary. */
static void write_mm_vec(FILE *f, LIS_VECTOR v, LIS_INT format)
{
    LIS_SCALAR *hv = lisd_vec_host_view(v, 1);
    if (hv == NULL) return;
    for (LIS_INT i = 0; i < v->n; i++) {
        if (format == LIS_FMT_MM) fprintf(f, "%d %28.20e\n", (int)(v->is + i + 1), (double)hv[i]);
        else { mmb_vec_t r; memset(&r, 0, sizeof(r)); r.i = v->is + i + 1; r.value = hv[i]; fwrite(&r, sizeof(r), 1, f); }
    }
    lisd_vec_host_done(v, hv, 0);
}

LIS_INT lis_output(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, LIS_INT format, char *path)
{
    LIS_INT err = lis_host_matrix_check_input(A);
    if (err) return err;
    if (format != LIS_FMT_MM && format != LIS_FMT_MMB) return LIS_SUCCESS;      /* the reference writes nothing either */
    if (A->nprocs > 1) { LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "matrix output is single-process only\n"); return LIS_ERR_NOT_IMPLEMENTED; }
    LIS_MATRIX Cm = A, T = NULL;
    if (A->matrix_type != LIS_MATRIX_CSR || A->is_splited) {
        err = lis_matrix_duplicate(A, &T);
        if (err) return err;
        lis_matrix_set_type(T, LIS_MATRIX_CSR);
        err = lis_matrix_convert(A, T);
        if (err) { lis_matrix_destroy(T); return err; }
        Cm = T;
    }
    const int isb = b != NULL && !lis_vector_is_null(b), isx = x != NULL && !lis_vector_is_null(x);
    FILE *f = fopen(path, format == LIS_FMT_MM ? "w" : "wb");
    if (f == NULL) { if (T) lis_matrix_destroy(T); LIS_SETERR1(LIS_ERR_FILE_IO, "cannot open file %s\n", path); return LIS_ERR_FILE_IO; }
    fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n");
    if (format == LIS_FMT_MMB) fprintf(f, "%d %d %d %d %d %d\n", (int)Cm->gn, (int)Cm->gn, (int)Cm->nnz, isb, isx, host_little_endian() + 1);
    else if (!isb && !isx) fprintf(f, "%d %d %d\n", (int)Cm->gn, (int)Cm->gn, (int)Cm->nnz);
    else fprintf(f, "%d %d %d %d %d\n", (int)Cm->gn, (int)Cm->gn, (int)Cm->nnz, isb, isx);
    for (LIS_INT i = 0; i < Cm->n; i++)
        for (LIS_INT j = Cm->ptr[i]; j < Cm->ptr[i + 1]; j++) {
            if (format == LIS_FMT_MM) fprintf(f, "%d %d %28.20e\n", (int)(i + 1), (int)(Cm->index[j] + 1), (double)Cm->value[j]);
            else { mmb_mat_t r; memset(&r, 0, sizeof(r)); r.i = i + 1; r.j = Cm->index[j] + 1; r.value = Cm->value[j]; fwrite(&r, sizeof(r), 1, f); }
        }
    if (isb) write_mm_vec(f, b, format);
    if (isx) write_mm_vec(f, x, format);
    fclose(f);
    if (T) lis_matrix_destroy(T);
    return LIS_SUCCESS;
}

LIS_INT lis_output_matrix(LIS_MATRIX A, LIS_INT format, char *path) { return lis_output(A, NULL, NULL, format, path); }
