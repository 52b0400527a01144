/*
 * lis_array.c -- the small dense helpers of the public API (lis_array_*, include/lis.h:1016-1045 of
 * the reference; src/array/lis_array.c).  Host C on plain arrays, column-major n x n matrices; used by
 * the eigensolvers (lis_esolver.c) and by the reference's test6.c / etest7.c drivers.  Every loop has
 * the reference's operation order -- including its written-out n = 1, 2, 3 cases, which differ from
 * the general loop in the sign of a zero sum -- so results are bit-identical
 * (tests/test_host_logic.py::test_lis_array_matches_reference).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lislib.h"
#include "lis_host.h"

LIS_INT lis_array_swap(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y)
{ for (LIS_INT i = 0; i < n; i++) { const LIS_SCALAR t = y[i]; y[i] = x[i]; x[i] = t; } return LIS_SUCCESS; }
LIS_INT lis_array_copy(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y)
{ for (LIS_INT i = 0; i < n; i++) y[i] = x[i]; return LIS_SUCCESS; }
LIS_INT lis_array_axpy(LIS_INT n, LIS_SCALAR alpha, LIS_SCALAR *x, LIS_SCALAR *y)
{ for (LIS_INT i = 0; i < n; i++) y[i] = alpha * x[i] + y[i]; return LIS_SUCCESS; }
LIS_INT lis_array_xpay(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR alpha, LIS_SCALAR *y)
{ for (LIS_INT i = 0; i < n; i++) y[i] = x[i] + alpha * y[i]; return LIS_SUCCESS; }
LIS_INT lis_array_axpyz(LIS_INT n, LIS_SCALAR alpha, LIS_SCALAR *x, LIS_SCALAR *y, LIS_SCALAR *z)
{ for (LIS_INT i = 0; i < n; i++) z[i] = alpha * x[i] + y[i]; return LIS_SUCCESS; }
LIS_INT lis_array_scale(LIS_INT n, LIS_SCALAR alpha, LIS_SCALAR *x)
{ for (LIS_INT i = 0; i < n; i++) x[i] = alpha * x[i]; return LIS_SUCCESS; }
LIS_INT lis_array_pmul(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y, LIS_SCALAR *z)
{ for (LIS_INT i = 0; i < n; i++) z[i] = x[i] * y[i]; return LIS_SUCCESS; }
LIS_INT lis_array_pdiv(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y, LIS_SCALAR *z)
{ for (LIS_INT i = 0; i < n; i++) z[i] = x[i] / y[i]; return LIS_SUCCESS; }
LIS_INT lis_array_set_all(LIS_INT n, LIS_SCALAR alpha, LIS_SCALAR *x)
{ for (LIS_INT i = 0; i < n; i++) x[i] = alpha; return LIS_SUCCESS; }
LIS_INT lis_array_abs(LIS_INT n, LIS_SCALAR *x)
{ for (LIS_INT i = 0; i < n; i++) x[i] = fabs(x[i]); return LIS_SUCCESS; }
LIS_INT lis_array_reciprocal(LIS_INT n, LIS_SCALAR *x)
{ for (LIS_INT i = 0; i < n; i++) x[i] = 1 / x[i]; return LIS_SUCCESS; }
LIS_INT lis_array_conjugate(LIS_INT n, LIS_SCALAR *x) { (void)n; (void)x; return LIS_SUCCESS; }     /* real scalars */
LIS_INT lis_array_shift(LIS_INT n, LIS_SCALAR sigma, LIS_SCALAR *x)
{ for (LIS_INT i = 0; i < n; i++) x[i] = x[i] - sigma; return LIS_SUCCESS; }

LIS_INT lis_array_dot(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y, LIS_SCALAR *value)
{ *value = 0; for (LIS_INT i = 0; i < n; i++) *value = *value + x[i] * y[i]; return LIS_SUCCESS; }
LIS_INT lis_array_nhdot(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y, LIS_SCALAR *value)
{ *value = 0; for (LIS_INT i = 0; i < n; i++) *value = *value + x[i] * y[i]; return LIS_SUCCESS; }
LIS_INT lis_array_nrm1(LIS_INT n, LIS_SCALAR *x, LIS_REAL *value)
{ LIS_SCALAR t = 0.0; for (LIS_INT i = 0; i < n; i++) t += fabs(x[i]); *value = t; return LIS_SUCCESS; }
LIS_INT lis_array_nrm2(LIS_INT n, LIS_SCALAR *x, LIS_REAL *value)
{ LIS_SCALAR t = 0.0; for (LIS_INT i = 0; i < n; i++) t += x[i] * x[i]; *value = sqrt(t); return LIS_SUCCESS; }
LIS_INT lis_array_nrmi(LIS_INT n, LIS_SCALAR *x, LIS_REAL *value)
{ LIS_REAL t = 0.0; for (LIS_INT i = 0; i < n; i++) if (t < fabs(x[i])) t = fabs(x[i]); *value = t; return LIS_SUCCESS; }
LIS_INT lis_array_sum(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *value)
{ LIS_SCALAR t = 0.0; for (LIS_INT i = 0; i < n; i++) t += x[i]; *value = t; return LIS_SUCCESS; }

/* y (op)= A x, src/array/lis_array.c:429-530; A(i,j) = a[i + j*n].  T: transposed access a[i*n + j] (:532-633) */
#define MV_BODY(OP, A1, A2, A3)                                                                  \
    switch (n) {                                                                                 \
    case 1: y[0] OP a[0] * x[0]; break;                                                          \
    case 2: y[0] OP a[0] * x[0] + A2(0, 1) * x[1];                                               \
            y[1] OP A2(1, 0) * x[0] + a[3] * x[1]; break;                                        \
    case 3: y[0] OP a[0] * x[0] + A3(0, 1) * x[1] + A3(0, 2) * x[2];                             \
            y[1] OP A3(1, 0) * x[0] + a[4] * x[1] + A3(1, 2) * x[2];                             \
            y[2] OP A3(2, 0) * x[0] + A3(2, 1) * x[1] + a[8] * x[2]; break;                      \
    default:                                                                                     \
        for (LIS_INT i = 0; i < n; i++) {                                                        \
            LIS_SCALAR t = 0.0;                                                                  \
            for (LIS_INT j = 0; j < n; j++) t += A1(i, j) * x[j];                                \
            y[i] OP t;                                                                           \
        }                                                                                        \
        break;                                                                                   \
    }
#define AN(i, j) a[(i) + (j) * n]
#define A2N(i, j) a[(i) + (j) * 2]
#define A3N(i, j) a[(i) + (j) * 3]
#define AT(i, j) a[(i) * n + (j)]
#define A2T(i, j) a[(i) * 2 + (j)]
#define A3T(i, j) a[(i) * 3 + (j)]

LIS_INT lis_array_matvec(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *x, LIS_SCALAR *y, LIS_INT op)
{
    if (op == LIS_INS_VALUE) { MV_BODY(=, AN, A2N, A3N) }
    else if (op == LIS_SUB_VALUE) { MV_BODY(-=, AN, A2N, A3N) }
    else { MV_BODY(+=, AN, A2N, A3N) }
    return LIS_SUCCESS;
}

LIS_INT lis_array_matvech(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *x, LIS_SCALAR *y, LIS_INT op)
{
    if (op == LIS_INS_VALUE) { MV_BODY(=, AT, A2T, A3T) }
    else if (op == LIS_SUB_VALUE) { MV_BODY(-=, AT, A2T, A3T) }
    else { MV_BODY(+=, AT, A2T, A3T) }
    return LIS_SUCCESS;
}

/* m x n block with leading dimension lda (:635-714); any other op: the square n x n add of lis_array_matvec */
LIS_INT lis_array_matvec_ns(LIS_INT m, LIS_INT n, LIS_SCALAR *a, LIS_INT lda, LIS_SCALAR *x, LIS_SCALAR *y, LIS_INT op)
{
    if (op == LIS_INS_VALUE || op == LIS_SUB_VALUE || op == LIS_ADD_VALUE) {
        for (LIS_INT i = 0; i < m; i++) {
            LIS_SCALAR t = 0.0;
            for (LIS_INT j = 0; j < n; j++) t += a[i + j * lda] * x[j];
            if (op == LIS_INS_VALUE) y[i] = t; else if (op == LIS_SUB_VALUE) y[i] -= t; else y[i] += t;
        }
    } else { MV_BODY(+=, AN, A2N, A3N) }
    return LIS_SUCCESS;
}

/* C (op)= A B, n x n (:716-847) */
#define MM_SMALL(OP)                                                                             \
    case 1: c[0] OP a[0] * b[0]; break;                                                          \
    case 2: c[0] OP a[0] * b[0] + a[2] * b[1]; c[1] OP a[1] * b[0] + a[3] * b[1];                \
            c[2] OP a[0] * b[2] + a[2] * b[3]; c[3] OP a[1] * b[2] + a[3] * b[3]; break;         \
    case 3: for (int q = 0; q < 3; q++) {                                                        \
                c[3 * q + 0] OP a[0] * b[3 * q] + a[3] * b[3 * q + 1] + a[6] * b[3 * q + 2];     \
                c[3 * q + 1] OP a[1] * b[3 * q] + a[4] * b[3 * q + 1] + a[7] * b[3 * q + 2];     \
                c[3 * q + 2] OP a[2] * b[3 * q] + a[5] * b[3 * q + 1] + a[8] * b[3 * q + 2];     \
            } break;

LIS_INT lis_array_matmat(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *b, LIS_SCALAR *c, LIS_INT op)
{
    LIS_INT i, j, l;
    if (op == LIS_INS_VALUE) {
        switch (n) {
        MM_SMALL(=)
        default:
            for (j = 0; j < n; j++) {
                for (i = 0; i < n; i++) c[i + j * n] = 0.0;
                for (l = 0; l < n; l++) for (i = 0; i < n; i++) c[i + j * n] += a[i + l * n] * b[l + j * n];
            }
            break;
        }
    } else if (op == LIS_SUB_VALUE) {
        switch (n) {
        MM_SMALL(-=)
        default:
            for (j = 0; j < n; j++) for (l = 0; l < n; l++) for (i = 0; i < n; i++) c[i + j * n] -= a[i + l * n] * b[l + j * n];
            break;
        }
    } else {
        switch (n) {
        MM_SMALL(+=)
        default:
            for (j = 0; j < n; j++) for (l = 0; l < n; l++) for (i = 0; i < n; i++) c[i + j * n] += a[i + l * n] * b[l + j * n];
            break;
        }
    }
    return LIS_SUCCESS;
}

/* C (op)= A B, A l x n (lda), B n x m (ldb), C l x m (ldc) (:849-905) */
LIS_INT lis_array_matmat_ns(LIS_INT l, LIS_INT m, LIS_INT n, LIS_SCALAR *a, LIS_INT lda, LIS_SCALAR *b, LIS_INT ldb,
                            LIS_SCALAR *c, LIS_INT ldc, LIS_INT op)
{
    LIS_INT i, j, k;
    if (op == LIS_INS_VALUE) {
        for (j = 0; j < m; j++) {
            for (i = 0; i < l; i++) c[i + j * ldc] = 0.0;
            for (k = 0; k < n; k++) for (i = 0; i < l; i++) c[i + j * ldc] += a[i + k * lda] * b[k + j * ldb];
        }
    } else if (op == LIS_SUB_VALUE) {
        for (j = 0; j < m; j++) for (k = 0; k < n; k++) for (i = 0; i < l; i++) c[i + j * ldc] -= a[i + k * lda] * b[k + j * ldb];
    } else {
        for (j = 0; j < m; j++) for (k = 0; k < n; k++) for (i = 0; i < l; i++) c[i + j * ldc] += a[i + k * lda] * b[k + j * ldb];
    }
    return LIS_SUCCESS;
}

/* A <- A^-1 by Gaussian elimination without pivoting (:907-958) */
LIS_INT lis_array_ge(LIS_INT n, LIS_SCALAR *a)
{
    LIS_INT i, j, k;
    LIS_SCALAR t;
    LIS_SCALAR *lu = (LIS_SCALAR *)malloc((size_t)(n > 0 ? n * n : 1) * sizeof(LIS_SCALAR));
    if (lu == NULL) { LIS_SETERR_MEM(n * n * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    memcpy(lu, a, (size_t)n * n * sizeof(LIS_SCALAR));
    for (k = 0; k < n; k++) {
        lu[k + k * n] = 1.0 / lu[k + k * n];
        for (i = k + 1; i < n; i++) {
            t = lu[i + k * n] * lu[k + k * n];
            for (j = k + 1; j < n; j++) lu[i + j * n] -= t * lu[k + j * n];
            lu[i + k * n] = t;
        }
    }
    for (k = 0; k < n; k++) {
        for (i = 0; i < n; i++) {
            t = (i == k);
            for (j = 0; j < i; j++) t -= lu[i + j * n] * a[j + k * n];
            a[i + k * n] = t;
        }
        for (i = n - 1; i >= 0; i--) {
            t = a[i + k * n];
            for (j = i + 1; j < n; j++) t -= lu[i + j * n] * a[j + k * n];
            a[k * n + i] = t * lu[i + i * n];
        }
    }
    free(lu);
    return LIS_SUCCESS;
}

/* x = A^-1 b on a work copy w of A (:960-1027; n = 1 and 2 written out) */
LIS_INT lis_array_solve(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *b, LIS_SCALAR *x, LIS_SCALAR *w)
{
    LIS_INT i, j, k;
    LIS_SCALAR t;
    for (i = 0; i < n * n; i++) w[i] = a[i];
    switch (n) {
    case 1:
        x[0] = b[0] / w[0];
        break;
    case 2:
        w[0] = 1.0 / w[0];
        w[1] *= w[0];
        w[3] -= w[1] * w[2];
        w[3] = 1.0 / w[3];
        x[0] = b[0];
        x[1] = b[1] - w[1] * x[0];
        x[1] *= w[3];
        x[0] -= w[2] * x[1];
        x[0] *= w[0];
        break;
    default:
        for (k = 0; k < n; k++) {
            w[k + k * n] = 1.0 / w[k + k * n];
            for (i = k + 1; i < n; i++) {
                t = w[i + k * n] * w[k + k * n];
                for (j = k + 1; j < n; j++) w[i + j * n] -= t * w[k + j * n];
                w[i + k * n] = t;
            }
        }
        for (i = 0; i < n; i++) {
            x[i] = b[i];
            for (j = 0; j < i; j++) x[i] -= w[i + j * n] * x[j];
        }
        for (i = n - 1; i >= 0; i--) {
            for (j = i + 1; j < n; j++) x[i] -= w[i + j * n] * x[j];
            x[i] *= w[i + i * n];
        }
        break;
    }
    return LIS_SUCCESS;
}

/* classical Gram-Schmidt QR (:1029-1082) */
LIS_INT lis_array_cgs(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *q, LIS_SCALAR *r)
{
    const LIS_REAL tol = 1e-12;
    LIS_INT i, j, k;
    LIS_REAL nrm2;
    LIS_SCALAR *a_k = (LIS_SCALAR *)malloc((size_t)(n > 0 ? n : 1) * sizeof(LIS_SCALAR));
    if (a_k == NULL) { LIS_SETERR_MEM(n * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    for (i = 0; i < n * n; i++) { q[i] = 0.0; r[i] = 0.0; }
    for (k = 0; k < n; k++) {
        for (i = 0; i < n; i++) a_k[i] = a[i + k * n];
        for (j = 0; j < k; j++) {
            r[j + k * n] = 0;
            for (i = 0; i < n; i++) r[j + k * n] += q[i + j * n] * a[i + k * n];
            for (i = 0; i < n; i++) a_k[i] -= r[j + k * n] * q[i + j * n];
        }
        lis_array_nrm2(n, a_k, &nrm2);
        r[k + k * n] = nrm2;
        if (nrm2 < tol) break;
        for (i = 0; i < n; i++) q[i + k * n] = a_k[i] / nrm2;
    }
    free(a_k);
    return LIS_SUCCESS;
}

/* modified Gram-Schmidt QR; overwrites a (:1084-1134) */
LIS_INT lis_array_mgs(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *q, LIS_SCALAR *r)
{
    const LIS_REAL tol = 1e-12;
    LIS_INT i, j, k;
    LIS_REAL nrm2;
    LIS_SCALAR *a_j = (LIS_SCALAR *)malloc((size_t)(n > 0 ? n : 1) * sizeof(LIS_SCALAR));
    if (a_j == NULL) { LIS_SETERR_MEM(n * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    for (i = 0; i < n * n; i++) { q[i] = 0.0; r[i] = 0.0; }
    for (j = 0; j < n; j++) {
        for (i = 0; i < n; i++) a_j[i] = a[i + j * n];
        lis_array_nrm2(n, a_j, &nrm2);
        r[j + j * n] = nrm2;
        for (i = 0; i < n; i++) {
            if (nrm2 < tol) break;
            q[i + j * n] = a_j[i] / nrm2;
        }
        for (k = j + 1; k < n; k++) {
            r[j + k * n] = 0;
            for (i = 0; i < n; i++) r[j + k * n] += q[i + j * n] * a[i + k * n];
            for (i = 0; i < n; i++) a[i + k * n] -= r[j + k * n] * q[i + j * n];
        }
    }
    free(a_j);
    return LIS_SUCCESS;
}

/* unshifted QR iteration A <- R Q until |a[1]| < 1e-12 (:1136-1175) */
LIS_INT lis_array_qr(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *q, LIS_SCALAR *r, LIS_INT *qriter, LIS_REAL *qrerr)
{
    const LIS_INT maxiter = 100000;
    const LIS_REAL tol = 1e-12;
    LIS_INT i, j, k, iter = 0;
    LIS_REAL err = 0.0;
    while (iter < maxiter) {
        iter = iter + 1;
        lis_array_cgs(n, a, q, r);
        for (j = 0; j < n; j++)
            for (i = 0; i < n; i++) {
                a[i + j * n] = 0;
                for (k = 0; k < n; k++) a[i + j * n] += r[i + k * n] * q[k + j * n];
            }
        err = fabs(a[1]);
        if (err < tol) break;
    }
    *qriter = iter;
    *qrerr = err;
    return LIS_SUCCESS;
}
