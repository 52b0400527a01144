/*
 * lis_shim.c -- one test driver, written against the PUBLIC Lis API only (lis.h + the lislib.h
 * "friend" API that the reference's own spmvtest drivers use), compiled several times:
 *   - against the reference sources        -> oracle/_ref/libref_shim_{serial,omp}.so
 *   - against lis_b200 (include/ + liblis) -> lis_b200/_lib/liblis_b200_shim.so
 * Tests load the builds side by side with ctypes and hand them the same plain arrays, so a
 * parity test reads like the reference's test/spmvtest*.c and test/test3.c: build CSR with
 * lis_matrix_set_csr, convert with lis_matrix_convert, run lis_matvec / lis_solve, compare.
 *
 * TEST INFRASTRUCTURE: nothing in the product links this file.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef HAVE_CONFIG_H
#include "lis_config.h"
#endif
#ifdef _OPENMP
#include <omp.h>
#endif
#include <unistd.h>
#include <fcntl.h>
#include "lislib.h"

#define EXPORT __attribute__((visibility("default")))

/* "-print mem" makes the reference print its solver banner on stdout; keep test logs quiet */
static int quiet_begin(void)
{
    if (getenv("SHIM_VERBOSE")) return -1;
    fflush(stdout);
    const int saved = dup(1), devnull = open("/dev/null", O_WRONLY);
    if (saved < 0 || devnull < 0) return -1;
    dup2(devnull, 1);
    close(devnull);
    return saved;
}
static void quiet_end(int saved)
{
    if (saved < 0) return;
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
}

static int g_started = 0;

/* args: "-name value ..." forwarded to lis_initialize as a synthetic argv */
EXPORT int shim_begin(const char *args)
{
    static char buf[1024];
    static char *argvv[64];
    char **argv = argvv;
    int argc = 1;
    if (g_started) return 0;
    argvv[0] = (char *)"shim";
    if (args) {
        strncpy(buf, args, sizeof(buf) - 1);
        for (char *t = strtok(buf, " "); t && argc < 63; t = strtok(NULL, " ")) argvv[argc++] = t;
    }
    argvv[argc] = NULL;
    int err = (int)lis_initialize(&argc, &argv);
    if (!err) g_started = 1;
    return err;
}

EXPORT int shim_end(void)
{
    if (!g_started) return 0;
    g_started = 0;
    return (int)lis_finalize();
}

/* 1 = this build is lis_b200, 0 = the reference */
EXPORT int shim_is_b200(void)
{
#ifdef LIS_B200_LIS_H
    return 1;
#else
    return 0;
#endif
}

EXPORT int shim_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* thread count of the OpenMP reference == SSOR block count / reduction chunking there;
 * for lis_b200 the emulated count (SSOR blocks) */
EXPORT int shim_set_threads(int n)
{
#ifdef LIS_B200_LIS_H
    lis_b200_set_num_threads(n);
#elif defined(_OPENMP)
    omp_set_num_threads(n);
#else
    (void)n;
#endif
    return 0;
}

/* ------------------------------------------------------------------ helpers */
static LIS_INT make_csr(int n, const int *ptr, const int *idx, const double *val, int sort_rows, LIS_MATRIX *out)
{
    LIS_MATRIX A;
    LIS_INT err, *p, *ix;
    LIS_SCALAR *v;
    const int nnz = ptr[n];
    err = lis_matrix_create(LIS_COMM_WORLD, &A); if (err) return err;
    err = lis_matrix_set_size(A, 0, n); if (err) return err;
    p = (LIS_INT *)malloc(sizeof(LIS_INT) * ((size_t)n + 1));
    ix = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(nnz > 0 ? nnz : 1));
    v = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(nnz > 0 ? nnz : 1));
    if (!p || !ix || !v) return LIS_OUT_OF_MEMORY;
    for (int i = 0; i <= n; i++) p[i] = ptr[i];
    for (int j = 0; j < nnz; j++) { ix[j] = idx[j]; v[j] = val[j]; }
    err = lis_matrix_set_csr(nnz, p, ix, v, A); if (err) return err;
    err = lis_matrix_assemble(A); if (err) return err;
    if (sort_rows)          /* like test/spmvtest3.c:192-195 */
        for (int i = 0; i < n; i++) lis_sort_id(A->ptr[i], A->ptr[i + 1] - 1, A->index, A->value);
    *out = A;
    return LIS_SUCCESS;
}

static LIS_INT convert_to(LIS_MATRIX A0, int fmt, int bnr, int bnc, LIS_MATRIX *out)
{
    LIS_MATRIX A;
    LIS_INT err = lis_matrix_duplicate(A0, &A); if (err) return err;
    err = lis_matrix_set_type(A, fmt); if (err) return err;
    if ((fmt == LIS_MATRIX_BSR || fmt == LIS_MATRIX_BSC) && bnr > 0) { err = lis_matrix_set_blocksize(A, bnr, bnc, NULL, NULL); if (err) return err; }
    err = lis_matrix_convert(A0, A); if (err) return err;
    *out = A;
    return LIS_SUCCESS;
}

static LIS_INT make_vec(LIS_MATRIX A, const double *src, LIS_VECTOR *out)
{
    LIS_VECTOR v;
    LIS_INT err = lis_vector_duplicate(A, &v); if (err) return err;
    if (src && A->n > 0) { err = lis_vector_scatter((LIS_SCALAR *)src, v); if (err) return err; }
    *out = v;
    return LIS_SUCCESS;
}

static LIS_INT make_vec_n(int n, const double *src, LIS_VECTOR *out)
{
    LIS_VECTOR v;
    LIS_INT err = lis_vector_create(LIS_COMM_WORLD, &v); if (err) return err;
    err = lis_vector_set_size(v, 0, n); if (err) return err;
    if (src && n > 0) { err = lis_vector_scatter((LIS_SCALAR *)src, v); if (err) return err; }
    *out = v;
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ SpMV
 * fmt = LIS_MATRIX_* ; split=1 runs lis_matrix_split first (the D,L,U summation order);
 * iters >= 1 calls of lis_matvec are timed with lis_wtime like test/spmvtest1.c:217-222 */
EXPORT int shim_spmv(int fmt, int n, const int *ptr, const int *idx, const double *val, int bnr, int bnc,
                     int sort_rows, int split, const double *x, double *y, int iters, double *seconds)
{
    LIS_MATRIX A0 = NULL, A = NULL;
    LIS_VECTOR vx = NULL, vy = NULL;
    LIS_INT err;
    double t = 0.0;
    err = make_csr(n, ptr, idx, val, sort_rows, &A0); if (err) return (int)err;
    err = convert_to(A0, fmt, bnr, bnc, &A); if (err) return (int)err;
    if (split) { err = lis_matrix_split(A); if (err) return (int)err; }
    err = make_vec(A, x, &vx); if (err) return (int)err;
    err = make_vec(A, NULL, &vy); if (err) return (int)err;
    err = lis_matvec(A, vx, vy); if (err) return (int)err;      /* warm-up: first-use upload */
    for (int k = 0; k < iters; k++) {
        const double t0 = lis_wtime();
        err = lis_matvec(A, vx, vy);
        t += lis_wtime() - t0;
        if (err) return (int)err;
    }
    if (seconds) *seconds = t;
    if (n > 0) { err = lis_vector_gather(vy, y); if (err) return (int)err; }
    lis_vector_destroy(vx); lis_vector_destroy(vy);
    lis_matrix_destroy(A); lis_matrix_destroy(A0);
    return 0;
}

/* lis_output(A, b, x, format, path) for the matrix in storage format `fmt`; b / x may be NULL (then
 * an unset vector is passed, like lis_output_matrix does); vformat > 0 also writes b with
 * lis_output_vector(b, vformat, vpath) */
EXPORT int shim_output(int fmt, int n, const int *ptr, const int *idx, const double *val, const double *b, const double *x,
                       int format, const char *path, int vformat, const char *vpath)
{
    LIS_MATRIX A0 = NULL, A = NULL;
    LIS_VECTOR vb = NULL, vx = NULL;
    LIS_INT err;
    err = make_csr(n, ptr, idx, val, 0, &A0); if (err) return (int)err;
    err = convert_to(A0, fmt, 2, 2, &A); if (err) return (int)err;
    if (b) { err = make_vec(A, b, &vb); } else { err = lis_vector_create(LIS_COMM_WORLD, &vb); } if (err) return (int)err;
    if (x) { err = make_vec(A, x, &vx); } else { err = lis_vector_create(LIS_COMM_WORLD, &vx); } if (err) return (int)err;
    err = lis_output(A, vb, vx, format, (char *)path); if (err) return (int)err;
    if (vformat > 0 && b) { err = lis_output_vector(vb, vformat, (char *)vpath); if (err) return (int)err; }
    lis_vector_destroy(vb); lis_vector_destroy(vx);
    lis_matrix_destroy(A); lis_matrix_destroy(A0);
    return 0;
}

/* y = A^H x (lis_matvech), optionally on the split matrix */
EXPORT int shim_matvech(int fmt, int n, const int *ptr, const int *idx, const double *val, int split, const double *x, double *y)
{
    LIS_MATRIX A0 = NULL, A = NULL;
    LIS_VECTOR vx = NULL, vy = NULL;
    LIS_INT err;
    err = make_csr(n, ptr, idx, val, 0, &A0); if (err) return (int)err;
    err = convert_to(A0, fmt, 0, 0, &A); if (err) return (int)err;
    if (split) { err = lis_matrix_split(A); if (err) return (int)err; }
    err = make_vec(A, x, &vx); if (err) return (int)err;
    err = make_vec(A, NULL, &vy); if (err) return (int)err;
    err = lis_matvech(A, vx, vy); if (err) return (int)err;
    err = lis_vector_gather(vy, y); if (err) return (int)err;
    lis_vector_destroy(vx); lis_vector_destroy(vy); lis_matrix_destroy(A); lis_matrix_destroy(A0);
    return 0;
}

/* ------------------------------------------------------------------ converted layouts */
static LIS_MATRIX g_conv[16];
static LIS_MATRIX g_conv0[16];

EXPORT int shim_convert_open(int fmt, int n, const int *ptr, const int *idx, const double *val, int bnr, int bnc, int sort_rows)
{
    int h;
    for (h = 0; h < 16 && g_conv[h]; h++) ;
    if (h == 16) return -1;
    if (make_csr(n, ptr, idx, val, sort_rows, &g_conv0[h])) return -2;
    if (convert_to(g_conv0[h], fmt, bnr, bnc, &g_conv[h])) return -3;
    return h;
}

/* CSR -> fmt -> CSR; the handle holds the final CSR matrix */
EXPORT int shim_roundtrip_open(int fmt, int n, const int *ptr, const int *idx, const double *val, int bnr, int bnc)
{
    int h;
    LIS_MATRIX A0, A1;
    for (h = 0; h < 16 && g_conv[h]; h++) ;
    if (h == 16) return -1;
    if (make_csr(n, ptr, idx, val, 0, &A0)) return -2;
    if (convert_to(A0, fmt, bnr, bnc, &A1)) return -3;
    if (convert_to(A1, LIS_MATRIX_CSR, bnr, bnc, &g_conv[h])) return -4;
    lis_matrix_destroy(A1);
    g_conv0[h] = A0;
    return h;
}

/* lis_input (Matrix Market) into a handle; b/x presence flags returned in has[0..1] and the
 * vectors copied to bx (2*n doubles) when present */
EXPORT int shim_input_open(const char *path, int fmt, int *has, double *bx, int bx_cap)
{
    int h;
    LIS_MATRIX A;
    LIS_VECTOR b, x;
    for (h = 0; h < 16 && g_conv[h]; h++) ;
    if (h == 16) return -1;
    if (lis_matrix_create(LIS_COMM_WORLD, &A)) return -2;
    if (fmt != LIS_MATRIX_CSR && lis_matrix_set_type(A, fmt)) return -2;
    if (lis_vector_create(LIS_COMM_WORLD, &b) || lis_vector_create(LIS_COMM_WORLD, &x)) return -2;
    { const int q = quiet_begin(); const LIS_INT err = lis_input(A, b, x, (char *)path); quiet_end(q); if (err) return -100 - (int)err; }
    has[0] = !lis_vector_is_null(b); has[1] = !lis_vector_is_null(x);
    if (has[0] && bx_cap >= A->n) lis_vector_gather(b, bx);
    if (has[1] && bx_cap >= 2 * A->n) lis_vector_gather(x, bx + A->n);
    lis_vector_destroy(b); lis_vector_destroy(x);
    g_conv[h] = A; g_conv0[h] = NULL;
    return h;
}

/* row-wise assembly with lis_matrix_set_value (flag per entry), then lis_matrix_assemble */
EXPORT int shim_assemble_open(int n, int count, const int *rows, const int *cols, const double *vals, const int *flags, int fmt)
{
    int h;
    LIS_MATRIX A;
    for (h = 0; h < 16 && g_conv[h]; h++) ;
    if (h == 16) return -1;
    if (lis_matrix_create(LIS_COMM_WORLD, &A) || lis_matrix_set_size(A, 0, n)) return -2;
    if (fmt != LIS_MATRIX_CSR && lis_matrix_set_type(A, fmt)) return -2;
    for (int k = 0; k < count; k++) {
        const LIS_INT err = lis_matrix_set_value(flags[k], rows[k], cols[k], vals[k], A);
        if (err) return -100 - (int)err;
    }
    if (lis_matrix_assemble(A)) return -3;
    g_conv[h] = A; g_conv0[h] = NULL;
    return h;
}

EXPORT void shim_sort_id(int n, int *keys, double *vals) { lis_sort_id(0, n - 1, keys, vals); }

/* option parsing: returns solver, precon, maxiter, restart, storage, output, conv_cond, initx_zeros in o[8],
 * tol and ssor_omega in d[2] */
EXPORT int shim_parse_options(const char *text, int *o, double *d)
{
    LIS_SOLVER s;
    LIS_INT err = lis_solver_create(&s); if (err) return (int)err;
    err = lis_solver_set_option((char *)text, s);
    o[0] = s->options[LIS_OPTIONS_SOLVER]; o[1] = s->options[LIS_OPTIONS_PRECON]; o[2] = s->options[LIS_OPTIONS_MAXITER];
    o[3] = s->options[LIS_OPTIONS_RESTART]; o[4] = s->options[LIS_OPTIONS_STORAGE]; o[5] = s->options[LIS_OPTIONS_OUTPUT];
    o[6] = s->options[LIS_OPTIONS_CONV_COND]; o[7] = s->options[LIS_OPTIONS_INITGUESS_ZEROS];
    d[0] = s->params[LIS_PARAMS_RESID - LIS_OPTIONS_LEN]; d[1] = s->params[LIS_PARAMS_SSOR_OMEGA - LIS_OPTIONS_LEN];
    lis_solver_destroy(s);
    return (int)err;
}

/* dims: n, nnz, maxnzr, nnd, nr, bnr, bnc, bnnz, matrix_type, nc, ndz */
EXPORT int shim_convert_dims(int h, int *dims)
{
    LIS_MATRIX A = g_conv[h];
    dims[0] = A->n; dims[1] = A->nnz; dims[2] = A->maxnzr; dims[3] = A->nnd; dims[4] = A->nr;
    dims[5] = A->bnr; dims[6] = A->bnc; dims[7] = A->bnnz; dims[8] = A->matrix_type;
    dims[9] = A->nc; dims[10] = A->ndz;
    return 0;
}

/* which: 0 ptr, 1 index, 2 value, 3 row(perm), 4 bptr, 5 bindex, 6 col; count elements copied */
EXPORT int shim_convert_copy(int h, int which, void *dst, int count)
{
    LIS_MATRIX A = g_conv[h];
    const void *src = NULL;
    size_t sz = sizeof(LIS_INT);
    switch (which) {
    case 0: src = A->ptr; break;
    case 1: src = A->index; break;
    case 2: src = A->value; sz = sizeof(LIS_SCALAR); break;
    case 3: src = A->row; break;
    case 4: src = A->bptr; break;
    case 5: src = A->bindex; break;
    case 6: src = A->col; break;
    default: return -1;
    }
    if (src == NULL) return -2;
    memcpy(dst, src, sz * (size_t)count);
    return 0;
}

EXPORT int shim_convert_close(int h)
{
    lis_matrix_destroy(g_conv[h]);
    if (g_conv0[h]) lis_matrix_destroy(g_conv0[h]);
    g_conv[h] = NULL; g_conv0[h] = NULL;
    return 0;
}

/* ------------------------------------------------------------------ BLAS-1
 * op: 0 axpy(y+=a*x) 1 xpay(y=x+a*y) 2 axpyz(z=a*x+y) 3 scale(x=a*x) 4 copy(y=x) 5 set_all(x=a)
 *     6 pmul(z=x*y) 7 pdiv(z=x/y) 8 reciprocal(x=1/x) 9 abs 10 shift(x-=a) 11 swap
 *     20 dot 21 nrm2 22 nrm1 23 nrmi 24 sum
 * out_a / out_b receive the vectors the op wrote (may be NULL), scalar the reduction */
EXPORT int shim_vec_op(int op, int n, double alpha, const double *x, const double *y,
                       double *out_a, double *out_b, double *scalar)
{
    LIS_VECTOR vx = NULL, vy = NULL, vz = NULL;
    LIS_INT err;
    LIS_SCALAR s = 0.0;
    LIS_REAL r = 0.0;
    err = make_vec_n(n, x, &vx); if (err) return (int)err;
    err = make_vec_n(n, y, &vy); if (err) return (int)err;
    err = make_vec_n(n, NULL, &vz); if (err) return (int)err;
    LIS_VECTOR ra = NULL, rb = NULL;
    switch (op) {
    case 0: err = lis_vector_axpy(alpha, vx, vy); ra = vy; break;
    case 1: err = lis_vector_xpay(vx, alpha, vy); ra = vy; break;
    case 2: err = lis_vector_axpyz(alpha, vx, vy, vz); ra = vz; break;
    case 3: err = lis_vector_scale(alpha, vx); ra = vx; break;
    case 4: err = lis_vector_copy(vx, vy); ra = vy; break;
    case 5: err = lis_vector_set_all(alpha, vx); ra = vx; break;
    case 6: err = lis_vector_pmul(vx, vy, vz); ra = vz; break;
    case 7: err = lis_vector_pdiv(vx, vy, vz); ra = vz; break;
    case 8: err = lis_vector_reciprocal(vx); ra = vx; break;
    case 9: err = lis_vector_abs(vx); ra = vx; break;
    case 10: err = lis_vector_shift(alpha, vx); ra = vx; break;
    case 11: err = lis_vector_swap(vx, vy); ra = vx; rb = vy; break;
    case 20: err = lis_vector_dot(vx, vy, &s); break;
    case 21: err = lis_vector_nrm2(vx, &r); s = r; break;
    case 22: err = lis_vector_nrm1(vx, &r); s = r; break;
    case 23: err = lis_vector_nrmi(vx, &r); s = r; break;
    case 24: err = lis_vector_sum(vx, &s); break;
    default: err = LIS_ERR_ILL_ARG;
    }
    if (err) return (int)err;
    if (scalar) *scalar = s;
    if (ra && out_a && n > 0) { err = lis_vector_gather(ra, out_a); if (err) return (int)err; }
    if (rb && out_b && n > 0) { err = lis_vector_gather(rb, out_b); if (err) return (int)err; }
    lis_vector_destroy(vx); lis_vector_destroy(vy); lis_vector_destroy(vz);
    return 0;
}

/* mismatched lengths must fail with LIS_ERR_ILL_ARG (src/vector/lis_vector_opv.c:158-163) */
EXPORT int shim_vec_mismatch(int op)
{
    LIS_VECTOR a, b;
    LIS_SCALAR s;
    LIS_INT err;
    if (make_vec_n(8, NULL, &a) || make_vec_n(9, NULL, &b)) return -1;
    if (op == 0) err = lis_vector_axpy(1.0, a, b);
    else if (op == 1) err = lis_vector_xpay(a, 1.0, b);
    else if (op == 4) err = lis_vector_copy(a, b);
    else err = lis_vector_dot(a, b, &s);
    lis_vector_destroy(a); lis_vector_destroy(b);
    return (int)err;
}

/* ------------------------------------------------------------------ diagonal / preconditioner apply */
EXPORT int shim_get_diagonal(int fmt, int n, const int *ptr, const int *idx, const double *val, int bnr, int bnc, double *d)
{
    LIS_MATRIX A0, A;
    LIS_VECTOR v;
    LIS_INT err;
    err = make_csr(n, ptr, idx, val, 0, &A0); if (err) return (int)err;
    err = convert_to(A0, fmt, bnr, bnc, &A); if (err) return (int)err;
    err = make_vec(A, NULL, &v); if (err) return (int)err;
    err = lis_matrix_get_diagonal(A, v); if (err) return (int)err;
    err = lis_vector_gather(v, d); if (err) return (int)err;
    lis_vector_destroy(v); lis_matrix_destroy(A); lis_matrix_destroy(A0);
    return 0;
}

/* x = M^-1 b with the preconditioner selected by `options` ("-p ssor -ssor_omega 1.2" ...) */
static int psolve_any(int transposed, int n, const int *ptr, const int *idx, const double *val, const char *options,
                      const double *b, double *x);
EXPORT int shim_psolve(int n, const int *ptr, const int *idx, const double *val, const char *options,
                       const double *b, double *x)
{
    return psolve_any(0, n, ptr, idx, val, options, b, x);
}
/* x = M^-H b (what BiCG / BiCR call) */
EXPORT int shim_psolveh(int n, const int *ptr, const int *idx, const double *val, const char *options,
                        const double *b, double *x)
{
    return psolve_any(1, n, ptr, idx, val, options, b, x);
}
static int psolve_any(int transposed, int n, const int *ptr, const int *idx, const double *val, const char *options,
                      const double *b, double *x)
{
    LIS_MATRIX A;
    LIS_VECTOR vb, vx;
    LIS_SOLVER solver;
    LIS_PRECON precon;
    LIS_INT err;
    err = make_csr(n, ptr, idx, val, 0, &A); if (err) return (int)err;
    err = make_vec(A, b, &vb); if (err) return (int)err;
    err = make_vec(A, NULL, &vx); if (err) return (int)err;
    err = lis_solver_create(&solver); if (err) return (int)err;
    err = lis_solver_set_option((char *)options, solver); if (err) return (int)err;
    solver->A = A;
    err = lis_precon_create(solver, &precon); if (err) return (int)err;
    solver->precon = precon;
    err = transposed ? lis_psolveh(solver, vb, vx) : lis_psolve(solver, vb, vx); if (err) return (int)err;
    err = lis_vector_gather(vx, x); if (err) return (int)err;
    lis_precon_destroy(precon);
    solver->precon = NULL;
    lis_solver_destroy(solver);
    lis_vector_destroy(vb); lis_vector_destroy(vx); lis_matrix_destroy(A);
    return 0;
}

/* ------------------------------------------------------------------ lis_solve
 * out_i: iter, retcode(solver status), lis_solve return value, rhistory length
 * out_d: resid, time, itime, ptime */
EXPORT int shim_solve(int fmt, int n, const int *ptr, const int *idx, const double *val, const double *b,
                      double *x, const char *options, int *out_i, double *out_d, double *rhistory, int rh_cap)
{
    LIS_MATRIX A0, A;
    LIS_VECTOR vb, vx, vh;
    LIS_SOLVER solver;
    LIS_INT err, iter = 0, status = 0;
    LIS_REAL resid = 0.0;
    double time = 0, itime = 0, ptime = 0, pc = 0, pi = 0;
    err = make_csr(n, ptr, idx, val, 0, &A0); if (err) return (int)err;
    err = convert_to(A0, fmt, 0, 0, &A); if (err) return (int)err;
    lis_matrix_destroy(A0);
    err = make_vec(A, b, &vb); if (err) return (int)err;
    err = make_vec(A, x, &vx); if (err) return (int)err;
    err = lis_solver_create(&solver); if (err) return (int)err;
    err = lis_solver_set_option((char *)"-print mem", solver); if (err) return (int)err;
    err = lis_solver_set_option((char *)options, solver); if (err) return (int)err;
    { const int q = quiet_begin(); err = lis_solve(A, vb, vx, solver); quiet_end(q); }
    out_i[2] = (int)err;
    lis_solver_get_iter(solver, &iter);
    lis_solver_get_status(solver, &status);
    lis_solver_get_residualnorm(solver, &resid);
    lis_solver_get_timeex(solver, &time, &itime, &ptime, &pc, &pi);
    out_i[0] = (int)iter; out_i[1] = (int)status;
    out_d[0] = resid; out_d[1] = time; out_d[2] = itime; out_d[3] = ptime;
    out_i[3] = 0;
    if (!err && rhistory && rh_cap > 0) {
        int len = (int)iter + 1;
        if (status != LIS_SUCCESS) len--;
        if (len > rh_cap) len = rh_cap;
        if (len > 0) {
            if (make_vec_n(len, NULL, &vh) == 0) {
                lis_solver_get_rhistory(solver, vh);
                lis_vector_gather(vh, rhistory);
                lis_vector_destroy(vh);
                out_i[3] = len;
            }
        }
    }
    if (!err && n > 0) lis_vector_gather(vx, x);
    lis_solver_destroy(solver);
    lis_vector_destroy(vb); lis_vector_destroy(vx); lis_matrix_destroy(A);
    return (int)err;
}

/* ------------------------------------------------------------------ lis_esolve
 * out_i: iter[0], status, lis_esolve return value, rhistory length, subspace size
 * out_d: evalue0, resid[0]; x: initial vector in (used with -initx_ones false), eigenvector out;
 * evals/eresid/eiter: per-mode arrays for the subspace / Lanczos solvers (ss entries), else untouched */
EXPORT int shim_esolve(int fmt, int n, const int *ptr, const int *idx, const double *val, double *x, const char *options,
                       int *out_i, double *out_d, double *rhistory, int rh_cap, double *evals, double *eresid, int *eiter, int ecap)
{
    LIS_MATRIX A0, A;
    LIS_VECTOR vx, vh;
    LIS_ESOLVER esolver;
    LIS_INT err, iter = 0, status = 0, nesol = 0;
    LIS_REAL resid = 0.0;
    LIS_SCALAR evalue0 = 0.0;
    err = make_csr(n, ptr, idx, val, 0, &A0); if (err) return (int)err;
    err = convert_to(A0, fmt, 0, 0, &A); if (err) return (int)err;
    lis_matrix_destroy(A0);
    err = make_vec(A, x, &vx); if (err) return (int)err;
    err = lis_esolver_create(&esolver); if (err) return (int)err;
    err = lis_esolver_set_option((char *)"-eprint mem", esolver); if (err) return (int)err;
    err = lis_esolver_set_option((char *)options, esolver); if (err) return (int)err;
    { const int q = quiet_begin(); err = lis_esolve(A, vx, &evalue0, esolver); quiet_end(q); }
    out_i[2] = (int)err; out_i[3] = 0;
    if (!err) {
        lis_esolver_get_iter(esolver, &iter);
        lis_esolver_get_status(esolver, &status);
        lis_esolver_get_residualnorm(esolver, &resid);
        lis_esolver_get_esolver(esolver, &nesol);
        out_i[0] = (int)iter; out_i[1] = (int)status; out_i[4] = (int)esolver->options[LIS_EOPTIONS_SUBSPACE];
        out_d[0] = evalue0; out_d[1] = resid;
        if (rhistory && rh_cap > 0) {
            int len = (int)iter + 1;
            if (status != LIS_SUCCESS) len--;
            if (len > rh_cap) len = rh_cap;
            if (len > 0 && make_vec_n(len, NULL, &vh) == 0) {
                lis_esolver_get_rhistory(esolver, vh);
                lis_vector_gather(vh, rhistory);
                lis_vector_destroy(vh);
                out_i[3] = len;
            }
        }
        if (nesol == LIS_ESOLVER_SI || nesol == LIS_ESOLVER_LI || nesol == LIS_ESOLVER_AI)
            for (int m = 0; m < out_i[4] && m < ecap; m++) {
                LIS_SCALAR ev = 0.0; LIS_REAL er = 0.0; LIS_INT ei = 0;
                lis_esolver_get_specific_evalue(esolver, m, &ev);
                lis_esolver_get_specific_residualnorm(esolver, m, &er);
                lis_esolver_get_specific_iter(esolver, m, &ei);
                evals[m] = ev; eresid[m] = er; eiter[m] = (int)ei;
            }
        if (n > 0) lis_vector_gather(vx, x);
    }
    lis_esolver_destroy(esolver);
    lis_vector_destroy(vx); lis_matrix_destroy(A);
    return (int)err;
}

/* ------------------------------------------------------------------ synthetic matrices for bench.py
 * Rows [i0*m*n, i1*m*n) of the 7-point Poisson matrix of an l x m x n grid, lexicographic
 * ii = i*m*n + j*n + k (test/spmvtest3.c:142-157 values: 6 on the diagonal, -1 off it), global
 * column indices.  sorted != 0: ascending columns (what spmvtest3.c:192-195 leaves after
 * lis_sort_id); sorted == 0: test/test3.c:116-126 order (-mn, +mn, -n, +n, -1, +1, diagonal last).
 * pass ptr == NULL to get the entry count only.  Returns nnz. */
EXPORT long long shim_poisson7(int l, int m, int n, int i0, int i1, int sorted, int *ptr, int *idx, double *val)
{
    const long long mn = (long long)m * n;
    long long k = 0, row = 0;
    for (int i = i0; i < i1; i++)
        for (int j = 0; j < m; j++)
            for (int q = 0; q < n; q++, row++) {
                const long long ii = (long long)i * mn + (long long)j * n + q;
                if (ptr) ptr[row] = (int)k;
                if (!ptr) { k += 1 + (i > 0) + (i < l - 1) + (j > 0) + (j < m - 1) + (q > 0) + (q < n - 1); continue; }
                if (sorted) {
                    if (i > 0)     { idx[k] = (int)(ii - mn); val[k++] = -1.0; }
                    if (j > 0)     { idx[k] = (int)(ii - n);  val[k++] = -1.0; }
                    if (q > 0)     { idx[k] = (int)(ii - 1);  val[k++] = -1.0; }
                    idx[k] = (int)ii; val[k++] = 6.0;
                    if (q < n - 1) { idx[k] = (int)(ii + 1);  val[k++] = -1.0; }
                    if (j < m - 1) { idx[k] = (int)(ii + n);  val[k++] = -1.0; }
                    if (i < l - 1) { idx[k] = (int)(ii + mn); val[k++] = -1.0; }
                } else {
                    if (i > 0)     { idx[k] = (int)(ii - mn); val[k++] = -1.0; }
                    if (i < l - 1) { idx[k] = (int)(ii + mn); val[k++] = -1.0; }
                    if (j > 0)     { idx[k] = (int)(ii - n);  val[k++] = -1.0; }
                    if (j < m - 1) { idx[k] = (int)(ii + n);  val[k++] = -1.0; }
                    if (q > 0)     { idx[k] = (int)(ii - 1);  val[k++] = -1.0; }
                    if (q < n - 1) { idx[k] = (int)(ii + 1);  val[k++] = -1.0; }
                    idx[k] = (int)ii; val[k++] = 6.0;
                }
            }
    if (ptr) ptr[row] = (int)k;
    return k;
}

/* Rows [i0*m*n, i1*m*n) of the 27-point stencil of test/spmvtest3b.c:148-163 / test/test3b.c:113-135
 * on an l x m x n grid (26 on the diagonal, -1 elsewhere, the driver's loop order = ascending
 * columns), global column indices.  ptr == NULL: entry count only. */
EXPORT long long shim_poisson27(int l, int m, int n, int i0, int i1, int *ptr, int *idx, double *val)
{
    const long long mn = (long long)m * n;
    long long ctr = 0, row = 0;
    for (int i = i0; i < i1; i++)
        for (int j = 0; j < m; j++)
            for (int k = 0; k < n; k++, row++) {
                const long long ii = (long long)i * mn + (long long)j * n + k;
                if (ptr) ptr[row] = (int)ctr;
                for (int si = -1; si <= 1; si++) {
                    if (i + si < 0 || i + si >= l) continue;
                    for (int sj = -1; sj <= 1; sj++) {
                        if (j + sj < 0 || j + sj >= m) continue;
                        for (int sk = -1; sk <= 1; sk++) {
                            if (k + sk < 0 || k + sk >= n) continue;
                            if (ptr) {
                                const long long jj = ii + si * mn + (long long)sj * n + sk;
                                idx[ctr] = (int)jj;
                                val[ctr] = jj == ii ? 26.0 : -1.0;
                            }
                            ctr++;
                        }
                    }
                }
            }
    if (ptr) ptr[row] = (int)ctr;
    return ctr;
}

/* BASELINE.json config 4: rows [i0, i1) of a seeded unsymmetric banded
 * matrix with `per_row` stored entries per row -- the diagonal at a random position of the row and
 * per_row-1 off-diagonals in random storage order, columns within |i-j| <= band (reflected at the
 * ends), values uniform(-1,1), diagonal = dshift + dom * sum|off-diagonals| (dom = 1, dshift = 1:
 * strictly dominant; smaller dom: harder).  Every row is generated from its own counter-based
 * stream (splitmix64 of seed and row number), so any row partition yields the same matrix. */
static unsigned long long shim_sm64(unsigned long long *s)
{
    unsigned long long z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
EXPORT long long shim_banded_rows(int n, int i0, int i1, int per_row, int band, double dom, double dshift,
                                  unsigned long long seed, int *ptr, int *idx, double *val)
{
    long long ctr = 0;
    if (band > n - 1) band = n - 1;
    for (int i = i0; i < i1; i++) {
        unsigned long long s = seed ^ ((unsigned long long)(i + 1) * 0xD1B54A32D192ED03ull);
        const int pos = (int)(shim_sm64(&s) % (unsigned long long)per_row);
        double sumabs = 0.0;
        ptr[i - i0] = (int)ctr;
        for (int q = 0; q < per_row; q++, ctr++) {
            if (q == pos) { idx[ctr] = i; val[ctr] = 0.0; continue; }
            const long long off = 1 + (long long)(shim_sm64(&s) % (unsigned long long)(band > 0 ? band : 1));
            long long col = (shim_sm64(&s) & 1) ? i + off : i - off;
            if (col < 0 || col >= n) col = 2LL * i - col;             /* reflect at the ends */
            if (col < 0) col = 0;
            if (col >= n) col = n - 1;
            if (col == i) col = (i + 1) % n;
            const double v = (double)(shim_sm64(&s) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
            idx[ctr] = (int)col; val[ctr] = v;
            sumabs += v < 0 ? -v : v;
        }
        val[ptr[i - i0] + pos] = dshift + dom * sumabs;
    }
    ptr[i1 - i0] = (int)ctr;
    return ctr;
}

/* ------------------------------------------------------------------ bench handles
 * One matrix + x + y kept alive across steps (bench.py): open once, then time steps.
 * step_e2e is what a user with HOST buffers does per product: scatter x in, lis_matvec,
 * gather y out.  Public API only, so the same code times the reference's CPU path. */
static struct { LIS_MATRIX A; LIS_VECTOR x, y; } g_mv[8];

EXPORT int shim_mv_open(int fmt, int n, int *ptr, int *idx, double *val, int bnr, int bnc, int adopt)
{
    int h;
    LIS_MATRIX A0 = NULL, A = NULL;
    LIS_INT err;
    for (h = 0; h < 8 && g_mv[h].A; h++) ;
    if (h == 8) return -1;
    if (adopt) {
        /* take the caller's malloc'ed arrays as they are (lis_matrix_set_csr semantics) */
        err = lis_matrix_create(LIS_COMM_WORLD, &A0); if (err) return -2;
        err = lis_matrix_set_size(A0, 0, n); if (err) return -2;
        err = lis_matrix_set_csr(ptr[n], ptr, idx, val, A0); if (err) return -2;
        err = lis_matrix_assemble(A0); if (err) return -2;
    } else {
        if (make_csr(n, ptr, idx, val, 0, &A0)) return -2;
    }
    if (fmt == LIS_MATRIX_CSR) A = A0;
    else {
        if (convert_to(A0, fmt, bnr, bnc, &A)) return -3;
        lis_matrix_destroy(A0);
    }
    if (make_vec(A, NULL, &g_mv[h].x) || make_vec(A, NULL, &g_mv[h].y)) return -4;
    g_mv[h].A = A;
    return h;
}

/* row-partitioned: this rank hands over its n_local rows with GLOBAL column indices
 * (lis_matrix_set_size(A, n_local, 0), like an MPI rank of the reference would) */
EXPORT int shim_mv_open_dist(int fmt, int n_local, int *ptr, int *idx, double *val, int adopt)
{
    int h;
    LIS_MATRIX A0 = NULL, A = NULL;
    LIS_INT err, *p, *ix;
    LIS_SCALAR *v;
    const int nnz = ptr[n_local];
    for (h = 0; h < 8 && g_mv[h].A; h++) ;
    if (h == 8) return -1;
    err = lis_matrix_create(LIS_COMM_WORLD, &A0); if (err) return -2;
    err = lis_matrix_set_size(A0, n_local, 0); if (err) return -2;
    if (adopt) { p = ptr; ix = idx; v = val; }          /* malloc'ed by the caller, handed over */
    else {
        p = (LIS_INT *)malloc(sizeof(LIS_INT) * ((size_t)n_local + 1));
        ix = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(nnz > 0 ? nnz : 1));
        v = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(nnz > 0 ? nnz : 1));
        if (!p || !ix || !v) return -2;
        memcpy(p, ptr, sizeof(LIS_INT) * ((size_t)n_local + 1));
        memcpy(ix, idx, sizeof(LIS_INT) * (size_t)nnz);
        memcpy(v, val, sizeof(LIS_SCALAR) * (size_t)nnz);
    }
    err = lis_matrix_set_csr(nnz, p, ix, v, A0); if (err) return -2;
    err = lis_matrix_assemble(A0); if (err) return -2;
    if (fmt == LIS_MATRIX_CSR) A = A0;
    else {
        if (convert_to(A0, fmt, 0, 0, &A)) return -3;
        lis_matrix_destroy(A0);
    }
    if (make_vec(A, NULL, &g_mv[h].x) || make_vec(A, NULL, &g_mv[h].y)) return -4;
    g_mv[h].A = A;
    return h;
}

/* local slices in and out (each rank passes / receives only its own rows) */
EXPORT int shim_mv_set_x_local(int h, double *x_local)
{
    LIS_VECTOR x = g_mv[h].x;
    return (int)lis_vector_set_values2(LIS_INS_VALUE, x->is, x->n, x_local, x);
}
EXPORT int shim_mv_get_y_local(int h, double *y_local)
{
    LIS_VECTOR y = g_mv[h].y;
    return (int)lis_vector_get_values(y, y->is, y->n, y_local);
}
EXPORT int shim_mv_matvec(int h) { return (int)lis_matvec(g_mv[h].A, g_mv[h].x, g_mv[h].y); }
/* `iters` products enqueued back to back, one synchronisation at the end (lis_b200 only; the reference is synchronous anyway) */
EXPORT int shim_mv_matvec_queue(int h, int iters)
{
    LIS_INT err = 0;
#ifdef LIS_B200_LIS_H
    for (int k = 0; k < iters && !err; k++) err = lis_b200_matvec_async(g_mv[h].A, g_mv[h].x, g_mv[h].y);
    if (!err) err = lis_b200_sync();
#else
    for (int k = 0; k < iters && !err; k++) err = lis_matvec(g_mv[h].A, g_mv[h].x, g_mv[h].y);
#endif
    return (int)err;
}
EXPORT int shim_mv_p2p_release(int h)
{
#ifdef LIS_B200_LIS_H
    return (int)lis_b200_p2p_release(g_mv[h].A);
#else
    (void)h; return 0;
#endif
}
EXPORT int shim_mv_matvech(int h) { return (int)lis_matvech(g_mv[h].A, g_mv[h].x, g_mv[h].y); }
EXPORT int shim_mv_dot_xy(int h, double *out) { LIS_SCALAR s = 0; LIS_INT e = lis_vector_dot(g_mv[h].x, g_mv[h].y, &s); *out = s; return (int)e; }

EXPORT int shim_mv_step_e2e(int h, double *host_x, double *host_y)
{
    LIS_INT err = lis_vector_scatter(host_x, g_mv[h].x); if (err) return (int)err;
    err = lis_matvec(g_mv[h].A, g_mv[h].x, g_mv[h].y); if (err) return (int)err;
    return (int)lis_vector_gather(g_mv[h].y, host_y);
}

/* the same step through lis_b200's overlapped host-buffer product (lis_b200 builds only; the
 * reference has one address space and nothing to overlap) */
EXPORT int shim_mv_step_e2e_pipelined(int h, double *host_x, double *host_y)
{
#ifdef LIS_B200_LIS_H
    return (int)lis_b200_matvec_host(g_mv[h].A, host_x, g_mv[h].x, g_mv[h].y, host_y);
#else
    return shim_mv_step_e2e(h, host_x, host_y);
#endif
}
EXPORT int shim_mv_host_plan(int h, int cap, int *rows, int *need)
{
#ifdef LIS_B200_LIS_H
    return (int)lis_b200_matvec_host_plan(g_mv[h].A, cap, rows, need);
#else
    (void)h; (void)cap; (void)rows; (void)need;
    return 0;
#endif
}
EXPORT int shim_mv_get_xy(int h, double *x_out, double *y_out)
{
    LIS_INT err = lis_vector_gather(g_mv[h].x, x_out);
    if (!err) err = lis_vector_gather(g_mv[h].y, y_out);
    return (int)err;
}

/* `iters` products with resident vectors, wall seconds by lis_wtime (the drivers' own timing) */
EXPORT int shim_mv_run(int h, int iters, double *seconds, double *nrm2)
{
    LIS_INT err = 0;
    LIS_REAL nr = 0.0;
    const double t0 = lis_wtime();
    for (int k = 0; k < iters && !err; k++) err = lis_matvec(g_mv[h].A, g_mv[h].x, g_mv[h].y);
    *seconds = lis_wtime() - t0;
    if (!err) err = lis_vector_nrm2(g_mv[h].y, &nr);
    *nrm2 = nr;
    return (int)err;
}

EXPORT int shim_mv_set_x(int h, double *host_x) { return (int)lis_vector_scatter(host_x, g_mv[h].x); }

/* lis_solve on the handle's matrix with b = A*1 (test/test3.c:150-151); returns like shim_solve */
EXPORT int shim_mv_solve(int h, const char *options, int *out_i, double *out_d, double *host_x)
{
    LIS_MATRIX A = g_mv[h].A;
    LIS_VECTOR u, b, x;
    LIS_SOLVER solver;
    LIS_INT err, iter = 0, status = 0;
    LIS_REAL resid = 0.0;
    double time = 0, itime = 0, ptime = 0, pc = 0, pi = 0;
    if (make_vec(A, NULL, &u) || make_vec(A, NULL, &b) || make_vec(A, NULL, &x)) return -1;
    err = lis_vector_set_all(1.0, u); if (err) return (int)err;
    err = lis_matvec(A, u, b); if (err) return (int)err;
    err = lis_solver_create(&solver); if (err) return (int)err;
    err = lis_solver_set_option((char *)options, solver); if (err) return (int)err;
    const double t0 = lis_wtime();
    err = lis_solve(A, b, x, solver);
    if (!err && host_x) err = lis_vector_gather(x, host_x);
    out_d[4] = lis_wtime() - t0;
    out_i[2] = (int)err;
    lis_solver_get_iter(solver, &iter);
    lis_solver_get_status(solver, &status);
    lis_solver_get_residualnorm(solver, &resid);
    lis_solver_get_timeex(solver, &time, &itime, &ptime, &pc, &pi);
    out_i[0] = (int)iter; out_i[1] = (int)status;
    out_d[0] = resid; out_d[1] = time; out_d[2] = itime; out_d[3] = ptime;
    lis_solver_destroy(solver);
    lis_vector_destroy(u); lis_vector_destroy(b); lis_vector_destroy(x);
    return (int)err;
}

/* lis_solve on the handle's matrix with b = A*1, x0 = 0 (test/test3.c:150-151); residual history and the
 * times of lis_solver_get_timeex returned; out_d: resid, time, itime, ptime, wall, max|x-1| over the local rows */
EXPORT int shim_mv_solve_ones(int h, const char *options, int *out_i, double *out_d, double *rhistory, int rh_cap)
{
    LIS_MATRIX A = g_mv[h].A;
    LIS_VECTOR u, b, x;
    LIS_SOLVER solver;
    LIS_INT err, iter = 0, status = 0;
    LIS_REAL resid = 0.0, dev = 0.0;
    double time = 0, itime = 0, ptime = 0, pc = 0, pi = 0;
    if (make_vec(A, NULL, &u) || make_vec(A, NULL, &b) || make_vec(A, NULL, &x)) return -1;
    err = lis_vector_set_all(1.0, u); if (err) return (int)err;
    err = lis_matvec(A, u, b); if (err) return (int)err;
    err = lis_solver_create(&solver); if (err) return (int)err;
    err = lis_solver_set_option((char *)"-print mem", solver); if (err) return (int)err;
    err = lis_solver_set_option((char *)options, solver); if (err) return (int)err;
    const double t0 = lis_wtime();
    { const int q = quiet_begin(); err = lis_solve(A, b, x, solver); quiet_end(q); }
    out_d[4] = lis_wtime() - t0;
    out_i[2] = (int)err;
    lis_solver_get_iter(solver, &iter);
    lis_solver_get_status(solver, &status);
    lis_solver_get_residualnorm(solver, &resid);
    lis_solver_get_timeex(solver, &time, &itime, &ptime, &pc, &pi);
    out_i[0] = (int)iter; out_i[1] = (int)status;
    out_d[0] = resid; out_d[1] = time; out_d[2] = itime; out_d[3] = ptime;
    int len = (int)iter + 1 - (status != LIS_SUCCESS ? 1 : 0);
    if (len > rh_cap) len = rh_cap;
    out_i[3] = 0;
    if (!err && rhistory && len > 0 && solver->rhistory) { memcpy(rhistory, solver->rhistory, sizeof(double) * (size_t)len); out_i[3] = len; }
    if (!err) { err = lis_vector_axpy(-1.0, x, u); if (!err) err = lis_vector_nrmi(u, &dev); }      /* u = 1 - x */
    out_d[5] = dev;
    lis_solver_destroy(solver);
    lis_vector_destroy(u); lis_vector_destroy(b); lis_vector_destroy(x);
    return (int)err;
}

/* lis_solve with a caller-given right-hand side slice; x_local and rhistory returned */
EXPORT int shim_mv_solve_b(int h, const char *options, const double *b_local, double *x_local, int *out_i, double *out_d,
                           double *rhistory, int rh_cap)
{
    LIS_MATRIX A = g_mv[h].A;
    LIS_VECTOR b, x;
    LIS_SOLVER solver;
    LIS_INT err, iter = 0, status = 0;
    LIS_REAL resid = 0.0;
    if (make_vec(A, NULL, &b) || make_vec(A, NULL, &x)) return -1;
    err = lis_vector_set_values2(LIS_INS_VALUE, b->is, b->n, (LIS_SCALAR *)b_local, b); if (err) return (int)err;
    err = lis_solver_create(&solver); if (err) return (int)err;
    err = lis_solver_set_option((char *)"-print mem", solver); if (err) return (int)err;
    err = lis_solver_set_option((char *)options, solver); if (err) return (int)err;
    { const int q = quiet_begin(); err = lis_solve(A, b, x, solver); quiet_end(q); }
    out_i[2] = (int)err;
    lis_solver_get_iter(solver, &iter);
    lis_solver_get_status(solver, &status);
    lis_solver_get_residualnorm(solver, &resid);
    out_i[0] = (int)iter; out_i[1] = (int)status; out_d[0] = resid;
    { double time = 0, itime = 0, ptime = 0, pc = 0, pi = 0;          /* out_d needs 4 slots */
      lis_solver_get_timeex(solver, &time, &itime, &ptime, &pc, &pi);
      out_d[1] = time; out_d[2] = itime; out_d[3] = ptime; }
    int len = (int)iter + 1 - (status != LIS_SUCCESS ? 1 : 0);
    if (len > rh_cap) len = rh_cap;
    out_i[3] = 0;
    if (!err && len > 0 && solver->rhistory) { memcpy(rhistory, solver->rhistory, sizeof(double) * (size_t)len); out_i[3] = len; }
    if (!err) err = lis_vector_get_values(x, x->is, x->n, x_local);
    lis_solver_destroy(solver);
    lis_vector_destroy(b); lis_vector_destroy(x);
    return (int)err;
}

/* a new handle holding handle `src`'s matrix converted to storage format `fmt`
 * (lis_matrix_convert, timed with lis_wtime); x is copied from the source handle */
EXPORT int shim_mv_convert(int src, int fmt, int bnr, int bnc, double *seconds)
{
    int h;
    LIS_MATRIX A = NULL;
    for (h = 0; h < 8 && g_mv[h].A; h++) ;
    if (h == 8) return -1;
    const double t0 = lis_wtime();
    if (convert_to(g_mv[src].A, fmt, bnr, bnc, &A)) return -3;
    if (seconds) *seconds = lis_wtime() - t0;
    if (make_vec(A, NULL, &g_mv[h].x) || make_vec(A, NULL, &g_mv[h].y)) return -4;
    if (lis_vector_copy(g_mv[src].x, g_mv[h].x)) return -5;
    g_mv[h].A = A;
    return h;
}

EXPORT int shim_mv_close(int h)
{
    lis_vector_destroy(g_mv[h].x); lis_vector_destroy(g_mv[h].y); lis_matrix_destroy(g_mv[h].A);
    g_mv[h].A = NULL; g_mv[h].x = NULL; g_mv[h].y = NULL;
    return 0;
}
