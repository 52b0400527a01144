/*
 * lis_precon.c -- preconditioner setup and apply for the hot path: none, Jacobi, SSOR and
 * user-registered ones (the reference's plugin API, lis_precon_register).
 *
 * Reference: src/precon/lis_precon.c:58-159 (dispatch tables, create), :410-460 (register),
 * lis_precon_jacobi.c:60-147, lis_precon_ssor.c:57-115, and the triangular sweep
 * lis_matrix_solve_csr(..., LIS_MATRIX_SSOR), src/matrix/lis_matrix_csr.c:1572-1630.
 *
 * SSOR on the GPU.  The reference's sweep is a sequential dependency chain per block, where
 * a block is the row range of one OpenMP thread (LIS_GET_ISIE(thread, nthreads, n)) and
 * couplings that leave the block are dropped; the serial build is the one-block case.  Here
 * the block count is the emulated thread count (`-omp_num_threads N` on the command line,
 * lis_initialize; default 1 = the serial reference), and inside the blocks rows are
 * level-scheduled on the host once per matrix: all rows of a level are independent, one
 * kernel launch per level, each row still subtracts its products in storage order, so the
 * result is bit-identical to the CPU sweep with the same block partition.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lis_device.h"
#include "lis_host.h"
#include "lis_b200_kernels.h"

/* ------------------------------------------------------------------ registry */
typedef struct {
    LIS_INT precon_type;
    char name[LIS_PRECONNAME_MAX + 1];
    LIS_PRECON_CREATE_XXX pcreate;
    LIS_PSOLVE_XXX psolve;
    LIS_PSOLVEH_XXX psolveh;
} lis_precon_reg_t;

static lis_precon_reg_t *g_reg = NULL;
static LIS_INT g_reg_type = LIS_PRECON_TYPE_USERDEF;

LIS_INT lis_host_precon_type_end(void) { return g_reg_type; }

LIS_INT lis_host_precon_lookup(const char *name)
{
    for (LIS_INT i = 0; g_reg && i < g_reg_type - LIS_PRECON_TYPE_USERDEF; i++)
        if (strcmp(name, g_reg[i].name) == 0) return g_reg[i].precon_type;
    return -1;
}

LIS_INT lis_precon_register(char *name, LIS_PRECON_CREATE_XXX pcreate, LIS_PSOLVE_XXX psolve, LIS_PSOLVEH_XXX psolveh)
{
    if (g_reg == NULL) {
        g_reg = (lis_precon_reg_t *)lis_calloc(LIS_PRECON_REGISTER_MAX * sizeof(lis_precon_reg_t), "lis_precon_register::top");
        if (g_reg == NULL) { LIS_SETERR_MEM(sizeof(lis_precon_reg_t)); return LIS_OUT_OF_MEMORY; }
    }
    const LIS_INT k = g_reg_type - LIS_PRECON_TYPE_USERDEF;
    if (k == LIS_PRECON_REGISTER_MAX) {
        LIS_SETERR(LIS_FAILS, "lis_precon_resister is max\n");
        return LIS_FAILS;
    }
    g_reg[k].pcreate = pcreate; g_reg[k].psolve = psolve; g_reg[k].psolveh = psolveh;
    g_reg[k].precon_type = g_reg_type;
    strncpy(g_reg[k].name, name, LIS_PRECONNAME_MAX);
    g_reg[k].name[LIS_PRECONNAME_MAX] = '\0';
    for (char *p = g_reg[k].name; *p; p++) if (*p >= 'A' && *p <= 'Z') *p = (char)(*p - 'A' + 'a');
    g_reg_type++;
    return LIS_SUCCESS;
}

LIS_INT lis_precon_register_free(void)
{
    if (g_reg) { lis_free(g_reg); g_reg = NULL; }
    g_reg_type = LIS_PRECON_TYPE_USERDEF;
    return LIS_SUCCESS;
}

LIS_INT lis_host_ssor_prepare(LIS_MATRIX A);
static double setup_tick(const char *what, double t0);

/* WD = 1/(scale*D), remembered under `tag` (A->use_wd) like the reference does
 * (lis_precon_ssor.c:79-90, lis_solver_gs.c, lis_solver_sor.c) */
LIS_INT lis_host_set_wd(LIS_MATRIX A, LIS_SCALAR scale, int do_scale, LIS_INT tag)
{
    LIS_INT err;
    if (A->use_wd == tag) return LIS_SUCCESS;
    if (!A->WD) {
        err = lis_host_diag_create(A, &A->WD);
        if (err) return err;
    }
    for (LIS_INT i = 0; i < A->n; i++) {
        LIS_SCALAR v = A->D->value[i];
        if (do_scale) v = scale * v;
        A->WD->value[i] = 1.0 / v;
    }
    A->use_wd = tag;
    return lisd_matrix_refresh_wd(A);
}

/* ------------------------------------------------------------------ create / destroy */
static LIS_INT create_none(LIS_SOLVER solver, LIS_PRECON precon) { (void)solver; (void)precon; return LIS_SUCCESS; }

/* D = 1/diag(A): src/precon/lis_precon_jacobi.c:60-86 */
static LIS_INT create_jacobi(LIS_SOLVER solver, LIS_PRECON precon)
{
    LIS_INT err = lis_vector_duplicate(solver->A, &precon->D);
    if (err) return err;
    err = lis_matrix_get_diagonal(solver->A, precon->D);
    if (err) return err;
    return lis_vector_reciprocal(precon->D);
}

/* split A in place, WD = 1/(omega*D): src/precon/lis_precon_ssor.c:57-95,
 * src/matrix/lis_matrix_diag.c:663-672 (scale: value = alpha*value), :775-783 (inverse: 1.0/value) */
static LIS_INT create_ssor(LIS_SOLVER solver, LIS_PRECON precon)
{
    LIS_MATRIX A = solver->A;
    const LIS_SCALAR w = solver->params[LIS_PARAMS_SSOR_OMEGA - LIS_OPTIONS_LEN];
    LIS_INT err = lis_matrix_convert_self(solver);
    if (err) return err;
    A = solver->A;
    if (A->matrix_type != LIS_MATRIX_CSR) {
        /* The reference splits and sweeps in the storage format itself (lis_matrix_split_<fmt>, lis_matrix_solve_<fmt>).
         * For the scalar formats that is the same D, L, U and the same sweep as on the CSR form of the matrix, so the
         * sweeps run on a private CSR copy while the products stay in the chosen format (unsplit order: the
         * reference's products switch to the D + L + U order once split -- envelope parity, not bits).  The block
         * formats are a different preconditioner there (block SSOR on the dense diagonal blocks): not offered. */
        LIS_MATRIX C;
        if (A->matrix_type == LIS_MATRIX_BSR || A->matrix_type == LIS_MATRIX_BSC || A->matrix_type == LIS_MATRIX_VBR) {
            LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "SSOR on block storage (BSR/BSC/VBR) is block SSOR in Lis and is not available; use a scalar format\n");
            return LIS_ERR_NOT_IMPLEMENTED;
        }
        err = lis_matrix_duplicate(A, &C);
        if (err) return err;
        err = lis_matrix_set_type(C, LIS_MATRIX_CSR);
        if (!err) err = lis_matrix_convert(A, C);
        if (!err) err = lis_matrix_split(C);
        if (!err) err = lis_host_set_wd(C, w, 1, LIS_SOLVER_SOR);
        if (!err) err = lis_host_ssor_prepare(C);
        if (err) { lis_matrix_destroy(C); return err; }
        precon->A = C;
        precon->is_copy = LIS_TRUE;
        return LIS_SUCCESS;
    }
    double tk = setup_tick(NULL, 0.0);
    err = lis_matrix_split(A);
    if (err) return err;
    tk = setup_tick("split", tk);
    err = lis_host_set_wd(A, w, 1, LIS_SOLVER_SOR);
    if (err) return err;
    precon->A = A;
    precon->is_copy = LIS_FALSE;
    err = lis_host_ssor_prepare(A);       /* upload D/L/U and build the level schedule now, not in the first sweep */
    setup_tick("WD + mirror + schedule", tk);
    return err;
}

/* -p hybrid: M^-1 b = a few steps of another solver on the same matrix (src/precon/lis_precon_hybrid.c:54-135).
 * The inner solver takes -hybrid_i / _maxiter / _tol / _omega / _ell / _restart / _p, runs its loop directly
 * (no lis_solve front end: no scaling, no initial-vector copy beyond the one below) and keeps its work vectors,
 * iterate and preconditioner for the lifetime of the outer preconditioner. */
static LIS_INT create_hybrid(LIS_SOLVER solver, LIS_PRECON precon)
{
    LIS_SOLVER ps;
    LIS_PRECON pp = NULL;
    LIS_VECTOR xx = NULL;
    LIS_INT (*work)(LIS_SOLVER), (*run)(LIS_SOLVER);
    LIS_INT err = lis_solver_create(&ps);
    if (err) return err;
    ps->params[LIS_PARAMS_RESID - LIS_OPTIONS_LEN] = solver->params[LIS_PARAMS_PRESID - LIS_OPTIONS_LEN];
    ps->params[LIS_PARAMS_SSOR_OMEGA - LIS_OPTIONS_LEN] = solver->params[LIS_PARAMS_POMEGA - LIS_OPTIONS_LEN];
    ps->options[LIS_OPTIONS_MAXITER] = solver->options[LIS_OPTIONS_PMAXITER];
    ps->options[LIS_OPTIONS_ELL] = solver->options[LIS_OPTIONS_PELL];
    ps->options[LIS_OPTIONS_RESTART] = solver->options[LIS_OPTIONS_PRESTART];
    ps->options[LIS_OPTIONS_OUTPUT] = 0;
    ps->options[LIS_OPTIONS_SOLVER] = solver->options[LIS_OPTIONS_PSOLVER];
    ps->options[LIS_OPTIONS_PRECON] = solver->options[LIS_OPTIONS_PPRECON];
    ps->options[LIS_OPTIONS_INITGUESS_ZEROS] = solver->options[LIS_OPTIONS_INITGUESS_ZEROS];
    ps->options[LIS_OPTIONS_PRECISION] = solver->options[LIS_OPTIONS_PRECISION];
    ps->A = solver->A; ps->Ah = solver->Ah; ps->precision = solver->precision;
    err = lis_host_solver_entry(ps->options[LIS_OPTIONS_SOLVER], &work, &run);
    if (!err && ps->options[LIS_OPTIONS_PRECON] == LIS_PRECON_TYPE_HYBRID) { LIS_SETERR_IMP; err = LIS_ERR_NOT_IMPLEMENTED; }
    if (!err) err = lis_vector_duplicate(solver->A, &xx);
    if (!err) {
        ps->rhistory = (LIS_REAL *)lis_malloc(((size_t)ps->options[LIS_OPTIONS_MAXITER] + 2 + 64) * sizeof(LIS_REAL), "lis_precon_create_hybrid::rhistory");
        if (ps->rhistory == NULL) { LIS_SETERR_MEM(ps->options[LIS_OPTIONS_MAXITER]); err = LIS_OUT_OF_MEMORY; }
    }
    if (!err) err = lis_precon_create(ps, &pp);
    if (!err) err = work(ps);
    if (err) { if (pp) lis_precon_destroy(pp); if (xx) lis_vector_destroy(xx); lis_solver_destroy(ps); return err; }
    ps->x = xx;
    ps->precon = pp;
    precon->solver = ps;
    return LIS_SUCCESS;
}

/* src/precon/lis_precon_hybrid.c:138-200 */
LIS_INT lis_psolve_hybrid(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    LIS_SOLVER ps = solver->precon->solver;
    LIS_INT (*work)(LIS_SOLVER), (*run)(LIS_SOLVER);
    LIS_INT err = lis_host_solver_entry(ps->options[LIS_OPTIONS_SOLVER], &work, &run);
    if (err) return err;
    ps->b = b;
    err = ps->options[LIS_OPTIONS_INITGUESS_ZEROS] ? lisd_set_all(0.0, ps->x) : lisd_copy(b, ps->x);
    if (err) return err;
    err = run(ps);                                      /* MAXITER after -hybrid_maxiter steps is the normal way out */
    if (err == LIS_ERR_DEVICE || err == LIS_ERR_OUT_OF_MEMORY || err == LIS_ERR_NOT_IMPLEMENTED || err == LIS_ERR_ILL_ARG) return err;
    return lisd_copy(ps->x, x);
}

/* -p is (I+S, src/precon/lis_precon_is.c:54-100 and :417-460) at its default level with a non-stationary solver:
 * M^-1 = I - alpha*S, S = the first (is_m + 1) stored entries of every row of the strict upper part.  The matrix is
 * brought to CSR and split like there (so the products switch to the D + L + U order); S is kept as a small CSR
 * matrix of its own, so the apply is one product on the CSR kernel and one axpyz:
 *   y = x - alpha * (S x)     ->  t = S x ;  y = (-alpha)*t + x      (same bits: the sign change is exact)
 * The transposed apply uses S^T through lis_matvech: it sums S^T b and then scales, where the reference updates
 * y[jj] -= w*u*t in place -- equal to rounding only (the forward apply is bit for bit).  Level 0 (which rewrites the
 * system as (I+S)A) and the stationary solvers are not carried over. */
static LIS_INT create_is(LIS_SOLVER solver, LIS_PRECON precon)
{
    LIS_MATRIX A = solver->A;
    const LIS_INT nsol = solver->options[LIS_OPTIONS_SOLVER];
    LIS_INT err;
    const LIS_INT st = solver->options[LIS_OPTIONS_STORAGE];
    if (solver->options[LIS_OPTIONS_ISLEVEL] == 0 || (nsol >= LIS_SOLVER_JACOBI && nsol <= LIS_SOLVER_SOR) || A->nprocs > 1 ||
        (st > 0 && st != LIS_MATRIX_CSR)) {
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "-p is: only -is_level != 0 with a Krylov solver, CSR storage and one process is available\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    if (A->matrix_type != LIS_MATRIX_CSR) {            /* the reference converts the solver's matrix to CSR in place here */
        LIS_MATRIX B;
        err = lis_matrix_duplicate(A, &B);
        if (err) return err;
        lis_matrix_set_type(B, LIS_MATRIX_CSR);
        err = lis_matrix_convert(A, B);
        if (err) { lis_matrix_destroy(B); return err; }
        lis_host_matrix_adopt(A, B);
    }
    err = lis_matrix_split(A);
    if (err) return err;
    precon->work = (LIS_VECTOR *)lis_calloc(sizeof(LIS_VECTOR), "lis_precon_create_is::work");
    if (precon->work == NULL) { LIS_SETERR_MEM(sizeof(LIS_VECTOR)); return LIS_OUT_OF_MEMORY; }
    err = lis_vector_duplicate(A, &precon->work[0]);
    if (!err) precon->worklen = 1;
    return err;
}

/* S is cut out of the split matrix at the first apply: lis_solve scales the system to a unit diagonal AFTER the
 * preconditioner is created (src/solver/lis_solver.c:613-641), and the reference's apply reads A->U as it is then */
static LIS_INT is_build(LIS_SOLVER solver, LIS_PRECON precon)
{
    LIS_MATRIX A = solver->A, S = NULL;
    LIS_INT err;
    if (A->matrix_type != LIS_MATRIX_CSR || !A->is_splited || A->U == NULL) {
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "-p is needs the solver's matrix in (split) CSR storage: do not combine it with -storage\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    const LIS_INT n = A->n, m = solver->options[LIS_OPTIONS_M] + 1;
    LIS_INT nnz = 0, *ptr, *index;
    LIS_SCALAR *value;
    for (LIS_INT i = 0; i < n; i++) { const LIS_INT len = A->U->ptr[i + 1] - A->U->ptr[i]; nnz += len < m ? len : m; }
    err = lis_matrix_malloc_csr(n, nnz, &ptr, &index, &value);
    if (err) return err;
    ptr[0] = 0;
    for (LIS_INT i = 0, k = 0; i < n; i++) {
        const LIS_INT len = A->U->ptr[i + 1] - A->U->ptr[i], take = len < m ? len : m;
        for (LIS_INT j = 0; j < take; j++, k++) { index[k] = A->U->index[A->U->ptr[i] + j]; value[k] = A->U->value[A->U->ptr[i] + j]; }
        ptr[i + 1] = k;
    }
    err = lis_matrix_create(A->comm, &S);
    if (!err) err = lis_matrix_set_size(S, n, 0);
    if (!err) err = lis_matrix_set_csr(nnz, ptr, index, value, S);
    if (err) { lis_free2(3, ptr, index, value); if (S) lis_matrix_destroy(S); return err; }
    err = lis_matrix_assemble(S);
    if (err) { lis_matrix_destroy(S); return err; }
    precon->Ah = S;
    return LIS_SUCCESS;
}

LIS_INT lis_psolve_is(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    LIS_PRECON precon = solver->precon;
    LIS_INT err = precon->Ah ? LIS_SUCCESS : is_build(solver, precon);
    if (!err) err = lisd_matvec(precon->Ah, b, precon->work[0]);
    if (err) return err;
    return lisd_axpyz(-solver->params[LIS_PARAMS_ALPHA - LIS_OPTIONS_LEN], precon->work[0], b, x);
}

LIS_INT lis_psolveh_is(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    LIS_PRECON precon = solver->precon;
    LIS_INT err = precon->Ah ? LIS_SUCCESS : is_build(solver, precon);
    if (!err) err = lisd_matvech(precon->Ah, b, precon->work[0]);
    if (err) return err;
    return lisd_axpyz(-solver->params[LIS_PARAMS_ALPHA - LIS_OPTIONS_LEN], precon->work[0], b, x);
}

static LIS_INT create_unsupported(LIS_SOLVER solver, LIS_PRECON precon)
{
    (void)solver; (void)precon;
    LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "this preconditioner is not available (none, jacobi, ssor, ilu, ilut, hybrid, is and registered ones are)\n");
    return LIS_ERR_NOT_IMPLEMENTED;
}

LIS_INT lis_precon_create(LIS_SOLVER solver, LIS_PRECON *precon)
{
    const LIS_INT type = solver->options[LIS_OPTIONS_PRECON];
    LIS_INT err;
    *precon = (LIS_PRECON)lis_calloc(sizeof(struct LIS_PRECON_STRUCT), "lis_precon_create::precon");
    if (*precon == NULL) { LIS_SETERR_MEM(sizeof(struct LIS_PRECON_STRUCT)); return LIS_OUT_OF_MEMORY; }
    (*precon)->precon_type = type;
    if (type >= LIS_PRECON_TYPE_USERDEF) {
        if (type >= g_reg_type || g_reg == NULL) err = create_unsupported(solver, *precon);
        else err = g_reg[type - LIS_PRECON_TYPE_USERDEF].pcreate(solver, *precon);
    } else {
        switch (type) {
        case LIS_PRECON_TYPE_NONE: err = create_none(solver, *precon); break;
        case LIS_PRECON_TYPE_JACOBI: err = create_jacobi(solver, *precon); break;
        case LIS_PRECON_TYPE_SSOR: err = create_ssor(solver, *precon); break;
        case LIS_PRECON_TYPE_ILU: case LIS_PRECON_TYPE_ILUT: err = lis_host_ilu_create(solver, *precon); break;
        case LIS_PRECON_TYPE_HYBRID: err = create_hybrid(solver, *precon); break;
        case LIS_PRECON_TYPE_IS: err = create_is(solver, *precon); break;
        default: err = create_unsupported(solver, *precon); break;
        }
        if (!err && type && solver->options[LIS_OPTIONS_ADDS]) {
            /* -adds true: the preconditioner becomes the inner solve of an additive Schwarz /
             * Richardson loop (src/precon/lis_precon.c:142-146, lis_precon_ads.c:52-100): two work
             * vectors, the matrix the solver iterates on */
            LIS_VECTOR *work = (LIS_VECTOR *)lis_calloc(2 * sizeof(LIS_VECTOR), "lis_precon_create_adds::work");
            if (work == NULL) { LIS_SETERR_MEM(2 * sizeof(LIS_VECTOR)); err = LIS_OUT_OF_MEMORY; }
            else if ((*precon)->work) { lis_free(work); LIS_SETERR_IMP; err = LIS_ERR_NOT_IMPLEMENTED; }
            else {
                (*precon)->work = work;
                for (LIS_INT i = 0; i < 2 && !err; i++) { err = lis_vector_duplicate(solver->A, &work[i]); if (!err) (*precon)->worklen = i + 1; }
                if (!err) {
                    /* the Richardson loop multiplies by solver->A; an SSOR inner solve keeps its (possibly private) split matrix */
                    if ((*precon)->precon_type != LIS_PRECON_TYPE_SSOR) {
                        if ((*precon)->is_copy && (*precon)->A && (*precon)->A != solver->A) lis_matrix_destroy((*precon)->A);
                        (*precon)->A = solver->A;
                        (*precon)->is_copy = LIS_FALSE;
                    }
                    (*precon)->precon_type = LIS_PRECON_TYPE_ADDS;
                }
            }
        }
    }
    if (err) { lis_precon_destroy(*precon); *precon = NULL; return err; }
    return LIS_SUCCESS;
}

LIS_INT lis_precon_destroy(LIS_PRECON precon)
{
    if (precon) {
        if (precon->is_copy && precon->A) lis_matrix_destroy(precon->A);
        if (precon->Ah) lis_matrix_destroy(precon->Ah);  /* I+S: the truncated strict upper part */
        if (precon->solver) {                           /* hybrid: the inner solver with its iterate and preconditioner */
            if (precon->solver->x) lis_vector_destroy(precon->solver->x);
            lis_precon_destroy(precon->solver->precon);
            lis_solver_destroy(precon->solver);
        }
        if (precon->D) lis_vector_destroy(precon->D);
        if (precon->b200_ilu) lis_host_ilu_free(precon->b200_ilu);
        if (precon->work) {
            for (LIS_INT i = 0; i < precon->worklen; i++) lis_vector_destroy(precon->work[i]);
            lis_free(precon->work);
        }
        lis_free(precon);
    }
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ apply */
LIS_INT lis_psolve_none(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    (void)solver;
    return lisd_copy(b, x);
}

/* x = b .* D (a multiply, not a divide): src/precon/lis_precon_jacobi.c:119-126 */
LIS_INT lis_psolve_jacobi(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    return lisd_pmul(b, solver->precon->D, x);
}

LIS_INT lis_psolve_ssor(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    return lis_matrix_solve(solver->precon->A, b, x, LIS_MATRIX_SSOR);
}

static LIS_INT psolve_type(LIS_INT type, LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);
static LIS_INT psolveh_type(LIS_INT type, LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);

/* additive Schwarz wrapper, src/precon/lis_precon_ads.c:104-150 (transposed: :196-245):
 *   x = 0; r = b; repeat adds_iter+1 times: w = M^-1 r; x += w; (not after the last) r = b - A x
 * On one process this is adds_iter steps of preconditioned Richardson.  "x += w" and "r = b - r" are
 * axpy(1.0) and xpay(-1.0): multiplications by +-1 are exact, the bits are those of the reference's loops.
 * The reference also zeroes the halo part of r before each inner solve; the inner solves here never
 * read beyond the owned rows. */
static LIS_INT psolve_adds(LIS_SOLVER solver, LIS_VECTOR B, LIS_VECTOR X, int transposed)
{
    LIS_PRECON precon = solver->precon;
    LIS_VECTOR W = precon->work[0], R = precon->work[1];
    const LIS_INT iter = solver->options[LIS_OPTIONS_ADDS_ITER], ptype = solver->options[LIS_OPTIONS_PRECON];
    LIS_INT err = lisd_set_all(0.0, X);
    if (!err) err = lisd_copy(B, R);
    for (LIS_INT k = 0; k < iter + 1 && !err; k++) {
        err = transposed ? psolveh_type(ptype, solver, R, W) : psolve_type(ptype, solver, R, W);
        if (!err) err = lisd_axpy(1.0, W, X);
        if (!err && k != iter) {
            err = transposed ? lisd_matvech(solver->A, X, R) : lisd_matvec(solver->A, X, R);
            if (!err) err = lisd_xpay(B, -1.0, R);
        }
    }
    return err;
}

/* the lis_psolve macro of the reference (include/lis_precon.h:32) as a function; async */
LIS_INT lis_psolve(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    if (solver->precon->precon_type == LIS_PRECON_TYPE_ADDS) return psolve_adds(solver, b, x, 0);
    return psolve_type(solver->precon->precon_type, solver, b, x);
}

static LIS_INT psolve_type(LIS_INT type, LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    switch (type) {
    case LIS_PRECON_TYPE_NONE: return lis_psolve_none(solver, b, x);
    case LIS_PRECON_TYPE_JACOBI: return lis_psolve_jacobi(solver, b, x);
    case LIS_PRECON_TYPE_SSOR: return lis_psolve_ssor(solver, b, x);
    case LIS_PRECON_TYPE_ILU: case LIS_PRECON_TYPE_ILUT: return lis_psolve_iluk(solver, b, x);
    case LIS_PRECON_TYPE_HYBRID: return lis_psolve_hybrid(solver, b, x);
    case LIS_PRECON_TYPE_IS: return lis_psolve_is(solver, b, x);
    default:
        if (type >= LIS_PRECON_TYPE_USERDEF && type < g_reg_type && g_reg)
            return g_reg[type - LIS_PRECON_TYPE_USERDEF].psolve(solver, b, x);
        LIS_SETERR_IMP;
        return LIS_ERR_NOT_IMPLEMENTED;
    }
}

/* M^-H b for BiCG / BiCR: none and Jacobi are self-adjoint for real scalars (lis_precon_jacobi.c
 * lis_psolveh_jacobi: x = b*conj(d)); SSOR and ILU run their transposed sweeps */
LIS_INT lis_psolveh(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    if (solver->precon->precon_type == LIS_PRECON_TYPE_ADDS) return psolve_adds(solver, b, x, 1);
    return psolveh_type(solver->precon->precon_type, solver, b, x);
}

static LIS_INT psolveh_type(LIS_INT type, LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x)
{
    switch (type) {
    case LIS_PRECON_TYPE_NONE: return lis_psolve_none(solver, b, x);
    case LIS_PRECON_TYPE_JACOBI: return lis_psolve_jacobi(solver, b, x);
    case LIS_PRECON_TYPE_SSOR: return lis_matrix_solveh(solver->precon->A, b, x, LIS_MATRIX_SSOR);
    case LIS_PRECON_TYPE_ILU: case LIS_PRECON_TYPE_ILUT: return lis_psolveh_iluk(solver, b, x);
    case LIS_PRECON_TYPE_IS: return lis_psolveh_is(solver, b, x);
    default:
        if (type >= LIS_PRECON_TYPE_USERDEF && type < g_reg_type && g_reg && g_reg[type - LIS_PRECON_TYPE_USERDEF].psolveh)
            return g_reg[type - LIS_PRECON_TYPE_USERDEF].psolveh(solver, b, x);
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "this preconditioner has no transposed solve\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
}

/* ------------------------------------------------------------------ SSOR level schedule */
typedef struct lisd_sweep {
    int n, nblocks;
    int nlev_f, nlev_b;
    int *h_fptr, *h_bptr;          /* host: level -> [start,end) in rows arrays */
    int *d_frows, *d_brows;        /* device: rows ordered by level */
    int *d_blk_start, *d_blk_end;  /* device: per row, the owning block's range */
    lisd_perm pf, pb;              /* one-launch variant */
    double *d_w;                   /* forward-sweep result (input of the backward sweep) */
    unsigned int *d_ticket;
    lisd_tri *tUT, *tLT;           /* transposed sweep (lis_matrix_solveh), built on first use */
} lisd_sweep;

void lisd_perm_free(lisd_perm *p)
{
    lisd_free(p->d_order); lisd_free(p->d_wptr); lisd_free(p->d_plen); lisd_free(p->d_wdep); lisd_free(p->d_sidx); lisd_free(p->d_sval);
    lisd_free(p->d_slots);
    lisd_free(p->d_rptr); lisd_free(p->d_rdep);
    memset(p, 0, sizeof(*p));
}

/* rows[] is level-ordered, lptr[l] its level pointers.  Builds, on the device, what the one-launch sweep
 * kernel reads (kernels/sweep.cu): the padded slot order, and the triangular part (ptr/idx/val, host)
 * as SELL-32 slices in that order -- per warp of 32 slots, entry q of lane l at wptr[w] + 32*q + l,
 * each row's entries in their storage order, the column of an entry replaced by the SLOT of that row (the
 * kernel publishes and polls in slot order).  blk_lo/blk_hi (or NULL): per row, the range of columns
 * the row keeps; couplings outside are the ones the block sweep drops (src/matrix/lis_matrix_csr.c:1590,
 * 1601) and are left out here.  wdep[w]: of all neighbours the warp's rows read, the slot latest in slot order. */
int lisd_sweep_ahead(int maxlen);
static double setup_tick(const char *what, double t0);

/* the two passes of lisd_perm_build over a range of warps (32 slots each); disjoint outputs per warp */
typedef struct {
    const int *order; int *plen; const int *slot_of; const int *wptr; int *wdep; int *width; int *sidx; double *sval;
    const LIS_INT *ptr, *idx; const LIS_SCALAR *val; const int *blk_lo, *blk_hi; int unused;
} perm_ctx;

typedef struct { int *order; int *slot_of; const int *lptr; const int *rows; const size_t *koff; int nlev; } slot_ctx;

static void perm_slot_order(size_t k0, size_t k1, void *ctx)
{
    slot_ctx *c = (slot_ctx *)ctx;
    int lo = 0, hi = c->nlev;                       /* the level that owns slot k0 */
    while (hi - lo > 1) { const int mid = lo + (hi - lo) / 2; if (c->koff[mid] <= k0) lo = mid; else hi = mid; }
    int l = lo;
    for (size_t k = k0; k < k1; k++) {
        while (k >= c->koff[l + 1]) l++;
        const size_t r = k - c->koff[l];
        const int cnt = c->lptr[l + 1] - c->lptr[l];
        if (r < (size_t)cnt) { const int row = c->rows[c->lptr[l] + (int)r]; c->order[k] = row; c->slot_of[row] = (int)k; }
        else c->order[k] = -1;
    }
}

static void perm_count_warps(size_t w0, size_t w1, void *ctx)
{
    perm_ctx *c = (perm_ctx *)ctx;
    for (size_t w = w0; w < w1; w++) {
        int width = 0, latest = -1;
        for (int lane = 0; lane < 32; lane++) {
            const size_t k = w * 32 + (size_t)lane;
            const int i = c->order[k];
            int cnt = 0;
            if (i >= 0)
                for (LIS_INT j = c->ptr[i]; j < c->ptr[i + 1]; j++) {
                    const int jj = c->idx[j];
                    if (c->blk_lo && (jj < c->blk_lo[i] || jj >= c->blk_hi[i])) continue;
                    cnt++;
                    if (c->slot_of[jj] > latest) latest = c->slot_of[jj];
                }
            c->plen[k] = cnt;
            if (cnt > width) width = cnt;
        }
        c->width[w] = width;
        c->wdep[w] = latest;
    }
}

static void perm_fill_warps(size_t w0, size_t w1, void *ctx)
{
    perm_ctx *c = (perm_ctx *)ctx;
    for (size_t w = w0; w < w1; w++) {
        const int width = (c->wptr[w + 1] - c->wptr[w]) / 32;
        for (int lane = 0; lane < 32; lane++) {
            const int i = c->order[w * 32 + (size_t)lane];
            int q = 0;
            if (i >= 0)
                for (LIS_INT j = c->ptr[i]; j < c->ptr[i + 1]; j++) {
                    const int jj = c->idx[j];
                    if (c->blk_lo && (jj < c->blk_lo[i] || jj >= c->blk_hi[i])) continue;
                    c->sidx[(size_t)c->wptr[w] + 32 * (size_t)q + (size_t)lane] = c->slot_of[jj];
                    c->sval[(size_t)c->wptr[w] + 32 * (size_t)q + (size_t)lane] = c->val[j];
                    q++;
                }
            for (; q < width; q++) {
                c->sidx[(size_t)c->wptr[w] + 32 * (size_t)q + (size_t)lane] = (int)(w * 32 + (size_t)lane);
                c->sval[(size_t)c->wptr[w] + 32 * (size_t)q + (size_t)lane] = 0.0;
            }
        }
    }
}

LIS_INT lisd_perm_build(lisd_perm *P, int n, int nlev, const int *lptr, const int *rows,
                        const LIS_INT *ptr, const LIS_INT *idx, const LIS_SCALAR *val, const int *blk_lo, const int *blk_hi)
{
    size_t nslots = 0;
    for (int l = 0; l < nlev; l++) nslots += (size_t)((lptr[l + 1] - lptr[l] + 31) & ~31);
    if (nslots > 0x7fffff00u) { LIS_SETERR(LIS_ERR_OUT_OF_MEMORY, "sweep schedule too large\n"); return LIS_ERR_OUT_OF_MEMORY; }
    const size_t nw = nslots / 32;
    int *order = (int *)malloc(sizeof(int) * (nslots ? nslots : 1));
    int *plen = (int *)malloc(sizeof(int) * (nslots ? nslots : 1));
    int *slot_of = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *wptr = (int *)malloc(sizeof(int) * (nw + 1));
    int *wdep = (int *)malloc(sizeof(int) * (nw ? nw : 1));
    int *sidx = NULL;
    double *sval = NULL;
    LIS_INT err = LIS_OUT_OF_MEMORY;
    double tk = setup_tick(NULL, 0.0);
    if (!order || !plen || !slot_of || !wptr || !wdep) { LIS_SETERR_MEM(nslots * 12); goto done; }
    {
        /* slot order: level l owns the slots [koff[l], koff[l+1]), its rows first, then -1 up to the multiple of 32 */
        size_t *koff = (size_t *)malloc(sizeof(size_t) * ((size_t)nlev + 1));
        if (!koff) { LIS_SETERR_MEM(nlev * 8); goto done; }
        koff[0] = 0;
        for (int l = 0; l < nlev; l++) koff[l + 1] = koff[l] + (size_t)((lptr[l + 1] - lptr[l] + 31) & ~31);
        slot_ctx c = { order, slot_of, lptr, rows, koff, nlev };
        lis_host_parallel_for(nslots, 65536, perm_slot_order, &c);
        free(koff);
    }
    tk = setup_tick("  slot order", tk);
    /* pass 1: kept entries per row, slice widths, the warp's latest neighbour (host worker threads over the warps) */
    int maxlen = 0;
    {
        int *width = (int *)malloc(sizeof(int) * (nw ? nw : 1));
        if (!width) { LIS_SETERR_MEM(nw * 4); goto done; }
        perm_ctx c = { order, plen, slot_of, wptr, wdep, width, NULL, NULL, ptr, idx, val, blk_lo, blk_hi, 0 };
        lis_host_parallel_for(nw, 2048, perm_count_warps, &c);
        size_t total = 0;
        for (size_t w = 0; w < nw; w++) {
            wptr[w] = (int)total;
            if (width[w] > maxlen) maxlen = width[w];
            total += (size_t)32 * (size_t)width[w];
            if (total > 0x7fffff00u) { free(width); LIS_SETERR(LIS_ERR_OUT_OF_MEMORY, "sweep schedule too large\n"); err = LIS_ERR_OUT_OF_MEMORY; goto done; }
        }
        free(width);
        wptr[nw] = (int)total;
        sidx = (int *)malloc(sizeof(int) * (total ? total : 1));
        sval = (double *)malloc(sizeof(double) * (total ? total : 1));
        if (!sidx || !sval) { LIS_SETERR_MEM(total * 12); goto done; }
    }
    tk = setup_tick("  pass 1 (count)", tk);
    /* How far ahead of its last neighbour a warp leaves the cheap one-address wait.  0: it waits for wdep itself (the
     * latest neighbour), then needs one more L2 round trip to collect the neighbours that were missing at its first
     * poll -- two dependent round trips per level.  a > 0: it waits for the latest neighbour of the warp that holds
     * its latest neighbour (applied a times), which is published about a levels earlier, and from then on polls its
     * own missing neighbours directly: the result of the level before is seen by the first poll that reaches L2 after
     * it, one round trip per level, at the price of 3-8x the polling traffic for the warps at the sweep front only.
     * The kernel is the same -- the wait address is a hint, the loop that collects the neighbours decides.
     * Measured (CG + SSOR, 256^3 7-point, 363 iterations, profiles/r02_sweep_ahead.txt): a = 0: 2.494, a = 1: 2.349,
     * a = 2: 2.354 ms per iteration, same iteration count and residual bits.  Default: 1 for factors with short rows
     * (<= 4 kept entries, the stencil case that was measured), 0 otherwise; LIS_B200_SWEEP_AHEAD overrides. */
    {
        const int ahead = lisd_sweep_ahead(maxlen);
        if (ahead > 0 && nw > 0) {
            int *chain = (int *)malloc(sizeof(int) * nw);
            if (!chain) { LIS_SETERR_MEM(nw * 4); goto done; }
            memcpy(chain, wdep, sizeof(int) * nw);
            for (int a = 0; a < ahead; a++)
                for (size_t w = 0; w < nw; w++)
                    if (wdep[w] >= 0) wdep[w] = chain[(size_t)wdep[w] >> 5];
            free(chain);
        }
    }
    /* long rows (tens of kept entries: the banded matrix of config 4, ILU factors with fill): a warp per row on a
     * CSR-by-slot copy instead of SELL slices and a thread per row (kernels/sweep.cu sweep_rowwarp_kernel) */
    {
        size_t kept = 0, nrows = 0;
        for (size_t k = 0; k < nslots; k++) if (order[k] >= 0) { kept += (size_t)plen[k]; nrows++; }
        /* opt-in only (LIS_B200_SWEEP_KERNEL=rows): measured on the 10 M x 70 banded matrix it LOSES to the thread-per-row
         * kernel -- 37 vs 25 ms per BiCGSTAB+SSOR iteration at 16 blocks, 329 vs 261 at one (profiles/r02_configs_n1c.jsonl) --
         * because a warp works through its rows one after the other and only ~4.7 k rows are in flight, which no longer
         * hides the DRAM latency of each row's own loads; kept as a tested alternative */
        const char *force = getenv("LIS_B200_SWEEP_KERNEL");
        const int want = force && strcmp(force, "rows") == 0 && nrows > 0;
        if (want && kept <= 0x7fffff00u) {
            int *rptr = (int *)malloc(sizeof(int) * (nslots + 1)), *rdep = (int *)malloc(sizeof(int) * (nslots ? nslots : 1));
            int *ridx = (int *)malloc(sizeof(int) * (kept ? kept : 1));
            double *rval = (double *)malloc(sizeof(double) * (kept ? kept : 1));
            err = LIS_OUT_OF_MEMORY;
            if (rptr && rdep && ridx && rval) {
                size_t q = 0;
                for (size_t k = 0; k < nslots; k++) {
                    const int i = order[k];
                    int latest = -1;
                    rptr[k] = (int)q;
                    if (i >= 0)
                        for (LIS_INT j = ptr[i]; j < ptr[i + 1]; j++) {
                            const int jj = idx[j];
                            if (blk_lo && (jj < blk_lo[i] || jj >= blk_hi[i])) continue;
                            ridx[q] = slot_of[jj]; rval[q] = val[j]; q++;
                            if (slot_of[jj] > latest) latest = slot_of[jj];
                        }
                    rdep[k] = latest;
                }
                rptr[nslots] = (int)q;
                P->nslots = (int)nslots; P->row_warp = 1; P->short_rows = 0;
                err = lisd_malloc((void **)&P->d_order, sizeof(int) * (nslots ? nslots : 1));
                if (!err) err = lisd_upload(P->d_order, order, sizeof(int) * nslots);
                if (!err) err = lisd_malloc((void **)&P->d_rptr, sizeof(int) * (nslots + 1));
                if (!err) err = lisd_upload(P->d_rptr, rptr, sizeof(int) * (nslots + 1));
                if (!err) err = lisd_malloc((void **)&P->d_rdep, sizeof(int) * (nslots ? nslots : 1));
                if (!err) err = lisd_upload(P->d_rdep, rdep, sizeof(int) * nslots);
                if (!err) err = lisd_malloc((void **)&P->d_sidx, sizeof(int) * (kept ? kept : 1));
                if (!err) err = lisd_upload(P->d_sidx, ridx, sizeof(int) * kept);
                if (!err) err = lisd_malloc((void **)&P->d_sval, sizeof(double) * (kept ? kept : 1));
                if (!err) err = lisd_upload(P->d_sval, rval, sizeof(double) * kept);
                if (!err) err = lisd_malloc((void **)&P->d_slots, sizeof(double) * 2 * (nslots ? nslots : 1));
            } else LIS_SETERR_MEM(kept * 12);
            free(rptr); free(rdep); free(ridx); free(rval);
            goto done;
        }
    }
    /* pass 2: fill the slices; unused positions point at the row itself with a zero (never read) */
    {
        perm_ctx c = { order, plen, slot_of, wptr, wdep, NULL, sidx, sval, ptr, idx, val, blk_lo, blk_hi, 0 };
        lis_host_parallel_for(nw, 2048, perm_fill_warps, &c);
    }
    tk = setup_tick("  pass 2 (fill)", tk);
    P->nslots = (int)nslots;
    P->short_rows = maxlen <= 4;
    {
        const size_t total = (size_t)wptr[nw];
        err = lisd_malloc((void **)&P->d_order, sizeof(int) * (nslots ? nslots : 1));
        if (!err) err = lisd_upload(P->d_order, order, sizeof(int) * nslots);
        if (!err) err = lisd_malloc((void **)&P->d_plen, sizeof(int) * (nslots ? nslots : 1));
        if (!err) err = lisd_upload(P->d_plen, plen, sizeof(int) * nslots);
        if (!err) err = lisd_malloc((void **)&P->d_wptr, sizeof(int) * (nw + 1));
        if (!err) err = lisd_upload(P->d_wptr, wptr, sizeof(int) * (nw + 1));
        if (!err) err = lisd_malloc((void **)&P->d_wdep, sizeof(int) * (nw ? nw : 1));
        if (!err) err = lisd_upload(P->d_wdep, wdep, sizeof(int) * nw);
        if (!err) err = lisd_malloc((void **)&P->d_sidx, sizeof(int) * (total ? total : 1));
        if (!err) err = lisd_upload(P->d_sidx, sidx, sizeof(int) * total);
        if (!err) err = lisd_malloc((void **)&P->d_sval, sizeof(double) * (total ? total : 1));
        if (!err) err = lisd_upload(P->d_sval, sval, sizeof(double) * total);
        if (!err) err = lisd_malloc((void **)&P->d_slots, sizeof(double) * 2 * (nslots ? nslots : 1));     /* slot-ordered results + wd */
    }
    tk = setup_tick("  upload", tk);
done:
    free(order); free(plen); free(slot_of); free(wptr); free(wdep); free(sidx); free(sval);
    return err;
}

/* LIS_B200_SWEEP_AHEAD (see lisd_perm_build) */
int lisd_sweep_ahead(int maxlen)
{
    const char *e = getenv("LIS_B200_SWEEP_AHEAD");          /* read per schedule build: once per preconditioner */
    int v = e && e[0] >= '0' && e[0] <= '9' ? atoi(e) : (maxlen <= 4 ? 1 : 0);
    return v > 8 ? 8 : v;
}

int lisd_sweep_ctas(void);
/* one sweep on a prepared factor; mode as in lisb200_sweep_sell; async on the library stream */
int lisd_perm_sweep(const lisd_perm *P, int mode, int n, const double *d_wd, const double *d_in, double *d_out, unsigned int *d_ticket)
{
    if (P->row_warp)
        return lisb200_sweep_rows(mode, n, P->nslots, P->d_order, P->d_rptr, P->d_rdep, P->d_sidx, P->d_sval,
                                  d_wd, d_in, d_out, P->d_slots, d_ticket, lisd_sweep_ctas(), lisd_stream());
    return lisb200_sweep_sell(mode, n, P->nslots, P->d_order, P->d_wptr, P->d_plen, P->d_wdep, P->d_sidx, P->d_sval,
                              d_wd, d_in, d_out, P->d_slots, d_ticket, lisd_sweep_ctas() | (P->short_rows ? 0x100 : 0), lisd_stream());
}

/* how many CTAs per SM the persistent sweep grid gets (LIS_B200_SWEEP_CTAS=1..9; default: as many as fit, 9 for short-row factors, 6 otherwise) */
int lisd_sweep_ctas(void)
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("LIS_B200_SWEEP_CTAS"); v = e ? atoi(e) : 0; if (v < 1 || v > 9) v = 0; }
    return v;
}

void lisd_sweep_free(void *p)
{
    lisd_sweep *S = (lisd_sweep *)p;
    if (S == NULL) return;
    free(S->h_fptr); free(S->h_bptr);
    lisd_free(S->d_frows); lisd_free(S->d_brows); lisd_free(S->d_blk_start); lisd_free(S->d_blk_end);
    lisd_perm_free(&S->pf); lisd_perm_free(&S->pb);
    lisd_free(S->d_w); lisd_free(S->d_ticket);
    lisd_tri_free(S->tUT); lisd_tri_free(S->tLT);
    free(S);
}

/* counting sort of rows by level; returns level pointers */
int *lisd_order_by_level(int n, const int *lvl, int nlev, int *rows)
{
    int *ptr = (int *)calloc((size_t)nlev + 2, sizeof(int));
    if (!ptr) return NULL;
    for (int i = 0; i < n; i++) ptr[lvl[i] + 1]++;
    for (int l = 0; l < nlev; l++) ptr[l + 1] += ptr[l];
    int *cur = (int *)malloc(((size_t)nlev + 1) * sizeof(int));
    if (!cur) { free(ptr); return NULL; }
    memcpy(cur, ptr, ((size_t)nlev + 1) * sizeof(int));
    for (int i = 0; i < n; i++) rows[cur[lvl[i]]++] = i;
    free(cur);
    return ptr;
}

/* SSOR block count = emulated OpenMP thread count, at most one block per row */
static int sweep_blocks(int n)
{
    int nb = lis_host_num_threads();
    if (nb < 1) nb = 1;
    if (nb > n && n > 0) nb = n;
    return nb;
}

/* LIS_B200_TRACE_SETUP=1: wall time of each phase of the schedule build on stderr */
static double setup_tick(const char *what, double t0)
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("LIS_B200_TRACE_SETUP"); on = e && e[0] == '1'; }
    const double t = lis_wtime();
    if (on && what) fprintf(stderr, "lis_b200 setup: %-28s %8.3f s\n", what, t - t0);
    return t;
}

static LIS_INT sweep_build(LIS_MATRIX A, int nb, lisd_sweep **out)
{
    const int n = A->n;
    double tk = setup_tick(NULL, 0.0);
    lisd_sweep *S = (lisd_sweep *)calloc(1, sizeof(lisd_sweep));
    int *bs = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *be = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *lvl = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int *rows = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    LIS_INT err = LIS_OUT_OF_MEMORY;
    if (!S || !bs || !be || !lvl || !rows) goto fail;
    S->n = n; S->nblocks = nb;
    for (int k = 0; k < nb; k++) {
        LIS_INT is, ie;
        LIS_GET_ISIE(k, nb, n, is, ie);
        for (LIS_INT i = is; i < ie; i++) { bs[i] = is; be[i] = ie; }
    }
    /* forward: row i waits for every L neighbour inside its block */
    int nlev = 0;
    for (int i = 0; i < n; i++) {
        int l = 0;
        for (LIS_INT j = A->L->ptr[i]; j < A->L->ptr[i + 1]; j++) {
            const int jj = A->L->index[j];
            if (jj < bs[i]) continue;
            if (lvl[jj] + 1 > l) l = lvl[jj] + 1;
        }
        lvl[i] = l;
        if (l + 1 > nlev) nlev = l + 1;
    }
    S->nlev_f = nlev;
    tk = setup_tick("forward levels", tk);
    S->h_fptr = lisd_order_by_level(n, lvl, nlev, rows);
    if (!S->h_fptr) goto fail;
    tk = setup_tick("forward counting sort", tk);
    err = lisd_malloc((void **)&S->d_frows, sizeof(int) * (size_t)(n > 0 ? n : 1));
    if (!err) err = lisd_upload(S->d_frows, rows, sizeof(int) * (size_t)n);
    if (!err) err = lisd_perm_build(&S->pf, n, nlev, S->h_fptr, rows, A->L->ptr, A->L->index, A->L->value, bs, be);
    if (err) goto fail;
    tk = setup_tick("forward slices + upload", tk);
    /* backward: row i waits for every U neighbour inside its block */
    nlev = 0;
    for (int i = n - 1; i >= 0; i--) {
        int l = 0;
        for (LIS_INT j = A->U->ptr[i]; j < A->U->ptr[i + 1]; j++) {
            const int jj = A->U->index[j];
            if (jj < bs[i] || jj >= be[i]) continue;
            if (lvl[jj] + 1 > l) l = lvl[jj] + 1;
        }
        lvl[i] = l;
        if (l + 1 > nlev) nlev = l + 1;
    }
    S->nlev_b = nlev;
    S->h_bptr = lisd_order_by_level(n, lvl, nlev, rows);
    tk = setup_tick("backward levels + sort", tk);
    err = LIS_OUT_OF_MEMORY;
    if (!S->h_bptr) goto fail;
    err = lisd_malloc((void **)&S->d_brows, sizeof(int) * (size_t)(n > 0 ? n : 1));
    if (!err) err = lisd_upload(S->d_brows, rows, sizeof(int) * (size_t)n);
    if (!err) err = lisd_perm_build(&S->pb, n, nlev, S->h_bptr, rows, A->U->ptr, A->U->index, A->U->value, bs, be);
    tk = setup_tick("backward slices + upload", tk);
    if (!err) err = lisd_malloc((void **)&S->d_w, sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (!err) err = lisd_malloc((void **)&S->d_ticket, 64);
    if (!err) err = lisd_malloc((void **)&S->d_blk_start, sizeof(int) * (size_t)(n > 0 ? n : 1));
    if (!err) err = lisd_upload(S->d_blk_start, bs, sizeof(int) * (size_t)n);
    if (!err) err = lisd_malloc((void **)&S->d_blk_end, sizeof(int) * (size_t)(n > 0 ? n : 1));
    if (!err) err = lisd_upload(S->d_blk_end, be, sizeof(int) * (size_t)n);
    if (err) goto fail;
    free(bs); free(be); free(lvl); free(rows);
    *out = S;
    return LIS_SUCCESS;
fail:
    free(bs); free(be); free(lvl); free(rows);
    lisd_sweep_free(S);
    if (err == LIS_OUT_OF_MEMORY) LIS_SETERR_MEM(n);
    return err;
}

/* device mirror of the split matrix + the level schedule for the current block count; built
 * once per matrix (setup cost, like the reference's lis_matrix_split in lis_precon_create_ssor) */
static LIS_INT sweep_prepare(LIS_MATRIX A, lisd_matrix **Mout, lisd_sweep **Sout)
{
    lisd_matrix *M;
    LIS_INT err = lisd_matrix_get(A, &M);
    if (err) return err;
    if (M->wd == NULL) { err = lisd_matrix_refresh_wd(A); if (err) return err; }
    if (M->sweep == NULL || ((lisd_sweep *)M->sweep)->nblocks != sweep_blocks(A->n)) {
        lisd_sweep *S;
        if (M->sweep) { lisd_sweep_free(M->sweep); M->sweep = NULL; }
        err = sweep_build(A, sweep_blocks(A->n), &S);
        if (err) return err;
        M->sweep = S;
    }
    *Mout = M; *Sout = (lisd_sweep *)M->sweep;
    return LIS_SUCCESS;
}

/* the plain triangular solve (LIS_MATRIX_LOWER) is one global sweep whatever the thread count
 * (src/matrix/lis_matrix_csr.c:1553-1562): its own single-block schedule when the SSOR one is blocked */
static LIS_INT sweep_prepare_global(LIS_MATRIX A, lisd_matrix **Mout, lisd_sweep **Sout)
{
    LIS_INT err = sweep_prepare(A, Mout, Sout);
    if (err || (*Sout)->nblocks == 1) return err;
    lisd_matrix *M = *Mout;
    if (M->sweep_global == NULL) {
        lisd_sweep *S;
        err = sweep_build(A, 1, &S);
        if (err) return err;
        M->sweep_global = S;
    }
    *Sout = (lisd_sweep *)M->sweep_global;
    return LIS_SUCCESS;
}

LIS_INT lis_host_ssor_prepare(LIS_MATRIX A)
{
    lisd_matrix *M; lisd_sweep *S;
    if (!lisd_available()) return LIS_SUCCESS;         /* host-only use: the sweep itself will report the missing device */
    return sweep_prepare(A, &M, &S);
}

/* x = M^-1 b; async on the library stream */
LIS_INT lis_matrix_solve(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, LIS_INT flag)
{
    LIS_INT err = lisd_require("lis_matrix_solve");
    if (err) return err;
    if (flag != LIS_MATRIX_SSOR && flag != LIS_MATRIX_LOWER) {
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "lis_matrix_solve: the SSOR sweep and the lower triangular solve are available\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    if (!A->is_splited) { err = lis_matrix_split(A); if (err) return err; }
    if (A->WD == NULL) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "lis_matrix_solve: the scaled diagonal WD is not set up\n");
        return LIS_ERR_ILL_ARG;
    }
    lisd_matrix *M;
    lisd_sweep *S;
    if (flag == LIS_MATRIX_LOWER) {
        /* x[i] = (b[i] - sum_L L*x[jj]) * WD[i]: the forward sweep alone, written straight into x */
        if (b == x || b->value == x->value) { LIS_SETERR(LIS_ERR_ILL_ARG, "lis_matrix_solve: b and x must not alias\n"); return LIS_ERR_ILL_ARG; }
        err = sweep_prepare_global(A, &M, &S);
        if (err) return err;
        err = lisd_vec_device(b);
        if (!err) err = lisd_vec_device(x);
        if (err) return err;
        lisd_mark_busy();
        return lisd_check(lisd_perm_sweep(&S->pf, 0, S->n, M->wd, b->value, x->value, S->d_ticket), "lower triangular solve");
    }
    err = sweep_prepare(A, &M, &S);
    if (err) return err;
    err = lisd_vec_device(b);
    if (!err) err = lisd_vec_device(x);
    if (err) return err;
    void *st = lisd_stream();
    lisd_mark_busy();
    {
        /* default: one launch per direction; LIS_B200_SSOR=levels keeps the launch-per-level path */
        const char *e = getenv("LIS_B200_SSOR");
        if (!(e && strcmp(e, "levels") == 0)) {
            err = lisd_check(lisd_perm_sweep(&S->pf, 0, S->n, M->wd, b->value, S->d_w, S->d_ticket), "SSOR forward sweep");
            if (err) return err;
            return lisd_check(lisd_perm_sweep(&S->pb, 3, S->n, M->wd, S->d_w, x->value, S->d_ticket), "SSOR backward sweep");
        }
    }
    for (int l = 0; l < S->nlev_f; l++) {
        const int s = S->h_fptr[l], cnt = S->h_fptr[l + 1] - s;
        err = lisd_check(lisb200_ssor_forward_level(cnt, S->d_frows + s, M->L.ptr, M->L.idx, M->L.val, M->wd,
                                                    S->d_blk_start, b->value, x->value, st), "SSOR forward sweep");
        if (err) return err;
    }
    for (int l = 0; l < S->nlev_b; l++) {
        const int s = S->h_bptr[l], cnt = S->h_bptr[l + 1] - s;
        err = lisd_check(lisb200_ssor_backward_level(cnt, S->d_brows + s, M->U.ptr, M->U.idx, M->U.val, M->wd,
                                                     S->d_blk_start, S->d_blk_end, x->value, st), "SSOR backward sweep");
        if (err) return err;
    }
    return LIS_SUCCESS;
}

/* R^T restricted to couplings inside the owning block, as CSR; a row lists its entries by
 * ascending (or descending) source row = the order the reference's column-oriented loops
 * subtract them in */
static LIS_INT core_transpose(LIS_INT n, LIS_MATRIX_CORE R, int nb, int descending, LIS_INT **optr, LIS_INT **oidx, LIS_SCALAR **oval)
{
    const size_t nnz = (size_t)R->ptr[n];
    LIS_INT *ptr = (LIS_INT *)calloc((size_t)n + 1, sizeof(LIS_INT));
    LIS_INT *idx = (LIS_INT *)malloc(sizeof(LIS_INT) * (nnz ? nnz : 1));
    LIS_SCALAR *val = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (nnz ? nnz : 1));
    LIS_INT *cur = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n > 0 ? n : 1));
    LIS_INT *bs = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n > 0 ? n : 1));
    LIS_INT *be = (LIS_INT *)malloc(sizeof(LIS_INT) * (size_t)(n > 0 ? n : 1));
    if (!ptr || !idx || !val || !cur || !bs || !be) {
        free(ptr); free(idx); free(val); free(cur); free(bs); free(be);
        LIS_SETERR_MEM(nnz * 12);
        return LIS_OUT_OF_MEMORY;
    }
    for (int k = 0; k < nb; k++) {
        LIS_INT is, ie;
        LIS_GET_ISIE(k, nb, n, is, ie);
        for (LIS_INT i = is; i < ie; i++) { bs[i] = is; be[i] = ie; }
    }
    for (LIS_INT i = 0; i < n; i++)
        for (LIS_INT j = R->ptr[i]; j < R->ptr[i + 1]; j++) {
            const LIS_INT c = R->index[j];
            if (c >= bs[i] && c < be[i]) ptr[c + 1]++;
        }
    for (LIS_INT i = 0; i < n; i++) ptr[i + 1] += ptr[i];
    for (LIS_INT i = 0; i < n; i++) cur[i] = ptr[i];
    for (LIS_INT s = 0; s < n; s++) {
        const LIS_INT i = descending ? n - 1 - s : s;
        for (LIS_INT j = R->ptr[i]; j < R->ptr[i + 1]; j++) {
            const LIS_INT c = R->index[j];
            if (c < bs[i] || c >= be[i]) continue;
            idx[cur[c]] = i; val[cur[c]++] = R->value[j];
        }
    }
    free(cur); free(bs); free(be);
    *optr = ptr; *oidx = idx; *oval = val;
    return LIS_SUCCESS;
}

/* x = M^-T b for the SSOR splitting: src/matrix/lis_matrix_csr.c:1804-1855.  The reference runs
 * two column-oriented loops,  t = x[i]*WD[i]; x[jj] -= U[i][jj]*t  (i ascending, x[i] itself left
 * unscaled) and  x[i] = t = x[i]*WD[i]; x[jj] -= L[i][jj]*t  (i descending).  Row-wise these are
 *   z[i] = b[i] - sum_k U^T[i][k] * (z[k]*WD[k])      k ascending
 *   x[i] = (z[i] - sum_k L^T[i][k] * x[k]) * WD[i]    k descending
 * i.e. two triangular solves on the transposed parts with the same products in the same order. */
LIS_INT lis_matrix_solveh(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, LIS_INT flag)
{
    LIS_INT err = lisd_require("lis_matrix_solveh");
    if (err) return err;
    if (flag != LIS_MATRIX_SSOR) {
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "lis_matrix_solveh: the transposed SSOR sweep is available\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    if (!A->is_splited) { err = lis_matrix_split(A); if (err) return err; }
    if (A->WD == NULL) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "lis_matrix_solveh: the scaled diagonal WD is not set up\n");
        return LIS_ERR_ILL_ARG;
    }
    lisd_matrix *M;
    lisd_sweep *S;
    err = sweep_prepare(A, &M, &S);
    if (err) return err;
    if (S->tUT == NULL) {
        LIS_INT *ptr, *idx;
        LIS_SCALAR *val;
        err = core_transpose(A->n, A->U, S->nblocks, 0, &ptr, &idx, &val);
        if (err) return err;
        err = lisd_tri_build(A->n, ptr, idx, val, &S->tUT);
        free(ptr); free(idx); free(val);
        if (err) return err;
        err = core_transpose(A->n, A->L, S->nblocks, 1, &ptr, &idx, &val);
        if (err) return err;
        err = lisd_tri_build(A->n, ptr, idx, val, &S->tLT);
        free(ptr); free(idx); free(val);
        if (err) return err;
    }
    err = lisd_vec_device(b);
    if (!err) err = lisd_vec_device(x);
    if (err) return err;
    err = lisd_tri_solve(S->tUT, 2, M->wd, b->value, S->d_w, "transposed SSOR sweep (U^T)");
    if (err) return err;
    return lisd_tri_solve(S->tLT, 0, M->wd, S->d_w, x->value, "transposed SSOR sweep (L^T)");
}
