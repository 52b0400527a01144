/*
 * lis_esolver.c -- the eigensolver drivers of the reference on the B200 kernels:
 * lis_esolver_* / lis_esolve (src/esolver/lis_esolver.c:142-1385) and the standard eigensolvers
 *   power (lis_esolver_pi.c:127-225), inverse (lis_esolver_ii.c:127-300), Rayleigh quotient
 *   (lis_esolver_rqi.c:124-260), CG and CR (lis_esolver_cg.c, LOBPCG-style / Suetomi-Sekimoto),
 *   subspace (lis_esolver_si.c), Lanczos (lis_esolver_li.c), Arnoldi (lis_esolver_ai.c).
 * They add no kernels: each is the reference's sequence of lis_matvec / lis_vector_* / lis_solve_kernel
 * calls with the same scalar arithmetic on the host, so on the mock device (tests/hostcheck) eigenvalue,
 * iteration count, residual history and eigenvector equal the serial reference bit for bit.  The small
 * dense helpers they use (3x3 Rayleigh-Ritz, QR iteration on the tridiagonal matrix) are restated from
 * src/array/lis_array.c in the operation order that file has.
 * Not carried: the generalized (A x = lambda B x) variants, quad precision.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <ctype.h>
#include "lis_device.h"
#include "lis_host.h"

static const char *k_esolver_atoi[] = {"pi", "ii", "rqi", "cg", "cr", "si", "li", "ai", "gpi", "gii", "grqi", "gcg", "gcr", "gsi", "gli", "gai"};
static const char *k_eprint_atoi[] = {"none", "mem", "out", "all"};
static const char *k_etruefalse_atoi[] = {"false", "true"};
static const char *k_estorage_atoi[] = {"csr", "csc", "msr", "dia", "ell", "jad", "bsr", "bsc", "vbr", "coo", "dns"};
static const char *k_eprecision_atoi[] = {"double", "quad", "switch"};
static const char *k_esolvername[] = {"", "Power", "Inverse", "Rayleigh Quotient", "CG", "CR", "Subspace", "Lanczos", "Arnoldi",
                                      "Generalized Power", "Generalized Inverse", "Generalized Rayleigh Quotient", "Generalized CG",
                                      "Generalized CR", "Generalized Subspace", "Generalized Lanczos", "Generalized Arnoldi"};
static const char *k_estoragename[] = {"CSR", "CSC", "MSR", "DIA", "ELL", "JAD", "BSR", "BSC", "VBR", "COO", "DNS"};
static const char *k_ereturncode[] = {"LIS_SUCCESS", "LIS_ILL_OPTION", "LIS_BREAKDOWN", "LIS_OUT_OF_MEMORY", "LIS_MAXITER",
                                      "LIS_NOT_IMPLEMENTED", "LIS_ERR_FILE_IO"};

static const struct { const char *name; int slot; } k_eoptions[] = {
    {"-emaxiter", LIS_EOPTIONS_MAXITER}, {"-etol", LIS_EPARAMS_RESID}, {"-e", LIS_EOPTIONS_ESOLVER}, {"-ss", LIS_EOPTIONS_SUBSPACE},
    {"-m", LIS_EOPTIONS_MODE}, {"-shift", LIS_EPARAMS_SHIFT}, {"-shift_im", LIS_EPARAMS_SHIFT_IM}, {"-eprint", LIS_EOPTIONS_OUTPUT},
    {"-initx_ones", LIS_EOPTIONS_INITGUESS_ONES}, {"-ie", LIS_EOPTIONS_INNER_ESOLVER}, {"-ige", LIS_EOPTIONS_INNER_GENERALIZED_ESOLVER},
    {"-estorage", LIS_EOPTIONS_STORAGE}, {"-estorage_block", LIS_EOPTIONS_STORAGE_BLOCK}, {"-ef", LIS_EOPTIONS_PRECISION},
    {"-rval", LIS_EOPTIONS_RVAL},
};
#define NWORDS(a) ((int)(sizeof(a) / sizeof((a)[0])))

/* ------------------------------------------------------------------ create / destroy / options */
LIS_INT lis_esolver_create(LIS_ESOLVER *esolver)
{
    LIS_ESOLVER e = (LIS_ESOLVER)lis_calloc(sizeof(struct LIS_ESOLVER_STRUCT), "lis_esolver_create::esolver");
    *esolver = NULL;
    if (e == NULL) { LIS_SETERR_MEM(sizeof(struct LIS_ESOLVER_STRUCT)); return LIS_OUT_OF_MEMORY; }
    e->eprecision = LIS_PRECISION_DOUBLE;
    e->options[LIS_EOPTIONS_ESOLVER] = LIS_ESOLVER_CR;
    e->options[LIS_EOPTIONS_MAXITER] = 1000;
    e->options[LIS_EOPTIONS_SUBSPACE] = 1;
    e->options[LIS_EOPTIONS_MODE] = 0;
    e->options[LIS_EOPTIONS_OUTPUT] = LIS_FALSE;
    e->options[LIS_EOPTIONS_INITGUESS_ONES] = LIS_TRUE;
    e->options[LIS_EOPTIONS_INNER_ESOLVER] = LIS_ESOLVER_II;
    e->options[LIS_EOPTIONS_INNER_GENERALIZED_ESOLVER] = LIS_ESOLVER_GII;
    e->options[LIS_EOPTIONS_STORAGE] = 0;
    e->options[LIS_EOPTIONS_STORAGE_BLOCK] = 2;
    e->options[LIS_EOPTIONS_PRECISION] = LIS_PRECISION_DOUBLE;
    e->options[LIS_EOPTIONS_RVAL] = LIS_FALSE;
    e->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN] = 1.0e-12;
    e->params[LIS_EPARAMS_SHIFT - LIS_EOPTIONS_LEN] = 0.0;
    e->params[LIS_EPARAMS_SHIFT_IM - LIS_EOPTIONS_LEN] = 0.0;
    *esolver = e;
    return LIS_SUCCESS;
}

LIS_INT lis_esolver_work_destroy(LIS_ESOLVER esolver)
{
    if (esolver && esolver->work) {
        for (LIS_INT i = 0; i < esolver->worklen; i++) lis_vector_destroy(esolver->work[i]);
        lis_free(esolver->work);
        esolver->work = NULL;
        esolver->worklen = 0;
    }
    return LIS_SUCCESS;
}

static int keeps_evectors(LIS_INT nesolver) { return nesolver == LIS_ESOLVER_SI || nesolver == LIS_ESOLVER_LI || nesolver == LIS_ESOLVER_AI; }

static void evectors_free(LIS_ESOLVER esolver)
{
    if (esolver->evector) {
        for (LIS_INT i = 0; i < esolver->nevector; i++) if (esolver->evector[i]) lis_vector_destroy(esolver->evector[i]);
        lis_free(esolver->evector);
        esolver->evector = NULL;
        esolver->nevector = 0;
    }
}

LIS_INT lis_esolver_destroy(LIS_ESOLVER esolver)
{
    if (esolver == NULL) return LIS_SUCCESS;
    lis_esolver_work_destroy(esolver);
    if (esolver->rhistory) lis_free(esolver->rhistory);
    if (esolver->evalue) lis_free(esolver->evalue);
    if (esolver->resid) lis_free(esolver->resid);
    if (esolver->iter) lis_free(esolver->iter);
    if (esolver->iter2) lis_free(esolver->iter2);
    evectors_free(esolver);
    lis_free(esolver);
    return LIS_SUCCESS;
}

static LIS_INT ekeyword(const char *arg, const char **words, int nwords, int base, char maxdigit, LIS_INT *dst, const char *what)
{
    if (arg[0] >= '0' && arg[0] <= maxdigit) { int v = 0; sscanf(arg, "%d", &v); *dst = v; return LIS_SUCCESS; }
    for (int i = 0; i < nwords; i++)
        if (strcmp(arg, words[i]) == 0) { *dst = i + base; return LIS_SUCCESS; }
    LIS_SETERR1(LIS_ERR_ILL_ARG, "Parameter %s is not correct\n", what);
    return LIS_ERR_ILL_ARG;
}

static LIS_INT eset_option2(const char *name, const char *value, LIS_ESOLVER esolver)
{
    LIS_INT err = LIS_SUCCESS;
    for (size_t k = 0; k < sizeof(k_eoptions) / sizeof(k_eoptions[0]); k++) {
        if (strcmp(name, k_eoptions[k].name) != 0) continue;
        const int slot = k_eoptions[k].slot;
        switch (slot) {
        case LIS_EOPTIONS_ESOLVER: err = ekeyword(value, k_esolver_atoi, NWORDS(k_esolver_atoi), 1, '9', &esolver->options[slot], "LIS_EOPTIONS_ESOLVER"); break;
        case LIS_EOPTIONS_INNER_ESOLVER: err = ekeyword(value, k_esolver_atoi, NWORDS(k_esolver_atoi), 1, '9', &esolver->options[slot], "LIS_EOPTIONS_INNER_ESOLVER"); break;
        case LIS_EOPTIONS_INNER_GENERALIZED_ESOLVER: err = ekeyword(value, k_esolver_atoi, NWORDS(k_esolver_atoi), 1, '9', &esolver->options[slot], "LIS_EOPTIONS_INNER_GENERALIZED_ESOLVER"); break;
        case LIS_EOPTIONS_OUTPUT: err = ekeyword(value, k_eprint_atoi, NWORDS(k_eprint_atoi), 0, '3', &esolver->options[slot], "LIS_EOPTIONS_OUTPUT"); break;
        case LIS_EOPTIONS_INITGUESS_ONES: case LIS_EOPTIONS_RVAL:
            err = ekeyword(value, k_etruefalse_atoi, NWORDS(k_etruefalse_atoi), 0, '1', &esolver->options[slot], "LIS_EOPTIONS_TRUEFALSE"); break;
        case LIS_EOPTIONS_STORAGE: err = ekeyword(value, k_estorage_atoi, NWORDS(k_estorage_atoi), 1, '9', &esolver->options[slot], "LIS_EOPTIONS_STORAGE"); break;
        case LIS_EOPTIONS_PRECISION: err = ekeyword(value, k_eprecision_atoi, NWORDS(k_eprecision_atoi), 0, '1', &esolver->options[slot], "LIS_EOPTIONS_PRECISION"); break;
        default:
            if (slot < LIS_EOPTIONS_LEN) { int v = esolver->options[slot]; sscanf(value, "%d", &v); esolver->options[slot] = v; }
            else { double dv = esolver->params[slot - LIS_EOPTIONS_LEN]; sscanf(value, "%lg", &dv); esolver->params[slot - LIS_EOPTIONS_LEN] = dv; }
            break;
        }
        if (err) { lis_esolver_work_destroy(esolver); esolver->retcode = err; return err; }
    }
    return LIS_SUCCESS;
}

LIS_INT lis_esolver_set_option(char *text, LIS_ESOLVER esolver)
{
    if (text == NULL) return LIS_SUCCESS;
    char *buf = (char *)malloc(strlen(text) + 1);
    if (buf == NULL) { LIS_SETERR_MEM(strlen(text) + 1); return LIS_OUT_OF_MEMORY; }
    strcpy(buf, text);
    for (char *p = buf; *p; p++) *p = (char)tolower((unsigned char)*p);
    char *save = NULL, *name = NULL;
    LIS_INT err = LIS_SUCCESS;
    for (char *tok = strtok_r(buf, " \t\r\n", &save); tok; tok = strtok_r(NULL, " \t\r\n", &save)) {
        if (name == NULL) { if (tok[0] == '-') name = tok; continue; }
        err = eset_option2(name, tok, esolver);
        name = NULL;
        if (err) break;
    }
    free(buf);
    return err;
}

LIS_INT lis_esolver_set_optionC(LIS_ESOLVER esolver)
{
    int count = 0;
    const lis_arg_t *args = lis_host_args(&count);
    char name[256];
    for (int i = 0; i < count; i++) {
        snprintf(name, sizeof(name), "-%s", args[i].name);
        LIS_INT err = eset_option2(name, args[i].value, esolver);
        if (err) return err;
    }
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ getters */
LIS_INT lis_esolver_get_iter(LIS_ESOLVER e, LIS_INT *iter) { *iter = e->iter[0]; return LIS_SUCCESS; }
LIS_INT lis_esolver_get_iterex(LIS_ESOLVER e, LIS_INT *iter, LIS_INT *iter_double, LIS_INT *iter_quad)
{ *iter = e->iter[0]; *iter_double = e->iter2[0]; *iter_quad = e->iter[0] - e->iter2[0]; return LIS_SUCCESS; }
LIS_INT lis_esolver_get_time(LIS_ESOLVER e, double *time) { *time = e->time; return LIS_SUCCESS; }
LIS_INT lis_esolver_get_timeex(LIS_ESOLVER e, double *time, double *itime, double *ptime, double *p_c_time, double *p_i_time)
{
    *time = e->time;
    if (itime) *itime = e->itime;
    if (ptime) *ptime = e->ptime;
    if (p_c_time) *p_c_time = e->p_c_time;
    if (p_i_time) *p_i_time = e->p_i_time;
    return LIS_SUCCESS;
}
LIS_INT lis_esolver_get_residualnorm(LIS_ESOLVER e, LIS_REAL *residual) { *residual = e->resid[0]; return LIS_SUCCESS; }
LIS_INT lis_esolver_get_status(LIS_ESOLVER e, LIS_INT *status) { *status = e->retcode; return LIS_SUCCESS; }
LIS_INT lis_esolver_get_esolver(LIS_ESOLVER e, LIS_INT *nesol) { *nesol = e->options[LIS_EOPTIONS_ESOLVER]; return LIS_SUCCESS; }
LIS_INT lis_esolver_get_esolvername(LIS_INT esolver, char *name)
{
    if (esolver < 1 || esolver > LIS_ESOLVER_LEN) { LIS_SETERR(LIS_ERR_ILL_ARG, "esolver number is out of range\n"); return LIS_ERR_ILL_ARG; }
    strcpy(name, k_esolvername[esolver]);
    return LIS_SUCCESS;
}

static LIS_INT need_subspace_solver(LIS_ESOLVER e)
{
    if (!keeps_evectors(e->options[LIS_EOPTIONS_ESOLVER])) {
        LIS_SETERR1(LIS_ERR_ILL_ARG, "Parameter LIS_EOPTIONS_ESOLVER is %D (Set Subspace, Lanczos, or Arnoldi)\n", e->options[LIS_EOPTIONS_ESOLVER]);
        return LIS_ERR_ILL_ARG;
    }
    return LIS_SUCCESS;
}

LIS_INT lis_esolver_get_rhistory(LIS_ESOLVER e, LIS_VECTOR v)
{
    LIS_INT maxiter = e->iter[0] + 1;
    if (e->retcode != LIS_SUCCESS) maxiter--;
    const LIS_INT n = v->n < maxiter ? v->n : maxiter;
    return n > 0 ? lis_vector_set_values2(LIS_INS_VALUE, v->is + v->origin, n, e->rhistory, v) : LIS_SUCCESS;
}

static LIS_INT fill_from_array(LIS_ESOLVER e, LIS_VECTOR v, int which)
{
    LIS_INT err = need_subspace_solver(e);
    if (err) return err;
    const LIS_INT ss = e->options[LIS_EOPTIONS_SUBSPACE];
    if (lis_vector_is_null(v)) { err = lis_vector_set_size(v, 0, ss); if (err) return err; }
    for (LIS_INT i = 0; i < ss && i < v->gn; i++) {
        const LIS_SCALAR val = which == 0 ? e->evalue[i] : which == 1 ? (LIS_SCALAR)e->resid[i] : (LIS_SCALAR)e->iter[i];
        if (i >= v->is && i < v->ie) { err = lis_vector_set_value(LIS_INS_VALUE, i + v->origin, val, v); if (err) return err; }
    }
    return LIS_SUCCESS;
}
LIS_INT lis_esolver_get_evalues(LIS_ESOLVER e, LIS_VECTOR v) { return fill_from_array(e, v, 0); }
LIS_INT lis_esolver_get_residualnorms(LIS_ESOLVER e, LIS_VECTOR v) { return fill_from_array(e, v, 1); }
LIS_INT lis_esolver_get_iters(LIS_ESOLVER e, LIS_VECTOR v) { return fill_from_array(e, v, 2); }
LIS_INT lis_esolver_get_specific_evalue(LIS_ESOLVER e, LIS_INT mode, LIS_SCALAR *evalue)
{ LIS_INT err = need_subspace_solver(e); if (!err) *evalue = e->evalue[mode]; return err; }
LIS_INT lis_esolver_get_specific_residualnorm(LIS_ESOLVER e, LIS_INT mode, LIS_REAL *residual)
{ LIS_INT err = need_subspace_solver(e); if (!err) *residual = e->resid[mode]; return err; }
LIS_INT lis_esolver_get_specific_iter(LIS_ESOLVER e, LIS_INT mode, LIS_INT *iter)
{ LIS_INT err = need_subspace_solver(e); if (!err) *iter = e->iter[mode]; return err; }
LIS_INT lis_esolver_get_specific_evector(LIS_ESOLVER e, LIS_INT mode, LIS_VECTOR x)
{ LIS_INT err = need_subspace_solver(e); if (!err) err = lis_vector_copy(e->evector[mode], x); return err; }

/* the eigenvectors as the columns of M (n x ss entries, row by row; the reference assembles them as
 * COO, src/esolver/lis_esolver.c:1201-1239 -- here CSR, same entries in the same order) */
LIS_INT lis_esolver_get_evectors(LIS_ESOLVER e, LIS_MATRIX M)
{
    LIS_INT err = need_subspace_solver(e), n, gn, is, ie, js = 0;
    if (err) return err;
    const LIS_INT ss = e->options[LIS_EOPTIONS_SUBSPACE];
    err = lis_matrix_set_size(M, 0, e->evector[0]->gn);
    if (err) return err;
    lis_matrix_get_size(M, &n, &gn);
    lis_matrix_get_range(M, &is, &ie);
    if (e->evector[0]->origin) { is++; js++; }
    LIS_SCALAR *col = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(n > 0 ? n : 1));
    if (col == NULL) { LIS_SETERR_MEM(n * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
    for (LIS_INT j = 0; j < ss && !err; j++) {
        err = lis_vector_get_values(e->evector[j], e->evector[j]->is + e->evector[j]->origin, n, col);
        for (LIS_INT i = 0; i < n && !err; i++) err = lis_matrix_set_value(LIS_INS_VALUE, i + is, j + js, col[i], M);
    }
    free(col);
    if (err) return err;
    lis_matrix_set_type(M, LIS_MATRIX_CSR);
    return lis_matrix_assemble(M);
}

LIS_INT lis_esolver_output_rhistory(LIS_ESOLVER esolver, char *filename)
{
    LIS_INT maxiter = esolver->iter[0] + 1;
    if (esolver->retcode != LIS_SUCCESS) maxiter--;
    if (esolver->rhistory == NULL) { LIS_SETERR(LIS_FAILS, "eigensolver's residual history is empty\n"); return LIS_FAILS; }
    if (lisd_rank() != 0) return LIS_SUCCESS;
    FILE *f = fopen(filename, "w");
    if (f == NULL) { LIS_SETERR1(LIS_ERR_FILE_IO, "cannot open file %s\n", filename); return LIS_ERR_FILE_IO; }
    for (LIS_INT i = 0; i < maxiter; i++) fprintf(f, "%e\n", (double)esolver->rhistory[i]);
    fclose(f);
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ helpers shared by the algorithms */
#define ECHK(e) do { LIS_INT e_ = (e); if (e_) return e_; } while (0)

static LIS_INT ework(LIS_ESOLVER esolver, LIS_INT worklen)
{
    LIS_VECTOR *work = (LIS_VECTOR *)lis_malloc((size_t)worklen * sizeof(LIS_VECTOR), "lis_esolver::work");
    LIS_INT i, err = LIS_SUCCESS;
    if (work == NULL) { LIS_SETERR_MEM(worklen * sizeof(LIS_VECTOR)); return LIS_ERR_OUT_OF_MEMORY; }
    for (i = 0; i < worklen; i++) { err = lis_vector_duplicate(esolver->A, &work[i]); if (err) break; }
    if (i < worklen) { for (LIS_INT j = 0; j < i; j++) lis_vector_destroy(work[j]); lis_free(work); return err; }
    esolver->worklen = worklen;
    esolver->work = work;
    return LIS_SUCCESS;
}

static void erecord(LIS_ESOLVER esolver, LIS_INT output, LIS_INT iter, LIS_REAL resid)
{
    if (output) {
        if (output & LIS_EPRINT_MEM) esolver->rhistory[iter] = resid;
        if (output & LIS_EPRINT_OUT) lis_host_print_rhistory(iter, resid);
    }
}

/* the shift the algorithms subtract from the diagonal: -shift, overridden by the inner shift the
 * Lanczos / Arnoldi refinement passes down (lis_esolver_pi.c:163-164 and alike) */
static LIS_SCALAR eshift(LIS_ESOLVER esolver)
{
    LIS_SCALAR oshift = esolver->params[LIS_EPARAMS_SHIFT - LIS_EOPTIONS_LEN];
    if (esolver->ishift != 0.0) oshift = esolver->ishift;
    return oshift;
}

/* the inner linear solver: "-i <default> -p none", then the command-line options on top
 * (lis_esolver_ii.c:181-196) */
static LIS_INT inner_solver(LIS_ESOLVER esolver, const char *defaults, LIS_SOLVER *out)
{
    LIS_SOLVER solver;
    char text[64], solvername[128], preconname[128];
    LIS_INT nsol, precon_type;
    ECHK(lis_solver_create(&solver));
    strcpy(text, defaults);
    lis_solver_set_option(text, solver);
    { LIS_INT err = lis_solver_set_optionC(solver); if (err) { lis_solver_destroy(solver); return err; } }
    lis_solver_get_solver(solver, &nsol);
    lis_solver_get_precon(solver, &precon_type);
    lis_solver_get_solvername(nsol, solvername);
    lis_solver_get_preconname(precon_type, preconname);
    if (esolver->options[LIS_EOPTIONS_OUTPUT]) {
        lis_printf(LIS_COMM_WORLD, "linear solver         : %s\n", solvername);
        lis_printf(LIS_COMM_WORLD, "preconditioner        : %s\n", preconname);
    }
    *out = solver;
    return LIS_SUCCESS;
}

static void add_solver_times(LIS_ESOLVER esolver, LIS_SOLVER solver)
{
    esolver->ptime += solver->ptime;
    esolver->itime += solver->itime;
    esolver->p_c_time += solver->p_c_time;
    esolver->p_i_time += solver->p_i_time;
}

static LIS_INT normalize(LIS_VECTOR v)
{
    LIS_REAL nrm2;
    ECHK(lis_vector_nrm2(v, &nrm2));
    return lis_vector_scale(1.0 / nrm2, v);
}

/* ------------------------------------------------------------------ power iteration
 * src/esolver/lis_esolver_pi.c:127-225 */
static LIS_INT lis_epi(LIS_ESOLVER esolver)
{
    LIS_MATRIX A = esolver->A;
    LIS_VECTOR v = esolver->x, y = esolver->work[0], q = esolver->work[1];
    const LIS_INT emaxiter = esolver->options[LIS_EOPTIONS_MAXITER], output = esolver->options[LIS_EOPTIONS_OUTPUT];
    const LIS_REAL tol = esolver->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN];
    const LIS_SCALAR oshift = eshift(esolver);
    LIS_SCALAR theta = 0.0;
    LIS_REAL resid = 0.0;
    LIS_INT iter = 0, ret = LIS_MAXITER;
    if (esolver->options[LIS_EOPTIONS_INITGUESS_ONES]) ECHK(lis_vector_set_all(1.0, v));
    if (oshift != 0.0) ECHK(lis_matrix_shift_diagonal(A, oshift));
    if (output) lis_printf(LIS_COMM_WORLD, "shift                 : %e\n", (double)oshift);
    while (iter < emaxiter) {
        iter = iter + 1;
        ECHK(normalize(v));                                  /* v = v / ||v||_2 */
        ECHK(lis_matvec(A, v, y));                           /* y = A v */
        ECHK(lis_vector_dot(v, y, &theta));                  /* theta = <v,y> */
        ECHK(lis_vector_axpyz(-theta, v, y, q));             /* resid = ||y - theta v||_2 / |theta| */
        ECHK(lis_vector_nrm2(q, &resid));
        resid = resid / fabs(theta);
        ECHK(lis_vector_copy(y, v));
        erecord(esolver, output, iter, resid);
        if (tol >= resid) { ret = LIS_SUCCESS; break; }
    }
    esolver->retcode = ret;
    esolver->iter[0] = iter;
    esolver->resid[0] = resid;
    esolver->evalue[0] = theta + oshift;
    ECHK(normalize(v));
    if (oshift != 0.0) ECHK(lis_matrix_shift_diagonal(A, -oshift));
    return ret;
}

/* ------------------------------------------------------------------ inverse iteration
 * src/esolver/lis_esolver_ii.c:127-300 */
static LIS_INT lis_eii(LIS_ESOLVER esolver)
{
    LIS_MATRIX A = esolver->A;
    LIS_VECTOR v = esolver->x, y = esolver->work[0], q = esolver->work[1];
    const LIS_INT emaxiter = esolver->options[LIS_EOPTIONS_MAXITER], output = esolver->options[LIS_EOPTIONS_OUTPUT];
    const LIS_REAL tol = esolver->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN];
    const LIS_SCALAR oshift = eshift(esolver);
    LIS_SCALAR theta = 0.0;
    LIS_REAL resid = 0.0;
    LIS_INT iter = 0, ret = LIS_MAXITER, err;
    LIS_SOLVER solver;
    LIS_PRECON precon;
    if (esolver->options[LIS_EOPTIONS_INITGUESS_ONES]) ECHK(lis_vector_set_all(1.0, v));
    if (oshift != 0.0) ECHK(lis_matrix_shift_diagonal(A, oshift));
    if (output) lis_printf(LIS_COMM_WORLD, "shift                 : %e\n", (double)oshift);
    ECHK(inner_solver(esolver, "-i bicg -p none", &solver));
    solver->A = A;
    err = lis_precon_create(solver, &precon);
    if (err) { lis_solver_destroy(solver); return err; }
    while (iter < emaxiter) {
        iter = iter + 1;
        err = normalize(v);
        if (!err) err = lis_solve_kernel(A, v, y, solver, precon);        /* y = A^-1 v */
        if (err) { lis_precon_destroy(precon); lis_solver_destroy(solver); return err; }
        err = lis_vector_dot(v, y, &theta);
        if (!err) err = lis_vector_axpyz(-theta, v, y, q);
        if (!err) err = lis_vector_nrm2(q, &resid);
        if (!err) err = lis_vector_copy(y, v);
        if (err) { lis_precon_destroy(precon); lis_solver_destroy(solver); return err; }
        resid = resid / fabs(theta);
        add_solver_times(esolver, solver);
        erecord(esolver, output, iter, resid);
        if (tol >= resid) { ret = LIS_SUCCESS; break; }
    }
    esolver->retcode = ret;
    esolver->iter[0] = iter;
    esolver->resid[0] = resid;
    esolver->evalue[0] = 1.0 / theta + oshift;
    err = normalize(v);
    if (!err && oshift != 0.0) err = lis_matrix_shift_diagonal(A, -oshift);
    lis_precon_destroy(precon);
    lis_solver_destroy(solver);
    return err ? err : ret;
}

/* ------------------------------------------------------------------ Rayleigh quotient iteration
 * src/esolver/lis_esolver_rqi.c:124-260 */
static LIS_INT lis_erqi(LIS_ESOLVER esolver)
{
    LIS_MATRIX A = esolver->A;
    LIS_VECTOR v = esolver->x, y = esolver->work[0], q = esolver->work[1];
    const LIS_INT emaxiter = esolver->options[LIS_EOPTIONS_MAXITER], output = esolver->options[LIS_EOPTIONS_OUTPUT];
    const LIS_REAL tol = esolver->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN];
    LIS_SCALAR theta = 0.0, dotvy = 0.0, rho = 0.0;
    LIS_REAL resid = 0.0, ynrm = 0.0;
    LIS_INT iter = 0, ret = LIS_MAXITER, err;
    LIS_SOLVER solver;
    LIS_PRECON precon;
    if (esolver->options[LIS_EOPTIONS_INITGUESS_ONES]) ECHK(lis_vector_set_all(1.0, v));
    ECHK(inner_solver(esolver, "-i bicg -p none", &solver));
    solver->A = A;
    err = lis_precon_create(solver, &precon);
    if (err) { lis_solver_destroy(solver); return err; }
    err = normalize(v);
    if (!err) err = lis_matvec(A, v, y);                                    /* rho = <v,Av> / <v,v> */
    if (!err) err = lis_vector_dot(v, y, &rho);
    while (!err && iter < emaxiter) {
        iter = iter + 1;
        err = lis_matrix_shift_diagonal(A, rho);                             /* y = (A - rho I)^-1 v */
        if (!err) err = lis_solve_kernel(A, v, y, solver, precon);
        if (err) break;
        err = lis_matrix_shift_diagonal(A, -rho);
        if (!err) err = lis_vector_nrm2(y, &ynrm);                           /* theta = ||y||_2 */
        theta = ynrm;
        if (!err) err = lis_vector_dot(v, y, &dotvy);
        if (err) break;
        rho = rho + dotvy / (theta * theta);
        err = lis_vector_axpyz(-dotvy, v, y, q);                             /* resid = ||y - <v,y> v||_2 / |<v,y>| */
        if (!err) err = lis_vector_nrm2(q, &resid);
        resid = resid / fabs(dotvy);
        if (!err) err = lis_vector_scale(1.0 / theta, y);                    /* v = y / theta */
        if (!err) err = lis_vector_copy(y, v);
        if (err) break;
        erecord(esolver, output, iter, resid);
        add_solver_times(esolver, solver);
        if (tol >= resid) { ret = LIS_SUCCESS; break; }
    }
    if (err) { lis_precon_destroy(precon); lis_solver_destroy(solver); return err; }
    esolver->retcode = ret;
    esolver->iter[0] = iter;
    esolver->resid[0] = resid;
    esolver->evalue[0] = rho;
    err = normalize(v);
    lis_precon_destroy(precon);
    lis_solver_destroy(solver);
    return err ? err : ret;
}

/* ------------------------------------------------------------------ dense helpers (src/array/lis_array.c,
 * column-major n x n arrays; same loops, same accumulation order) */
static LIS_REAL arr_nrm2(LIS_INT n, const LIS_SCALAR *x)
{
    LIS_SCALAR t = 0.0;
    for (LIS_INT i = 0; i < n; i++) t += x[i] * x[i];
    return sqrt(t);
}
static LIS_SCALAR arr_dot(LIS_INT n, const LIS_SCALAR *x, const LIS_SCALAR *y)
{
    LIS_SCALAR t = 0.0;
    for (LIS_INT i = 0; i < n; i++) t += x[i] * y[i];
    return t;
}
/* y = A x, :429-465 (n = 3 is written out there) */
static void arr_matvec3(const LIS_SCALAR *a, const LIS_SCALAR *x, LIS_SCALAR *y)
{
    y[0] = a[0] * x[0] + a[3] * x[1] + a[6] * x[2];
    y[1] = a[1] * x[0] + a[4] * x[1] + a[7] * x[2];
    y[2] = a[2] * x[0] + a[5] * x[1] + a[8] * x[2];
}
/* x = A^-1 b by Gaussian elimination without pivoting on a copy w, :960-1024 (general-n branch) */
static void arr_solve(LIS_INT n, const LIS_SCALAR *a, const LIS_SCALAR *b, LIS_SCALAR *x, LIS_SCALAR *w)
{
    LIS_INT i, j, k;
    for (i = 0; i < n * n; i++) w[i] = a[i];
    for (k = 0; k < n; k++) {
        w[k + k * n] = 1.0 / w[k + k * n];
        for (i = k + 1; i < n; i++) {
            const LIS_SCALAR t = w[i + k * n] * w[k + k * n];
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
}
/* classical Gram-Schmidt QR, :1029-1080 */
static void arr_cgs(LIS_INT n, const LIS_SCALAR *a, LIS_SCALAR *q, LIS_SCALAR *r, LIS_SCALAR *a_k)
{
    const LIS_REAL tol = 1e-12;
    LIS_INT i, j, k;
    for (i = 0; i < n * n; i++) { q[i] = 0.0; r[i] = 0.0; }
    for (k = 0; k < n; k++) {
        for (i = 0; i < n; i++) a_k[i] = a[i + k * n];
        for (j = 0; j < k; j++) {
            r[j + k * n] = 0;
            for (i = 0; i < n; i++) r[j + k * n] += q[i + j * n] * a[i + k * n];
            for (i = 0; i < n; i++) a_k[i] -= r[j + k * n] * q[i + j * n];
        }
        const LIS_REAL nrm2 = arr_nrm2(n, a_k);
        r[k + k * n] = nrm2;
        if (nrm2 < tol) break;
        for (i = 0; i < n; i++) q[i + k * n] = a_k[i] / nrm2;
    }
}
/* unshifted QR iteration A <- R Q until |a[1]| < 1e-12, :1136-1175 */
static void arr_qr(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *q, LIS_SCALAR *r, LIS_INT *qriter, LIS_REAL *qrerr)
{
    const LIS_INT maxiter = 100000;
    const LIS_REAL tol = 1e-12;
    LIS_INT iter = 0;
    LIS_REAL err = 0.0;
    LIS_SCALAR *a_k = (LIS_SCALAR *)malloc(sizeof(LIS_SCALAR) * (size_t)(n > 0 ? n : 1));
    while (iter < maxiter) {
        iter = iter + 1;
        arr_cgs(n, a, q, r, a_k);
        for (LIS_INT j = 0; j < n; j++)
            for (LIS_INT i = 0; i < n; i++) {
                a[i + j * n] = 0;
                for (LIS_INT k = 0; k < n; k++) a[i + j * n] += r[i + k * n] * q[k + j * n];
            }
        err = fabs(a[1]);
        if (err < tol) break;
    }
    free(a_k);
    *qriter = iter;
    *qrerr = err;
}

/* ------------------------------------------------------------------ CG (locally optimal, Rayleigh-Ritz on
 * span{w, x, p}), src/esolver/lis_esolver_cg.c lis_ecg */
static LIS_INT lis_ecg(LIS_ESOLVER esolver)
{
    LIS_MATRIX A = esolver->A;
    LIS_VECTOR x = esolver->x;
    LIS_VECTOR r = esolver->work[0], w = esolver->work[1], p = esolver->work[2], Ax = esolver->work[3], Aw = esolver->work[4], Ap = esolver->work[5];
    const LIS_INT emaxiter = esolver->options[LIS_EOPTIONS_MAXITER], output = esolver->options[LIS_EOPTIONS_OUTPUT];
    const LIS_REAL tol = esolver->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN];
    const LIS_SCALAR oshift = eshift(esolver);
    LIS_SCALAR lambda = 0.0, A3[9], B3[9], W3[9], v3[3], z3[3], q3[3], B3v3[3], mu3;
    LIS_REAL nrm2, resid = 0.0, resid3;
    LIS_INT iter = 0, iter3, err;
    LIS_SOLVER solver;
    LIS_PRECON precon;
    double ptime = 0.0, time;
    if (esolver->options[LIS_EOPTIONS_INITGUESS_ONES]) ECHK(lis_vector_set_all(1.0, x));
    if (oshift != 0.0) ECHK(lis_matrix_shift_diagonal(A, oshift));
    if (output) lis_printf(LIS_COMM_WORLD, "shift                 : %e\n", (double)oshift);
    ECHK(normalize(x));
    ECHK(lis_matvec(A, x, Ax));
    ECHK(inner_solver(esolver, "-i cg -p none", &solver));
    err = lis_solve(A, x, p, solver);                                          /* p = A^-1 x */
    if (!err) err = lis_vector_copy(x, Ap);
    if (!err) err = lis_precon_create(solver, &precon);
    if (err) { lis_solver_destroy(solver); return err; }
    solver->precon = precon;
#define CGCHK(e) do { err = (e); if (err) goto fail; } while (0)
    while (iter < emaxiter) {
        iter = iter + 1;
        CGCHK(lis_vector_dot(x, Ax, &lambda));                                   /* mu = <x,x>/<x,Ax> = 1/lambda */
        CGCHK(lis_vector_axpyz(-1.0 / lambda, Ax, x, r));                        /* r = x - mu A x */
        CGCHK(lis_vector_nrm2(r, &nrm2));
        resid = nrm2;
        erecord(esolver, output, iter, resid);
        if (resid < tol) break;
        time = lis_wtime();
        CGCHK(lis_psolve(solver, r, w));                                         /* w = M^-1 r */
        ptime += lis_wtime() - time;
        CGCHK(normalize(w));
        CGCHK(lis_matvec(A, w, Aw));
        CGCHK(lis_vector_dot(w, Aw, &A3[0]));
        CGCHK(lis_vector_dot(x, Aw, &A3[3]));
        CGCHK(lis_vector_dot(p, Aw, &A3[6]));
        A3[1] = A3[3];
        CGCHK(lis_vector_dot(x, Ax, &A3[4]));
        CGCHK(lis_vector_dot(p, Ax, &A3[7]));
        A3[2] = A3[6];
        A3[5] = A3[7];
        CGCHK(lis_vector_dot(p, Ap, &A3[8]));
        CGCHK(lis_vector_dot(w, w, &B3[0]));
        CGCHK(lis_vector_dot(x, w, &B3[3]));
        CGCHK(lis_vector_dot(p, w, &B3[6]));
        B3[1] = B3[3];
        CGCHK(lis_vector_dot(x, x, &B3[4]));
        CGCHK(lis_vector_dot(p, x, &B3[7]));
        B3[2] = B3[6];
        B3[5] = B3[7];
        CGCHK(lis_vector_dot(p, p, &B3[8]));
        /* eigenvector v3 of the 3x3 pencil by inverse iteration */
        v3[0] = v3[1] = v3[2] = 1.0;
        iter3 = 0;
        while (iter3 < emaxiter) {
            iter3 = iter3 + 1;
            nrm2 = arr_nrm2(3, v3);
            for (int k = 0; k < 3; k++) v3[k] = (1.0 / nrm2) * v3[k];
            arr_matvec3(B3, v3, B3v3);
            arr_solve(3, A3, B3v3, z3, W3);
            mu3 = arr_dot(3, B3v3, z3);
            for (int k = 0; k < 3; k++) q3[k] = -mu3 * B3v3[k] + z3[k];
            resid3 = arr_nrm2(3, q3);
            if (resid3 < tol) break;
            for (int k = 0; k < 3; k++) v3[k] = z3[k];
        }
        /* update x, p and A x, A p */
        CGCHK(lis_vector_scale(v3[0], w));
        CGCHK(lis_vector_axpy(v3[2], p, w));
        CGCHK(lis_vector_xpay(w, v3[1], x));
        CGCHK(lis_vector_copy(w, p));
        CGCHK(lis_vector_scale(v3[0], Aw));
        CGCHK(lis_vector_axpy(v3[2], Ap, Aw));
        CGCHK(lis_vector_xpay(Aw, v3[1], Ax));
        CGCHK(lis_vector_copy(Aw, Ap));
        CGCHK(lis_vector_nrm2(x, &nrm2));
        CGCHK(lis_vector_scale(1.0 / nrm2, x));
        CGCHK(lis_vector_scale(1.0 / nrm2, Ax));
        CGCHK(lis_vector_nrm2(p, &nrm2));
        CGCHK(lis_vector_scale(1.0 / nrm2, p));
        CGCHK(lis_vector_scale(1.0 / nrm2, Ap));
    }
    esolver->iter[0] = iter;
    esolver->resid[0] = resid;
    esolver->evalue[0] = lambda + oshift;
    esolver->ptime = ptime;
    esolver->itime = solver->itime;
    esolver->p_c_time = solver->p_c_time;
    esolver->p_i_time = solver->p_i_time;
    lis_precon_destroy(precon);
    solver->precon = NULL;
    lis_solver_destroy(solver);
    if (oshift != 0.0) ECHK(lis_matrix_shift_diagonal(A, -oshift));
    esolver->retcode = resid < tol ? LIS_SUCCESS : LIS_MAXITER;
    return esolver->retcode;
fail:
    lis_precon_destroy(precon);
    solver->precon = NULL;
    lis_solver_destroy(solver);
    return err;
#undef CGCHK
}

/* ------------------------------------------------------------------ CR, src/esolver/lis_esolver_cg.c lis_ecr
 * (the default eigensolver) */
static LIS_INT lis_ecr(LIS_ESOLVER esolver)
{
    LIS_MATRIX A = esolver->A;
    LIS_VECTOR x = esolver->x;
    LIS_VECTOR r = esolver->work[0], p = esolver->work[1], w = esolver->work[2], Ax = esolver->work[3], Ap = esolver->work[4], Aw = esolver->work[5];
    const LIS_INT emaxiter = esolver->options[LIS_EOPTIONS_MAXITER], output = esolver->options[LIS_EOPTIONS_OUTPUT];
    const LIS_REAL tol = esolver->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN];
    const LIS_SCALAR oshift = eshift(esolver);
    LIS_SCALAR lambda = 0.0, alpha, beta, rAp, rp, ApAp, pAp, pp, AwAp, pAw, wAp, wp;
    LIS_REAL nrm2, resid = 0.0;
    LIS_INT iter = 0, err;
    LIS_SOLVER solver;
    LIS_PRECON precon;
    double ptime = 0.0, time;
    if (esolver->options[LIS_EOPTIONS_INITGUESS_ONES]) ECHK(lis_vector_set_all(1.0, x));
    if (oshift != 0.0) ECHK(lis_matrix_shift_diagonal(A, oshift));
    if (output) lis_printf(LIS_COMM_WORLD, "shift                 : %e\n", (double)oshift);
    ECHK(normalize(x));
    ECHK(lis_matvec(A, x, Ax));
    ECHK(lis_vector_set_all(0.0, p));
    ECHK(lis_vector_set_all(0.0, Ap));
    ECHK(inner_solver(esolver, "-i bicg -p none", &solver));
    err = lis_solve_setup(A, solver);                                          /* the solver only serves as the preconditioner's owner */
    if (!err) err = lis_precon_create(solver, &precon);
    if (err) { lis_solver_destroy(solver); return err; }
    solver->precon = precon;
#define CRCHK(e) do { err = (e); if (err) goto fail; } while (0)
    CRCHK(lis_vector_dot(x, Ax, &lambda));                                       /* lambda = <Ax,x>/<x,x> */
    CRCHK(lis_vector_axpyz(-lambda, x, Ax, r));                                  /* r = lambda x - A x */
    CRCHK(lis_vector_scale(-1.0, r));
    CRCHK(lis_vector_copy(r, p));
    CRCHK(lis_matvec(A, p, Ap));
    while (iter < emaxiter) {
        iter = iter + 1;
        CRCHK(lis_vector_dot(r, Ap, &rAp));
        CRCHK(lis_vector_dot(r, p, &rp));
        CRCHK(lis_vector_dot(Ap, Ap, &ApAp));
        CRCHK(lis_vector_dot(p, Ap, &pAp));
        CRCHK(lis_vector_dot(p, p, &pp));
        alpha = (rAp - lambda * rp) / (ApAp - 2.0 * lambda * pAp + lambda * lambda * pp);
        CRCHK(lis_vector_axpy(alpha, p, x));
        CRCHK(lis_matvec(A, x, Ax));
        CRCHK(lis_vector_dot(x, Ax, &lambda));
        CRCHK(lis_vector_nrm2(x, &nrm2));
        lambda = lambda / (nrm2 * nrm2);
        CRCHK(lis_vector_axpyz(-lambda, x, Ax, r));
        CRCHK(lis_vector_scale(-1.0, r));
        time = lis_wtime();
        CRCHK(lis_psolve(solver, r, w));
        ptime += lis_wtime() - time;
        CRCHK(lis_matvec(A, w, Aw));
        CRCHK(lis_vector_dot(Aw, Ap, &AwAp));
        CRCHK(lis_vector_dot(p, Aw, &pAw));
        CRCHK(lis_vector_dot(w, Ap, &wAp));
        CRCHK(lis_vector_dot(w, p, &wp));
        beta = -(AwAp - lambda * (pAw + wAp) + lambda * lambda * wp) / (ApAp - 2.0 * lambda * pAp + lambda * lambda * pp);
        CRCHK(lis_vector_xpay(w, beta, p));
        CRCHK(lis_vector_xpay(Aw, beta, Ap));
        CRCHK(lis_vector_nrm2(r, &nrm2));
        resid = nrm2 / fabs(lambda);
        erecord(esolver, output, iter, resid);
        if (resid < tol) break;
    }
    esolver->iter[0] = iter;
    esolver->resid[0] = resid;
    esolver->evalue[0] = lambda + oshift;
    CRCHK(normalize(x));
    esolver->ptime = ptime;
    esolver->itime = solver->itime;
    esolver->p_c_time = solver->p_c_time;
    esolver->p_i_time = solver->p_i_time;
    lis_precon_destroy(precon);
    solver->precon = NULL;
    lis_solver_destroy(solver);
    if (oshift != 0.0) ECHK(lis_matrix_shift_diagonal(A, -oshift));
    esolver->retcode = resid < tol ? LIS_SUCCESS : LIS_MAXITER;
    return esolver->retcode;
fail:
    lis_precon_destroy(precon);
    solver->precon = NULL;
    lis_solver_destroy(solver);
    return err;
#undef CRCHK
}

/* ------------------------------------------------------------------ subspace iteration, src/esolver/lis_esolver_si.c
 * work[0] = r, work[1] = q, v = &work[2] (v[1..ss]); inner eigensolver: power or inverse */
static LIS_INT lis_esi(LIS_ESOLVER esolver)
{
    LIS_MATRIX A = esolver->A;
    const LIS_INT ss = esolver->options[LIS_EOPTIONS_SUBSPACE], emaxiter = esolver->options[LIS_EOPTIONS_MAXITER];
    const LIS_INT output = esolver->options[LIS_EOPTIONS_OUTPUT], niesolver = esolver->options[LIS_EOPTIONS_INNER_ESOLVER];
    const LIS_REAL tol = esolver->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN];
    const LIS_SCALAR oshift = eshift(esolver);
    LIS_VECTOR r = esolver->work[0], q = esolver->work[1], *v = &esolver->work[2];
    LIS_SCALAR theta = 0.0, dot;
    LIS_REAL nrm2, resid = 0.0;
    LIS_INT iter = 0, j, k, err = LIS_SUCCESS;
    LIS_SOLVER solver = NULL;
    LIS_PRECON precon = NULL;
    char esolvername[128];
    if (niesolver != LIS_ESOLVER_PI && niesolver != LIS_ESOLVER_II) {
        LIS_SETERR1(LIS_ERR_ILL_ARG, "Parameter LIS_EOPTIONS_INNER_ESOLVER is %D (Set Power or Inverse for Subspace)\n", niesolver);
        return LIS_ERR_ILL_ARG;
    }
    ECHK(lis_vector_set_all(1.0, r));
    ECHK(normalize(r));
    if (oshift != 0.0) ECHK(lis_matrix_shift_diagonal(A, oshift));
    if (output) lis_printf(LIS_COMM_WORLD, "shift                 : %e\n", (double)oshift);
    lis_esolver_get_esolvername(niesolver, esolvername);
    if (output) lis_printf(LIS_COMM_WORLD, "inner eigensolver     : %s\n", esolvername);
    if (niesolver == LIS_ESOLVER_II) ECHK(inner_solver(esolver, "-i bicg -p none", &solver));
    if (output) {
        lis_printf(LIS_COMM_WORLD, "size of subspace      : %D\n\n", ss);
        lis_printf(LIS_COMM_WORLD, "compute eigenpairs in subspace:\n\n");
    }
#define SCHK(e) do { err = (e); if (err) goto done; } while (0)
    j = 0;
    while (j < ss) {
        const double etime0 = lis_wtime();
        SCHK(lis_vector_duplicate(A, &esolver->evector[j]));
        j = j + 1;
        SCHK(lis_vector_copy(r, v[j]));
        if (niesolver == LIS_ESOLVER_II) {
            if (precon) lis_precon_destroy(precon);             /* the reference creates one per eigenpair and leaks all but the last */
            precon = NULL;
            solver->A = A;
            SCHK(lis_precon_create(solver, &precon));
        }
        iter = 0;
        while (iter < emaxiter) {
            iter = iter + 1;
            for (k = 1; k < j; k++) {                            /* orthogonalise against the converged vectors */
                SCHK(lis_vector_dot(v[j], v[k], &dot));
                SCHK(lis_vector_axpy(-dot, v[k], v[j]));
            }
            if (niesolver == LIS_ESOLVER_PI) SCHK(lis_matvec(A, v[j], r));
            else SCHK(lis_solve_kernel(A, v[j], r, solver, precon));
            if (j == 1 && niesolver == LIS_ESOLVER_II) add_solver_times(esolver, solver);
            SCHK(lis_vector_nrm2(r, &nrm2));
            SCHK(lis_vector_dot(v[j], r, &theta));
            SCHK(lis_vector_axpyz(-theta, v[j], r, q));
            SCHK(lis_vector_nrm2(q, &resid));
            resid = resid / fabs(theta);
            SCHK(lis_vector_scale(1.0 / nrm2, r));
            SCHK(lis_vector_copy(r, v[j]));
            if (j == 1) {
                if (output & LIS_EPRINT_MEM) esolver->rhistory[iter] = resid;
                esolver->iter[j - 1] = iter;
            }
            if (output & LIS_EPRINT_OUT) lis_host_print_rhistory(iter, resid);
            if (tol > resid) break;
        }
        esolver->evalue[j - 1] = (niesolver == LIS_ESOLVER_PI ? theta : 1 / theta) + oshift;
        esolver->resid[j - 1] = resid;
        esolver->iter[j - 1] = iter;
        SCHK(lis_vector_copy(v[j], esolver->evector[j - 1]));
        if (output && ss > 1) {
            lis_printf(LIS_COMM_WORLD, "Subspace: mode number          = %D\n", j - 1);
            lis_printf(LIS_COMM_WORLD, "Subspace: eigenvalue           = %e\n", (double)esolver->evalue[j - 1]);
            lis_printf(LIS_COMM_WORLD, "Subspace: elapsed time         = %e sec.\n", lis_wtime() - etime0);
            lis_printf(LIS_COMM_WORLD, "Subspace: number of iterations = %D\n", iter);
            lis_printf(LIS_COMM_WORLD, "Subspace: relative residual    = %e\n\n", (double)resid);
        }
    }
    if (oshift != 0.0) SCHK(lis_matrix_shift_diagonal(A, -oshift));
    SCHK(lis_vector_copy(esolver->evector[0], esolver->x));
done:
    if (precon) lis_precon_destroy(precon);
    if (solver) lis_solver_destroy(solver);
    return err;
#undef SCHK
}

/* Lanczos / Arnoldi, second half: every Ritz value becomes the shift of one run of the inner
 * eigensolver (lis_esolver_li.c:300-370, lis_esolver_ai.c:330-395); mode 0 also supplies the
 * residual history and the times */
static LIS_INT refine_ritz_pairs(LIS_ESOLVER esolver, const char *who)
{
    LIS_MATRIX A = esolver->A;
    const LIS_INT ss = esolver->options[LIS_EOPTIONS_SUBSPACE], output = esolver->options[LIS_EOPTIONS_OUTPUT];
    LIS_ESOLVER esolver2;
    LIS_SCALAR evalue = 0.0;
    LIS_INT err = LIS_SUCCESS;
    if (output) lis_printf(LIS_COMM_WORLD, "computing refined eigenpairs using inner eigensolver:\n\n");
    ECHK(lis_esolver_create(&esolver2));
    esolver2->options[LIS_EOPTIONS_ESOLVER] = esolver->options[LIS_EOPTIONS_INNER_ESOLVER];
    esolver2->options[LIS_EOPTIONS_SUBSPACE] = 1;
    esolver2->options[LIS_EOPTIONS_MAXITER] = esolver->options[LIS_EOPTIONS_MAXITER];
    esolver2->options[LIS_EOPTIONS_OUTPUT] = output;
    esolver2->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN] = esolver->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN];
    for (LIS_INT i = 0; i < ss && !err; i++) {
        err = lis_vector_duplicate(A, &esolver->evector[i]);
        if (err) break;
        esolver2->ishift = esolver->evalue[i];
        err = lis_esolve(A, esolver->evector[i], &evalue, esolver2);
        if (err) break;
        lis_esolver_work_destroy(esolver2);
        esolver->evalue[i] = evalue;
        esolver->iter[i] = esolver2->iter[0];
        esolver->resid[i] = esolver2->resid[0];
        if (i == 0) {
            if (output & LIS_EPRINT_MEM) for (LIS_INT ic = 0; ic < esolver2->iter[0] + 1; ic++) esolver->rhistory[ic] = esolver2->rhistory[ic];
            esolver->ptime += esolver2->ptime;
            esolver->itime += esolver2->itime;
            esolver->p_c_time += esolver2->p_c_time;
            esolver->p_i_time += esolver2->p_i_time;
        }
        if (output) {
            lis_printf(LIS_COMM_WORLD, "%s: mode number          = %D\n", who, i);
            lis_printf(LIS_COMM_WORLD, "%s: eigenvalue           = %e\n", who, (double)esolver->evalue[i]);
            lis_printf(LIS_COMM_WORLD, "%s: elapsed time         = %e sec.\n", who, esolver2->time);
            lis_printf(LIS_COMM_WORLD, "%s: number of iterations = %D\n", who, esolver2->iter[0]);
            lis_printf(LIS_COMM_WORLD, "%s: relative residual    = %e\n\n", who, (double)esolver2->resid[0]);
        }
    }
    lis_esolver_destroy(esolver2);
    if (!err) err = lis_vector_copy(esolver->evector[0], esolver->x);
    return err;
}

/* ------------------------------------------------------------------ Lanczos, src/esolver/lis_esolver_li.c
 * work[0] = r, v = &work[1]; Ritz values of the ss x ss tridiagonal matrix by QR iteration, then each
 * refined by the inner eigensolver with the Ritz value as shift */
static LIS_INT lis_eli(LIS_ESOLVER esolver)
{
    LIS_MATRIX A = esolver->A;
    const LIS_INT ss = esolver->options[LIS_EOPTIONS_SUBSPACE];
    const LIS_INT output = esolver->options[LIS_EOPTIONS_OUTPUT], niesolver = esolver->options[LIS_EOPTIONS_INNER_ESOLVER];
    const LIS_INT rval = esolver->options[LIS_EOPTIONS_RVAL];
    const LIS_REAL tol = esolver->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN];
    LIS_VECTOR r = esolver->work[0], *v = &esolver->work[1];
    LIS_SCALAR *t, *tq, *tr, dot;
    LIS_REAL nrm2, qrerr, beta;
    LIS_INT i, j, k, qriter, err = LIS_SUCCESS;
    char esolvername[128];
    if (niesolver < LIS_ESOLVER_PI || niesolver > LIS_ESOLVER_CR) {
        LIS_SETERR1(LIS_ERR_NOT_IMPLEMENTED, "inner eigensolver %D is not available for Lanczos in lis_b200\n", niesolver);
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    t = (LIS_SCALAR *)lis_malloc((size_t)ss * ss * sizeof(LIS_SCALAR), "lis_eli::t");
    tq = (LIS_SCALAR *)lis_malloc((size_t)ss * ss * sizeof(LIS_SCALAR), "lis_eli::tq");
    tr = (LIS_SCALAR *)lis_malloc((size_t)ss * ss * sizeof(LIS_SCALAR), "lis_eli::tr");
    if (!t || !tq || !tr) { lis_free2(3, t, tq, tr); LIS_SETERR_MEM(ss * ss * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
#define LCHK(e) do { err = (e); if (err) goto done; } while (0)
    LCHK(lis_vector_set_all(0.0, v[0]));
    LCHK(lis_vector_set_all(1.0, r));
    LCHK(lis_vector_nrm2(r, &nrm2));
    {
        /* the reference creates (and prints) a linear solver here that the loop never uses */
        LIS_SOLVER solver;
        lis_esolver_get_esolvername(niesolver, esolvername);
        if (output) lis_printf(LIS_COMM_WORLD, "inner eigensolver     : %s\n", esolvername);
        LCHK(inner_solver(esolver, "-i bicg -p none", &solver));
        lis_solver_destroy(solver);
    }
    for (i = 0; i < ss * ss; i++) t[i] = 0.0;
    j = 0;
    while (j < ss - 1) {
        j = j + 1;
        LCHK(lis_vector_copy(r, v[j]));
        if (j == 1) {
            LCHK(lis_vector_scale(1.0 / nrm2, v[j]));
            LCHK(lis_matvec(A, v[j], r));
        } else {
            LCHK(lis_vector_scale(1.0 / t[(j - 2) * ss + j - 1], v[j]));
            LCHK(lis_matvec(A, v[j], r));
            LCHK(lis_vector_axpy(-t[(j - 2) * ss + j - 1], v[j - 1], r));
        }
        LCHK(lis_vector_dot(v[j], r, &t[(j - 1) * ss + j - 1]));                  /* alpha(j) */
        LCHK(lis_vector_axpy(-t[(j - 1) * ss + j - 1], v[j], r));
        for (k = 1; k < j; k++) {                                                 /* reorthogonalisation */
            LCHK(lis_vector_dot(v[j], v[k], &dot));
            LCHK(lis_vector_axpy(-dot, v[k], v[j]));
        }
        LCHK(lis_vector_nrm2(r, &beta));                                          /* beta(j) */
        t[(j - 1) * ss + j] = beta;
        if (fabs(t[(j - 1) * ss + j]) < tol) break;
        t[j * ss + j - 1] = t[(j - 1) * ss + j];
    }
    {
        const double time0 = lis_wtime();
        arr_qr(ss, t, tq, tr, &qriter, &qrerr);
        for (i = 0; i < ss; i++) esolver->evalue[i] = t[i * ss + i];
        if (output) {
            lis_printf(LIS_COMM_WORLD, "size of subspace      : %D\n\n", ss);
            lis_printf(LIS_COMM_WORLD, "Ritz values:\n\n");
            for (i = 0; i < ss; i++) {
                lis_printf(LIS_COMM_WORLD, "Lanczos: mode number          = %D\n", i);
                lis_printf(LIS_COMM_WORLD, "Lanczos: Ritz value           = %e\n", (double)esolver->evalue[i]);
            }
            lis_printf(LIS_COMM_WORLD, "Lanczos: elapsed time         = %e sec.\n\n", lis_wtime() - time0);
        }
    }
    if (rval) goto done;
    LCHK(refine_ritz_pairs(esolver, "Lanczos"));
done:
    lis_free2(3, t, tq, tr);
    return err;
#undef LCHK
}

/* ------------------------------------------------------------------ Arnoldi, src/esolver/lis_esolver_ai.c
 * work[0] = w, v = &work[1]; Ritz values from the ss x ss Hessenberg matrix by QR iteration (1x1 and 2x2
 * diagonal blocks; a complex pair contributes its real part), then refined like Lanczos' */
static LIS_INT lis_eai(LIS_ESOLVER esolver)
{
    LIS_MATRIX A = esolver->A;
    const LIS_INT ss = esolver->options[LIS_EOPTIONS_SUBSPACE];
    const LIS_INT output = esolver->options[LIS_EOPTIONS_OUTPUT], niesolver = esolver->options[LIS_EOPTIONS_INNER_ESOLVER];
    const LIS_INT rval = esolver->options[LIS_EOPTIONS_RVAL];
    const LIS_REAL tol = esolver->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN];
    LIS_VECTOR w = esolver->work[0], *v = &esolver->work[1];
    LIS_SCALAR *h, *hq, *hr;
    LIS_REAL hqrerr, D, nrm;
    LIS_INT i, j, hqriter, err = LIS_SUCCESS;
    char esolvername[128];
    if (niesolver < LIS_ESOLVER_PI || niesolver > LIS_ESOLVER_CR) {
        LIS_SETERR1(LIS_ERR_NOT_IMPLEMENTED, "inner eigensolver %D is not available for Arnoldi in lis_b200\n", niesolver);
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    /* one spare entry: the reference's last h(j+1,j) lands one past its ss*ss array */
    h = (LIS_SCALAR *)lis_malloc(((size_t)ss * ss + 1) * sizeof(LIS_SCALAR), "lis_eai::h");
    hq = (LIS_SCALAR *)lis_malloc((size_t)ss * ss * sizeof(LIS_SCALAR), "lis_eai::hq");
    hr = (LIS_SCALAR *)lis_malloc((size_t)ss * ss * sizeof(LIS_SCALAR), "lis_eai::hr");
    if (!h || !hq || !hr) { lis_free2(3, h, hq, hr); LIS_SETERR_MEM(ss * ss * sizeof(LIS_SCALAR)); return LIS_OUT_OF_MEMORY; }
#define ACHK(e) do { err = (e); if (err) goto done; } while (0)
    ACHK(lis_vector_set_all(1.0, v[0]));
    ACHK(normalize(v[0]));
    {
        LIS_SOLVER solver;
        lis_esolver_get_esolvername(niesolver, esolvername);
        if (output) lis_printf(LIS_COMM_WORLD, "inner eigensolver     : %s\n", esolvername);
        ACHK(inner_solver(esolver, "-i bicg -p none", &solver));
        lis_solver_destroy(solver);
    }
    for (i = 0; i < ss * ss; i++) h[i] = 0.0;
    j = -1;
    while (j < ss - 1) {
        j = j + 1;
        ACHK(lis_matvec(A, v[j], w));
        for (i = 0; i <= j; i++) {
            ACHK(lis_vector_dot(v[i], w, &h[i + j * ss]));
            ACHK(lis_vector_axpy(-h[i + j * ss], v[i], w));
        }
        ACHK(lis_vector_nrm2(w, &nrm));
        h[j + 1 + j * ss] = nrm;
        if (fabs(h[j + 1 + j * ss]) < tol) break;
        ACHK(lis_vector_scale(1 / h[j + 1 + j * ss], w));
        ACHK(lis_vector_copy(w, v[j + 1]));
    }
    {
        const double time0 = lis_wtime();
        arr_qr(ss, h, hq, hr, &hqriter, &hqrerr);
        const double time = lis_wtime() - time0;
        if (output) {
            lis_printf(LIS_COMM_WORLD, "size of subspace      : %D\n\n", ss);
            lis_printf(LIS_COMM_WORLD, "Ritz values:\n\n");
        }
        i = 0;
        while (i < ss) {
            i = i + 1;
            if (ss == i || fabs(h[i + (i - 1) * ss]) < tol) {
                if (output) {
                    lis_printf(LIS_COMM_WORLD, "Arnoldi: mode number          = %D\n", i - 1);
                    lis_printf(LIS_COMM_WORLD, "Arnoldi: Ritz value           = %e\n", (double)(h[i - 1 + (i - 1) * ss]));
                }
                esolver->evalue[i - 1] = h[i - 1 + (i - 1) * ss];
            } else {
                D = (h[i - 1 + (i - 1) * ss] + h[i + i * ss]) * (h[i - 1 + (i - 1) * ss] + h[i + i * ss])
                  - 4 * (h[i - 1 + (i - 1) * ss] * h[i + i * ss] - h[i - 1 + i * ss] * h[i + (i - 1) * ss]);
                if (D < 0) {
                    if (output) {
                        lis_printf(LIS_COMM_WORLD, "Arnoldi: mode number          = %D\n", i - 1);
                        lis_printf(LIS_COMM_WORLD, "Arnoldi: Ritz value           = (%e, %e)\n", (double)((h[i - 1 + (i - 1) * ss] + h[i + i * ss]) / 2), (double)sqrt(-D) / 2);
                        lis_printf(LIS_COMM_WORLD, "Arnoldi: mode number          = %D\n", i);
                        lis_printf(LIS_COMM_WORLD, "Arnoldi: Ritz value           = (%e, %e)\n", (double)((h[i - 1 + (i - 1) * ss] + h[i + i * ss]) / 2), (double)-sqrt(-D) / 2);
                    }
                    esolver->evalue[i - 1] = (h[i - 1 + (i - 1) * ss] + h[i + i * ss]) / 2;
                    esolver->evalue[i] = (h[i - 1 + (i - 1) * ss] + h[i + i * ss]) / 2;
                    i = i + 1;
                } else {
                    if (output) {
                        lis_printf(LIS_COMM_WORLD, "Arnoldi: mode number          = %D\n", i - 1);
                        lis_printf(LIS_COMM_WORLD, "Arnoldi: Ritz value           = %e\n", (double)(h[i - 1 + (i - 1) * ss]));
                    }
                    esolver->evalue[i - 1] = h[i - 1 + (i - 1) * ss];
                }
            }
        }
        lis_printf(LIS_COMM_WORLD, "Arnoldi: elapsed time         = %e sec.\n\n", time);     /* unconditional there too */
    }
    if (rval) goto done;
    ACHK(refine_ritz_pairs(esolver, "Arnoldi"));
done:
    lis_free2(3, h, hq, hr);
    return err;
#undef ACHK
}

/* ------------------------------------------------------------------ lis_esolve (lis_gesolve with B = NULL,
 * src/esolver/lis_esolver.c:286-665) */
typedef struct { LIS_INT (*run)(LIS_ESOLVER); LIS_INT nwork; int per_ss; } esolver_entry_t;
static const esolver_entry_t k_esolvers[LIS_ESOLVER_LEN + 1] = {
    {NULL, 0, 0},
    {lis_epi, 2, 0}, {lis_eii, 2, 0}, {lis_erqi, 2, 0}, {lis_ecg, 6, 0}, {lis_ecr, 6, 0},
    {lis_esi, 4, 1},           /* lis_esi_malloc_work: 4 + ss */
    {lis_eli, 2, 1},           /* lis_eli_malloc_work: 2 + ss */
    {lis_eai, 2, 1},           /* lis_eai_malloc_work: 2 + ss */
    {NULL, 0, 0}, {NULL, 0, 0}, {NULL, 0, 0}, {NULL, 0, 0}, {NULL, 0, 0}, {NULL, 0, 0}, {NULL, 0, 0}, {NULL, 0, 0},
};

LIS_INT lis_esolve(LIS_MATRIX A, LIS_VECTOR x, LIS_SCALAR *evalue0, LIS_ESOLVER esolver)
{
    LIS_INT err = lis_host_matrix_check_input(A);
    if (err) return err;
    if (x == NULL) { LIS_SETERR(LIS_ERR_ILL_ARG, "vector x is undefined\n"); return LIS_ERR_ILL_ARG; }
    if (A->n != x->n) return LIS_ERR_ILL_ARG;
    if (A->gn <= 0) { LIS_SETERR1(LIS_ERR_ILL_ARG, "Size n(=%D) of matrix A is less than 0\n", A->gn); return LIS_ERR_ILL_ARG; }
    const LIS_INT nesolver = esolver->options[LIS_EOPTIONS_ESOLVER];
    const LIS_INT ss = esolver->options[LIS_EOPTIONS_SUBSPACE], mode = esolver->options[LIS_EOPTIONS_MODE];
    const LIS_INT emaxiter = esolver->options[LIS_EOPTIONS_MAXITER], output = esolver->options[LIS_EOPTIONS_OUTPUT];
    const LIS_INT estorage = esolver->options[LIS_EOPTIONS_STORAGE], eblock = esolver->options[LIS_EOPTIONS_STORAGE_BLOCK];
    const LIS_INT eprecision = esolver->options[LIS_EOPTIONS_PRECISION];
    esolver->eprecision = eprecision;
    if (nesolver < 1 || nesolver > LIS_ESOLVER_LEN) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "Parameter LIS_EOPTIONS_ESOLVER is %D (Set between 1 to %D)\n", nesolver, LIS_ESOLVER_LEN);
        return LIS_ERR_ILL_ARG;
    }
    if (k_esolvers[nesolver].run == NULL) {
        LIS_SETERR1(LIS_ERR_NOT_IMPLEMENTED, "eigensolver %s is not part of lis_b200 (the generalized eigensolvers are not carried)\n", k_esolvername[nesolver]);
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    if (eprecision != LIS_PRECISION_DOUBLE) { LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "quad precision is not part of lis_b200\n"); return LIS_ERR_NOT_IMPLEMENTED; }
    if (k_esolvers[nesolver].per_ss && ss > A->gn) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "Parameter LIS_EOPTIONS_SUBSPACE is %D (Set less than or equal to matrix size %D)\n", ss, A->gn);
        return LIS_ERR_ILL_ARG;
    }
    if (k_esolvers[nesolver].per_ss && mode >= ss) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "Parameter LIS_EOPTIONS_MODE is %D (Set less than subspace size %D)\n", mode, ss);
        return LIS_ERR_ILL_ARG;
    }
    if (ss < 1) { LIS_SETERR1(LIS_ERR_ILL_ARG, "Parameter LIS_EOPTIONS_SUBSPACE is %D (Set 1 or more)\n", ss); return LIS_ERR_ILL_ARG; }

    /* result arrays: ss + 2 entries each, the residual history emaxiter + 2 */
    if (esolver->evalue) lis_free(esolver->evalue);
    if (esolver->resid) lis_free(esolver->resid);
    if (esolver->iter) lis_free(esolver->iter);
    if (esolver->iter2) lis_free(esolver->iter2);
    if (esolver->rhistory) lis_free(esolver->rhistory);
    evectors_free(esolver);
    esolver->evalue = (LIS_SCALAR *)lis_calloc((size_t)(ss + 2) * sizeof(LIS_SCALAR), "lis_esolve::evalue");
    esolver->resid = (LIS_REAL *)lis_calloc((size_t)(ss + 2) * sizeof(LIS_REAL), "lis_esolve::resid");
    esolver->iter = (LIS_INT *)lis_calloc((size_t)(ss + 2) * sizeof(LIS_SCALAR), "lis_esolve::iter");
    esolver->iter2 = (LIS_INT *)lis_calloc((size_t)(ss + 2) * sizeof(LIS_SCALAR), "lis_esolve::iter2");
    esolver->evector = (LIS_VECTOR *)lis_calloc((size_t)(ss + 2) * sizeof(LIS_VECTOR), "lis_esolve::evector");
    esolver->rhistory = (LIS_REAL *)lis_calloc((size_t)(emaxiter + 2) * sizeof(LIS_REAL), "lis_esolve::rhistory");
    if (!esolver->evalue || !esolver->resid || !esolver->iter || !esolver->iter2 || !esolver->evector || !esolver->rhistory) {
        LIS_SETERR_MEM((ss + 2) * sizeof(LIS_SCALAR));
        esolver->retcode = LIS_OUT_OF_MEMORY;
        return LIS_OUT_OF_MEMORY;
    }
    esolver->nevector = keeps_evectors(nesolver) ? ss + 2 : 0;
    esolver->evalue[0] = 1.0;
    esolver->evalue[ss - 1] = 1.0;
    esolver->rhistory[0] = 1.0;

    LIS_VECTOR xx;
    err = lis_vector_duplicate(A, &xx);
    if (err) { esolver->retcode = err; return err; }
    if (esolver->options[LIS_EOPTIONS_INITGUESS_ONES]) {
        if (output) lis_printf(LIS_COMM_WORLD, "initial vector x      : all components set to 1\n");
        err = lis_vector_set_all(1.0, xx);
    } else {
        if (output) lis_printf(LIS_COMM_WORLD, "initial vector x      : user defined\n");
        err = lis_vector_copy(x, xx);
    }
    if (err) { lis_vector_destroy(xx); esolver->retcode = err; return err; }

    if (estorage > 0 && A->matrix_type != estorage) {           /* -estorage: converts A in place, like -storage */
        LIS_MATRIX A0;
        err = lis_matrix_duplicate(A, &A0);
        if (err) { lis_vector_destroy(xx); return err; }
        lis_matrix_set_blocksize(A0, eblock, eblock, NULL, NULL);
        lis_matrix_set_type(A0, estorage);
        err = lis_matrix_convert(A, A0);
        if (err) { lis_matrix_destroy(A0); lis_vector_destroy(xx); return err; }
        lis_host_matrix_adopt(A, A0);
    }
    esolver->A = A;
    esolver->B = NULL;
    if (output) {
        lis_printf(LIS_COMM_WORLD, "precision             : %s\n", k_eprecision_atoi[eprecision]);
        lis_printf(LIS_COMM_WORLD, "eigensolver           : %s\n", k_esolvername[nesolver]);
        lis_printf(LIS_COMM_WORLD, "convergence condition : ||lx-(B^-1)Ax||_2 <= %6.1e * ||lx||_2\n", (double)esolver->params[LIS_EPARAMS_RESID - LIS_EOPTIONS_LEN]);
        if (A->matrix_type == LIS_MATRIX_BSR) lis_printf(LIS_COMM_WORLD, "matrix storage format : %s(%D x %D)\n", k_estoragename[A->matrix_type - 1], eblock, eblock);
        else lis_printf(LIS_COMM_WORLD, "matrix storage format : %s\n", k_estoragename[A->matrix_type - 1]);
    }
    const double time = lis_wtime();
    esolver->ptime = 0; esolver->itime = 0; esolver->p_c_time = 0; esolver->p_i_time = 0;
    err = ework(esolver, k_esolvers[nesolver].nwork + (k_esolvers[nesolver].per_ss ? ss : 0));
    if (err) { lis_vector_destroy(xx); esolver->retcode = err; return err; }
    esolver->x = xx;
    esolver->xx = x;

    err = k_esolvers[nesolver].run(esolver);
    esolver->retcode = err;
    if (err == LIS_ERR_DEVICE || err == LIS_ERR_OUT_OF_MEMORY || err == LIS_ERR_NOT_IMPLEMENTED || err == LIS_ERR_ILL_ARG) {
        lis_vector_destroy(xx);
        esolver->x = NULL;
        return err;
    }
    *evalue0 = esolver->evalue[0];
    lis_vector_copy(esolver->x, x);
    esolver->time = lis_wtime() - time;
    if (output) {
        if (err) lis_printf(LIS_COMM_WORLD, "eigensolver status    : %s(code=%D)\n\n", k_ereturncode[err], err);
        else lis_printf(LIS_COMM_WORLD, "eigensolver status    : normal end\n\n");
    }
    esolver->iter2[mode] = esolver->iter[mode];
    lis_vector_destroy(xx);
    esolver->x = NULL;
    return LIS_SUCCESS;
}

/* The reference's lis_esolve is lis_gesolve(A, NULL, ...) (src/esolver/lis_esolver.c:263-282).  The generalized
 * problem A x = lambda B x (its eight g-solvers) is not carried over: B must be NULL. */
LIS_INT lis_gesolve(LIS_MATRIX A, LIS_MATRIX B, LIS_VECTOR x, LIS_SCALAR *evalue0, LIS_ESOLVER esolver)
{
    if (B != NULL) {
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "generalized eigenproblems (A x = lambda B x) are not available; pass B = NULL for the standard problem\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    return lis_esolve(A, x, evalue0, esolver);
}
