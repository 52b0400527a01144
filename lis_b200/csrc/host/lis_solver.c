/*
 * lis_solver.c -- LIS_SOLVER objects, option parsing, lis_solve / lis_solve_kernel and the
 * shared pieces of the Krylov drivers (initial residual, residual norm, work vectors).
 *
 * Host C restatement of the control flow in src/solver/lis_solver.c of the reference:
 *   defaults :242-284, option tables :175-199, set_option :1095-1243, lis_solve :367-406,
 *   lis_solve_kernel :441-952, lis_solver_get_initial_residual :957-1090, residual functions
 *   :1792-1813, shadow residual :1817-1875, getters :1640-1790.
 * Every vector/matrix operation inside is a CUDA kernel launch; scalars stay on the host.
 * Not carried over (outside the hot path, rejected with an error instead of ignored):
 * quad/switch precision, -use_at, I+S preconditioning.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <math.h>
#include "lis_device.h"
#include "lis_host.h"

/* ------------------------------------------------------------------ tables */
static const char *k_solver_atoi[] = {"cg", "bicg", "cgs", "bicgstab", "bicgstabl", "gpbicg", "tfqmr", "orthomin", "gmres",
    "jacobi", "gs", "sor", "bicgsafe", "cr", "bicr", "crs", "bicrstab", "gpbicr", "bicrsafe", "fgmres", "idrs", "idr1",
    "minres", "cocg", "cocr"};
static const char *k_precon_atoi[] = {"none", "jacobi", "ilu", "ssor", "hybrid", "is", "sainv", "saamg", "iluc", "ilut", "bjacobi", ""};
static const char *k_storage_atoi[] = {"csr", "csc", "msr", "dia", "ell", "jad", "bsr", "bsc", "vbr", "coo", "dns"};
static const char *k_print_atoi[] = {"none", "mem", "out", "all"};
static const char *k_scale_atoi[] = {"none", "jacobi", "symm_diag"};
static const char *k_truefalse_atoi[] = {"false", "true"};
static const char *k_precision_atoi[] = {"double", "quad", "switch"};
static const char *k_conv_cond_atoi[] = {"nrm2_r", "nrm2_b", "nrm1_b"};

static const char *k_solvername[] = {"", "CG", "BiCG", "CGS", "BiCGSTAB", "BiCGSTAB(l)", "GPBiCG", "TFQMR", "Orthomin", "GMRES",
    "Jacobi", "Gauss-Seidel", "SOR", "BiCGSafe", "CR", "BiCR", "CRS", "BiCRSTAB", "GPBiCR", "BiCRSafe", "FGMRES", "IDR(s)",
    "IDR(1)", "MINRES", "COCG", "COCR"};
static const char *k_preconname[] = {"none", "Jacobi", "ILU", "SSOR", "Hybrid", "I+S", "SAINV", "SAAMG", "Crout ILU", "ILUT", "Block Jacobi"};
static const char *k_returncode[] = {"LIS_SUCCESS", "LIS_ILL_OPTION", "LIS_BREAKDOWN", "LIS_OUT_OF_MEMORY", "LIS_MAXITER",
    "LIS_NOT_IMPLEMENTED", "LIS_ERR_FILE_IO", "LIS_ERR_DEVICE"};
static const char *k_storagename[] = {"CSR", "CSC", "MSR", "DIA", "ELL", "JAD", "BSR", "BSC", "VBR", "COO", "DNS"};

/* option name -> slot; slots >= LIS_OPTIONS_LEN address params[] */
#define OPT_FILE (-1)
typedef struct { const char *name; int slot; } lis_optdef_t;
static const lis_optdef_t k_options[] = {
    {"-maxiter", LIS_OPTIONS_MAXITER}, {"-tol", LIS_PARAMS_RESID}, {"-print", LIS_OPTIONS_OUTPUT}, {"-scale", LIS_OPTIONS_SCALE},
    {"-ssor_omega", LIS_PARAMS_SSOR_OMEGA}, {"-ilu_fill", LIS_OPTIONS_FILL}, {"-ilu_relax", LIS_PARAMS_RELAX},
    {"-is_alpha", LIS_PARAMS_ALPHA}, {"-is_level", LIS_OPTIONS_ISLEVEL}, {"-is_m", LIS_OPTIONS_M},
    {"-hybrid_maxiter", LIS_OPTIONS_PMAXITER}, {"-hybrid_ell", LIS_OPTIONS_PELL}, {"-hybrid_restart", LIS_OPTIONS_PRESTART},
    {"-hybrid_tol", LIS_PARAMS_PRESID}, {"-hybrid_omega", LIS_PARAMS_POMEGA}, {"-hybrid_i", LIS_OPTIONS_PSOLVER},
    {"-sainv_drop", LIS_PARAMS_DROP}, {"-ric2s_tau", LIS_PARAMS_TAU}, {"-ric2s_sigma", LIS_PARAMS_SIGMA},
    {"-ric2s_gamma", LIS_PARAMS_GAMMA}, {"-restart", LIS_OPTIONS_RESTART}, {"-ell", LIS_OPTIONS_ELL}, {"-omega", LIS_PARAMS_OMEGA},
    {"-i", LIS_OPTIONS_SOLVER}, {"-p", LIS_OPTIONS_PRECON}, {"-f", LIS_OPTIONS_PRECISION}, {"-h", OPT_FILE}, {"-ver", OPT_FILE},
    {"-hybrid_p", LIS_OPTIONS_PPRECON}, {"-initx_zeros", LIS_OPTIONS_INITGUESS_ZEROS}, {"-adds", LIS_OPTIONS_ADDS},
    {"-adds_iter", LIS_OPTIONS_ADDS_ITER}, {"-use_at", LIS_OPTIONS_USE_AT}, {"-switch_tol", LIS_PARAMS_SWITCH_RESID},
    {"-switch_maxiter", LIS_OPTIONS_SWITCH_MAXITER}, {"-saamg_unsym", LIS_OPTIONS_SAAMG_UNSYM}, {"-iluc_drop", LIS_PARAMS_DROP},
    {"-iluc_gamma", LIS_PARAMS_GAMMA}, {"-iluc_rate", LIS_PARAMS_RATE}, {"-storage", LIS_OPTIONS_STORAGE},
    {"-storage_block", LIS_OPTIONS_STORAGE_BLOCK}, {"-conv_cond", LIS_OPTIONS_CONV_COND}, {"-tol_w", LIS_PARAMS_RESID_WEIGHT},
    {"-saamg_theta", LIS_PARAMS_SAAMG_THETA}, {"-irestart", LIS_OPTIONS_IDRS_RESTART},
};

/* ------------------------------------------------------------------ create / destroy */
static void solver_init(LIS_SOLVER s)
{
    memset(s, 0, sizeof(struct LIS_SOLVER_STRUCT));
    s->precision = LIS_PRECISION_DOUBLE;
    s->options[LIS_OPTIONS_SOLVER] = LIS_SOLVER_BICG;
    s->options[LIS_OPTIONS_PRECON] = LIS_PRECON_TYPE_NONE;
    s->options[LIS_OPTIONS_OUTPUT] = LIS_FALSE;
    s->options[LIS_OPTIONS_MAXITER] = 1000;
    s->options[LIS_OPTIONS_RESTART] = 40;
    s->options[LIS_OPTIONS_ELL] = 2;
    s->options[LIS_OPTIONS_SCALE] = LIS_SCALE_NONE;
    s->options[LIS_OPTIONS_FILL] = 0;
    s->options[LIS_OPTIONS_M] = 3;
    s->options[LIS_OPTIONS_PSOLVER] = LIS_SOLVER_SOR;
    s->options[LIS_OPTIONS_PMAXITER] = 25;
    s->options[LIS_OPTIONS_PRESTART] = 40;
    s->options[LIS_OPTIONS_PELL] = 2;
    s->options[LIS_OPTIONS_PPRECON] = LIS_PRECON_TYPE_NONE;
    s->options[LIS_OPTIONS_ISLEVEL] = 1;
    s->options[LIS_OPTIONS_INITGUESS_ZEROS] = LIS_TRUE;
    s->options[LIS_OPTIONS_ADDS] = LIS_FALSE;
    s->options[LIS_OPTIONS_ADDS_ITER] = 1;
    s->options[LIS_OPTIONS_PRECISION] = LIS_PRECISION_DOUBLE;
    s->options[LIS_OPTIONS_USE_AT] = LIS_FALSE;
    s->options[LIS_OPTIONS_SWITCH_MAXITER] = -1;
    s->options[LIS_OPTIONS_SAAMG_UNSYM] = LIS_FALSE;
    s->options[LIS_OPTIONS_STORAGE] = 0;
    s->options[LIS_OPTIONS_STORAGE_BLOCK] = 2;
    s->options[LIS_OPTIONS_CONV_COND] = 0;
    s->options[LIS_OPTIONS_INIT_SHADOW_RESID] = LIS_RESID;
    s->options[LIS_OPTIONS_IDRS_RESTART] = 2;
    s->params[LIS_PARAMS_RESID - LIS_OPTIONS_LEN] = 1.0e-12;
    s->params[LIS_PARAMS_RESID_WEIGHT - LIS_OPTIONS_LEN] = 1.0;
    s->params[LIS_PARAMS_OMEGA - LIS_OPTIONS_LEN] = 1.9;
    s->params[LIS_PARAMS_SSOR_OMEGA - LIS_OPTIONS_LEN] = 1.0;
    s->params[LIS_PARAMS_RELAX - LIS_OPTIONS_LEN] = 1.0;
    s->params[LIS_PARAMS_DROP - LIS_OPTIONS_LEN] = 0.05;
    s->params[LIS_PARAMS_ALPHA - LIS_OPTIONS_LEN] = 1.0;
    s->params[LIS_PARAMS_TAU - LIS_OPTIONS_LEN] = 0.05;
    s->params[LIS_PARAMS_SIGMA - LIS_OPTIONS_LEN] = 2.0;
    s->params[LIS_PARAMS_GAMMA - LIS_OPTIONS_LEN] = 1.0;
    s->params[LIS_PARAMS_PRESID - LIS_OPTIONS_LEN] = 1.0e-3;
    s->params[LIS_PARAMS_POMEGA - LIS_OPTIONS_LEN] = 1.5;
    s->params[LIS_PARAMS_SWITCH_RESID - LIS_OPTIONS_LEN] = 1.0e-12;
    s->params[LIS_PARAMS_RATE - LIS_OPTIONS_LEN] = 5.0;
    s->params[LIS_PARAMS_SAAMG_THETA - LIS_OPTIONS_LEN] = 0.05;
    s->setup = LIS_FALSE;
}

LIS_INT lis_solver_create(LIS_SOLVER *solver)
{
    *solver = (LIS_SOLVER)lis_malloc(sizeof(struct LIS_SOLVER_STRUCT), "lis_solver_create::solver");
    if (*solver == NULL) { LIS_SETERR_MEM(sizeof(struct LIS_SOLVER_STRUCT)); return LIS_OUT_OF_MEMORY; }
    solver_init(*solver);
    return LIS_SUCCESS;
}

LIS_INT lis_solver_work_destroy(LIS_SOLVER solver)
{
    if (solver && solver->work) {
        for (LIS_INT i = 0; i < solver->worklen; i++)
            if (solver->work[i]) lis_vector_destroy(solver->work[i]);
        lis_free(solver->work);
        solver->work = NULL;
        solver->worklen = 0;
    }
    return LIS_SUCCESS;
}

LIS_INT lis_solver_destroy(LIS_SOLVER solver)
{
    if (solver) {
        lis_solver_work_destroy(solver);
        if (solver->d) lis_vector_destroy(solver->d);
        if (solver->rhistory) lis_free(solver->rhistory);
        lis_free(solver);
    }
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ options */
static LIS_INT set_keyword(const char *arg, const char **words, int nwords, int base, char maxdigit,
                           LIS_INT *dst, const char *what)
{
    if (arg[0] >= '0' && arg[0] <= maxdigit) { int v = 0; sscanf(arg, "%d", &v); *dst = v; return LIS_SUCCESS; }
    for (int i = 0; i < nwords; i++)
        if (strcmp(arg, words[i]) == 0) { *dst = i + base; return LIS_SUCCESS; }
    LIS_SETERR1(LIS_ERR_ILL_ARG, "Parameter %s is not correct\n", what);
    return LIS_ERR_ILL_ARG;
}

static LIS_INT set_precon(const char *arg, LIS_INT *dst, const char *what)
{
    if (arg[0] >= '0' && arg[0] <= '9') { int v = 0; sscanf(arg, "%d", &v); *dst = v; return LIS_SUCCESS; }
    for (int i = 0; i < LIS_PRECON_TYPE_LEN - 1; i++)
        if (strcmp(arg, k_precon_atoi[i]) == 0) { *dst = i; return LIS_SUCCESS; }
    const LIS_INT reg = lis_host_precon_lookup(arg);
    if (reg >= 0) { *dst = reg; return LIS_SUCCESS; }
    LIS_SETERR1(LIS_ERR_ILL_ARG, "Parameter %s is not correct\n", what);
    return LIS_ERR_ILL_ARG;
}

#define NWORDS(a) ((int)(sizeof(a) / sizeof((a)[0])))

static LIS_INT set_option2(const char *name, const char *value, LIS_SOLVER solver)
{
    LIS_INT err = LIS_SUCCESS;
    for (size_t k = 0; k < sizeof(k_options) / sizeof(k_options[0]); k++) {
        if (strcmp(name, k_options[k].name) != 0) continue;
        const int slot = k_options[k].slot;
        switch (slot) {
        case OPT_FILE: break;
        case LIS_OPTIONS_SOLVER:
            err = set_keyword(value, k_solver_atoi, NWORDS(k_solver_atoi), 1, '9', &solver->options[slot], "LIS_OPTIONS_SOLVER"); break;
        case LIS_OPTIONS_PSOLVER:
            err = set_keyword(value, k_solver_atoi, NWORDS(k_solver_atoi), 1, '9', &solver->options[slot], "LIS_OPTIONS_PSOLVER"); break;
        case LIS_OPTIONS_PRECON: err = set_precon(value, &solver->options[slot], "LIS_OPTIONS_PRECON"); break;
        case LIS_OPTIONS_PPRECON: err = set_precon(value, &solver->options[slot], "LIS_OPTIONS_PPRECON"); break;
        case LIS_OPTIONS_SCALE:
            err = set_keyword(value, k_scale_atoi, NWORDS(k_scale_atoi), 0, '2', &solver->options[slot], "LIS_OPTIONS_SCALE"); break;
        case LIS_OPTIONS_OUTPUT:
            err = set_keyword(value, k_print_atoi, NWORDS(k_print_atoi), 0, '3', &solver->options[slot], "LIS_OPTIONS_OUTPUT"); break;
        case LIS_OPTIONS_INITGUESS_ZEROS: case LIS_OPTIONS_ADDS: case LIS_OPTIONS_USE_AT: case LIS_OPTIONS_SAAMG_UNSYM:
            err = set_keyword(value, k_truefalse_atoi, NWORDS(k_truefalse_atoi), 0, '1', &solver->options[slot], "LIS_OPTIONS_TRUEFALSE");
            if (!err && slot == LIS_OPTIONS_SAAMG_UNSYM && solver->options[slot])
                solver->params[LIS_PARAMS_SAAMG_THETA - LIS_OPTIONS_LEN] = 0.12;
            break;
        case LIS_OPTIONS_PRECISION:
            err = set_keyword(value, k_precision_atoi, NWORDS(k_precision_atoi), 0, '1', &solver->options[slot], "LIS_OPTIONS_PRECISION"); break;
        case LIS_OPTIONS_STORAGE:
            err = set_keyword(value, k_storage_atoi, NWORDS(k_storage_atoi), 1, '9', &solver->options[slot], "LIS_OPTIONS_STORAGE"); break;
        case LIS_OPTIONS_CONV_COND:
            err = set_keyword(value, k_conv_cond_atoi, NWORDS(k_conv_cond_atoi), 0, '3', &solver->options[slot], "LIS_OPTIONS_CONV_COND"); break;
        default:
            if (slot < LIS_OPTIONS_LEN) { int v = solver->options[slot]; sscanf(value, "%d", &v); solver->options[slot] = v; }
            else { double dv = solver->params[slot - LIS_OPTIONS_LEN]; sscanf(value, "%lg", &dv); solver->params[slot - LIS_OPTIONS_LEN] = dv; }
            break;
        }
        if (err) { lis_solver_work_destroy(solver); solver->retcode = err; return err; }
    }
    return LIS_SUCCESS;
}

/* "-name value -name value ..." ; names and values are lower-cased (src/system/lis_init.c:248-316) */
LIS_INT lis_solver_set_option(char *text, LIS_SOLVER solver)
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
        err = set_option2(name, tok, solver);
        name = NULL;
        if (err) break;
    }
    free(buf);
    return err;
}

LIS_INT lis_solver_set_optionC(LIS_SOLVER solver)
{
    int count = 0;
    const lis_arg_t *args = lis_host_args(&count);
    char name[256];
    for (int i = 0; i < count; i++) {
        snprintf(name, sizeof(name), "-%s", args[i].name);
        LIS_INT err = set_option2(name, args[i].value, solver);
        if (err) return err;
    }
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ getters */
LIS_INT lis_solver_get_iter(LIS_SOLVER solver, LIS_INT *iter) { *iter = solver->iter; return LIS_SUCCESS; }
LIS_INT lis_solver_get_iterex(LIS_SOLVER solver, LIS_INT *iter, LIS_INT *iter_double, LIS_INT *iter_quad)
{
    *iter = solver->iter; *iter_double = solver->iter2; *iter_quad = solver->iter - solver->iter2;
    return LIS_SUCCESS;
}
LIS_INT lis_solver_get_time(LIS_SOLVER solver, double *time) { *time = solver->time; return LIS_SUCCESS; }
LIS_INT lis_solver_get_timeex(LIS_SOLVER solver, double *time, double *itime, double *ptime, double *p_c_time, double *p_i_time)
{
    *time = solver->time; *itime = solver->itime; *ptime = solver->ptime;
    *p_c_time = solver->p_c_time; *p_i_time = solver->p_i_time;
    return LIS_SUCCESS;
}
LIS_INT lis_solver_get_residualnorm(LIS_SOLVER solver, LIS_REAL *residual) { *residual = solver->resid; return LIS_SUCCESS; }
LIS_INT lis_solver_get_solver(LIS_SOLVER solver, LIS_INT *nsol) { *nsol = solver->options[LIS_OPTIONS_SOLVER]; return LIS_SUCCESS; }
LIS_INT lis_solver_get_precon(LIS_SOLVER solver, LIS_INT *precon_type) { *precon_type = solver->options[LIS_OPTIONS_PRECON]; return LIS_SUCCESS; }
LIS_INT lis_solver_get_status(LIS_SOLVER solver, LIS_INT *status) { *status = solver->retcode; return LIS_SUCCESS; }

LIS_INT lis_solver_get_rhistory(LIS_SOLVER solver, LIS_VECTOR v)
{
    LIS_INT maxiter = solver->iter + 1;
    if (solver->retcode != LIS_SUCCESS) maxiter--;
    if (solver->rhistory == NULL) { LIS_SETERR(LIS_FAILS, "residual history is empty\n"); return LIS_FAILS; }
    const LIS_INT n = _min(v->n, maxiter);
    if (n <= 0) return LIS_SUCCESS;
    return lis_vector_set_values2(LIS_INS_VALUE, v->is + v->origin, n, solver->rhistory, v);
}

LIS_INT lis_solver_get_solvername(LIS_INT solver, char *solvername)
{
    if (solver < 1 || solver > LIS_SOLVER_LEN) return LIS_FAILS;
    strcpy(solvername, k_solvername[solver]);
    return LIS_SUCCESS;
}

LIS_INT lis_solver_get_preconname(LIS_INT precon_type, char *preconname)
{
    if (precon_type < 0 || precon_type > LIS_PRECON_TYPE_LEN - 2) return LIS_FAILS;
    strcpy(preconname, k_preconname[precon_type]);
    return LIS_SUCCESS;
}

/* %e, one value per line: src/system/lis_output.c:570-640 */
LIS_INT lis_solver_output_rhistory(LIS_SOLVER solver, char *filename)
{
    LIS_INT maxiter = solver->iter + 1;
    if (solver->retcode != LIS_SUCCESS) maxiter--;
    if (solver->rhistory == NULL) { LIS_SETERR(LIS_FAILS, "residual history is empty\n"); return LIS_FAILS; }
    if (lisd_rank() != 0) return LIS_SUCCESS;
    FILE *f = fopen(filename, "w");
    if (f == NULL) { LIS_SETERR1(LIS_ERR_FILE_IO, "cannot open file %s\n", filename); return LIS_ERR_FILE_IO; }
    for (LIS_INT i = 0; i < maxiter; i++) fprintf(f, "%e\n", (double)solver->rhistory[i]);
    fclose(f);
    return LIS_SUCCESS;
}

void lis_host_print_rhistory(LIS_INT iter, LIS_REAL resid)
{
    lis_printf(LIS_COMM_WORLD, "iteration: %5d  relative residual = %E\n", (int)iter, (double)resid);
}

/* ------------------------------------------------------------------ shared solver pieces */
LIS_INT lis_host_solver_malloc_work(LIS_SOLVER solver, LIS_INT worklen, LIS_INT first)
{
    LIS_VECTOR *work = (LIS_VECTOR *)lis_calloc((size_t)worklen * sizeof(LIS_VECTOR), "lis_solver_malloc_work::work");
    if (work == NULL) { LIS_SETERR_MEM(worklen * sizeof(LIS_VECTOR)); return LIS_ERR_OUT_OF_MEMORY; }
    for (LIS_INT i = first; i < worklen; i++) {
        LIS_INT err = lis_vector_duplicate(solver->A, &work[i]);
        if (err) {
            for (LIS_INT j = first; j < i; j++) lis_vector_destroy(work[j]);
            lis_free(work);
            return err;
        }
    }
    solver->worklen = worklen;
    solver->work = work;
    return LIS_SUCCESS;
}

/* ||r|| * bnrm (nrm2_r, nrm2_b) or ||r||_1 (nrm1_b): src/solver/lis_solver.c:1792-1813 */
LIS_INT lis_host_solver_residual(LIS_SOLVER solver, LIS_VECTOR r, LIS_REAL *res)
{
    LIS_INT err;
    if (solver->options[LIS_OPTIONS_CONV_COND] == LIS_CONV_COND_NRM1_B) return lis_vector_nrm1(r, res);
    err = lis_vector_nrm2(r, res);
    *res = *res * solver->bnrm;
    return err;
}

/* rs0 = conj(r0) (default) or MT19937 uniforms: src/solver/lis_solver.c:1817-1875 */
LIS_INT lis_host_solver_shadow_residual(LIS_SOLVER solver, LIS_VECTOR r0, LIS_VECTOR rs0)
{
    if (solver->options[LIS_OPTIONS_INIT_SHADOW_RESID] == LIS_RANDOM) return lis_host_fill_mt19937(1, solver->A->n, &rs0);
    return lisd_copy(r0, rs0);
}

/* returns LIS_FAILS when the initial guess already satisfies the criterion (iter = 1) */
LIS_INT lis_solver_get_initial_residual(LIS_SOLVER solver, LIS_PRECON M, LIS_VECTOR t, LIS_VECTOR r, LIS_REAL *bnrm2)
{
    LIS_MATRIX A = solver->A;
    LIS_VECTOR b = solver->b, x = solver->x, p;
    const LIS_INT conv = solver->options[LIS_OPTIONS_CONV_COND];
    const LIS_REAL tol = solver->params[LIS_PARAMS_RESID - LIS_OPTIONS_LEN];
    const LIS_REAL tol_w = solver->params[LIS_PARAMS_RESID_WEIGHT - LIS_OPTIONS_LEN];
    const LIS_REAL tol_switch = solver->params[LIS_PARAMS_SWITCH_RESID - LIS_OPTIONS_LEN];
    LIS_REAL nrm2 = 0.0;
    LIS_INT err;

    p = (M == NULL) ? r : t;
    if (!solver->options[LIS_OPTIONS_INITGUESS_ZEROS]) {
        err = lisd_matvec(A, x, p);                    /* p = Ax    */
        if (!err) err = lisd_xpay(b, -1.0, p);         /* p = b - p */
    } else {
        err = lisd_copy(b, p);
    }
    if (err) return err;
    switch (conv) {
    case LIS_CONV_COND_NRM2_B:
        err = lis_vector_nrm2(p, &nrm2);
        if (!err) err = lis_vector_nrm2(b, bnrm2);
        solver->tol = tol; solver->tol_switch = tol_switch;
        break;
    case LIS_CONV_COND_NRM1_B:
        err = lis_vector_nrm1(p, &nrm2);
        if (!err) err = lis_vector_nrm1(b, bnrm2);
        solver->tol = *bnrm2 * tol_w + tol; solver->tol_switch = *bnrm2 * tol_w + tol_switch;
        break;
    default:
        err = lis_vector_nrm2(p, &nrm2);
        *bnrm2 = nrm2;
        solver->tol = tol; solver->tol_switch = tol_switch;
        break;
    }
    if (err) return err;
    if (*bnrm2 == 0.0) *bnrm2 = 1.0; else *bnrm2 = 1.0 / *bnrm2;
    solver->bnrm = *bnrm2;
    nrm2 = nrm2 * *bnrm2;
    if (nrm2 <= fabs(tol)) {
        solver->retcode = LIS_SUCCESS;
        solver->iter = 1;
        solver->resid = nrm2;
        return LIS_FAILS;
    }
    if (M != NULL) { err = lis_psolve(solver, p, r); if (err) return err; }
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ dispatch */
typedef LIS_INT (*lis_solver_fn)(LIS_SOLVER);

static LIS_INT work_cg(LIS_SOLVER s) { return lis_host_solver_malloc_work(s, 4, 0); }
static LIS_INT work_bicgstab(LIS_SOLVER s) { return lis_host_solver_malloc_work(s, 7, 0); }
static LIS_INT work_bicg(LIS_SOLVER s) { return lis_host_solver_malloc_work(s, 6, 0); }
static LIS_INT work_n(LIS_SOLVER s, LIS_INT n) { return lis_host_solver_malloc_work(s, n, 0); }
static LIS_INT work_3(LIS_SOLVER s) { return work_n(s, 3); }
static LIS_INT work_4(LIS_SOLVER s) { return work_n(s, 4); }
static LIS_INT work_6(LIS_SOLVER s) { return work_n(s, 6); }
static LIS_INT work_7(LIS_SOLVER s) { return work_n(s, 7); }
static LIS_INT work_10(LIS_SOLVER s) { return work_n(s, 10); }
static LIS_INT work_9(LIS_SOLVER s) { return work_n(s, 9); }
static LIS_INT work_12(LIS_SOLVER s) { return work_n(s, 12); }
static LIS_INT work_13(LIS_SOLVER s) { return work_n(s, 13); }
static LIS_INT work_14(LIS_SOLVER s) { return work_n(s, 14); }
/* work[0] of the reference is the (restart+1)-vector s of the least-squares problem; here
 * that short vector is a host array owned by lis_gmres, so slot 0 stays empty */
static LIS_INT work_gmres(LIS_SOLVER s) { return lis_host_solver_malloc_work(s, 4 + s->options[LIS_OPTIONS_RESTART] + 1, 1); }

static LIS_INT work_orthomin(LIS_SOLVER s) { return work_n(s, 3 + 3 * (s->options[LIS_OPTIONS_RESTART] + 1)); }
/* FGMRES: slot 0 (the short vector s of the reference) stays empty, like GMRES */
static LIS_INT work_fgmres(LIS_SOLVER s) { return lis_host_solver_malloc_work(s, 4 + 2 * s->options[LIS_OPTIONS_RESTART] + 1, 1); }
static LIS_INT work_bicgstabl(LIS_SOLVER s) { return work_n(s, 4 + 2 * (s->options[LIS_OPTIONS_ELL] + 1)); }
static LIS_INT check_bicgstabl(LIS_SOLVER s)
{
    const LIS_INT ell = s->options[LIS_OPTIONS_ELL];
    if (ell < 1) {
        LIS_SETERR1(LIS_ERR_ILL_ARG, "Parameter LIS_OPTIONS_ELL(=%D) is less than 1\n", ell);
        return LIS_ERR_ILL_ARG;
    }
    return LIS_SUCCESS;
}
static LIS_INT work_idrs(LIS_SOLVER s) { return work_n(s, 4 + 3 * s->options[LIS_OPTIONS_IDRS_RESTART]); }
static LIS_INT work_idr1(LIS_SOLVER s) { return work_n(s, 4 + 3 * (s->options[LIS_OPTIONS_IDRS_RESTART] > 1 ? s->options[LIS_OPTIONS_IDRS_RESTART] : 1)); }
static LIS_INT check_idrs(LIS_SOLVER s)
{
    if (s->options[LIS_OPTIONS_IDRS_RESTART] < 1) {
        LIS_SETERR1(LIS_ERR_ILL_ARG, "Parameter LIS_OPTIONS_IDRS_RESTART(=%D) is less than 1\n", s->options[LIS_OPTIONS_IDRS_RESTART]);
        return LIS_ERR_ILL_ARG;
    }
    return LIS_SUCCESS;
}
static LIS_INT check_none(LIS_SOLVER s) { (void)s; return LIS_SUCCESS; }
static LIS_INT check_gmres(LIS_SOLVER s)
{
    const LIS_INT restart = s->options[LIS_OPTIONS_RESTART];
    if (restart < 0) {
        LIS_SETERR1(LIS_ERR_ILL_ARG, "Parameter LIS_OPTIONS_RESTART(=%D) is less than 0\n", restart);
        return LIS_ERR_ILL_ARG;
    }
    return LIS_SUCCESS;
}

typedef struct { lis_solver_fn check, work, run; int conv_cond_ok; } lis_solver_entry;
static lis_solver_entry solver_entry(LIS_INT nsolver)
{
    lis_solver_entry e = {NULL, NULL, NULL, 0};
    switch (nsolver) {
    case LIS_SOLVER_CG: e.check = check_none; e.work = work_cg; e.run = lis_cg; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_BICG: e.check = check_none; e.work = work_bicg; e.run = lis_bicg; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_BICGSTAB: e.check = check_none; e.work = work_bicgstab; e.run = lis_bicgstab; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_CGS: e.check = check_none; e.work = work_7; e.run = lis_cgs; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_CRS: e.check = check_none; e.work = work_6; e.run = lis_crs; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_CR: e.check = check_none; e.work = work_6; e.run = lis_cr; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_COCR: e.check = check_none; e.work = work_6; e.run = lis_cocr; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_COCG: e.check = check_none; e.work = work_4; e.run = lis_cocg; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_BICR: e.check = check_none; e.work = work_10; e.run = lis_bicr; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_BICRSTAB: e.check = check_none; e.work = work_9; e.run = lis_bicrstab; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_TFQMR: e.check = check_none; e.work = work_9; e.run = lis_tfqmr; e.conv_cond_ok = 0; break;
    case LIS_SOLVER_GPBICG: e.check = check_none; e.work = work_14; e.run = lis_gpbicg; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_GPBICR: e.check = check_none; e.work = work_14; e.run = lis_gpbicr; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_BICGSAFE: e.check = check_none; e.work = work_12; e.run = lis_bicgsafe; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_BICRSAFE: e.check = check_none; e.work = work_13; e.run = lis_bicrsafe; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_ORTHOMIN: e.check = check_gmres; e.work = work_orthomin; e.run = lis_orthomin; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_MINRES: e.check = check_none; e.work = work_7; e.run = lis_minres; e.conv_cond_ok = 0; break;
    case LIS_SOLVER_FGMRES: e.check = check_gmres; e.work = work_fgmres; e.run = lis_fgmres; e.conv_cond_ok = 0; break;
    case LIS_SOLVER_BICGSTABL: e.check = check_bicgstabl; e.work = work_bicgstabl; e.run = lis_bicgstabl; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_JACOBI: e.check = check_none; e.work = work_4; e.run = lis_jacobi; e.conv_cond_ok = 0; break;
    case LIS_SOLVER_GS: e.check = check_none; e.work = work_3; e.run = lis_gs; e.conv_cond_ok = 0; break;
    case LIS_SOLVER_SOR: e.check = check_none; e.work = work_3; e.run = lis_sor; e.conv_cond_ok = 0; break;
    case LIS_SOLVER_IDRS: e.check = check_idrs; e.work = work_idrs; e.run = lis_idrs; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_IDR1: e.check = check_none; e.work = work_idr1; e.run = lis_idr1; e.conv_cond_ok = 1; break;
    case LIS_SOLVER_GMRES: e.check = check_gmres; e.work = work_gmres; e.run = lis_gmres; e.conv_cond_ok = 0; break;
    default: break;
    }
    return e;
}

/* work-vector allocator and iteration loop of solver `nsolver` (the reference's lis_solver_malloc_work[] /
 * lis_solver_execute[] tables, which its hybrid preconditioner indexes directly) */
LIS_INT lis_host_solver_entry(LIS_INT nsolver, LIS_INT (**work)(LIS_SOLVER), LIS_INT (**run)(LIS_SOLVER))
{
    const lis_solver_entry e = solver_entry(nsolver);
    if (e.run == NULL) { LIS_SETERR_IMP; return LIS_ERR_NOT_IMPLEMENTED; }
    *work = e.work; *run = e.run;
    return LIS_SUCCESS;
}

LIS_INT lis_solve(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, LIS_SOLVER solver)
{
    LIS_INT err;
    LIS_PRECON precon;
    solver->A = A;
    if (solver->options[LIS_OPTIONS_PRECON] < 0 || solver->options[LIS_OPTIONS_PRECON] >= lis_host_precon_type_end()) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "Parameter LIS_OPTIONS_PRECON is %D (Set between 0 to %D)\n",
                    solver->options[LIS_OPTIONS_PRECON], lis_host_precon_type_end() - 1);
        return LIS_ERR_ILL_ARG;
    }
    err = lis_precon_create(solver, &precon);
    if (err) { lis_solver_work_destroy(solver); solver->retcode = err; return err; }
    err = lis_solve_kernel(A, b, x, solver, precon);
    if (err) {
        lis_solver_work_destroy(solver);
        lis_precon_destroy(precon);
        solver->precon = NULL;
        solver->retcode = err;
        return err;
    }
    lis_precon_destroy(precon);
    solver->precon = NULL;
    return LIS_SUCCESS;
}

LIS_INT lis_solver_set_matrix(LIS_MATRIX A, LIS_SOLVER solver) { solver->A = A; return LIS_SUCCESS; }

/* run lis_solve up to, but not including, the solver loop: option checks, -storage conversion, work
 * vectors -- what the CR eigensolver needs before it borrows the preconditioner
 * (src/solver/lis_solver.c:408-437) */
LIS_INT lis_solve_setup(LIS_MATRIX A, LIS_SOLVER solver)
{
    LIS_VECTOR b, x;
    LIS_INT err = lis_vector_duplicate(A, &b);
    if (err) return err;
    err = lis_vector_duplicate(A, &x);
    if (err) { lis_vector_destroy(b); return err; }
    solver->setup = LIS_TRUE;
    err = lis_solve(A, b, x, solver);
    if (err) { lis_solver_work_destroy(solver); solver->retcode = err; }
    lis_vector_destroy(b);
    lis_vector_destroy(x);
    return err;
}

LIS_INT lis_solve_kernel(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, LIS_SOLVER solver, LIS_PRECON precon)
{
    const LIS_Comm comm = LIS_COMM_WORLD;
    const LIS_INT nsolver = solver->options[LIS_OPTIONS_SOLVER];
    const LIS_INT precon_type = solver->options[LIS_OPTIONS_PRECON];
    const LIS_INT maxiter = solver->options[LIS_OPTIONS_MAXITER];
    const LIS_INT output = solver->options[LIS_OPTIONS_OUTPUT];
    LIS_INT scale = solver->options[LIS_OPTIONS_SCALE];
    const LIS_INT precision = solver->options[LIS_OPTIONS_PRECISION];
    const LIS_INT conv_cond = solver->options[LIS_OPTIONS_CONV_COND];
    const LIS_REAL tol = solver->params[LIS_PARAMS_RESID - LIS_OPTIONS_LEN];
    const LIS_REAL tol_w = solver->params[LIS_PARAMS_RESID_WEIGHT - LIS_OPTIONS_LEN];
    LIS_INT err;
    LIS_REAL *rhistory, nrm2;
    LIS_VECTOR xx, t;
    double p_c_time, p_i_time, itime;
    char buf[64];

    solver->precision = precision;
    if (nsolver < 1 || nsolver > LIS_SOLVER_LEN) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "Parameter LIS_OPTIONS_SOLVER is %D (Set between 1 to %D)\n", nsolver, LIS_SOLVER_LEN);
        return LIS_ERR_ILL_ARG;
    }
    lis_solver_entry entry = solver_entry(nsolver);
    if (entry.run == NULL) {
        LIS_SETERR1(LIS_ERR_NOT_IMPLEMENTED, "solver %s has no driver in this library\n", k_solvername[nsolver]);
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    if (precon_type < 0 || precon_type >= lis_host_precon_type_end()) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "Parameter LIS_OPTIONS_PRECON is %D (Set between 0 to %D)\n", precon_type, lis_host_precon_type_end() - 1);
        return LIS_ERR_ILL_ARG;
    }
    if (maxiter < 0) {
        LIS_SETERR1(LIS_ERR_ILL_ARG, "Parameter LIS_OPTIONS_MAXITER(=%D) is less than 0\n", maxiter);
        return LIS_ERR_ILL_ARG;
    }
    if (conv_cond > 0 && !entry.conv_cond_ok) {
        LIS_SETERR1(LIS_ERR_ILL_ARG, "Option conv_cond is not implemented for solver %s\n", k_solvername[nsolver]);
        return LIS_ERR_ILL_ARG;
    }
    if (precision != LIS_PRECISION_DOUBLE) {
        LIS_SETERR(LIS_ERR_ILL_ARG, "Quad precision is not enabled\n");
        return LIS_ERR_ILL_ARG;
    }
    if (solver->options[LIS_OPTIONS_USE_AT]) {
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "-use_at is outside the B200 hot path\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    if (scale != LIS_SCALE_NONE && solver->options[LIS_OPTIONS_STORAGE] == LIS_MATRIX_BSR && scale == LIS_SCALE_JACOBI) {
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "block Jacobi scaling of BSR matrices is outside the B200 hot path\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    err = entry.check(solver);
    if (err) { solver->retcode = err; return err; }

    solver->A = A;
    solver->b = b;

    /* initial vector */
    err = lis_vector_duplicate(A, &xx);
    if (err) { solver->retcode = err; return err; }
    if (solver->options[LIS_OPTIONS_INITGUESS_ZEROS]) {
        if (output) lis_printf(comm, "initial vector x      : all components set to 0\n");
        err = lisd_set_all(0.0, xx);
    } else {
        if (output) lis_printf(comm, "initial vector x      : user defined\n");
        err = lisd_copy(x, xx);
    }
    if (err) { lis_vector_destroy(xx); solver->retcode = err; return err; }

    /* residual history */
    if (solver->rhistory) lis_free(solver->rhistory);
    solver->rhistory = NULL;
    /* maxiter+2 entries like the reference, plus slack: BiCGSTAB(l) counts l steps per sweep and can
     * record up to iteration maxiter+l (the reference writes past its array there) */
    rhistory = (LIS_REAL *)lis_malloc(((size_t)maxiter + 2 + 64) * sizeof(LIS_REAL), "lis_solve::rhistory");
    if (rhistory == NULL) {
        LIS_SETERR_MEM((maxiter + 2) * sizeof(LIS_SCALAR));
        lis_vector_destroy(xx);
        solver->retcode = LIS_ERR_OUT_OF_MEMORY;
        return LIS_ERR_OUT_OF_MEMORY;
    }
    rhistory[0] = 1.0;

    p_c_time = lis_wtime();                    /* "matrix creation": the scaling block, as there (:612-743) */

    /* system scaling (src/solver/lis_solver.c:636-721): the stationary solvers with a preconditioner
     * always work on D^-1 A; -scale jacobi|symm_diag on request (CG turns jacobi into symm_diag to keep
     * the matrix symmetric).  A and b stay scaled afterwards, exactly as there. */
    const LIS_INT was_scaled = A->is_scaled;
    LIS_INT scaled_with = LIS_SCALE_JACOBI;
    if (precon_type == LIS_PRECON_TYPE_IS) {
        /* I+S works on the unit-diagonal system D^-1 A (:613-641) */
        if (solver->d == NULL) err = lis_vector_duplicate(A, &solver->d);
        if (!err && !A->is_scaled) err = lis_matrix_scale(A, b, solver->d, LIS_SCALE_JACOBI);
        else if (!err && !b->is_scaled) err = lis_vector_pmul(b, solver->d, b);
    } else if (nsolver >= LIS_SOLVER_JACOBI && nsolver <= LIS_SOLVER_SOR && precon_type != LIS_PRECON_TYPE_NONE) {
        if (solver->d == NULL) err = lis_vector_duplicate(A, &solver->d);
        if (!err && !A->is_scaled) err = lis_matrix_scale(A, b, solver->d, LIS_SCALE_JACOBI);
    } else if (scale) {
        if (solver->d == NULL) err = lis_vector_duplicate(A, &solver->d);
        if (scale == LIS_SCALE_JACOBI && nsolver == LIS_SOLVER_CG) scale = LIS_SCALE_SYMM_DIAG;
        scaled_with = scale;
        if (!err && !A->is_scaled) err = lis_matrix_scale(A, b, solver->d, scale);
        else if (!err && !b->is_scaled) err = lis_vector_pmul(b, solver->d, b);
    }
    /* SSOR / GS / SOR with -storage <scalar format> sweep on a private split CSR copy of the matrix made when the
     * preconditioner was created, i.e. before this scaling: the reference's split matrix IS the scaled one (with WD still
     * from the unscaled diagonal, as there), so the copy gets the same factors */
    if (!err && !was_scaled && A->is_scaled && precon && precon->is_copy && precon->A && precon->A != A)
        err = lis_host_matrix_scale_like(precon->A, solver->d, scaled_with);
    if (err) { lis_vector_destroy(xx); lis_free(rhistory); solver->retcode = err; return err; }
    p_c_time = lis_wtime() - p_c_time;
    itime = lis_wtime();

    /* -storage: converts A in place */
    err = lis_matrix_convert_self(solver);
    if (err) { lis_vector_destroy(xx); lis_free(rhistory); solver->retcode = err; return err; }

    if (output && A->my_rank == 0) {
        printf("precision             : %s\n", k_precision_atoi[precision]);
        printf("linear solver         : %s\n", k_solvername[nsolver]);
        /* src/solver/lis_solver.c:770-800: ILU carries its fill level ("Block" on block storage), -adds a suffix */
        if (precon_type == LIS_PRECON_TYPE_ILU)
            snprintf(buf, sizeof(buf), "%s%s(%d)", (A->matrix_type == LIS_MATRIX_BSR || A->matrix_type == LIS_MATRIX_VBR) ? "Block " : "",
                     k_preconname[precon_type], (int)solver->options[LIS_OPTIONS_FILL]);
        else if (precon_type < LIS_PRECON_TYPE_LEN - 1) snprintf(buf, sizeof(buf), "%s", k_preconname[precon_type]);
        else snprintf(buf, sizeof(buf), "user defined");
        if (solver->options[LIS_OPTIONS_ADDS] && precon_type) printf("preconditioner        : %s + Additive Schwarz\n", buf);
        else printf("preconditioner        : %s\n", buf);
    }
    switch (conv_cond) {
    case LIS_CONV_COND_NRM2_R:
        if (output) lis_printf(comm, "convergence condition : ||b-Ax||_2 <= %6.1e * ||b-Ax_0||_2\n", (double)tol);
        break;
    case LIS_CONV_COND_NRM2_B:
        lis_vector_nrm2(b, &nrm2);
        nrm2 = nrm2 * tol;
        if (output) lis_printf(comm, "convergence condition : ||b-Ax||_2 <= %6.1e*||b||_2 = %6.1e\n", (double)tol, (double)nrm2);
        break;
    case LIS_CONV_COND_NRM1_B:
        lis_vector_nrm1(b, &nrm2);
        nrm2 = nrm2 * tol_w + tol;
        if (output) lis_printf(comm, "convergence condition : ||b-Ax||_1 <= %6.1e*||b||_1 + %6.1e = %6.1e\n", (double)tol_w, (double)tol, (double)nrm2);
        break;
    }
    if (output) {
        if (A->matrix_type == LIS_MATRIX_BSR)
            lis_printf(comm, "matrix storage format : %s(%D x %D)\n", k_storagename[A->matrix_type - 1], A->bnr, A->bnr);
        else
            lis_printf(comm, "matrix storage format : %s\n", k_storagename[A->matrix_type - 1]);
    }

    /* work vectors */
    err = entry.work(solver);
    if (err) { lis_vector_destroy(xx); lis_free(rhistory); solver->retcode = err; return err; }

    solver->x = xx;
    solver->xx = x;
    solver->precon = precon;
    solver->rhistory = rhistory;
    solver->ptime = 0.0;

    if (!solver->setup) {
        err = entry.run(solver);
        solver->retcode = err;
        if (err == LIS_ERR_DEVICE || err == LIS_ERR_OUT_OF_MEMORY || err == LIS_ERR_NOT_IMPLEMENTED || err == LIS_ERR_ILL_ARG) {
            lis_solver_work_destroy(solver);
            lis_vector_destroy(xx);
            solver->x = NULL;
            return err;
        }
    }
    {
        /* symmetric scaling solved for D^1/2 x: x = xx .* d -- not under I+S, whose unit-diagonal scaling is a left
         * scaling only, whatever -scale says (src/solver/lis_solver.c:877-886) */
        LIS_INT e2 = (scale == LIS_SCALE_SYMM_DIAG && precon_type != LIS_PRECON_TYPE_IS && solver->d) ? lisd_pmul(xx, solver->d, x)
                                                                                                       : lisd_copy(xx, x);
        if (!e2) e2 = lisd_sync();
        if (e2) { lis_solver_work_destroy(solver); lis_vector_destroy(xx); return e2; }
    }
    itime = lis_wtime() - itime - solver->ptime;
    p_i_time = solver->ptime;
    solver->ptime = p_c_time + p_i_time;
    solver->p_c_time = p_c_time;
    solver->p_i_time = p_i_time;
    solver->time = solver->ptime + itime;
    solver->itime = itime;
    lis_solver_work_destroy(solver);

    /* true residual b - A*xx: computed and discarded, like the reference (:910-924) */
    if (lis_vector_duplicate(A, &t) == LIS_SUCCESS) {
        if (lisd_matvec(A, xx, t) == LIS_SUCCESS && lisd_xpay(b, -1.0, t) == LIS_SUCCESS) lis_vector_nrm2(t, &nrm2);
        lis_vector_destroy(t);
    }

    if (output) {
        if (err) lis_printf(comm, "linear solver status  : %s(code=%D)\n\n", k_returncode[err < 8 ? err : 1], err);
        else lis_printf(comm, "linear solver status  : normal end\n\n");
    }
    solver->iter2 = solver->iter;
    lis_vector_destroy(xx);
    solver->x = NULL;
    return LIS_SUCCESS;
}
