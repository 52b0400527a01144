/*
 * lislib.h -- the "friend" interface of the library: what the reference's own drivers reach
 * for beyond lis.h (test/spmvtest1.c:44 includes lislib.h and calls lis_sort_id; the solver
 * layer calls the per-format kernels and lis_psolve directly).  Reference: include/lislib.h,
 * lis_matvec.h:76-205, lis_system.h:33-111, lis_precon.h:32, lis_solver.h.
 */
#ifndef LIS_B200_LISLIB_H
#define LIS_B200_LISLIB_H

#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "lis.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error reporting (reference: include/lis_system.h:33-40) ---- */
#ifndef __FUNC__
#define __FUNC__ "unknown"
#endif
LIS_INT lis_error(const char *file, const char *func, const LIS_INT line, const LIS_INT code, const char *mess, ...);
#define LIS_SETERR(code, mess)              lis_error(__FILE__, __func__, __LINE__, code, mess)
#define LIS_SETERR1(code, mess, a1)         lis_error(__FILE__, __func__, __LINE__, code, mess, a1)
#define LIS_SETERR2(code, mess, a1, a2)     lis_error(__FILE__, __func__, __LINE__, code, mess, a1, a2)
#define LIS_SETERR3(code, mess, a1, a2, a3) lis_error(__FILE__, __func__, __LINE__, code, mess, a1, a2, a3)
#define LIS_SETERR_MEM(sz)                  lis_error(__FILE__, __func__, __LINE__, LIS_ERR_OUT_OF_MEMORY, "malloc size = %d\n", (int)(sz))
#define LIS_SETERR_IMP                      lis_error(__FILE__, __func__, __LINE__, LIS_ERR_NOT_IMPLEMENTED, "not implemented\n")
#define LIS_SETERR_FIO                      lis_error(__FILE__, __func__, __LINE__, LIS_ERR_FILE_IO, "file i/o error\n")

/* ---- sorting helpers (reference: src/system/lis_sort.c) ---- */
void lis_sort_i(LIS_INT is, LIS_INT ie, LIS_INT *i1);
void lis_sort_id(LIS_INT is, LIS_INT ie, LIS_INT *i1, LIS_SCALAR *d1);
void lis_sort_ii(LIS_INT is, LIS_INT ie, LIS_INT *i1, LIS_INT *i2);
void lis_sortr_ii(LIS_INT is, LIS_INT ie, LIS_INT *i1, LIS_INT *i2);

/* ---- row partition (reference: src/system/lis_init.c:401-472) ---- */
LIS_INT lis_ranges_create(LIS_Comm comm, LIS_INT *local_n, LIS_INT *global_n, LIS_INT **ranges,
                          LIS_INT *is, LIS_INT *ie, LIS_INT *nprocs, LIS_INT *my_rank);

/* ---- per-format SpMV seam: raw x[], y[] like the reference (include/lis_matvec.h:76-205).
 * x and y must be device-accessible (vector storage of this library is). ---- */
typedef void (*LIS_MATVEC_FUNC)(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_csr(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_csc(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_ell(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_dia(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_jad(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_bsr(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_msr(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_coo(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_bsc(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_vbr(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
void lis_matvec_dns(LIS_MATRIX A, LIS_SCALAR x[], LIS_SCALAR y[]);
extern LIS_MATVEC_FUNC LIS_MATVEC;

/* halo exchange before a row-partitioned SpMV (reference: src/matrix/lis_matrix_mpi.c:834) */
LIS_INT lis_send_recv(LIS_COMMTABLE commtable, LIS_SCALAR x[]);
LIS_INT lis_reduce(LIS_COMMTABLE commtable, LIS_SCALAR x[]);      /* the reverse step: x[n..np) back to the owners, added (lis_matrix_mpi.c:958) */

/* ---- matrix internals the solver layer uses ---- */
LIS_INT lis_matrix_split(LIS_MATRIX A);
LIS_INT lis_matrix_merge(LIS_MATRIX A);
LIS_INT lis_matrix_sort_csr(LIS_MATRIX A);
LIS_INT lis_matrix_solve(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, LIS_INT flag);
LIS_INT lis_matrix_solveh(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, LIS_INT flag);   /* SSOR only */
LIS_INT lis_matrix_convert_self(LIS_SOLVER solver);
LIS_INT lis_matrix_storage_destroy(LIS_MATRIX A);
LIS_INT lis_matrix_DLU_destroy(LIS_MATRIX A);
LIS_INT lis_matrix_diag_destroy(LIS_MATRIX_DIAG D);

/* ---- preconditioner dispatch (reference: include/lis_precon.h:32, src/precon/lis_precon.c) ---- */
LIS_INT lis_precon_create(LIS_SOLVER solver, LIS_PRECON *precon);
LIS_INT lis_precon_destroy(LIS_PRECON precon);
LIS_INT lis_psolve(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);
LIS_INT lis_psolveh(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);
LIS_INT lis_psolve_none(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);
LIS_INT lis_psolve_jacobi(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);
LIS_INT lis_psolve_ssor(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);

/* ---- Krylov drivers (reference: src/solver/lis_solver_{cg,bicg,bicgstab,gmres}.c) ---- */
LIS_INT lis_cg(LIS_SOLVER solver);
LIS_INT lis_bicg(LIS_SOLVER solver);
LIS_INT lis_bicgstab(LIS_SOLVER solver);
LIS_INT lis_gmres(LIS_SOLVER solver);
LIS_INT lis_cgs(LIS_SOLVER solver);
LIS_INT lis_crs(LIS_SOLVER solver);
LIS_INT lis_cr(LIS_SOLVER solver);
LIS_INT lis_cocg(LIS_SOLVER solver);
LIS_INT lis_cocr(LIS_SOLVER solver);
LIS_INT lis_bicr(LIS_SOLVER solver);
LIS_INT lis_bicrstab(LIS_SOLVER solver);
LIS_INT lis_idrs(LIS_SOLVER solver);
LIS_INT lis_idr1(LIS_SOLVER solver);
LIS_INT lis_jacobi(LIS_SOLVER solver);
LIS_INT lis_gs(LIS_SOLVER solver);
LIS_INT lis_sor(LIS_SOLVER solver);
LIS_INT lis_bicgstabl(LIS_SOLVER solver);
LIS_INT lis_orthomin(LIS_SOLVER solver);
LIS_INT lis_minres(LIS_SOLVER solver);
LIS_INT lis_fgmres(LIS_SOLVER solver);
LIS_INT lis_tfqmr(LIS_SOLVER solver);
LIS_INT lis_gpbicg(LIS_SOLVER solver);
LIS_INT lis_gpbicr(LIS_SOLVER solver);
LIS_INT lis_bicgsafe(LIS_SOLVER solver);
LIS_INT lis_bicrsafe(LIS_SOLVER solver);
LIS_INT lis_solver_get_initial_residual(LIS_SOLVER solver, LIS_PRECON M, LIS_VECTOR t, LIS_VECTOR r, LIS_REAL *bnrm2);
LIS_INT lis_solver_work_destroy(LIS_SOLVER solver);
LIS_INT lis_matvech(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y);

#ifdef __cplusplus
}
#endif
#endif
