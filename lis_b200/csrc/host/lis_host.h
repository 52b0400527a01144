/* lis_host.h -- private host-side helpers shared by the lis_b200 host sources. */
#ifndef LIS_B200_HOST_H
#define LIS_B200_HOST_H
#include "lislib.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { char *name; char *value; } lis_arg_t;   /* "-name value" from argv, lower-cased */
const lis_arg_t *lis_host_args(int *count);

void lisd_comm_finalize(void);

/* object header initialisation / copy (label, sizes, partition) */
void lis_host_header_copy(const void *src, void *dst);

/* CSR-level host helpers used by conversion, split and the SSOR schedule */
LIS_INT lis_host_csr_from(LIS_MATRIX A, LIS_INT **ptr, LIS_INT **index, LIS_SCALAR **value, LIS_INT *owned);

LIS_INT lis_host_matrix_check_input(LIS_MATRIX A);
LIS_INT lis_host_matrix_check_set(LIS_MATRIX A);
/* MSR / COO / BSC / VBR / DNS (lis_formats_ext.c) */
LIS_INT lis_host_ext_from_csr(LIS_MATRIX Acsr, LIS_MATRIX Aout, int *handled);
LIS_INT lis_host_ext_to_csr(LIS_MATRIX Ain, LIS_MATRIX Aout, int *handled);
LIS_INT lis_host_ext_shift_diagonal(LIS_MATRIX A, LIS_SCALAR sigma, int *handled);
LIS_INT lis_host_vbr_partition(LIS_MATRIX Ain, LIS_INT *nblk, LIS_INT **row, LIS_INT **col);
LIS_INT lis_host_ext_get_diagonal(LIS_MATRIX A, LIS_SCALAR *d, int *handled);
LIS_INT lis_host_transposed_rows(LIS_MATRIX A, LIS_INT **ptr, LIS_INT **index, LIS_SCALAR **value);
LIS_INT lis_host_ordered_rows(LIS_MATRIX A, int keep_zeros, LIS_INT *nnz, LIS_INT **ptr, LIS_INT **index, LIS_SCALAR **value);
void    lis_host_matrix_adopt(LIS_MATRIX dst, LIS_MATRIX src);
LIS_INT lis_host_diag_create(LIS_MATRIX A, LIS_MATRIX_DIAG *Dout);
LIS_INT lis_host_transpose(LIS_INT n, LIS_INT ncols, const LIS_INT *ptr, const LIS_INT *index, const LIS_SCALAR *value,
                           LIS_INT **optr, LIS_INT **oindex, LIS_SCALAR **ovalue);

/* emulated OpenMP thread count of the reference (-omp_num_threads N): the SSOR block count */
int  lis_host_num_threads(void);
void lis_host_set_num_threads(int n);
int  lis_host_worker_count(void);                 /* real host threads for set-up passes (LIS_B200_HOST_THREADS) */
void lis_host_parallel_for(size_t count, size_t grain, void (*fn)(size_t lo, size_t hi, void *ctx), void *ctx);

/* preconditioner registry / solver helpers */
LIS_INT lis_host_precon_type_end(void);
LIS_INT lis_host_precon_lookup(const char *name);
void    lis_host_print_rhistory(LIS_INT iter, LIS_REAL resid);
LIS_INT lis_host_solver_entry(LIS_INT nsolver, LIS_INT (**work)(LIS_SOLVER), LIS_INT (**run)(LIS_SOLVER));
LIS_INT lis_host_solver_malloc_work(LIS_SOLVER solver, LIS_INT worklen, LIS_INT first);
LIS_INT lis_host_solver_residual(LIS_SOLVER solver, LIS_VECTOR r, LIS_REAL *res);
LIS_INT lis_host_mgs(LIS_VECTOR *v, LIS_INT i, LIS_SCALAR *hcol, LIS_REAL *nrm);       /* modified Gram-Schmidt of v[i], lis_krylov.c */
LIS_INT lis_host_ilu_create(LIS_SOLVER solver, LIS_PRECON precon);      /* lis_precon_ilu.c */
void    lis_host_ilu_free(void *factors);
LIS_INT lis_psolve_iluk(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);
LIS_INT lis_psolveh_iluk(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);
LIS_INT lis_host_fill_mt19937(LIS_INT s, LIS_INT n, LIS_VECTOR *P);
LIS_INT lis_host_solver_shadow_residual(LIS_SOLVER solver, LIS_VECTOR r0, LIS_VECTOR rs0);

LIS_INT lis_host_set_wd(LIS_MATRIX A, LIS_SCALAR scale, int do_scale, LIS_INT tag);

/* vectors */
LIS_INT lis_vector_check_same(LIS_VECTOR x, LIS_VECTOR y);

#ifdef __cplusplus
}
#endif
LIS_INT lis_host_matrix_scale_like(LIS_MATRIX C, LIS_VECTOR Dv, LIS_INT action);   /* lis_matrix.c */
#endif
