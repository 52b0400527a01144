/*
 * lis_oracle.h -- CPU restatement of the Lis 2.1.11 SpMV/Krylov hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (lis_b200/, include/) may include, link
 * or call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function below bit
 * for bit against the real reference compiled from /root/reference (oracle/_ref/, built by
 * oracle/Makefile) and against the committed fixtures in tests/golden/ generated from it.
 *
 * Every function takes plain arrays.  `nthreads` emulates, serially and deterministically,
 * what the reference's OpenMP build does with omp_get_max_threads()==nthreads (chunked
 * reduction order, thread-blocked DIA/JAD layouts, block-SSOR); nthreads==1 is the serial
 * build.  Citations are file:line in the reference tree.
 */
#ifndef LIS_ORACLE_H
#define LIS_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* row partition of the reference: LIS_GET_ISIE, include/lis.h:1067-1078 */
void orc_get_isie(int id, int nprocs, int n, int *is, int *ie);

/* ---- SpMV ---- */
void orc_spmv_csr(int n, const int *ptr, const int *idx, const double *val, const double *x, double *y);
void orc_spmv_csr_split(int n, const double *diag, const int *lptr, const int *lidx, const double *lval,
                        const int *uptr, const int *uidx, const double *uval, const double *x, double *y);
void orc_spmv_ell(int n, int maxnzr, const int *idx, const double *val, const double *x, double *y);
void orc_spmv_dia(int n, int nnd, const int *off, const double *val, const double *x, double *y, int nthreads);
void orc_spmv_jad(int n, int maxnzr, const int *jptr, const int *perm, const int *idx, const double *val,
                  const double *x, double *y, int nthreads);
void orc_spmv_bsr(int n, int nr, int bnr, int bnc, const int *bptr, const int *bidx, const double *val,
                  const double *x, double *y);
void orc_spmv_csc(int n, const int *ptr, const int *idx, const double *val, const double *x, double *y);

/* ---- format builders (two-pass: size query, then fill caller-allocated arrays) ---- */
void orc_sort_csr_rows(int n, const int *ptr, int *idx, double *val);   /* lis_matrix_sort_csr */
int  orc_csr2ell_maxnzr(int n, const int *ptr);
void orc_csr2ell(int n, const int *ptr, const int *idx, const double *val, int maxnzr, int *eidx, double *eval);
int  orc_csr2dia_nnd(int n, const int *ptr, const int *idx);            /* rows must be sorted */
void orc_csr2dia(int n, const int *ptr, const int *idx, const double *val, int nnd, int *off, double *dval, int nthreads);
int  orc_csr2jad_maxnzr(int n, const int *ptr);
void orc_csr2jad(int n, const int *ptr, const int *idx, const double *val, int maxnzr,
                 int *perm, int *jptr, int *jidx, double *jval, int nthreads);
int  orc_csr2bsr_bnnz(int n, const int *ptr, const int *idx, int bnr, int bnc, int *bptr);
void orc_csr2bsr(int n, const int *ptr, const int *idx, const double *val, int bnr, int bnc,
                 const int *bptr, int *bidx, double *bval);
void orc_csr2csc(int n, const int *ptr, const int *idx, const double *val, int *cptr, int *cidx, double *cval);

/* ---- BLAS-1 ---- */
double orc_dot(int n, const double *x, const double *y, int nthreads);
double orc_nrm2(int n, const double *x, int nthreads);
double orc_nrm1(int n, const double *x, int nthreads);
double orc_nrmi(int n, const double *x);
double orc_sum(int n, const double *x, int nthreads);
void orc_axpy(int n, double alpha, const double *x, double *y);
void orc_xpay(int n, const double *x, double alpha, double *y);
void orc_axpyz(int n, double alpha, const double *x, const double *y, double *z);
void orc_scale(int n, double alpha, double *x);
void orc_pmul(int n, const double *x, const double *y, double *z);
void orc_pdiv(int n, const double *x, const double *y, double *z);
void orc_reciprocal(int n, double *x);
void orc_shift(int n, double sigma, double *x);
void orc_abs(int n, double *x);

/* ---- preconditioner pieces ---- */
void orc_csr_get_diagonal(int n, const int *ptr, const int *idx, const double *val, double *d);
void orc_csr_split_count(int n, const int *ptr, const int *idx, int *nnzl, int *nnzu);
void orc_csr_split(int n, const int *ptr, const int *idx, const double *val,
                   int *lptr, int *lidx, double *lval, int *uptr, int *uidx, double *uval, double *diag);
/* x = M^-1 b, M = (D/w + L)(I + w D^-1 U); wd = 1/(w*D) */
void orc_ssor_sweep(int n, const int *lptr, const int *lidx, const double *lval,
                    const int *uptr, const int *uidx, const double *uval, const double *wd,
                    const double *b, double *x, int nthreads);

/* ---- Krylov drivers ---- */
typedef struct {
    int    precon;        /* 0 none, 1 jacobi, 3 ssor (LIS_PRECON_TYPE_*) */
    double ssor_omega;    /* default 1.0 */
    double tol;           /* default 1e-12 */
    int    maxiter;       /* default 1000 */
    int    restart;       /* GMRES, default 40 */
    int    nthreads;      /* OpenMP emulation, 1 = serial */
    /* outputs */
    int    iter;
    int    retcode;       /* 0 success, 2 breakdown, 4 maxiter */
    double resid;
    /* input: SSOR block count; 0 = nthreads (what the reference does).  Setting it apart from
     * nthreads lets a test vary only the reduction order while the preconditioner stays fixed */
    int    ssor_blocks;
} orc_solver_t;

/* x is the initial guess (the reference default zeroes it: pass zeros) and the result.
 * rhistory must hold maxiter+2 doubles.  A in CSR (unsorted allowed).  With precon==3 the
 * matrix-vector product uses the split order, as the reference does after SSOR setup. */
int orc_cg(int n, const int *ptr, const int *idx, const double *val, const double *b, double *x,
           orc_solver_t *s, double *rhistory);
int orc_bicgstab(int n, const int *ptr, const int *idx, const double *val, const double *b, double *x,
                 orc_solver_t *s, double *rhistory);
int orc_gmres(int n, const int *ptr, const int *idx, const double *val, const double *b, double *x,
              orc_solver_t *s, double *rhistory);

#ifdef __cplusplus
}
#endif
#endif
