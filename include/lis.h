/*
 * lis.h -- public C API of lis_b200, a B200-native implementation of the Lis SpMV/Krylov
 * hot path.  Source-compatible with the subset of Lis 2.1.11's lis.h that the reference
 * drivers test/test1.c, test3.c, test3b.c, test4.c, test5.c and test/spmvtest{1,2,2b,3,3b}.c
 * use (reference: include/lis.h:55-283 constants, :513-758 handle types, :824-1045 API,
 * :1052-1078 error codes and LIS_GET_ISIE), so those files compile unchanged against it.
 *
 * Differences a maintainer should know about (details in INTEGRATION.md):
 *   - LIS_SCALAR is double, LIS_INT is int (the reference's default build); no complex,
 *     long-double, quad or 64-bit-index variants.
 *   - vector storage (`v->value`) is CUDA managed memory: host code may read and write it
 *     between API calls exactly as with the reference, kernels run on it in HBM.
 *   - matrices keep the caller's host arrays (same ownership rules as the reference) plus a
 *     private device mirror built on first use.
 *   - every compute entry point runs on the GPU; without a CUDA device it fails with
 *     LIS_ERR_DEVICE instead of silently computing on the CPU.
 */
#ifndef LIS_B200_LIS_H
#define LIS_B200_LIS_H

#include <stdio.h>
#include <stddef.h>

#define LIS_VERSION "2.1.11-b200"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ scalar & index types */
typedef double LIS_SCALAR;
typedef double LIS_REAL;
#ifdef HAVE_COMPLEX_H                      /* set by lis_config.h for the drivers (test/test7.c uses _Complex_I, creal through lis.h) */
#include <complex.h>
#endif
typedef double _Complex LIS_COMPLEX;      /* include/lis.h:418 of the reference; only test/test7.c names it (real build: LIS_SCALAR stays double) */
typedef int LIS_INT;
typedef unsigned int LIS_UNSIGNED_INT;
typedef LIS_INT LIS_Comm;            /* no MPI: ranks are processes bootstrapped from the env */
#define LIS_COMM_WORLD ((LIS_Comm)0x1)
#ifndef conj
#define conj(x) (x)
#endif
#define _max(a, b) ((a) >= (b) ? (a) : (b))
#define _min(a, b) ((a) <= (b) ? (a) : (b))

/* ------------------------------------------------------------------ return codes */
#define LIS_TRUE 1
#define LIS_FALSE 0
#define LIS_FAILS (-1)
#define LIS_SUCCESS 0
#define LIS_ILL_OPTION 1
#define LIS_ERR_ILL_ARG 1
#define LIS_BREAKDOWN 2
#define LIS_OUT_OF_MEMORY 3
#define LIS_ERR_OUT_OF_MEMORY 3
#define LIS_MAXITER 4
#define LIS_ERR_NOT_IMPLEMENTED 5
#define LIS_ERR_FILE_IO 6
#define LIS_ERR_DEVICE 7             /* extension: CUDA device missing or a kernel failed */

/* ------------------------------------------------------------------ enumerations */
/* storage formats (matrix_type) */
#define LIS_MATRIX_ASSEMBLING 0
#define LIS_MATRIX_CSR 1
#define LIS_MATRIX_CSC 2
#define LIS_MATRIX_MSR 3
#define LIS_MATRIX_DIA 4
#define LIS_MATRIX_CDS 4
#define LIS_MATRIX_ELL 5
#define LIS_MATRIX_JAD 6
#define LIS_MATRIX_BSR 7
#define LIS_MATRIX_BSC 8
#define LIS_MATRIX_VBR 9
#define LIS_MATRIX_COO 10
#define LIS_MATRIX_DENSE 11
#define LIS_MATRIX_DNS 11
#define LIS_MATRIX_RCO 255
#define LIS_MATRIX_DECIDING_SIZE (-(LIS_MATRIX_RCO + 1))
#define LIS_MATRIX_NULL (-(LIS_MATRIX_RCO + 2))
#define LIS_MATRIX_DEFAULT LIS_MATRIX_CSR
#define LIS_MATRIX_POINT LIS_MATRIX_CSR
#define LIS_MATRIX_BLOCK LIS_MATRIX_BSR

/* triangular-solve selector of lis_matrix_solve */
#define LIS_MATRIX_LOWER 0
#define LIS_MATRIX_UPPER 1
#define LIS_MATRIX_SSOR 2

/* value insertion */
#define LIS_INS_VALUE 0
#define LIS_ADD_VALUE 1
#define LIS_SUB_VALUE 2

#define LIS_ORIGIN_0 0
#define LIS_ORIGIN_1 1

/* file formats */
#define LIS_FMT_AUTO 0
#define LIS_FMT_PLAIN 1
#define LIS_FMT_MM 2
#define LIS_FMT_LIS 3
#define LIS_FMT_LIS_ASCII 3
#define LIS_FMT_LIS_BINARY 4
#define LIS_FMT_FREE 5
#define LIS_FMT_ITBL 6
#define LIS_FMT_HB 7
#define LIS_FMT_MMB 8

/* solvers (the numbering of the reference; only CG, BiCG, BiCGSTAB and GMRES run here) */
#define LIS_SOLVER_LEN 25
#define LIS_SOLVER_CG 1
#define LIS_SOLVER_BICG 2
#define LIS_SOLVER_CGS 3
#define LIS_SOLVER_BICGSTAB 4
#define LIS_SOLVER_BICGSTABL 5
#define LIS_SOLVER_GPBICG 6
#define LIS_SOLVER_QMR 7
#define LIS_SOLVER_TFQMR 7
#define LIS_SOLVER_ORTHOMIN 8
#define LIS_SOLVER_GMRES 9
#define LIS_SOLVER_JACOBI 10
#define LIS_SOLVER_GS 11
#define LIS_SOLVER_SOR 12
#define LIS_SOLVER_BICGSAFE 13
#define LIS_SOLVER_CR 14
#define LIS_SOLVER_BICR 15
#define LIS_SOLVER_CRS 16
#define LIS_SOLVER_BICRSTAB 17
#define LIS_SOLVER_GPBICR 18
#define LIS_SOLVER_BICRSAFE 19
#define LIS_SOLVER_FGMRES 20
#define LIS_SOLVER_IDRS 21
#define LIS_SOLVER_IDR1 22
#define LIS_SOLVER_MINRES 23
#define LIS_SOLVER_COCG 24
#define LIS_SOLVER_COCR 25

/* preconditioners (only NONE, JACOBI, SSOR and user-registered ones run here) */
#define LIS_PRECON_TYPE_LEN 12
#define LIS_PRECON_TYPE_NONE 0
#define LIS_PRECON_TYPE_JACOBI 1
#define LIS_PRECON_TYPE_ILU 2
#define LIS_PRECON_TYPE_SSOR 3
#define LIS_PRECON_TYPE_HYBRID 4
#define LIS_PRECON_TYPE_IS 5
#define LIS_PRECON_TYPE_SAI 6
#define LIS_PRECON_TYPE_SAAMG 7
#define LIS_PRECON_TYPE_ILUC 8
#define LIS_PRECON_TYPE_ILUT 9
#define LIS_PRECON_TYPE_BJACOBI 10
#define LIS_PRECON_TYPE_ADDS 11
#define LIS_PRECON_TYPE_USERDEF LIS_PRECON_TYPE_LEN
#define LIS_PRECONNAME_MAX 10
#define LIS_PRECON_REGISTER_MAX 10

/* solver->options[] slots and solver->params[] slots (params are addressed as
 * LIS_PARAMS_x - LIS_OPTIONS_LEN, like in the reference) */
#define LIS_OPTIONS_LEN 27
#define LIS_OPTIONS_SOLVER 0
#define LIS_OPTIONS_PRECON 1
#define LIS_OPTIONS_MAXITER 2
#define LIS_OPTIONS_OUTPUT 3
#define LIS_OPTIONS_RESTART 4
#define LIS_OPTIONS_ELL 5
#define LIS_OPTIONS_SCALE 6
#define LIS_OPTIONS_FILL 7
#define LIS_OPTIONS_M 8
#define LIS_OPTIONS_PSOLVER 9
#define LIS_OPTIONS_PMAXITER 10
#define LIS_OPTIONS_PRESTART 11
#define LIS_OPTIONS_PELL 12
#define LIS_OPTIONS_PPRECON 13
#define LIS_OPTIONS_ISLEVEL 14
#define LIS_OPTIONS_INITGUESS_ZEROS 15
#define LIS_OPTIONS_ADDS 16
#define LIS_OPTIONS_ADDS_ITER 17
#define LIS_OPTIONS_PRECISION 18
#define LIS_OPTIONS_USE_AT 19
#define LIS_OPTIONS_SWITCH_MAXITER 20
#define LIS_OPTIONS_SAAMG_UNSYM 21
#define LIS_OPTIONS_STORAGE 22
#define LIS_OPTIONS_STORAGE_BLOCK 23
#define LIS_OPTIONS_CONV_COND 24
#define LIS_OPTIONS_INIT_SHADOW_RESID 25
#define LIS_OPTIONS_IDRS_RESTART 26

#define LIS_PARAMS_LEN 15
#define LIS_PARAMS_RESID (LIS_OPTIONS_LEN + 0)
#define LIS_PARAMS_OMEGA (LIS_OPTIONS_LEN + 1)
#define LIS_PARAMS_RELAX (LIS_OPTIONS_LEN + 2)
#define LIS_PARAMS_DROP (LIS_OPTIONS_LEN + 3)
#define LIS_PARAMS_ALPHA (LIS_OPTIONS_LEN + 4)
#define LIS_PARAMS_TAU (LIS_OPTIONS_LEN + 5)
#define LIS_PARAMS_SIGMA (LIS_OPTIONS_LEN + 6)
#define LIS_PARAMS_GAMMA (LIS_OPTIONS_LEN + 7)
#define LIS_PARAMS_SSOR_OMEGA (LIS_OPTIONS_LEN + 8)
#define LIS_PARAMS_PRESID (LIS_OPTIONS_LEN + 9)
#define LIS_PARAMS_POMEGA (LIS_OPTIONS_LEN + 10)
#define LIS_PARAMS_SWITCH_RESID (LIS_OPTIONS_LEN + 11)
#define LIS_PARAMS_RATE (LIS_OPTIONS_LEN + 12)
#define LIS_PARAMS_RESID_WEIGHT (LIS_OPTIONS_LEN + 13)
#define LIS_PARAMS_SAAMG_THETA (LIS_OPTIONS_LEN + 14)

#define LIS_PRINT_NONE 0
#define LIS_PRINT_MEM 1
#define LIS_PRINT_OUT 2
#define LIS_PRINT_ALL 3

#define LIS_SCALE_NONE 0
#define LIS_SCALE_JACOBI 1
#define LIS_SCALE_SYMM_DIAG 2

#define LIS_CONV_COND_DEFAULT 0
#define LIS_CONV_COND_NRM2_R 0
#define LIS_CONV_COND_NRM2_B 1
#define LIS_CONV_COND_NRM1_B 2

#define LIS_RESID 0
#define LIS_RANDOM 1

#define LIS_PRECISION_DEFAULT 0
#define LIS_PRECISION_DOUBLE 0
#define LIS_PRECISION_QUAD 1
#define LIS_PRECISION_SWITCH 2

#define LIS_LABEL_VECTOR 0
#define LIS_LABEL_MATRIX 1

#define LIS_VECTOR_NULL (-1)
#define LIS_VECTOR_ASSEMBLING 0
#define LIS_VECTOR_ASSEMBLED 1

#define LIS_MATRIX_OPTION_LEN 10

/* function-trace hooks of the reference's debug build: no-ops here */
#define LIS_DEBUG_FUNC_IN
#define LIS_DEBUG_FUNC_OUT

/* ------------------------------------------------------------------ handle types
 * Vectors and matrices start with the same header so that lis_vector_duplicate() can take
 * either (reference: src/vector/lis_vector.c:370-390 checks `label`). */
#define LIS_B200_OBJECT_HEADER                                                             \
    LIS_INT label;      /* LIS_LABEL_VECTOR / LIS_LABEL_MATRIX */                          \
    LIS_INT status;                                                                        \
    LIS_INT precision;                                                                     \
    LIS_INT gn;         /* global size */                                                  \
    LIS_INT n;          /* rows owned by this rank */                                      \
    LIS_INT np;         /* n + halo entries */                                             \
    LIS_INT pad;                                                                           \
    LIS_INT origin;                                                                        \
    LIS_INT is_copy;                                                                       \
    LIS_INT is_destroy;                                                                    \
    LIS_INT is_scaled;                                                                     \
    LIS_INT my_rank;                                                                       \
    LIS_INT nprocs;                                                                        \
    LIS_Comm comm;                                                                         \
    LIS_INT is;         /* first global row owned */                                       \
    LIS_INT ie;         /* one past the last global row owned */                           \
    LIS_INT *ranges;    /* nprocs+1 row offsets */

struct LIS_VECTOR_STRUCT {
    LIS_B200_OBJECT_HEADER
    LIS_SCALAR *value;  /* managed memory, np+pad entries */
    LIS_SCALAR *work;
    LIS_INT intvalue;
    /* private to lis_b200 */
    LIS_INT b200_resident;   /* 1 = pages were last prefetched to the device */
    size_t b200_capacity;    /* allocated entries */
    LIS_INT b200_managed;    /* value came from the device allocator */
};
typedef struct LIS_VECTOR_STRUCT *LIS_VECTOR;

/* strictly lower / upper part of a split matrix */
struct LIS_MATRIX_CORE_STRUCT {
    LIS_INT nnz, ndz, bnr, bnc, nr, nc, bnnz, nnd, maxnzr;
    LIS_INT *ptr, *row, *col, *index, *bptr, *bindex;
    LIS_SCALAR *value;
    LIS_SCALAR *work;
};
typedef struct LIS_MATRIX_CORE_STRUCT *LIS_MATRIX_CORE;

/* (block-)diagonal of a split matrix; scalar blocks only */
struct LIS_MATRIX_DIAG_STRUCT {
    LIS_B200_OBJECT_HEADER
    LIS_SCALAR *value;
    LIS_SCALAR *work;
    LIS_INT bn, nr;
    LIS_INT *bns, *ptr;
    LIS_SCALAR **v_value;
};
typedef struct LIS_MATRIX_DIAG_STRUCT *LIS_MATRIX_DIAG;

struct LIS_COMMTABLE_STRUCT;
typedef struct LIS_COMMTABLE_STRUCT *LIS_COMMTABLE;

struct LIS_MATRIX_STRUCT {
    LIS_B200_OBJECT_HEADER
    LIS_INT matrix_type;
    LIS_INT nnz;        /* CSR, CSC, JAD */
    LIS_INT ndz;
    LIS_INT bnr, bnc;   /* BSR block shape */
    LIS_INT nr, nc;     /* BSR block rows / columns */
    LIS_INT bnnz;       /* BSR blocks */
    LIS_INT nnd;        /* DIA diagonals */
    LIS_INT maxnzr;     /* ELL, JAD */
    LIS_INT *ptr;       /* CSR, CSC, JAD */
    LIS_INT *row;       /* JAD permutation */
    LIS_INT *col;
    LIS_INT *index;     /* CSR, CSC, DIA offsets, ELL, JAD */
    LIS_INT *bptr;      /* BSR */
    LIS_INT *bindex;    /* BSR */
    LIS_SCALAR *value;
    LIS_SCALAR *work;

    LIS_MATRIX_CORE L, U;
    LIS_MATRIX_DIAG D, WD;

    LIS_INT is_block, pad_comm, is_pmat, is_sorted, is_splited, is_save, is_comm, is_fallocated;
    LIS_INT use_wd;
    LIS_INT conv_bnr, conv_bnc;
    LIS_INT *conv_row, *conv_col;
    LIS_INT options[LIS_MATRIX_OPTION_LEN];

    /* row-wise assembly buffers filled by lis_matrix_set_value */
    LIS_INT w_annz;
    LIS_INT *w_nnz;
    LIS_INT *w_row;
    LIS_INT **w_index;
    LIS_SCALAR **w_value;

    LIS_INT *l2g_map;        /* halo slot -> global column */
    LIS_COMMTABLE commtable;

    void *b200_dev;          /* private: device mirror (lis_b200/csrc/host/lis_device.h) */
};
typedef struct LIS_MATRIX_STRUCT *LIS_MATRIX;

struct LIS_SOLVER_STRUCT;

struct LIS_PRECON_STRUCT {
    LIS_INT precon_type;
    LIS_MATRIX A;            /* SSOR: the split matrix */
    LIS_MATRIX Ah;
    LIS_VECTOR D;            /* Jacobi: 1/diag(A) */
    LIS_VECTOR *work;
    struct LIS_SOLVER_STRUCT *solver;
    LIS_INT worklen;
    LIS_INT is_copy;
    void *b200_sweep;        /* private: SSOR level schedule on the device */
    void *b200_ilu;          /* private: ILU(k) factors (host) and their device schedules */
};
typedef struct LIS_PRECON_STRUCT *LIS_PRECON;

struct LIS_SOLVER_STRUCT {
    LIS_MATRIX A, Ah;
    LIS_VECTOR b, x, xx, d;
    LIS_MATRIX_DIAG WD;
    LIS_PRECON precon;
    LIS_VECTOR *work;
    LIS_REAL *rhistory;
    LIS_INT worklen;
    LIS_INT options[LIS_OPTIONS_LEN];
    LIS_SCALAR params[LIS_PARAMS_LEN];
    LIS_INT retcode;
    LIS_INT iter;
    LIS_INT iter2;
    LIS_REAL resid;
    double time, itime, ptime, p_c_time, p_i_time;
    LIS_INT precision;
    LIS_REAL bnrm;
    LIS_REAL tol;
    LIS_REAL tol_switch;
    LIS_INT setup;
};
typedef struct LIS_SOLVER_STRUCT *LIS_SOLVER;

/* eigensolver handle: the reference's public struct (include/lis.h:760-785 there) */
#define LIS_EOPTIONS_LEN 13
#define LIS_EOPTIONS_ESOLVER 0
#define LIS_EOPTIONS_MAXITER 1
#define LIS_EOPTIONS_SUBSPACE 2
#define LIS_EOPTIONS_MODE 3
#define LIS_EOPTIONS_OUTPUT 4
#define LIS_EOPTIONS_INITGUESS_ONES 5
#define LIS_EOPTIONS_INNER_ESOLVER 6
#define LIS_EOPTIONS_INNER_GENERALIZED_ESOLVER 7
#define LIS_EOPTIONS_STORAGE 8
#define LIS_EOPTIONS_STORAGE_BLOCK 9
#define LIS_EOPTIONS_PRECISION 10
#define LIS_EOPTIONS_SWITCH_MAXITER 11
#define LIS_EOPTIONS_RVAL 12
#define LIS_EPARAMS_LEN 3
#define LIS_EPARAMS_RESID (LIS_EOPTIONS_LEN + 0)
#define LIS_EPARAMS_SHIFT (LIS_EOPTIONS_LEN + 1)
#define LIS_EPARAMS_SHIFT_IM (LIS_EOPTIONS_LEN + 2)
#define LIS_EPRINT_NONE 0
#define LIS_EPRINT_MEM 1
#define LIS_EPRINT_OUT 2
#define LIS_EPRINT_ALL 3
#define LIS_ESOLVER_LEN 16
#define LIS_ESOLVER_PI 1
#define LIS_ESOLVER_II 2
#define LIS_ESOLVER_RQI 3
#define LIS_ESOLVER_CG 4
#define LIS_ESOLVER_CR 5
#define LIS_ESOLVER_SI 6
#define LIS_ESOLVER_LI 7
#define LIS_ESOLVER_AI 8
#define LIS_ESOLVER_GPI 9
#define LIS_ESOLVER_GII 10
#define LIS_ESOLVER_GRQI 11
#define LIS_ESOLVER_GCG 12
#define LIS_ESOLVER_GCR 13
#define LIS_ESOLVER_GSI 14
#define LIS_ESOLVER_GLI 15
#define LIS_ESOLVER_GAI 16

struct LIS_ESOLVER_STRUCT {
    LIS_MATRIX A, B;
    LIS_VECTOR x, xx, d;
    LIS_SCALAR *evalue;
    LIS_VECTOR *evector;
    LIS_REAL *resid;
    LIS_VECTOR *work;
    LIS_REAL *rhistory;
    LIS_INT worklen;
    LIS_INT options[LIS_EOPTIONS_LEN];
    LIS_SCALAR params[LIS_EPARAMS_LEN];
    LIS_INT retcode;
    LIS_INT *iter;
    LIS_INT *iter2;
    double time;
    LIS_INT *nesol;
    double itime;
    double ptime;
    double p_c_time;
    double p_i_time;
    LIS_INT eprecision;
    LIS_SCALAR ishift;
    LIS_REAL nrm2;
    LIS_REAL tol;
    LIS_INT nevector;        /* private to lis_b200: eigenvector handles owned by evector[] */
};
typedef struct LIS_ESOLVER_STRUCT *LIS_ESOLVER;

typedef LIS_INT (*LIS_PRECON_CREATE_XXX)(LIS_SOLVER solver, LIS_PRECON precon);
typedef LIS_INT (*LIS_PSOLVE_XXX)(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);
typedef LIS_INT (*LIS_PSOLVEH_XXX)(LIS_SOLVER solver, LIS_VECTOR b, LIS_VECTOR x);

/* ------------------------------------------------------------------ library lifetime */
LIS_INT lis_initialize(int *argc, char **argv[]);
LIS_INT lis_finalize(void);
double lis_wtime(void);
void CHKERR(LIS_INT err);
LIS_INT lis_printf(LIS_Comm comm, const char *mess, ...);   /* rank 0 only; %D = LIS_INT */
void *lis_malloc(size_t size, char *tag);
void *lis_calloc(size_t size, char *tag);
void *lis_realloc(void *p, size_t size);
void lis_free(void *p);
void lis_free2(LIS_INT n, ...);
LIS_INT lis_is_malloc(void *p);
void lis_date(char *date);

/* ------------------------------------------------------------------ vectors */
LIS_INT lis_vector_create(LIS_Comm comm, LIS_VECTOR *vec);
LIS_INT lis_vector_set_size(LIS_VECTOR vec, LIS_INT local_n, LIS_INT global_n);
LIS_INT lis_vector_destroy(LIS_VECTOR vec);
LIS_INT lis_vector_duplicate(void *vin, LIS_VECTOR *vout);
LIS_INT lis_vector_get_size(LIS_VECTOR v, LIS_INT *local_n, LIS_INT *global_n);
LIS_INT lis_vector_get_range(LIS_VECTOR v, LIS_INT *is, LIS_INT *ie);
LIS_INT lis_vector_get_value(LIS_VECTOR v, LIS_INT i, LIS_SCALAR *value);
LIS_INT lis_vector_get_values(LIS_VECTOR v, LIS_INT start, LIS_INT count, LIS_SCALAR value[]);
LIS_INT lis_vector_set_value(LIS_INT flag, LIS_INT i, LIS_SCALAR value, LIS_VECTOR v);
LIS_INT lis_vector_set_values(LIS_INT flag, LIS_INT count, LIS_INT index[], LIS_SCALAR value[], LIS_VECTOR v);
LIS_INT lis_vector_set_values2(LIS_INT flag, LIS_INT start, LIS_INT count, LIS_SCALAR value[], LIS_VECTOR v);
LIS_INT lis_vector_print(LIS_VECTOR x);
LIS_INT lis_vector_scatter(LIS_SCALAR value[], LIS_VECTOR v);
LIS_INT lis_vector_gather(LIS_VECTOR v, LIS_SCALAR value[]);
LIS_INT lis_vector_is_null(LIS_VECTOR v);

LIS_INT lis_vector_swap(LIS_VECTOR vsrc, LIS_VECTOR vdst);
LIS_INT lis_vector_copy(LIS_VECTOR vsrc, LIS_VECTOR vdst);
LIS_INT lis_vector_axpy(LIS_SCALAR alpha, LIS_VECTOR vx, LIS_VECTOR vy);
LIS_INT lis_vector_xpay(LIS_VECTOR vx, LIS_SCALAR alpha, LIS_VECTOR vy);
LIS_INT lis_vector_axpyz(LIS_SCALAR alpha, LIS_VECTOR vx, LIS_VECTOR vy, LIS_VECTOR vz);
LIS_INT lis_vector_scale(LIS_SCALAR alpha, LIS_VECTOR vx);
LIS_INT lis_vector_pmul(LIS_VECTOR vx, LIS_VECTOR vy, LIS_VECTOR vz);
LIS_INT lis_vector_pdiv(LIS_VECTOR vx, LIS_VECTOR vy, LIS_VECTOR vz);
LIS_INT lis_vector_set_all(LIS_SCALAR alpha, LIS_VECTOR vx);
LIS_INT lis_vector_abs(LIS_VECTOR vx);
LIS_INT lis_vector_reciprocal(LIS_VECTOR vx);
LIS_INT lis_vector_conjugate(LIS_VECTOR vx);
LIS_INT lis_vector_shift(LIS_SCALAR sigma, LIS_VECTOR vx);
LIS_INT lis_vector_dot(LIS_VECTOR vx, LIS_VECTOR vy, LIS_SCALAR *value);
LIS_INT lis_vector_nhdot(LIS_VECTOR vx, LIS_VECTOR vy, LIS_SCALAR *value);
LIS_INT lis_vector_nrm1(LIS_VECTOR vx, LIS_REAL *value);
LIS_INT lis_vector_nrm2(LIS_VECTOR vx, LIS_REAL *value);
LIS_INT lis_vector_nrmi(LIS_VECTOR vx, LIS_REAL *value);
LIS_INT lis_vector_sum(LIS_VECTOR vx, LIS_SCALAR *value);

/* ------------------------------------------------------------------ matrices */
LIS_INT lis_matrix_create(LIS_Comm comm, LIS_MATRIX *Amat);
LIS_INT lis_matrix_destroy(LIS_MATRIX Amat);
LIS_INT lis_matrix_assemble(LIS_MATRIX A);
LIS_INT lis_matrix_is_assembled(LIS_MATRIX A);
LIS_INT lis_matrix_duplicate(LIS_MATRIX Ain, LIS_MATRIX *Aout);
LIS_INT lis_matrix_set_size(LIS_MATRIX A, LIS_INT local_n, LIS_INT global_n);
LIS_INT lis_matrix_get_size(LIS_MATRIX A, LIS_INT *local_n, LIS_INT *global_n);
LIS_INT lis_matrix_get_range(LIS_MATRIX A, LIS_INT *is, LIS_INT *ie);
LIS_INT lis_matrix_get_nnz(LIS_MATRIX A, LIS_INT *nnz);
LIS_INT lis_matrix_set_type(LIS_MATRIX A, LIS_INT matrix_type);
LIS_INT lis_matrix_get_type(LIS_MATRIX A, LIS_INT *matrix_type);
LIS_INT lis_matrix_set_value(LIS_INT flag, LIS_INT i, LIS_INT j, LIS_SCALAR value, LIS_MATRIX A);
LIS_INT lis_matrix_malloc(LIS_MATRIX A, LIS_INT nnz_row, LIS_INT nnz[]);
LIS_INT lis_matrix_get_diagonal(LIS_MATRIX A, LIS_VECTOR d);
LIS_INT lis_matrix_convert(LIS_MATRIX Ain, LIS_MATRIX Aout);
LIS_INT lis_matrix_copy(LIS_MATRIX Ain, LIS_MATRIX Aout);
LIS_INT lis_matrix_set_blocksize(LIS_MATRIX A, LIS_INT bnr, LIS_INT bnc, LIS_INT row[], LIS_INT col[]);
LIS_INT lis_matrix_unset(LIS_MATRIX A);

/* set_* adopt the caller's arrays without copying; lis_matrix_destroy frees them */
LIS_INT lis_matrix_malloc_csr(LIS_INT n, LIS_INT nnz, LIS_INT **ptr, LIS_INT **index, LIS_SCALAR **value);
LIS_INT lis_matrix_set_csr(LIS_INT nnz, LIS_INT *ptr, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A);
LIS_INT lis_matrix_malloc_csc(LIS_INT n, LIS_INT nnz, LIS_INT **ptr, LIS_INT **index, LIS_SCALAR **value);
LIS_INT lis_matrix_set_csc(LIS_INT nnz, LIS_INT *ptr, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A);
LIS_INT lis_matrix_malloc_bsr(LIS_INT n, LIS_INT bnr, LIS_INT bnc, LIS_INT bnnz, LIS_INT **bptr, LIS_INT **bindex, LIS_SCALAR **value);
LIS_INT lis_matrix_set_bsr(LIS_INT bnr, LIS_INT bnc, LIS_INT bnnz, LIS_INT *bptr, LIS_INT *bindex, LIS_SCALAR *value, LIS_MATRIX A);
LIS_INT lis_matrix_malloc_ell(LIS_INT n, LIS_INT maxnzr, LIS_INT **index, LIS_SCALAR **value);
LIS_INT lis_matrix_set_ell(LIS_INT maxnzr, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A);
LIS_INT lis_matrix_malloc_jad(LIS_INT n, LIS_INT nnz, LIS_INT maxnzr, LIS_INT **perm, LIS_INT **ptr, LIS_INT **index, LIS_SCALAR **value);
LIS_INT lis_matrix_set_jad(LIS_INT nnz, LIS_INT maxnzr, LIS_INT *perm, LIS_INT *ptr, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A);
/* in-place update of a stored entry of an assembled CSR matrix, and friends (include/lis.h:827-892, :1032-1037 of the reference) */
LIS_INT lis_matrix_psd_set_value(LIS_INT flag, LIS_INT i, LIS_INT j, LIS_SCALAR value, LIS_MATRIX A);
LIS_INT lis_matrix_psd_set_value_csr(LIS_INT flag, LIS_INT i, LIS_INT j, LIS_SCALAR value, LIS_MATRIX A);
LIS_INT lis_matrix_psd_reset_scale(LIS_MATRIX A);
LIS_INT lis_vector_psd_reset_scale(LIS_VECTOR vec);
LIS_INT lis_matrix_get_vbr_rowcol(LIS_MATRIX Ain, LIS_INT *nr, LIS_INT *nc, LIS_INT **row, LIS_INT **col);
void    lis_do_not_handle_mpi(void);
LIS_INT lis_debug_trace_func(LIS_INT flag, char *func);
/* the formats outside the named hot path (MSR, COO, BSC, VBR, DNS; /root/reference include/lis.h:898-914): conversion
 * from and to CSR, lis_matvec through a row-ordered device mirror with the reference's accumulation order */
LIS_INT lis_matrix_malloc_msr(LIS_INT n, LIS_INT nnz, LIS_INT ndz, LIS_INT **index, LIS_SCALAR **value);
LIS_INT lis_matrix_set_msr(LIS_INT nnz, LIS_INT ndz, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A);
LIS_INT lis_matrix_malloc_coo(LIS_INT nnz, LIS_INT **row, LIS_INT **col, LIS_SCALAR **value);
LIS_INT lis_matrix_set_coo(LIS_INT nnz, LIS_INT *row, LIS_INT *col, LIS_SCALAR *value, LIS_MATRIX A);
LIS_INT lis_matrix_malloc_bsc(LIS_INT n, LIS_INT bnr, LIS_INT bnc, LIS_INT bnnz, LIS_INT **bptr, LIS_INT **bindex, LIS_SCALAR **value);
LIS_INT lis_matrix_set_bsc(LIS_INT bnr, LIS_INT bnc, LIS_INT bnnz, LIS_INT *bptr, LIS_INT *bindex, LIS_SCALAR *value, LIS_MATRIX A);
LIS_INT lis_matrix_malloc_vbr(LIS_INT n, LIS_INT nnz, LIS_INT nr, LIS_INT nc, LIS_INT bnnz, LIS_INT **row, LIS_INT **col, LIS_INT **ptr,
                              LIS_INT **bptr, LIS_INT **bindex, LIS_SCALAR **value);
LIS_INT lis_matrix_set_vbr(LIS_INT nnz, LIS_INT nr, LIS_INT nc, LIS_INT bnnz, LIS_INT *row, LIS_INT *col, LIS_INT *ptr, LIS_INT *bptr,
                           LIS_INT *bindex, LIS_SCALAR *value, LIS_MATRIX A);
LIS_INT lis_matrix_malloc_dns(LIS_INT n, LIS_INT np, LIS_SCALAR **value);
LIS_INT lis_matrix_set_dns(LIS_SCALAR *value, LIS_MATRIX A);
LIS_INT lis_matrix_malloc_dia(LIS_INT n, LIS_INT nnd, LIS_INT **index, LIS_SCALAR **value);
LIS_INT lis_matrix_set_dia(LIS_INT nnd, LIS_INT *index, LIS_SCALAR *value, LIS_MATRIX A);

/* ------------------------------------------------------------------ matrix-vector product */
LIS_INT lis_matvec(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y);
LIS_INT lis_matvech(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y);

/* ------------------------------------------------------------------ linear solvers */
LIS_INT lis_solver_create(LIS_SOLVER *solver);
LIS_INT lis_solver_destroy(LIS_SOLVER solver);
LIS_INT lis_solver_get_iter(LIS_SOLVER solver, LIS_INT *iter);
LIS_INT lis_solver_get_iterex(LIS_SOLVER solver, LIS_INT *iter, LIS_INT *iter_double, LIS_INT *iter_quad);
LIS_INT lis_solver_get_time(LIS_SOLVER solver, double *time);
LIS_INT lis_solver_get_timeex(LIS_SOLVER solver, double *time, double *itime, double *ptime, double *p_c_time, double *p_i_time);
LIS_INT lis_solver_get_residualnorm(LIS_SOLVER solver, LIS_REAL *residual);
LIS_INT lis_solver_get_solver(LIS_SOLVER solver, LIS_INT *nsol);
LIS_INT lis_solver_get_precon(LIS_SOLVER solver, LIS_INT *precon_type);
LIS_INT lis_solver_get_status(LIS_SOLVER solver, LIS_INT *status);
LIS_INT lis_solver_get_rhistory(LIS_SOLVER solver, LIS_VECTOR v);
LIS_INT lis_solver_set_option(char *text, LIS_SOLVER solver);
LIS_INT lis_solver_set_optionC(LIS_SOLVER solver);
LIS_INT lis_solve(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, LIS_SOLVER solver);
LIS_INT lis_solve_kernel(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, LIS_SOLVER solver, LIS_PRECON precon);
LIS_INT lis_solve_setup(LIS_MATRIX A, LIS_SOLVER solver);
LIS_INT lis_solver_set_matrix(LIS_MATRIX A, LIS_SOLVER solver);
LIS_INT lis_matrix_set_values(LIS_INT flag, LIS_INT n, LIS_SCALAR value[], LIS_MATRIX A);
LIS_INT lis_matrix_shift_diagonal(LIS_MATRIX A, LIS_SCALAR sigma);
LIS_INT lis_matrix_scale(LIS_MATRIX A, LIS_VECTOR B, LIS_VECTOR D, LIS_INT action);

/* ------------------------------------------------------------------ eigensolvers (standard problem:
 * power, inverse, Rayleigh quotient, CG, CR, subspace, Lanczos, Arnoldi) */
LIS_INT lis_esolver_create(LIS_ESOLVER *esolver);
LIS_INT lis_esolver_destroy(LIS_ESOLVER esolver);
LIS_INT lis_esolver_work_destroy(LIS_ESOLVER esolver);
LIS_INT lis_esolver_set_option(char *text, LIS_ESOLVER esolver);
LIS_INT lis_esolver_set_optionC(LIS_ESOLVER esolver);
LIS_INT lis_esolve(LIS_MATRIX A, LIS_VECTOR x, LIS_SCALAR *evalue0, LIS_ESOLVER esolver);
LIS_INT lis_gesolve(LIS_MATRIX A, LIS_MATRIX B, LIS_VECTOR x, LIS_SCALAR *evalue0, LIS_ESOLVER esolver);   /* B must be NULL: standard problems only */
LIS_INT lis_esolver_get_iter(LIS_ESOLVER esolver, LIS_INT *iter);
LIS_INT lis_esolver_get_iterex(LIS_ESOLVER esolver, LIS_INT *iter, LIS_INT *iter_double, LIS_INT *iter_quad);
LIS_INT lis_esolver_get_time(LIS_ESOLVER esolver, double *time);
LIS_INT lis_esolver_get_timeex(LIS_ESOLVER esolver, double *time, double *itime, double *ptime, double *p_c_time, double *p_i_time);
LIS_INT lis_esolver_get_residualnorm(LIS_ESOLVER esolver, LIS_REAL *residual);
LIS_INT lis_esolver_get_status(LIS_ESOLVER esolver, LIS_INT *status);
LIS_INT lis_esolver_get_rhistory(LIS_ESOLVER esolver, LIS_VECTOR v);
LIS_INT lis_esolver_get_evalues(LIS_ESOLVER esolver, LIS_VECTOR v);
LIS_INT lis_esolver_get_specific_evalue(LIS_ESOLVER esolver, LIS_INT mode, LIS_SCALAR *evalue);
LIS_INT lis_esolver_get_evectors(LIS_ESOLVER esolver, LIS_MATRIX M);
LIS_INT lis_esolver_get_specific_evector(LIS_ESOLVER esolver, LIS_INT mode, LIS_VECTOR x);
LIS_INT lis_esolver_get_residualnorms(LIS_ESOLVER esolver, LIS_VECTOR v);
LIS_INT lis_esolver_get_specific_residualnorm(LIS_ESOLVER esolver, LIS_INT mode, LIS_REAL *residual);
LIS_INT lis_esolver_get_iters(LIS_ESOLVER esolver, LIS_VECTOR v);
LIS_INT lis_esolver_get_specific_iter(LIS_ESOLVER esolver, LIS_INT mode, LIS_INT *iter);
LIS_INT lis_esolver_get_esolver(LIS_ESOLVER esolver, LIS_INT *nesol);
LIS_INT lis_esolver_get_esolvername(LIS_INT esolver, char *esolvername);
LIS_INT lis_esolver_output_rhistory(LIS_ESOLVER esolver, char *filename);
LIS_INT lis_solver_get_solvername(LIS_INT solver, char *solvername);
LIS_INT lis_solver_get_preconname(LIS_INT precon_type, char *preconname);
LIS_INT lis_precon_register(char *name, LIS_PRECON_CREATE_XXX pcreate, LIS_PSOLVE_XXX psolve, LIS_PSOLVEH_XXX psolveh);
LIS_INT lis_precon_register_free(void);

/* ------------------------------------------------------------------ lis_b200 extensions */
/* drop the device mirror of A after editing A->value / A->index behind the library's back */
LIS_INT lis_matrix_b200_invalidate(LIS_MATRIX A);
/* emulated OpenMP thread count of the reference (= SSOR block count); same as passing
 * `-omp_num_threads N` to lis_initialize.  Returns the previous value. */
LIS_INT lis_b200_set_num_threads(LIS_INT nthreads);
/* join a process group explicitly (rank, size, 64-bit job token shared by all ranks) instead
 * of through RANK / WORLD_SIZE / MASTER_PORT in the environment */
LIS_INT lis_b200_comm_attach(LIS_INT rank, LIS_INT nranks, unsigned long long token);
/* rank-ordered sum of host scalars over the process group (what every dot/nrm2 does) */
LIS_INT lis_b200_allreduce_sum(double *vals, LIS_INT count);
/* sizes and lists of A's halo exchange: out = {halo entries, exported entries, neighbours} */
LIS_INT lis_b200_commtable_info(LIS_MATRIX A, LIS_INT *out, LIS_INT *import_ptr, LIS_INT *export_ptr, LIS_INT *export_index,
                                LIS_INT *l2g_map, LIS_INT cap);
/* y = A x for an application whose vectors live in host arrays: the same result as
 * lis_vector_scatter(host_x, x); lis_matvec(A, x, y); lis_vector_gather(y, host_y) -- x, y and
 * host_y end up with the same bits -- but copy-in, product and copy-out run chunk-wise
 * overlapped on three streams (pin host_x/host_y for the copies to be asynchronous).  On a
 * row-partitioned matrix host_x / host_y are the calling rank's n local entries. */
LIS_INT lis_b200_matvec_host(LIS_MATRIX A, LIS_SCALAR host_x[], LIS_VECTOR x, LIS_VECTOR y, LIS_SCALAR host_y[]);
LIS_INT lis_b200_matvec_host_plan(LIS_MATRIX A, LIS_INT cap, LIS_INT *rows, LIS_INT *need);
/* row-partitioned CSR products run the rows that read no halo entry on a second stream while the halo
 * exchange is in flight; so does CG's fused q = A p, <p,q> step (the dot is then the sum of the range
 * shares).  on = 1 both (default), 2 products only (LIS_B200_OVERLAP=spmv), 0 neither (LIS_B200_OVERLAP=0).
 * Returns the old setting. */
LIS_INT lis_b200_set_overlap(LIS_INT on);
/* partial scalars of dot/nrm2 across ranks: 0 = host control plane (default), 1 = ncclAllGather over NVLink */
LIS_INT lis_b200_set_reduce(LIS_INT nccl);
/* row-partitioned CSR products exchange their halo inside the SpMV kernel over peer memory (CUDA IPC + NVLink)
 * where every rank can map its neighbours' inbox (virtual-memory API, host/lis_peer.c): 1 (default) use it, 0 the NCCL
 * send/recv exchange (also LIS_B200_P2P=0).
 * Returns the old setting; call on every rank alike. */
LIS_INT lis_b200_set_p2p(LIS_INT on);
unsigned long long lis_b200_p2p_products(void);
LIS_INT lis_b200_p2p_release(LIS_MATRIX A);        /* unmap the neighbours' inboxes of A's communication table (collective) */   /* products so far that took that path (diagnostics) */
/* lis_matvec enqueued on the library stream without waiting for it; lis_b200_sync (or any host-synchronous call) waits */
LIS_INT lis_b200_matvec_async(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y);
LIS_INT lis_b200_sync(void);
/* the CUDA stream (cudaStream_t) all kernels of this process are enqueued on */
void *lis_b200_stream(void);

/* ------------------------------------------------------------------ small dense helpers (host arrays,
 * column-major n x n; src/array/lis_array.c) */
LIS_INT lis_array_swap(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y);
LIS_INT lis_array_copy(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y);
LIS_INT lis_array_axpy(LIS_INT n, LIS_SCALAR alpha, LIS_SCALAR *x, LIS_SCALAR *y);
LIS_INT lis_array_xpay(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR alpha, LIS_SCALAR *y);
LIS_INT lis_array_axpyz(LIS_INT n, LIS_SCALAR alpha, LIS_SCALAR *x, LIS_SCALAR *y, LIS_SCALAR *z);
LIS_INT lis_array_scale(LIS_INT n, LIS_SCALAR alpha, LIS_SCALAR *x);
LIS_INT lis_array_pmul(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y, LIS_SCALAR *z);
LIS_INT lis_array_pdiv(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y, LIS_SCALAR *z);
LIS_INT lis_array_set_all(LIS_INT n, LIS_SCALAR alpha, LIS_SCALAR *x);
LIS_INT lis_array_abs(LIS_INT n, LIS_SCALAR *x);
LIS_INT lis_array_reciprocal(LIS_INT n, LIS_SCALAR *x);
LIS_INT lis_array_conjugate(LIS_INT n, LIS_SCALAR *x);
LIS_INT lis_array_shift(LIS_INT n, LIS_SCALAR sigma, LIS_SCALAR *x);
LIS_INT lis_array_dot(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y, LIS_SCALAR *value);
LIS_INT lis_array_nhdot(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *y, LIS_SCALAR *value);
LIS_INT lis_array_nrm1(LIS_INT n, LIS_SCALAR *x, LIS_REAL *value);
LIS_INT lis_array_nrm2(LIS_INT n, LIS_SCALAR *x, LIS_REAL *value);
LIS_INT lis_array_nrmi(LIS_INT n, LIS_SCALAR *x, LIS_REAL *value);
LIS_INT lis_array_sum(LIS_INT n, LIS_SCALAR *x, LIS_SCALAR *value);
LIS_INT lis_array_matvec(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *x, LIS_SCALAR *y, LIS_INT op);
LIS_INT lis_array_matvech(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *x, LIS_SCALAR *y, LIS_INT op);
LIS_INT lis_array_matvec_ns(LIS_INT m, LIS_INT n, LIS_SCALAR *a, LIS_INT lda, LIS_SCALAR *x, LIS_SCALAR *y, LIS_INT op);
LIS_INT lis_array_matmat(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *b, LIS_SCALAR *c, LIS_INT op);
LIS_INT lis_array_matmat_ns(LIS_INT l, LIS_INT m, LIS_INT n, LIS_SCALAR *a, LIS_INT lda, LIS_SCALAR *b, LIS_INT ldb,
                            LIS_SCALAR *c, LIS_INT ldc, LIS_INT op);
LIS_INT lis_array_ge(LIS_INT n, LIS_SCALAR *a);
LIS_INT lis_array_solve(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *b, LIS_SCALAR *x, LIS_SCALAR *w);
LIS_INT lis_array_cgs(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *q, LIS_SCALAR *r);
LIS_INT lis_array_mgs(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *q, LIS_SCALAR *r);
LIS_INT lis_array_qr(LIS_INT n, LIS_SCALAR *a, LIS_SCALAR *q, LIS_SCALAR *r, LIS_INT *qriter, LIS_REAL *qrerr);

/* ------------------------------------------------------------------ file I/O */
LIS_INT lis_input(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, char *filename);
LIS_INT lis_input_matrix(LIS_MATRIX A, char *filename);
LIS_INT lis_input_vector(LIS_VECTOR v, char *filename);
LIS_INT lis_output_vector(LIS_VECTOR v, LIS_INT format, char *filename);
LIS_INT lis_output_matrix(LIS_MATRIX A, LIS_INT format, char *path);
LIS_INT lis_output(LIS_MATRIX A, LIS_VECTOR b, LIS_VECTOR x, LIS_INT format, char *path);
LIS_INT lis_solver_output_rhistory(LIS_SOLVER solver, char *filename);

#ifdef __cplusplus
}
#endif

/* contiguous 1-D row partition: rows [is, ie) of n belong to part `id` of `nprocs` */
#define LIS_GET_ISIE(id, nprocs, n, is, ie)                                                \
    if ((id) < (n) % (nprocs)) {                                                           \
        (ie) = (n) / (nprocs) + 1;                                                         \
        (is) = (ie) * (id);                                                                \
    } else {                                                                               \
        (ie) = (n) / (nprocs);                                                             \
        (is) = (ie) * (id) + (n) % (nprocs);                                               \
    }                                                                                      \
    (ie) = (ie) + (is);

#endif /* LIS_B200_LIS_H */
