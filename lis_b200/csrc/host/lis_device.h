/*
 * lis_device.h -- private device runtime of lis_b200: one CUDA stream per process (one
 * process per GPU), scratch for reductions, mapped pinned scalars for dot/nrm2 results,
 * managed-memory vector storage and device mirrors of the matrix formats.
 *
 * Replaces, for the GPU, what the reference gets for free from a single address space:
 * src/system/lis_memory.c (allocation) and the global reduction scratch lis_vec_tmp
 * (src/system/lis_init.c:66).
 */
#ifndef LIS_B200_DEVICE_H
#define LIS_B200_DEVICE_H

#include "lislib.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LISD_NSCALARS 64          /* mapped host scalars available to kernels */

/* ---- runtime ---- */
int   lisd_available(void);                         /* 1 when a CUDA device is usable */
LIS_INT lisd_require(const char *what);             /* LIS_SUCCESS or LIS_ERR_DEVICE (+message) */
void *lisd_stream(void);
void  lisd_mark_busy(void);                         /* a kernel was enqueued */
LIS_INT lisd_sync(void);                            /* wait for the stream if busy */
LIS_INT lisd_check(int cuda_rc, const char *what);  /* map a kernel-ABI return code */
void  lisd_shutdown(void);
int   lisd_device_id(void);

/* ---- memory ---- */
LIS_INT lisd_alloc_vector(size_t count, LIS_SCALAR **value, LIS_INT *managed);
void    lisd_free_vector(LIS_SCALAR *value, LIS_INT managed);
void    lisd_free_vector_bytes(LIS_SCALAR *value, LIS_INT managed, size_t count);   /* parks the block in the pool */
LIS_INT lisd_malloc(void **p, size_t bytes);        /* plain device memory */
void    lisd_free(void *p);
LIS_INT lisd_upload(void *dst, const void *src, size_t bytes);      /* H2D, synchronous */
LIS_INT lisd_download(void *dst, const void *src, size_t bytes);    /* D2H, synchronous */
LIS_INT lisd_memset(void *dst, int byte, size_t bytes);

/* ---- copy pipeline: H2D stream -> main stream -> D2H stream, chained per chunk by events ---- */
LIS_INT lisd_pipe_begin(int nchunks);                                   /* copy-in stream starts behind the main stream */
LIS_INT lisd_pipe_h2d(int c, void *dst, const void *src, size_t bytes); /* queue chunk c's input */
LIS_INT lisd_pipe_wait_in(int c);                                       /* main stream waits for chunk c's input */
LIS_INT lisd_pipe_d2h(int c, void *dst, const void *src, size_t bytes); /* after what the main stream holds now */
LIS_INT lisd_pipe_end(void);                                            /* main stream joins the copy-out stream */
LIS_INT lisd_pipe_staging(size_t xcount, size_t ycount, double **xs, double **ys);   /* cached cudaMalloc'ed staging vectors */
LIS_INT lisd_d2d(void *dst, const void *src, size_t bytes);             /* async on the main stream */

/* ---- second compute stream: fork behind the main stream's current work, join back ---- */
LIS_INT lisd_aux_fork(void **stream);
LIS_INT lisd_aux_join(void);

/* ---- vectors: residency tracking of managed storage ---- */
LIS_INT lisd_vec_device(LIS_VECTOR v);              /* make resident before a kernel touches it */
void    lisd_vec_host(LIS_VECTOR v);                /* host is about to read/write v->value */
LIS_SCALAR *lisd_vec_host_view(LIS_VECTOR v, int load);                      /* v->value, or a heap copy of device-only storage */
LIS_INT lisd_vec_host_done(LIS_VECTOR v, LIS_SCALAR *view, int store);       /* writes the copy back (store) and frees it */

/* ---- reductions ---- */
double       *lisd_partial(size_t slots);           /* device scratch, grown on demand */
unsigned int *lisd_counter(void);
double       *lisd_scalar_dev(int slot);            /* device alias of mapped scalar `slot` */
double        lisd_scalar_get(int slot);            /* host value (after lisd_sync) */
double       *lisd_dev_scalar(int slot);            /* device-resident scalar slot (stays on the GPU) */
LIS_INT       lisd_dev_scalars_fetch(int first, int count);   /* queue D2H of slots; valid after the next sync */
double        lisd_fetched(int slot);
/* <x,y> into device slot `slot` without waiting; y += (scale * slot) * x reading it back on the device */
LIS_INT       lisd_dot_to_slot(LIS_VECTOR x, LIS_VECTOR y, int slot);
LIS_INT       lisd_axpy_from_slot(int slot, double scale, LIS_VECTOR x, LIS_VECTOR y);
/* both in one pass: y += (scale * slot_in) * x; then <y,u> -> slot_out, or (u == NULL) ||y||_2 -> *nrm2 (waits) */
LIS_INT       lisd_mgs_step(int slot_in, double scale, LIS_VECTOR x, LIS_VECTOR y, LIS_VECTOR u, int slot_out, LIS_REAL *nrm2);

/* ---- matrix device mirror ---- */
#define LISD_SH_IDX 1u
#define LISD_SH_VAL 2u
#define LISD_SH_PERM 4u
#define LISD_SH_BIDX 8u
void *lisd_shared_alloc(size_t bytes);       /* managed block, resident on the device; NULL when unavailable */
int   lisd_is_shared(const void *p);
int   lisd_shared_release(void *p);          /* 1: p was such a block and has been freed */

typedef struct lisd_csr {
    int n, nnz;
    int *ptr, *idx;
    double *val;
    int tma_rows, tma_tile, tma_stages;   /* plan of the TMA row-block kernel; tma_rows == 0: product-tile kernel */
} lisd_csr;

typedef struct lisd_matrix {
    int type;                 /* LIS_MATRIX_* the mirror was built for */
    int n, np;
    lisd_csr csr;             /* CSR; for CSC: the row-major (transposed-storage) mirror */
    /* ELL / DIA / JAD / BSR */
    int maxnzr, nnd, ld, nr, bnr, bnc, bnnz;
    int *idx, *off, *jptr, *perm, *bptr, *bidx;
    double *val;
    /* split parts */
    int splited;
    lisd_csr L, U;
    double *diag;             /* D */
    double *wd;               /* WD (scaled + inverted diagonal), when present */
    void *sweep;              /* SSOR level schedule (lis_precon.c), built on first psolve */
    void *sweep_global;       /* one-block schedule for LIS_MATRIX_LOWER when `sweep` is blocked */
    /* row-partitioned CSR: rows [ov_lo, ov_hi) read no halo entry and run on a second stream while the
     * halo exchange is in flight (ov_built: 0 not looked at yet, 1 usable, -1 not worth it) */
    int ov_built, ov_lo, ov_hi;
    /* row chunks of the pipelined host-buffer product (lis_b200_matvec_host), built on first use:
     * rows [pipe_row[c], pipe_row[c+1]) read no x entry beyond chunk pipe_need[c] */
    int pipe_n;
    int *pipe_row, *pipe_need;
    /* transposed mirrors for lis_matvech (BiCG), built on first use */
    unsigned shared;          /* LISD_SH_*: pointers below that are the matrix's public (managed) arrays, not the mirror's own */
    int has_t;
    lisd_csr csrT, LT, UT;
} lisd_matrix;

LIS_INT lisd_matrix_get(LIS_MATRIX A, lisd_matrix **out);   /* build on first use */
void    lisd_matrix_drop(LIS_MATRIX A);                     /* invalidate / free */
LIS_INT lisd_matrix_refresh_wd(LIS_MATRIX A);               /* after WD changed on the host */
LIS_INT lisd_matrix_shift_diagonal(LIS_MATRIX A, LIS_SCALAR sigma);  /* same edit on the mirror (or drops it) */
void    lisd_mirror_free(lisd_matrix *M);                   /* a mirror that is not (yet) attached to a matrix */
/* CSR -> ELL/DIA/JAD/BSR by kernels (lis_convert_dev.c); *done = 0: not handled, run the host builder */
int     lisd_convert_on_device(void);
LIS_INT lisd_convert_from_csr(LIS_MATRIX Acsr, LIS_MATRIX Aout, int *done);

/* ---- internal async vector ops (no host sync; the public lis_vector_* wrap these) ---- */
LIS_INT lisd_copy(LIS_VECTOR x, LIS_VECTOR y);
LIS_INT lisd_axpy(LIS_SCALAR alpha, LIS_VECTOR x, LIS_VECTOR y);
LIS_INT lisd_xpay(LIS_VECTOR x, LIS_SCALAR alpha, LIS_VECTOR y);
LIS_INT lisd_axpyz(LIS_SCALAR alpha, LIS_VECTOR x, LIS_VECTOR y, LIS_VECTOR z);
LIS_INT lisd_scale(LIS_SCALAR alpha, LIS_VECTOR x);
LIS_INT lisd_set_all(LIS_SCALAR alpha, LIS_VECTOR x);
LIS_INT lisd_pmul(LIS_VECTOR x, LIS_VECTOR y, LIS_VECTOR z);
LIS_INT lisd_reduce(int kind, LIS_VECTOR x, LIS_VECTOR y, LIS_SCALAR *value);  /* syncs, allreduces */
LIS_INT lisd_matvec(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y);                 /* async */
LIS_INT lisd_matvech(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y);                /* y = A^H x, async */
void    lisd_sweep_free(void *sweep);

/* ---- triangular factors prepared for the one-launch ("sync-free") solve kernels ---- */
typedef struct lisd_perm {        /* rows in dependency-level order + the factor as SELL-32 slices in that order */
    int nslots;                   /* levels concatenated, each padded to a multiple of 32 */
    int short_rows;               /* no row keeps more than 4 entries: the narrow-batch kernel instantiation */
    int row_warp;                 /* long rows: CSR by slot (d_rptr, d_rdep, d_sidx, d_sval), a warp per row */
    int *d_rptr, *d_rdep;         /* row_warp: entries of slot k at [rptr[k], rptr[k+1]); its neighbour latest in slot order */
    int *d_order;                 /* slot -> row (or -1) */
    int *d_wptr;                  /* per warp of 32 slots: start of its slice (nslots/32 + 1 entries) */
    int *d_plen;                  /* per slot: kept entries of the row */
    int *d_wdep;                  /* per warp: the neighbour SLOT latest in slot order (-1: none) */
    int *d_sidx;                  /* slices: entry q of lane l at wptr[w] + 32*q + l; columns are slot numbers */
    double *d_sval;
    double *d_slots;              /* 2 * nslots doubles of scratch: slot-ordered results (the dependency signal) and wd */
} lisd_perm;
LIS_INT lisd_perm_build(lisd_perm *P, int n, int nlev, const int *lptr, const int *rows,
                        const LIS_INT *ptr, const LIS_INT *idx, const LIS_SCALAR *val, const int *blk_lo, const int *blk_hi);
int     lisd_perm_sweep(const lisd_perm *P, int mode, int n, const double *d_wd, const double *d_in, double *d_out, unsigned int *d_ticket);
void    lisd_perm_free(lisd_perm *p);
int    *lisd_order_by_level(int n, const int *lvl, int nlev, int *rows);   /* counting sort; returns level pointers */

typedef struct lisd_tri {         /* one strictly triangular CSR factor (lower or upper), lis_sptrsv.c */
    int n, nlev;
    lisd_perm p;
    unsigned int *d_ticket;
} lisd_tri;
/* ptr/idx/val: host CSR of the strict triangle, entries in the order the row sum must run */
LIS_INT lisd_tri_build(int n, const LIS_INT *ptr, const LIS_INT *idx, const LIS_SCALAR *val, lisd_tri **out);
void    lisd_tri_free(lisd_tri *T);
/* mode: 0 out=(in-sum)*wd, 1 out=in-sum, 2 out=in-sum over v*(out[jj]*wd[jj]); device pointers, async */
LIS_INT lisd_tri_solve(const lisd_tri *T, int mode, const double *d_wd, const double *d_in, double *d_out, const char *what);

/* ---- fused steps of the Krylov loops: one launch, one host wait, scalar(s) returned ---- */
LIS_INT lisd_jacobi_dot(LIS_VECTOR r, LIS_VECTOR dinv, LIS_VECTOR z, LIS_SCALAR *rho);        /* z=r.*dinv; <r,z> */
LIS_INT lisd_cg_update_jacobi(LIS_SCALAR alpha, LIS_VECTOR p, LIS_VECTOR q, LIS_VECTOR x, LIS_VECTOR r, LIS_VECTOR dinv, LIS_VECTOR z,
                              LIS_REAL *nrm2_r, LIS_SCALAR *rho, int *fused_out);
const struct lisb200_p2p *lisd_p2p_begin(LIS_MATRIX A, int local_ok, unsigned long long *epoch);   /* in-kernel halo exchange: table + epoch, or NULL */
int     lisd_p2p_error(void);
/* host/lis_peer.c: one block of device memory the neighbours' GPUs can store into (virtual-memory API, no peer access for
 * anything else) and the descriptor hand-over between the processes of the node */
int     lisd_peer_available(void);
int     lisd_peer_can_access(int peer_dev);
int     lisd_peer_alloc(size_t bytes, void **ptr, size_t *size, int *fd, unsigned long long *handle);
int     lisd_peer_import(int fd, size_t size, void **ptr);
void    lisd_peer_unmap(void *ptr, size_t size);
void    lisd_peer_free(void *ptr, size_t size, int fd, unsigned long long handle);
int     lisd_fd_socket(const char *job, int rank);
int     lisd_fd_send(int sock, const char *job, int to_rank, int my_rank, int fd);
int     lisd_fd_recv(int sock, int *from_rank, int *fd, int timeout_ms);
LIS_INT lisd_halo_reduce_raw(LIS_MATRIX A, double *d_y);                                         /* y[n..np) back to the owners, added in rank order */
LIS_INT lisd_matvec_dot(LIS_MATRIX A, LIS_VECTOR x, LIS_VECTOR y, LIS_SCALAR *dot_xy);        /* y=Ax; <x,y> */
LIS_INT lisd_cg_update(LIS_SCALAR alpha, LIS_VECTOR p, LIS_VECTOR q, LIS_VECTOR x, LIS_VECTOR r, LIS_REAL *nrm2_r);
LIS_INT lisd_dot2(LIS_VECTOR a, LIS_VECTOR b, LIS_SCALAR out[2]);                             /* <a,b>, <a,a> */
LIS_INT lisd_axpy_nrm2(LIS_SCALAR alpha, LIS_VECTOR x, LIS_VECTOR y, LIS_REAL *nrm2);
LIS_INT lisd_axpy_dot(LIS_SCALAR alpha, LIS_VECTOR x, LIS_VECTOR y, LIS_VECTOR u, LIS_SCALAR *dot);    /* y+=alpha x; <y,u> */         /* y+=alpha x; ||y|| */
LIS_INT lisd_bicgstab_p(LIS_SCALAR omega, LIS_SCALAR beta, LIS_VECTOR v, LIS_VECTOR r, LIS_VECTOR p);   /* p=r+beta(p-omega v) */
LIS_INT lisd_bicgstab_update(LIS_SCALAR alpha, LIS_SCALAR omega, LIS_VECTOR phat, LIS_VECTOR shat, LIS_VECTOR t,
                             LIS_VECTOR x, LIS_VECTOR r, LIS_REAL *nrm2);                      /* x,r updates; ||r|| */
/* finish a reduction whose kernel has been enqueued with result slot(s) 0..count-1:
 * wait, read the mapped scalars, combine across ranks in rank order */
LIS_INT lisd_reduce_finish(double *vals, int count, int is_max);

/* ---- process group (row-partitioned multi-GPU) ---- */
int     lisd_rank(void);
int     lisd_nranks(void);
LIS_INT lisd_comm_init(void);                       /* reads RANK/WORLD_SIZE/LOCAL_RANK */
int     lisd_reduce_uses_nccl(void);
double *lisd_reduce_dev_buffer(void);
LIS_INT lisd_reduce_nccl_finish(double *vals, int count, int is_max);
LIS_INT lisd_allreduce_sum(double *vals, int count);/* host scalars, in place, rank-ordered */
LIS_INT lisd_allreduce_max(double *vals, int count);
LIS_INT lisd_allgather_int(const int *mine, int count, int *all);
LIS_INT lisd_allgatherv_host(double *value, const LIS_INT *ranges, int nprocs);  /* in place */
LIS_INT lisd_commtable_create(LIS_MATRIX A);
LIS_INT lisd_commtable_duplicate(LIS_MATRIX Ain, LIS_MATRIX Aout);
LIS_INT lisd_matrix_g2l(LIS_MATRIX A);              /* global -> local+halo column numbering */
void    lisd_commtable_destroy(LIS_COMMTABLE t);
LIS_INT lisd_halo_exchange(LIS_MATRIX A, LIS_VECTOR x);   /* async on the stream */
LIS_INT lisd_halo_exchange_raw(LIS_MATRIX A, double *d_x); /* same on a raw device array of np entries */

#ifdef __cplusplus
}
#endif
#endif
