/*
 * lis_b200_kernels.h -- the thin C-ABI between the host C library (lis.h API) and the
 * hand-written sm_100a CUDA kernels.  Plain pointers and sizes only; every pointer named d_*
 * is a DEVICE (or managed) pointer, `stream` is a cudaStream_t passed as void*.
 *
 * Each entry point replaces one OpenMP loop of the reference (Lis 2.1.11); the reference
 * location is cited per function as  src/...:line  (relative to the reference tree).
 *
 * Arithmetic contract (what makes results bit-identical to the reference's CPU path):
 *   - IEEE fp64, multiply and add rounded separately (no FMA contraction; the reference is
 *     built -O3 without -march, configure.ac:505), division correctly rounded;
 *   - every row sum starts from +0.0 and adds products in STORAGE ORDER of the format;
 *   - dot/nrm2 are the only reassociated operations (fixed, run-to-run deterministic tree).
 *
 * All functions return 0 on success or a cudaError_t value (>0) on failure.
 * Vectors may alias exactly as the reference allows (e.g. axpy x==y is not supported there
 * either); x and y of an SpMV must not alias.
 */
#ifndef LIS_B200_KERNELS_H
#define LIS_B200_KERNELS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ------------------------------------------------------------------------- */
/* number of SMs of the current device (grid sizing), <=0 on error */
int  lisb200_sm_count(void);
const char *lisb200_error_string(int code);

/* ---- SpMV, y = A x ------------------------------------------------------------------- */
/* CSR, unsplit order.                                  src/matvec/lis_matvec_csr.c:90-110
 * rows [0,n); d_ptr has n+1 entries; d_idx/d_val must be readable up to nnz rounded up to a
 * multiple of 4 entries (the host library pads its device mirrors).                        */
int lisb200_spmv_csr(int n, const int *d_ptr, const int *d_idx, const double *d_val,
                     const double *d_x, double *d_y, void *stream);
/* CSR, unsplit order, short-row matrices: TMA-staged row blocks (cp.async.bulk of the
 * ptr/idx/val slices into shared memory by a producer warp, thread-per-row ordered walk).
 * Same result bits as lisb200_spmv_csr.  lisb200_spmv_csr_tma_plan inspects the HOST row
 * pointers once and returns 0 with (rows_per_block, tile, stages) when the matrix qualifies, 1 if
 * not (long or very ragged rows: use lisb200_spmv_csr).  d_ptr must be readable 16 bytes
 * past its n+1 entries.                                  src/matvec/lis_matvec_csr.c:90-110 */
int lisb200_spmv_csr_tma_plan(int n, const int *h_ptr, int *rows_per_block, int *tile, int *stages);
int lisb200_spmv_csr_tma(int n, int rows_per_block, int tile, int stages, const int *d_ptr, const int *d_idx,
                         const double *d_val, const double *d_x, double *d_y, void *stream);
/* ... fused with <x,y>; d_partial needs lisb200_reduce_slots() doubles (persistent grid) */
int lisb200_spmv_csr_tma_dot(int n, int rows_per_block, int tile, int stages, const int *d_ptr, const int *d_idx,
                             const double *d_val, const double *d_x, double *d_y, double *d_partial,
                             unsigned int *d_counter, double *d_result, void *stream);
/* ... on a row range (row-partitioned CG: interior rows during the halo exchange, boundary rows behind it):
 * d_ptr, d_y and d_dotx point at the first row of the range (d_dotx = d_x + first row), d_x at the whole
 * vector; the range's share of <x,y> lands in *d_result.  Concurrent launches need their own
 * d_partial / d_counter.                                                                      */
int lisb200_spmv_csr_tma_dot_rows(int n, int rows_per_block, int tile, int stages, const int *d_ptr, const int *d_idx,
                                  const double *d_val, const double *d_x, double *d_y, const double *d_dotx,
                                  double *d_partial, unsigned int *d_counter, double *d_result, void *stream);
/* Row-partitioned CSR product with the halo exchange INSIDE the kernel -- what replaces the reference's
 * LIS_MATVEC_SENDRECV (include/lis_matvec.h:31-44: lis_send_recv, src/matrix/lis_matrix_mpi.c:834-954, then the
 * local product) when every rank's GPU can map its neighbours' memory (one process per GPU, CUDA IPC, NVLink).
 * The table lives in DEVICE memory and is filled once per communication table by the host (host/lis_comm.c):
 * where each neighbour's segment of my export list lands in that neighbour's inbox, where my own inbox is, and
 * the arrival flags.  Per product: all CTAs push x[export_index[..]] into the neighbours' inboxes (buffer
 * epoch & 1), the last one to finish writes `epoch` into my flag in every neighbour; row blocks
 * [interior_lo, interior_hi) (multiples of rows_per_block; they read no column >= n) run first; before its first
 * other block a CTA waits until every neighbour's flag holds `epoch`; columns c >= n are read from
 * inbox[c - n].  epoch must grow by one per product on this table, starting at 1, on every rank alike.
 * Same products in the same order as lisb200_spmv_csr_tma: y has the same bits.  with_dot: also <x,y> as in
 * lisb200_spmv_csr_tma_dot (blocks are visited interior first, so its last bits differ from that call's). */
#define LISB200_P2P_MAX 16
typedef struct lisb200_p2p {
    int n_nbr;                                   /* neighbours (ranks I export to == ranks I import from) */
    int n_export;
    const int *export_index;                     /* device: local rows to send, neighbour segments back to back */
    int exp_start[LISB200_P2P_MAX + 1];          /* segment of neighbour s in export_index */
    int nbr_rank[LISB200_P2P_MAX];               /* rank of neighbour s */
    double *peer_inbox[LISB200_P2P_MAX];         /* mapped: where segment s starts in neighbour s's inbox, buffer 0 */
    long long peer_stride[LISB200_P2P_MAX];      /* doubles from buffer 0 to buffer 1 in that neighbour's inbox */
    unsigned long long *peer_flag[LISB200_P2P_MAX];   /* mapped: my arrival flag in neighbour s ([parity * LISB200_P2P_MAX]) */
    const double *inbox;                         /* my inbox, buffer 0: halo slot k at inbox[k] */
    long long inbox_stride;
    const unsigned long long *my_flag;           /* my flags: [parity * LISB200_P2P_MAX + sender rank] */
    unsigned int *push_count;                    /* device scratch, zero between products */
    int *error;                                  /* mapped host int: set to 1 if a neighbour's flag never arrived */
} lisb200_p2p;
int lisb200_spmv_csr_tma_p2p(int n, int rows_per_block, int tile, int stages, const int *d_ptr, const int *d_idx,
                             const double *d_val, const double *d_x, double *d_y, int with_dot, double *d_partial,
                             unsigned int *d_counter, double *d_result, const lisb200_p2p *d_table,
                             unsigned long long epoch, int interior_lo, int interior_hi, void *stream);
/* CSR, split order  t = D[i]*x[i]; t += L...; t += U...  src/matvec/lis_matvec_csr.c:64-87 */
int lisb200_spmv_csr_split(int n, const double *d_diag,
                           const int *d_lptr, const int *d_lidx, const double *d_lval,
                           const int *d_uptr, const int *d_uidx, const double *d_uval,
                           const double *d_x, double *d_y, void *stream);
/* CSR SpMV fused with the dot product <x,y> that follows it in CG (q=Ap; <p,q>).
 * d_partial: >= lisb200_spmv_csr_dot_slots(n) doubles of scratch (one per CTA); the reduced
 * scalar is written to *d_result (device or mapped-host pointer).  Same y bits as
 * lisb200_spmv_csr; the dot is a fixed (deterministic) tree over the rows, not the same
 * tree as lisb200_reduce, so the two may differ in the last bits.                          */
int lisb200_spmv_csr_dot_slots(int n);
int lisb200_spmv_csr_dot(int n, const int *d_ptr, const int *d_idx, const double *d_val,
                         const double *d_x, double *d_y, double *d_partial,
                         unsigned int *d_counter, double *d_result, void *stream);
/* ELL, column-major value[j*ld+i], index[j*ld+i], j<maxnzr  src/matvec/lis_matvec_ell.c:92-130 */
int lisb200_spmv_ell(int n, int maxnzr, int ld, const int *d_idx, const double *d_val,
                     const double *d_x, double *d_y, void *stream);
/* DIA, serial (nprocs=1) layout value[j*ld+i], offsets d_off[j] ascending
 *                                                      src/matvec/lis_matvec_dia.c:126-174
 * xlen = number of addressable x entries (n, or np in a row-partitioned matrix).          */
int lisb200_spmv_dia(int n, int xlen, int nnd, int ld, const int *d_off, const double *d_val,
                     const double *d_x, double *d_y, void *stream);
/* JAD, serial layout: d_jptr[maxnzr+1], d_perm[n] (y[perm[i]] = w[i])
 *                                                      src/matvec/lis_matvec_jad.c:144-198 */
int lisb200_spmv_jad(int n, int maxnzr, const int *d_jptr, const int *d_perm,
                     const int *d_idx, const double *d_val,
                     const double *d_x, double *d_y, void *stream);
/* BSR bnr x bnc, blocks column-major (value[bc*bnr*bnc + j*bnr + i])
 *                                      src/matvec/lis_matvec_bsr.c:57-150 and :152-858     */
int lisb200_spmv_bsr(int n, int nr, int bnr, int bnc, const int *d_bptr, const int *d_bidx,
                     const double *d_val, const double *d_x, double *d_y, void *stream);
/* ... for a row-partitioned matrix: x has ncols = n + halo entries, block columns cover [0, ncols) */
int lisb200_spmv_bsr_cols(int n, int ncols, int nr, int bnr, int bnc, const int *d_bptr,
                          const int *d_bidx, const double *d_val, const double *d_x, double *d_y, void *stream);

/* ---- BLAS-1 elementwise (bit-exact) -------------------- src/vector/lis_vector_opv.c ---- */
int lisb200_copy   (int n, const double *d_x, double *d_y, void *stream);               /* :136 */
int lisb200_axpy   (int n, double alpha, const double *d_x, double *d_y, void *stream); /* :176 y += a*x */
int lisb200_xpay   (int n, const double *d_x, double alpha, double *d_y, void *stream); /* :216 y = x + a*y */
/* axpy whose coefficient is scale * (*d_alpha), *d_alpha written by an earlier reduction on the
 * same stream (GMRES' Gram-Schmidt: t = <w,v_k>; w += (-t) v_k without a host round trip,
 * src/solver/lis_solver_gmres.c:225-232).  scale = -1 negates exactly => same bits as axpy(-t). */
int lisb200_axpy_dev(int n, const double *d_alpha, double scale, const double *d_x, double *d_y, void *stream);
int lisb200_axpyz  (int n, double alpha, const double *d_x, const double *d_y, double *d_z, void *stream); /* :256 */
int lisb200_scale  (int n, double alpha, double *d_x, void *stream);                    /* :287 */
int lisb200_pmul   (int n, const double *d_x, const double *d_y, double *d_z, void *stream); /* :328 (also Jacobi psolve, src/precon/lis_precon_jacobi.c:119-126) */
int lisb200_pdiv   (int n, const double *d_x, const double *d_y, double *d_z, void *stream); /* :368 */
int lisb200_set_all(int n, double alpha, double *d_x, void *stream);                    /* :399 */
int lisb200_abs    (int n, double *d_x, void *stream);                                  /* :430 */
int lisb200_reciprocal(int n, double *d_x, void *stream);                               /* :460 */
int lisb200_shift  (int n, double sigma, double *d_x, void *stream);                    /* :522 x -= sigma */
int lisb200_swap   (int n, double *d_x, double *d_y, void *stream);                     /* :94  */

/* ---- BLAS-1 reductions --------------------------------- src/vector/lis_vector_ops.c ---- */
/* Number of doubles of scratch a reduction needs in d_partial (per concurrently running
 * reduction), independent of n.                                                            */
int lisb200_reduce_slots(void);
/* kind: 0 dot(x,y) :58, 1 sum x*x (nrm2 before sqrt) :210, 2 nrm1 :278, 3 nrmi (max|x|) :344,
 * 4 sum :418.  The scalar lands in *d_result (device or mapped-host).  d_counter: one
 * zero-initialised unsigned int that the kernel resets before exiting.                      */
int lisb200_reduce(int kind, int n, const double *d_x, const double *d_y,
                   double *d_partial, unsigned int *d_counter, double *d_result, void *stream);
/* two dot products sharing one pass: r[0]=<a,b>, r[1]=<a,a>  (BiCGSTAB <t,s>,<t,t>,
 * src/solver/lis_solver_bicgstab.c:267-268).  Bits equal two separate lisb200_reduce calls. */
int lisb200_dot2(int n, const double *d_a, const double *d_b,
                 double *d_partial, unsigned int *d_counter, double *d_result2, void *stream);

/* ---- CG fused updates (bit-identical to the unfused call sequence) ---------------------- */
/* x += alpha*p ; r += (-alpha)*q ; rr = sum r*r          src/solver/lis_solver_cg.c:205-211 */
int lisb200_cg_update(int n, double alpha, const double *d_p, const double *d_q,
                      double *d_x, double *d_r,
                      double *d_partial, unsigned int *d_counter, double *d_rr, void *stream);
/* the same plus the Jacobi psolve and <r,z> of the NEXT iteration (src/solver/lis_solver_cg.c:171-177):
 * z = r*dinv ; d_rr_rho[0] = sum r*r ; d_rr_rho[1] = <r,z>.  Bits equal lisb200_cg_update followed by
 * lisb200_jacobi_dot.  Returns cudaErrorInvalidValue without launching when the 16-byte alignment of the
 * pointers is mixed (call the two separately then). */
int lisb200_cg_update_jacobi(int n, double alpha, const double *d_p, const double *d_q, double *d_x, double *d_r,
                             const double *d_dinv, double *d_z, double *d_partial, unsigned int *d_counter,
                             double *d_rr_rho, void *stream);
/* z = r .* dinv ; rho = <r,z>     src/solver/lis_solver_cg.c:173-177 with Jacobi psolve      */
int lisb200_jacobi_dot(int n, const double *d_r, const double *d_dinv, double *d_z,
                       double *d_partial, unsigned int *d_counter, double *d_rho, void *stream);

/* One link of the modified Gram-Schmidt chain of GMRES, src/solver/lis_solver_gmres.c:225-236:
 * w += (scale * *d_alpha) * v, then norm ? sum w*w : <w,u> into *d_result.  *d_alpha was written by
 * an earlier reduction on the same stream.  Same bits as lisb200_axpy_dev followed by
 * lisb200_reduce(0 or 1); all of v, w, u must be 16-byte aligned (else cudaErrorInvalidValue: launch
 * the two separately).  d_alpha == NULL: the coefficient is `scale` itself, i.e. axpy + norm / dot
 * in one pass (BiCGSTAB's s = r - alpha v; ||s||, src/solver/lis_solver_bicgstab.c:233-236).   */
int lisb200_mgs_step(int norm, int n, const double *d_alpha, double scale, const double *d_v, double *d_w, const double *d_u,
                     double *d_partial, unsigned int *d_counter, double *d_result, void *stream);

/* BiCGSTAB: p = r + beta*(p - omega*v)    src/solver/lis_solver_bicgstab.c:212-213 (axpy then xpay) */
int lisb200_bicgstab_p(int n, double omega, double beta, const double *d_v, const double *d_r, double *d_p, void *stream);
/* BiCGSTAB: x += alpha*phat; x += omega*shat; r += (-omega)*t; rr = sum r*r     :272-279           */
int lisb200_bicgstab_update(int n, double alpha, double omega, const double *d_phat, const double *d_shat, const double *d_t,
                            double *d_x, double *d_r, double *d_partial, unsigned int *d_counter, double *d_rr, void *stream);

/* ---- matrix helpers ---------------------------------------------------------------------- */
/* d[i] = first stored entry with index==i, else 0         src/matrix/lis_matrix_csr.c:540-553 */
int lisb200_csr_get_diagonal(int n, const int *d_ptr, const int *d_idx, const double *d_val,
                             double *d_d, void *stream);

/* ---- storage-format conversion on the device: CSR -> ELL / DIA / JAD / BSR --------------------
 * Same arrays as the reference's serial builders (and host/lis_convert.c) produce, entry for entry.
 * Tables the host needs anyway (DIA offsets, JAD pointers, BSR block pointers) are prefix-summed on
 * the host between the two launches of a conversion.                                            */
/* *d_out = longest row (ELL maxnzr, JAD maxnzr)            src/matrix/lis_matrix_ell.c:1000-1012 */
int lisb200_csr_max_row_len(int n, const int *d_ptr, int *d_out, void *stream);
/* *d_out = 1 when some row has its columns out of STRICTLY ascending order (the post-condition of lis_matrix_sort_csr,
 * src/matrix/lis_matrix_csr.c:1486-1521, for a matrix without repeated columns), else 0: lets the DIA conversion skip the
 * host sort of an input that is sorted; a repeated column takes the host sort, which decides which copy DIA keeps */
int lisb200_csr_rows_unsorted(int n, const int *d_ptr, const int *d_idx, int *d_out, void *stream);
/* ELL: d_eval[j*ld+i], d_eidx[j*ld+i], unused slots (0.0, i)  src/matrix/lis_matrix_ell.c:1035-1052 */
int lisb200_csr2ell(int n, int maxnzr, int ld, const int *d_ptr, const int *d_idx, const double *d_val,
                    int *d_eidx, double *d_eval, void *stream);
/* DIA (rows sorted by column).  mark: d_flags[col-row+n] (n+np bytes) and the number of distinct
 * offsets per segment of 1024 flags in d_seg_count[lisb200_dia_segments(n,np)].  The host sums
 * them (nnd) and passes the exclusive prefix as d_seg_base.  fill: ascending offsets into d_off,
 * d_dval[k*ld+i] for every diagonal and row (0.0 where nothing is stored).
 *                                                          src/matrix/lis_matrix_dia.c:1217-1300 */
int lisb200_dia_segments(int n, int np);
int lisb200_csr2dia_mark(int n, int np, const int *d_ptr, const int *d_idx,
                         unsigned char *d_flags, int *d_seg_count, void *stream);
int lisb200_csr2dia_fill(int n, int np, int nnd, int ld, const int *d_ptr, const int *d_idx, const double *d_val,
                         const unsigned char *d_flags, const int *d_seg_base, const int *d_seg_count,
                         int *d_off, double *d_dval, void *stream);
/* JAD, maxnzr < lisb200_jad_bins().  hist: rows per bin (bin = maxnzr - length) for each of the
 * lisb200_jad_ctas(n) CTAs, d_cta_bin[cta*bins + bin].  The host turns that into start positions
 * d_cta_base (bins in ascending order = descending row length, CTAs in order inside a bin) and
 * d_jptr[maxnzr+1].  fill: d_perm (rows by descending length, equal lengths in ascending row
 * order) and the jagged diagonals d_jidx/d_jval[jptr[j] + p] = j-th entry of row perm[p].
 *                                                          src/matrix/lis_matrix_jad.c:1682-1751 */
int lisb200_jad_ctas(int n);
int lisb200_jad_bins(void);
int lisb200_csr2jad_hist(int n, int maxnzr, const int *d_ptr, int *d_cta_bin, void *stream);
int lisb200_csr2jad_fill(int n, int maxnzr, const int *d_ptr, const int *d_idx, const double *d_val,
                         const int *d_cta_base, const int *d_jptr, int *d_perm, int *d_jidx, double *d_jval,
                         void *stream);
/* BSR bnr x bnc.  count: distinct block columns per block row in d_count[nr]; *d_overflow = 1 when a
 * block row holds more than lisb200_bsr_max_blocks() of them (convert on the host then).  The host
 * prefix-sums d_count into d_bptr.  fill: d_bidx in first-seen order, blocks column-major
 * d_bval[b*bnr*bnc + j*bnr + i], unset entries 0.0.       src/matrix/lis_matrix_bsr.c:411-540 */
int lisb200_bsr_max_blocks(void);
int lisb200_csr2bsr_count(int n, int nr, int bnr, int bnc, const int *d_ptr, const int *d_idx,
                          int *d_count, int *d_overflow, void *stream);
int lisb200_csr2bsr_fill(int n, int nr, int bnr, int bnc, const int *d_ptr, const int *d_idx, const double *d_val,
                         const int *d_bptr, int *d_bidx, double *d_bval, void *stream);

/* A <- A - sigma*I on a CSR mirror: the first stored diagonal entry of every row
 *                                                         src/matrix/lis_matrix_csr.c:565-603 */
int lisb200_csr_shift_diagonal(int n, const int *d_ptr, const int *d_idx, double *d_val, double sigma, void *stream);

/* ---- SSOR sweep (level-scheduled, block-per-"thread" like the reference's OpenMP path) --- */
/* forward:  x[i] = (b[i] - sum_{L, jj>=blk_start} L*x[jj]) * wd[i]
 * backward: x[i] -= (sum_{U, blk_start<=jj<blk_end} U*x[jj]) * wd[i]
 *                                                         src/matrix/lis_matrix_csr.c:1578-1628
 * Rows are processed level by level: d_lvl_rows lists the rows of level l in
 * [d_lvl_ptr[l], d_lvl_ptr[l+1]) (host array h_lvl_ptr drives the launches).
 * d_blk_of_row[i] gives [start,end) of the block that owns row i via d_blk_range[2*b..].     */
int lisb200_ssor_forward_level(int nrows, const int *d_rows,
                               const int *d_lptr, const int *d_lidx, const double *d_lval,
                               const double *d_wd, const int *d_rowblk_start,
                               const double *d_b, double *d_x, void *stream);
int lisb200_ssor_backward_level(int nrows, const int *d_rows,
                                const int *d_uptr, const int *d_uidx, const double *d_uval,
                                const double *d_wd, const int *d_rowblk_start,
                                const int *d_rowblk_end, double *d_x, void *stream);

/* One-launch ("sync-free") triangular sweep on a factor the host prepared once (host/lis_precon.c
 * lisd_perm_build): the rows of the strictly lower or strictly upper part listed level by level,
 * every level padded to a multiple of 32 slots with -1 (d_order, nslots entries); per warp of 32
 * slots a SELL slice -- entry q of lane l at d_wptr[w] + 32*q + l of d_sidx/d_sval, d_plen[slot]
 * entries per row, in the order the row sum must run, couplings that must be dropped already
 * removed, and the column of an entry given as the SLOT of that row -- and d_wdep[w], the slot
 * latest in slot order among everything the warp's rows read (-1: none).  d_slot_scratch: 2*nslots
 * doubles (results in slot order -- overwritten with a not-ready pattern first, this is what waiting
 * rows poll -- and, for mode 2, wd in slot order); d_ticket: one unsigned int.  d_out (n entries,
 * row order) must not alias d_in.
 *   mode 0: out[i] = (in[i] - sum v*out[jj]) * wd[i]     SSOR forward w = (D/w+L)^-1 b   src/matrix/lis_matrix_csr.c:1578-1592, 1610-1617
 *                                                        ILU's U solve                   src/precon/lis_precon_iluk.c:1040-1048
 *   mode 1: out[i] =  in[i] - sum v*out[jj]              ILU's unit-diagonal L solve     src/precon/lis_precon_iluk.c:1030-1037
 *   mode 2: out[i] =  in[i] - sum v*(out[jj]*wd[jj])     first half of the transposed SSOR sweep   src/matrix/lis_matrix_csr.c:1838-1845
 *   mode 3: out[i] =  in[i] - (sum v*out[jj]) * wd[i]    SSOR backward                   src/matrix/lis_matrix_csr.c:1593-1605, 1618-1628
 * Sums run in storage order, unfused: same result bits as the level-launched kernels and the
 * reference loops.  d_wd may be NULL for mode 1.  ctas_per_sm: low byte 1..9 bounds the persistent
 * grid (the number of rows waiting at any time; 0 = as many as fit); bit 8 set = no row has more than 4
 * entries: the narrow-batch instantiation (9 instead of 6 CTAs per SM). */
int lisb200_sweep_sell(int mode, int n, int nslots, const int *d_order, const int *d_wptr,
                       const int *d_plen, const int *d_wdep, const int *d_sidx, const double *d_sval,
                       const double *d_wd, const double *d_in, double *d_out, double *d_slot_scratch,
                       unsigned int *d_ticket, int ctas_per_sm, void *stream);

/* The same sweep for factors with LONG rows (tens of kept entries per row): a warp per row.  The factor is given as
 * CSR by slot -- entries of slot k at [d_rptr[k], d_rptr[k+1]) of d_ridx (neighbour slots) / d_rval, in the order the row
 * sum must run -- with d_rdep[k] the row's neighbour latest in slot order (-1: none).  All neighbours of a row are polled
 * at once; the products are then added in storage order (lane by lane): the bits of lisb200_sweep_sell and of the
 * reference loops.  Modes, scratch and ticket as there; nslots need not be padded to warps. */
int lisb200_sweep_rows(int mode, int n, int nslots, const int *d_order, const int *d_rptr, const int *d_rdep,
                       const int *d_ridx, const double *d_rval, const double *d_wd, const double *d_in,
                       double *d_out, double *d_slot_scratch, unsigned int *d_ticket, int ctas_per_sm, void *stream);

/* ---- halo pack (row-partitioned SpMV)                   src/matrix/lis_matrix_mpi.c:905-951 */
/* d_ws[i] = d_x[d_export_index[i]] */
int lisb200_gather(int count, const int *d_index, const double *d_x, double *d_out, void *stream);
/* the reverse step behind lis_matvech on a row-partitioned matrix (lis_reduce, src/matrix/lis_matrix_mpi.c:958-996):
 * d_y[d_index[i]] += d_src[i] for one neighbour's segment of the export list (no index twice inside a segment);
 * segments are applied one launch after the other in rank order, so the sums have a fixed order */
int lisb200_scatter_add(int count, const int *d_index, const double *d_src, double *d_y, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LIS_B200_KERNELS_H */
