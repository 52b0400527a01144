// sweep.cu -- the SSOR preconditioner sweep of lis_matrix_solve_csr(..., LIS_MATRIX_SSOR)
// (src/matrix/lis_matrix_csr.c:1572-1630 of the reference).
//
// The reference's OpenMP path is block-SSOR: thread k owns rows [is_k, ie_k) and drops every
// coupling that leaves its block.  Inside a block the sweep is a sequential dependency chain;
// on the GPU the rows of all blocks are level-scheduled (host/lis_precon.c builds the levels):
// rows of one level are mutually independent, so a level is one launch with a thread per row,
// and every row still subtracts its products in storage order => bit-identical to the CPU
// sweep with the same block partition.
#include "common.cuh"
#include "../../../include/lis_b200_kernels.h"

namespace lisb {

// forward: t = b[i]; for L entries (storage order) with jj >= blk_start: t -= L*x[jj];
//          x[i] = t * wd[i]
__global__ void __launch_bounds__(128)
ssor_fwd_kernel(int nrows, const int *__restrict__ rows,
                const int *__restrict__ lptr, const int *__restrict__ lidx, const double *__restrict__ lval,
                const double *__restrict__ wd, const int *__restrict__ blk_start,
                const double *__restrict__ b, double *x)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrows) return;
    const int i = rows[k];
    const int lo = blk_start[i];
    double t = b[i];
    const int e = lptr[i + 1];
    for (int j = lptr[i]; j < e; ++j) {
        const int jj = lidx[j];
        if (jj < lo) continue;
        t = sub(t, mul(lval[j], x[jj]));
    }
    x[i] = mul(t, wd[i]);
}

// backward: t = 0; for U entries with blk_start <= jj < blk_end: t += U*x[jj];
//           x[i] -= t * wd[i]
__global__ void __launch_bounds__(128)
ssor_bwd_kernel(int nrows, const int *__restrict__ rows,
                const int *__restrict__ uptr, const int *__restrict__ uidx, const double *__restrict__ uval,
                const double *__restrict__ wd, const int *__restrict__ blk_start,
                const int *__restrict__ blk_end, double *x)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrows) return;
    const int i = rows[k];
    const int lo = blk_start[i], hi = blk_end[i];
    double t = 0.0;
    const int e = uptr[i + 1];
    for (int j = uptr[i]; j < e; ++j) {
        const int jj = uidx[j];
        if (jj < lo || jj >= hi) continue;
        t = add(t, mul(uval[j], x[jj]));
    }
    x[i] = sub(x[i], mul(t, wd[i]));
}

}  // namespace lisb

using namespace lisb;

extern "C" int lisb200_ssor_forward_level(int nrows, const int *d_rows,
                                          const int *d_lptr, const int *d_lidx, const double *d_lval,
                                          const double *d_wd, const int *d_rowblk_start,
                                          const double *d_b, double *d_x, void *stream)
{
    if (nrows <= 0) return 0;
    ssor_fwd_kernel<<<(nrows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        nrows, d_rows, d_lptr, d_lidx, d_lval, d_wd, d_rowblk_start, d_b, d_x);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_ssor_backward_level(int nrows, const int *d_rows,
                                           const int *d_uptr, const int *d_uidx, const double *d_uval,
                                           const double *d_wd, const int *d_rowblk_start,
                                           const int *d_rowblk_end, double *d_x, void *stream)
{
    if (nrows <= 0) return 0;
    ssor_bwd_kernel<<<(nrows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        nrows, d_rows, d_uptr, d_uidx, d_uval, d_wd, d_rowblk_start, d_rowblk_end, d_x);
    LISB_CHECK_LAUNCH();
    return 0;
}
