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

// ---- one-launch ("sync-free") sweeps ---------------------------------------------------------
// Level-by-level launches pay a launch + drain per level (1534 levels for a 512^3 7-point grid,
// thousands for a random band).  Here the whole sweep is ONE kernel: slot k of `order` holds a
// row (levels concatenated, each padded to a multiple of 32 with -1 so that no warp straddles
// two levels => lanes of a warp never wait on each other), L/U are stored permuted in that order
// (coalesced), and a row simply waits for the "done" flag of each neighbour it reads before
// using it.  CTAs take a ticket at start and process slots in ticket order, so a waiting row's
// dependencies are always in CTAs that are already running or finished: no deadlock.  Each row
// still subtracts its products in storage order => same bits as the level-launched sweep and as
// the reference loop (src/matrix/lis_matrix_csr.c:1578-1628).
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_cg(const double *p) {      // L2-coherent read of x written by other SMs
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

template <bool kForward>
__global__ void __launch_bounds__(128)
ssor_syncfree_kernel(int nslots, const int *__restrict__ order,
                     const int *__restrict__ pptr, const int *__restrict__ pidx, const double *__restrict__ pval,
                     const double *__restrict__ wd, const int *__restrict__ blk_start, const int *__restrict__ blk_end,
                     const double *__restrict__ b, double *x, int *flag, int gen, unsigned int *ticket)
{
    __shared__ unsigned int vblock;
    if (threadIdx.x == 0) vblock = atomicAdd(ticket, 1u);
    __syncthreads();
    const int k = (int)vblock * blockDim.x + threadIdx.x;
    if (k >= nslots) return;
    const int i = order[k];
    if (i < 0) return;
    const int lo = blk_start[i], hi = blk_end[i];
    double t = kForward ? b[i] : 0.0;
    const int e = pptr[k + 1];
    for (int j = pptr[k]; j < e; ++j) {
        const int jj = pidx[j];
        if (kForward ? (jj < lo) : (jj < lo || jj >= hi)) continue;        // coupling leaves the block: dropped
        while (ld_acquire(flag + jj) != gen) { }
        const double xj = ld_cg(x + jj);
        t = kForward ? sub(t, mul(pval[j], xj)) : add(t, mul(pval[j], xj));
    }
    if (kForward) x[i] = mul(t, wd[i]);
    else x[i] = sub(ld_cg(x + i), mul(t, wd[i]));
    __threadfence();
    st_release(flag + i, gen);
}

}  // namespace lisb

using namespace lisb;

extern "C" int lisb200_ssor_sweep_syncfree(int forward, int nslots, const int *d_order,
                                           const int *d_pptr, const int *d_pidx, const double *d_pval,
                                           const double *d_wd, const int *d_rowblk_start, const int *d_rowblk_end,
                                           const double *d_b, double *d_x, int *d_flag, int gen,
                                           unsigned int *d_ticket, void *stream)
{
    if (nslots <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    const int grid = (nslots + 127) / 128;
    if (forward)
        ssor_syncfree_kernel<true><<<grid, 128, 0, st>>>(nslots, d_order, d_pptr, d_pidx, d_pval, d_wd, d_rowblk_start,
                                                        d_rowblk_end, d_b, d_x, d_flag, gen, d_ticket);
    else
        ssor_syncfree_kernel<false><<<grid, 128, 0, st>>>(nslots, d_order, d_pptr, d_pidx, d_pval, d_wd, d_rowblk_start,
                                                         d_rowblk_end, d_b, d_x, d_flag, gen, d_ticket);
    LISB_CHECK_LAUNCH();
    return 0;
}

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
