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
// thousands for a random band).  Here a whole sweep is ONE kernel: slot k of `order` holds a
// row (levels concatenated, each padded to a multiple of 32 with -1 so that no warp straddles
// two levels => lanes of a warp never wait on each other), L/U are stored permuted in that order
// (coalesced), and the output vector itself carries the "done" signal: it is pre-filled with a
// signalling-NaN pattern that no IEEE operation can produce, and a row polls each neighbour it
// reads until the pattern is gone (one L2 round trip per dependency hop, 8 neighbours polled
// at a time).  CTAs take a ticket at start and process slots in ticket order, so a waiting
// row's dependencies are always in CTAs that are already running or finished: no deadlock.
// Each row still subtracts its products in storage order => same bits as the level-launched
// sweep and as the reference loop (src/matrix/lis_matrix_csr.c:1578-1628).
//   forward : w[i] = (b[i] - sum_{L, in block} L*w[jj]) * wd[i]          (w pre-filled)
//   backward: x[i] = w[i] - (sum_{U, in block} U*x[jj]) * wd[i]          (x pre-filled)
constexpr unsigned long long kNotReady = 0x7ff4c0dedeadbeefull;     // sNaN payload: never a result

#ifndef LISB_EMU
__device__ __forceinline__ unsigned long long ld_poll(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_publish(double *p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}
#endif

__global__ void __launch_bounds__(256)
fill_not_ready_kernel(int n, double *x)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        reinterpret_cast<unsigned long long *>(x)[i] = kNotReady;
}

// kMode selects the row formula (all sums in storage order, unfused):
//   kSweepFwd   out[i] = (in[i] - sum v*out[jj]) * wd[i]        SSOR forward, (D/w+L)^-1, ILU's U solve
//   kSweepBwd   out[i] = in[i] - (sum v*out[jj]) * wd[i]        SSOR backward
//   kSweepPlain out[i] = in[i] - sum v*out[jj]                  unit-diagonal solve (ILU's L solve)
//   kSweepReadScaled out[i] = in[i] - sum v*(out[jj]*wd[jj])    first half of the transposed SSOR sweep
// kMasked: drop couplings outside the row's block (block-SSOR of the OpenMP reference); the
// unmasked instantiations take a triangular factor that was already filtered on the host.
enum { kSweepFwd = 0, kSweepBwd = 1, kSweepPlain = 2, kSweepReadScaled = 3 };

template <int kMode, bool kMasked>
__global__ void __launch_bounds__(128)
ssor_syncfree_kernel(int nslots, const int *__restrict__ order,
                     const int *__restrict__ pptr, const int *__restrict__ pidx, const double *__restrict__ pval,
                     const double *__restrict__ wd, const int *__restrict__ blk_start, const int *__restrict__ blk_end,
                     const double *__restrict__ in /* b (forward) or w (backward) */, double *out, unsigned int *ticket)
{
    __shared__ unsigned int vblock;
    if (threadIdx.x == 0) vblock = atomicAdd(ticket, 1u);
    __syncthreads();
    const int k = (int)vblock * blockDim.x + threadIdx.x;
    if (k >= nslots) return;
    const int i = order[k];
    if (i < 0) return;
    constexpr bool kForward = kMode != kSweepBwd;
    const int lo = kMasked ? blk_start[i] : 0, hi = kMasked ? blk_end[i] : 0x7fffffff;
    double t = kForward ? in[i] : 0.0;
    const int e = pptr[k + 1];
    constexpr int kBatch = 8;
    for (int j0 = pptr[k]; j0 < e; j0 += kBatch) {
        int jj[kBatch];
        double v[kBatch], xv[kBatch];
        unsigned int pending = 0, used = 0;
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
            const int j = min(j0 + q, e - 1);
            jj[q] = pidx[j];
            v[q] = pval[j];
            const bool inblk = !kMasked || (kForward ? (jj[q] >= lo) : (jj[q] >= lo && jj[q] < hi));   // else: coupling dropped
            if (j0 + q < e && inblk) used |= 1u << q;
            xv[q] = 0.0;
        }
        pending = used;
        while (pending) {
#pragma unroll
            for (int q = 0; q < kBatch; ++q)
                if (pending & (1u << q)) {
                    const unsigned long long bits = ld_poll(out + jj[q]);
                    if (bits != kNotReady) { xv[q] = __longlong_as_double((long long)bits); pending &= ~(1u << q); }
                }
        }
#pragma unroll
        for (int q = 0; q < kBatch; ++q)
            if (used & (1u << q)) {
                if (kMode == kSweepReadScaled) xv[q] = mul(xv[q], wd[jj[q]]);
                t = kForward ? sub(t, mul(v[q], xv[q])) : add(t, mul(v[q], xv[q]));
            }
    }
    st_publish(out + i, kMode == kSweepFwd ? mul(t, wd[i]) : kMode == kSweepBwd ? sub(in[i], mul(t, wd[i])) : t);
}

}  // namespace lisb

using namespace lisb;

extern "C" int lisb200_ssor_sweep_syncfree(int forward, int n, int nslots, const int *d_order,
                                           const int *d_pptr, const int *d_pidx, const double *d_pval,
                                           const double *d_wd, const int *d_rowblk_start, const int *d_rowblk_end,
                                           const double *d_in, double *d_out, unsigned int *d_ticket, void *stream)
{
    if (nslots <= 0 || n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    int fill_grid = (n + 255) / 256;
    if (fill_grid > 148 * 8) fill_grid = 148 * 8;
    fill_not_ready_kernel<<<fill_grid, 256, 0, st>>>(n, d_out);
    const int grid = (nslots + 127) / 128;
    if (forward)
        ssor_syncfree_kernel<kSweepFwd, true><<<grid, 128, 0, st>>>(nslots, d_order, d_pptr, d_pidx, d_pval, d_wd, d_rowblk_start,
                                                        d_rowblk_end, d_in, d_out, d_ticket);
    else
        ssor_syncfree_kernel<kSweepBwd, true><<<grid, 128, 0, st>>>(nslots, d_order, d_pptr, d_pidx, d_pval, d_wd, d_rowblk_start,
                                                         d_rowblk_end, d_in, d_out, d_ticket);
    LISB_CHECK_LAUNCH();
    return 0;
}

/* One-launch triangular solve on a host-prepared factor (levels concatenated, padded to warps,
 * entries permuted into slot order): mode 0 = scaled, 1 = plain, 2 = neighbours scaled on read. */
extern "C" int lisb200_sptrsv_syncfree(int mode, int n, int nslots, const int *d_order,
                                       const int *d_pptr, const int *d_pidx, const double *d_pval,
                                       const double *d_wd, const double *d_in, double *d_out,
                                       unsigned int *d_ticket, void *stream)
{
    if (nslots <= 0 || n <= 0) return 0;
    if (mode < 0 || mode > 2 || (mode != 1 && d_wd == nullptr)) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    int fill_grid = (n + 255) / 256;
    if (fill_grid > 148 * 8) fill_grid = 148 * 8;
    fill_not_ready_kernel<<<fill_grid, 256, 0, st>>>(n, d_out);
    const int grid = (nslots + 127) / 128;
    if (mode == 0)
        ssor_syncfree_kernel<kSweepFwd, false><<<grid, 128, 0, st>>>(nslots, d_order, d_pptr, d_pidx, d_pval, d_wd, nullptr,
                                                                     nullptr, d_in, d_out, d_ticket);
    else if (mode == 1)
        ssor_syncfree_kernel<kSweepPlain, false><<<grid, 128, 0, st>>>(nslots, d_order, d_pptr, d_pidx, d_pval, nullptr, nullptr,
                                                                       nullptr, d_in, d_out, d_ticket);
    else
        ssor_syncfree_kernel<kSweepReadScaled, false><<<grid, 128, 0, st>>>(nslots, d_order, d_pptr, d_pidx, d_pval, d_wd, nullptr,
                                                                            nullptr, d_in, d_out, d_ticket);
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
