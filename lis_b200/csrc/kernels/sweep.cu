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
// thousands for a random band).  Here a whole sweep is ONE kernel.  The host lists the rows level
// by level ("slots"; each level padded to a multiple of 32 with -1 so that no warp straddles two
// levels => lanes of a warp never wait on each other) and stores the triangular factor in that
// order as SELL-32: the 32 rows of a warp form a slice, entry q of lane l sits at
// slice_start + 32*q + l, so every load of the factor is coalesced.  Couplings the block sweep
// drops (OpenMP block-SSOR) are removed on the host.  Results are published twice: in row order
// (the output vector) and in SLOT order into a scratch vector that also carries the "done" signal --
// it is pre-filled with a signalling-NaN pattern that no IEEE operation can produce, and a row reads
// a neighbour (the factor's column indices are slot numbers) by polling until the pattern is gone.
// In slot order the rows a warp waits for sit next to each other (for a stencil the neighbours of 32
// consecutive rows of a level are 32 consecutive rows of the level before): a poll of the warp is
// 8 sectors instead of 32, and the part of the vector that is live -- a few levels -- is a compact
// window that stays in L2, where the row-ordered vector of the earlier versions was scattered over
// the whole array and every first touch went to DRAM.
//
// What bounds a sweep is the dependency depth times the time of one hop (publish -> L2 -> poll).
// Round 1's kernel let every waiting thread poll all of its neighbours: at 256^3, 22 GB of L2
// traffic for 1.3 GB of DRAM traffic, L2 at 71 % of peak and 3 us per hop
// (profiles/r02_ncu_ssor_v1.txt).  Now
//   * a row polls its own neighbours once (all polls of a batch in flight together; this also pulls
//     their sectors into L2 while the row is still levels ahead of the sweep front), then the warp
//     waits on ONE address -- the neighbour of its 32 rows that sits latest in slot order (found
//     by the host), all lanes the same sector -- and only then collects what was missing;
//   * everything that does not depend on a neighbour (slot -> row, row length, first batch of
//     the factor, in[i], wd[i]) is loaded before the first poll;
//   * the grid is persistent (CTAs take tickets of 128 slots in slot order), which bounds the
//     number of waiting warps; a waiting row's dependencies are always in CTAs that hold an
//     earlier ticket, i.e. are running or finished: no deadlock.
// Each row still subtracts its products in storage order => same bits as the level-launched
// sweep and as the reference loop (src/matrix/lis_matrix_csr.c:1578-1628).
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
//   kSweepPlain out[i] = in[i] - sum v*out[jj]                  unit-diagonal solve (ILU's L solve)
//   kSweepReadScaled out[i] = in[i] - sum v*(out[jj]*wd[jj])    first half of the transposed SSOR sweep
//   kSweepBwd   out[i] = in[i] - (sum v*out[jj]) * wd[i]        SSOR backward
enum { kSweepFwd = 0, kSweepPlain = 1, kSweepReadScaled = 2, kSweepBwd = 3 };
constexpr int kSweepThreads = 128;

template <int kMode>
__global__ void __launch_bounds__(kSweepThreads, 6)
sweep_sell_kernel(int nslots, const int *__restrict__ order, const int *__restrict__ wptr,
                  const int *__restrict__ plen, const int *__restrict__ wdep,
                  const int *__restrict__ sidx, const double *__restrict__ sval,
                  const double *__restrict__ wd, const double *__restrict__ wds /* kSweepReadScaled: wd in slot order */,
                  const double *__restrict__ in, double *__restrict__ out, double *pout, unsigned int *ticket)
{
    __shared__ unsigned int vblock[2];
    constexpr bool kSub = kMode != kSweepBwd;       // running value starts at in[i] and products are subtracted
    constexpr int kBatch = 8;
    if (threadIdx.x == 0) vblock[0] = atomicAdd(ticket, 1u);
    __syncthreads();
    for (int round = 0;; round ^= 1) {
        const long long k0 = (long long)vblock[round] * kSweepThreads;
        if (k0 >= nslots) break;
        // the next ticket is on its way while this block's rows are worked on
        unsigned int next_ticket = 0;
        if (threadIdx.x == 0) next_ticket = atomicAdd(ticket, 1u);
        const int k = (int)k0 + threadIdx.x;
        if (k < nslots) {                            // nslots is a multiple of 32: whole warps
            // first wave: nothing here depends on another load
            const int w = k >> 5, lane = k & 31;
            const int i = order[k];                  // -1: padding lane
            const int len = plen[k];
            const int dep = wdep[w];
            const size_t base = (size_t)wptr[w] + lane;
            if (i >= 0) {
                // second wave: the row's operands and its first batch of the factor
                const int *ci = sidx + base;
                const double *cv = sval + base;
                const double inv = in[i];
                const double wdv = (kMode == kSweepFwd || kMode == kSweepBwd) ? wd[i] : 0.0;
                int jj[kBatch];
                double v[kBatch];
#pragma unroll
                for (int q = 0; q < kBatch; ++q) {
                    const int qq = q < len ? q : 0;
                    jj[q] = len > 0 ? ci[32 * qq] : 0;
                    v[q] = len > 0 ? cv[32 * qq] : 0.0;
                }
                // long rows: touch the neighbours behind the first batch too (sectors into L2, values unused)
                for (int q = kBatch; q < len; ++q) (void)ld_poll(pout + ci[32 * (size_t)q]);
                double t = kSub ? inv : 0.0;
                bool waited = false;
                for (int q0 = 0; q0 < len; q0 += kBatch) {
                    double xv[kBatch];
                    unsigned int used = 0;
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) { if (q0 + q < len) used |= 1u << q; xv[q] = 0.0; }
                    unsigned int pending = used;
                    while (pending) {
                        // all polls of the batch are issued before the first answer is looked at: one
                        // round trip per batch, not one per neighbour.  The first round also brings the
                        // neighbours' sectors into L2 long before their values are published (this row
                        // is several levels ahead of the sweep front), so the round after the wait
                        // below is an L2 hit and not a DRAM fill.
                        unsigned long long bits[kBatch];
#pragma unroll
                        for (int q = 0; q < kBatch; ++q) bits[q] = (pending & (1u << q)) ? ld_poll(pout + jj[q]) : kNotReady;
#pragma unroll
                        for (int q = 0; q < kBatch; ++q)
                            if ((pending & (1u << q)) && bits[q] != kNotReady) { xv[q] = __longlong_as_double((long long)bits[q]); pending &= ~(1u << q); }
                        if (pending && !waited) {
                            // wait on ONE address for the whole warp -- the neighbour of its 32 rows that
                            // sits latest in slot order -- instead of every lane polling all of its own
                            waited = true;
                            if (dep >= 0) while (ld_poll(pout + dep) == kNotReady) { }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < kBatch; ++q)
                        if (used & (1u << q)) {
                            if (kMode == kSweepReadScaled) xv[q] = mul(xv[q], wds[jj[q]]);
                            t = kSub ? sub(t, mul(v[q], xv[q])) : add(t, mul(v[q], xv[q]));
                        }
                    if (q0 + kBatch < len) {
#pragma unroll
                        for (int q = 0; q < kBatch; ++q) {
                            const int qq = q0 + kBatch + q < len ? q0 + kBatch + q : q0 + kBatch;
                            jj[q] = ci[32 * (size_t)qq];
                            v[q] = cv[32 * (size_t)qq];
                        }
                    }
                }
                const double r = kMode == kSweepFwd ? mul(t, wdv) : kMode == kSweepBwd ? sub(inv, mul(t, wdv)) : t;
                st_publish(pout + k, r);             // slot order: what the waiting rows poll (contiguous per warp)
                out[i] = r;                          // row order: the result
            }
        }
        if (threadIdx.x == 0) vblock[round ^ 1] = next_ticket;
        __syncthreads();
    }
}

/* slot-ordered copy of a row-ordered vector (wd for the read-scaled mode) */
__global__ void __launch_bounds__(256)
gather_slots_kernel(int nslots, const int *__restrict__ order, const double *__restrict__ src, double *__restrict__ dst)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nslots) { const int i = order[k]; dst[k] = i >= 0 ? src[i] : 0.0; }
}

}  // namespace lisb
using namespace lisb;

extern "C" int lisb200_sweep_sell(int mode, int n, int nslots, const int *d_order, const int *d_wptr,
                                  const int *d_plen, const int *d_wdep, const int *d_sidx, const double *d_sval,
                                  const double *d_wd, const double *d_in, double *d_out, double *d_slot_scratch,
                                  unsigned int *d_ticket, int ctas_per_sm, void *stream)
{
    if (nslots <= 0 || n <= 0) return 0;
    if (mode < 0 || mode > 3 || (mode != kSweepPlain && d_wd == nullptr) || (nslots & 31) || d_slot_scratch == nullptr) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    static int sms = 0;
    if (sms <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    double *pout = d_slot_scratch, *wds = d_slot_scratch + nslots;      // scratch: 2 * nslots doubles
    int fill_grid = (nslots + 255) / 256;
    if (fill_grid > sms * 8) fill_grid = sms * 8;
    fill_not_ready_kernel<<<fill_grid, 256, 0, st>>>(nslots, pout);
    if (mode == kSweepReadScaled) gather_slots_kernel<<<(nslots + 255) / 256, 256, 0, st>>>(nslots, d_order, d_wd, wds);
    if (ctas_per_sm < 1 || ctas_per_sm > 6) ctas_per_sm = 6;
    int grid = (nslots + kSweepThreads - 1) / kSweepThreads;
    if (grid > sms * ctas_per_sm) grid = sms * ctas_per_sm;
    switch (mode) {
    case kSweepFwd: sweep_sell_kernel<kSweepFwd><<<grid, kSweepThreads, 0, st>>>(nslots, d_order, d_wptr, d_plen, d_wdep, d_sidx, d_sval, d_wd, wds, d_in, d_out, pout, d_ticket); break;
    case kSweepPlain: sweep_sell_kernel<kSweepPlain><<<grid, kSweepThreads, 0, st>>>(nslots, d_order, d_wptr, d_plen, d_wdep, d_sidx, d_sval, d_wd, wds, d_in, d_out, pout, d_ticket); break;
    case kSweepReadScaled: sweep_sell_kernel<kSweepReadScaled><<<grid, kSweepThreads, 0, st>>>(nslots, d_order, d_wptr, d_plen, d_wdep, d_sidx, d_sval, d_wd, wds, d_in, d_out, pout, d_ticket); break;
    default: sweep_sell_kernel<kSweepBwd><<<grid, kSweepThreads, 0, st>>>(nslots, d_order, d_wptr, d_plen, d_wdep, d_sidx, d_sval, d_wd, wds, d_in, d_out, pout, d_ticket); break;
    }
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
