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

// kBatch: neighbours polled together (registers: 8 -> 80, 4 -> ~56; the host picks 4 for factors whose
// rows are short -- stencils -- where the extra resident CTAs matter more than the batch width).
template <int kMode, int kBatch>
__global__ void __launch_bounds__(kSweepThreads, kBatch == 4 ? 9 : 6)
sweep_sell_kernel(int nslots, const int *__restrict__ order, const int *__restrict__ wptr,
                  const int *__restrict__ plen, const int *__restrict__ wdep,
                  const int *__restrict__ sidx, const double *__restrict__ sval,
                  const double *__restrict__ wd, const double *__restrict__ wds /* kSweepReadScaled: wd in slot order */,
                  const double *__restrict__ in, double *__restrict__ out, double *pout, unsigned int *ticket)
{
    // Tickets (128 slots each, in slot order) are drawn two rounds ahead, so that every thread knows the
    // NEXT block at the start of a round and its first wave of loads (slot -> row, row length, the warp's
    // latest neighbour, slice start) is in flight while this round waits for its neighbours.
    __shared__ unsigned int vblock[3];
    constexpr bool kSub = kMode != kSweepBwd;       // running value starts at in[i] and products are subtracted
    if (threadIdx.x == 0) { vblock[0] = atomicAdd(ticket, 1u); vblock[1] = atomicAdd(ticket, 1u); }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int i = -1, len = 0, dep = -1, base = 0, k = 0;
    {
        const long long k0 = (long long)vblock[0] * kSweepThreads + threadIdx.x;
        if (k0 < nslots) { k = (int)k0; i = order[k]; len = plen[k]; dep = wdep[k >> 5]; base = wptr[k >> 5]; }
    }
    for (int round = 0;; round = round == 2 ? 0 : round + 1) {
        if ((long long)vblock[round] * kSweepThreads >= nslots) break;
        const int rnext = round == 2 ? 0 : round + 1, rnext2 = rnext == 2 ? 0 : rnext + 1;
        unsigned int ticket2 = 0;
        if (threadIdx.x == 0) ticket2 = atomicAdd(ticket, 1u);          // for the round after the next
        // second wave of this round: the row's operands and its first batch of the factor
        const int *ci = sidx + (size_t)base + lane;
        const double *cv = sval + (size_t)base + lane;
        double inv = 0.0, wdv = 0.0;
        int jj[kBatch];
        double v[kBatch];
        if (i >= 0) {
            inv = in[i];
            if (kMode == kSweepFwd || kMode == kSweepBwd) wdv = wd[i];
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
                const int qq = q < len ? q : 0;
                jj[q] = len > 0 ? ci[32 * qq] : 0;
                v[q] = len > 0 ? cv[32 * qq] : 0.0;
            }
        }
        // first wave of the NEXT round
        int ni = -1, nlen = 0, ndep = -1, nbase = 0, nk = 0;
        {
            const long long k0 = (long long)vblock[rnext] * kSweepThreads + threadIdx.x;
            if (k0 < nslots) { nk = (int)k0; ni = order[nk]; nlen = plen[nk]; ndep = wdep[nk >> 5]; nbase = wptr[nk >> 5]; }
        }
        if (i >= 0) {                                // -1: padding lane (or past the end)
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
                    // is several levels ahead of the sweep front).
                    unsigned long long bits[kBatch];
#pragma unroll
                    for (int q = 0; q < kBatch; ++q) bits[q] = (pending & (1u << q)) ? ld_poll(pout + jj[q]) : kNotReady;
#pragma unroll
                    for (int q = 0; q < kBatch; ++q)
                        if ((pending & (1u << q)) && bits[q] != kNotReady) { xv[q] = __longlong_as_double((long long)bits[q]); pending &= ~(1u << q); }
                    if (pending && !waited) {
                        // wait on ONE address for the whole warp -- the slot latest in slot order among
                        // everything its 32 rows read -- instead of every lane polling all of its own
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
            st_publish(pout + k, r);                 // slot order: what the waiting rows poll (contiguous per warp)
            out[i] = r;                              // row order: the result
        }
        if (threadIdx.x == 0) vblock[rnext2] = ticket2;
        i = ni; len = nlen; dep = ndep; base = nbase; k = nk;
        __syncthreads();
    }
}

// ---- long rows: a WARP per row ------------------------------------------------------------------
// With tens of kept entries per row (the 70-per-row banded matrix of BASELINE config 4, ILU factors with
// fill) a thread per row needs several poll batches one after the other, and a level holds only a few
// thousand rows.  Here the 32 lanes of a warp hold the row's entries (the factor stays in CSR order by
// slot: coalesced), all neighbours are polled at once, and the products are then added in STORAGE order
// by walking the lanes with shuffles -- the same sequence of rounded operations as the thread-per-row
// kernel and the reference loop.  A CTA takes tickets of 32 slots, its 4 warps take every fourth slot.
constexpr int kRowWarpSlots = 32;

template <int kMode>
__global__ void __launch_bounds__(kSweepThreads, 8)
sweep_rowwarp_kernel(int nslots, const int *__restrict__ order, const int *__restrict__ rptr, const int *__restrict__ rdep,
                     const int *__restrict__ ridx, const double *__restrict__ rval,
                     const double *__restrict__ wd, const double *__restrict__ wds,
                     const double *__restrict__ in, double *__restrict__ out, double *pout, unsigned int *ticket)
{
    __shared__ unsigned int vblock[2];
    constexpr bool kSub = kMode != kSweepBwd;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) vblock[0] = atomicAdd(ticket, 1u);
    __syncthreads();
    for (int round = 0;; round ^= 1) {
        const long long k0 = (long long)vblock[round] * kRowWarpSlots;
        if (k0 >= nslots) break;
        unsigned int next_ticket = 0;
        if (threadIdx.x == 0) next_ticket = atomicAdd(ticket, 1u);
        for (int k = (int)k0 + warp; k < (int)k0 + kRowWarpSlots && k < nslots; k += kSweepThreads / 32) {
            const int i = order[k];
            if (i < 0) continue;                                   // padding slot (uniform in the warp)
            const int s = rptr[k], e = rptr[k + 1], dep = rdep[k];
            const double inv = in[i];
            const double wdv = (kMode == kSweepFwd || kMode == kSweepBwd) ? wd[i] : 0.0;
            double t = kSub ? inv : 0.0;
            bool waited = false;
            for (int c = s; c < e; c += 64) {                      // two entries per lane per chunk
                const int j0 = c + lane, j1 = c + 32 + lane;
                const bool u0 = j0 < e, u1 = j1 < e;
                const int d0 = u0 ? ridx[j0] : 0, d1 = u1 ? ridx[j1] : 0;
                const double v0 = u0 ? rval[j0] : 0.0, v1 = u1 ? rval[j1] : 0.0;
                double x0 = 0.0, x1 = 0.0;
                bool p0 = u0, p1 = u1;
                for (;;) {
                    const unsigned long long b0 = p0 ? ld_poll(pout + d0) : kNotReady, b1 = p1 ? ld_poll(pout + d1) : kNotReady;
                    if (p0 && b0 != kNotReady) { x0 = __longlong_as_double((long long)b0); p0 = false; }
                    if (p1 && b1 != kNotReady) { x1 = __longlong_as_double((long long)b1); p1 = false; }
                    if (!__any_sync(0xffffffffu, p0 || p1)) break;
                    if (!waited) {
                        // the whole warp waits on ONE address: the row's neighbour latest in slot order
                        waited = true;
                        if (dep >= 0) while (ld_poll(pout + dep) == kNotReady) { }
                    }
                }
                if (kMode == kSweepReadScaled) { if (u0) x0 = mul(x0, wds[d0]); if (u1) x1 = mul(x1, wds[d1]); }
                const double q0 = mul(v0, x0), q1 = mul(v1, x1);
                const int cnt = min(64, e - c);
                for (int q = 0; q < cnt; ++q) {                    // storage order: lane q of the first half, then the second
                    const double pq = __shfl_sync(0xffffffffu, q < 32 ? q0 : q1, q & 31);
                    t = kSub ? sub(t, pq) : add(t, pq);
                }
            }
            if (lane == 0) {
                const double r = kMode == kSweepFwd ? mul(t, wdv) : kMode == kSweepBwd ? sub(inv, mul(t, wdv)) : t;
                st_publish(pout + k, r);
                out[i] = r;
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
    // ctas_per_sm: bit 8 set = the factor's rows are short (<= 4 kept entries): the narrow-batch instantiation,
    // which fits 9 CTAs per SM instead of 6
    const bool short_rows = (ctas_per_sm & 0x100) != 0;
    ctas_per_sm &= 0xff;
    const int cap = short_rows ? 9 : 6;
    if (ctas_per_sm < 1 || ctas_per_sm > cap) ctas_per_sm = cap;
    int grid = (nslots + kSweepThreads - 1) / kSweepThreads;
    if (grid > sms * ctas_per_sm) grid = sms * ctas_per_sm;
#define LISB_SWEEP_LAUNCH(M, B) sweep_sell_kernel<M, B><<<grid, kSweepThreads, 0, st>>>(nslots, d_order, d_wptr, d_plen, d_wdep, d_sidx, d_sval, d_wd, wds, d_in, d_out, pout, d_ticket)
    if (short_rows) {
        switch (mode) {
        case kSweepFwd: LISB_SWEEP_LAUNCH(kSweepFwd, 4); break;
        case kSweepPlain: LISB_SWEEP_LAUNCH(kSweepPlain, 4); break;
        case kSweepReadScaled: LISB_SWEEP_LAUNCH(kSweepReadScaled, 4); break;
        default: LISB_SWEEP_LAUNCH(kSweepBwd, 4); break;
        }
    } else {
        switch (mode) {
        case kSweepFwd: LISB_SWEEP_LAUNCH(kSweepFwd, 8); break;
        case kSweepPlain: LISB_SWEEP_LAUNCH(kSweepPlain, 8); break;
        case kSweepReadScaled: LISB_SWEEP_LAUNCH(kSweepReadScaled, 8); break;
        default: LISB_SWEEP_LAUNCH(kSweepBwd, 8); break;
        }
    }
#undef LISB_SWEEP_LAUNCH
    LISB_CHECK_LAUNCH();
    return 0;
}

/* the same sweep for factors with long rows: a warp per row; see include/lis_b200_kernels.h */
extern "C" int lisb200_sweep_rows(int mode, int n, int nslots, const int *d_order, const int *d_rptr, const int *d_rdep,
                                  const int *d_ridx, const double *d_rval, const double *d_wd, const double *d_in,
                                  double *d_out, double *d_slot_scratch, unsigned int *d_ticket, int ctas_per_sm, void *stream)
{
    if (nslots <= 0 || n <= 0) return 0;
    if (mode < 0 || mode > 3 || (mode != kSweepPlain && d_wd == nullptr) || d_slot_scratch == nullptr) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    static int sms = 0;
    if (sms <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    double *pout = d_slot_scratch, *wds = d_slot_scratch + nslots;
    int fill_grid = (nslots + 255) / 256;
    if (fill_grid > sms * 8) fill_grid = sms * 8;
    fill_not_ready_kernel<<<fill_grid, 256, 0, st>>>(nslots, pout);
    if (mode == kSweepReadScaled) gather_slots_kernel<<<(nslots + 255) / 256, 256, 0, st>>>(nslots, d_order, d_wd, wds);
    ctas_per_sm &= 0xff;
    if (ctas_per_sm < 1 || ctas_per_sm > 8) ctas_per_sm = 8;
    int grid = (nslots + kRowWarpSlots - 1) / kRowWarpSlots;
    if (grid > sms * ctas_per_sm) grid = sms * ctas_per_sm;
    switch (mode) {
    case kSweepFwd: sweep_rowwarp_kernel<kSweepFwd><<<grid, kSweepThreads, 0, st>>>(nslots, d_order, d_rptr, d_rdep, d_ridx, d_rval, d_wd, wds, d_in, d_out, pout, d_ticket); break;
    case kSweepPlain: sweep_rowwarp_kernel<kSweepPlain><<<grid, kSweepThreads, 0, st>>>(nslots, d_order, d_rptr, d_rdep, d_ridx, d_rval, d_wd, wds, d_in, d_out, pout, d_ticket); break;
    case kSweepReadScaled: sweep_rowwarp_kernel<kSweepReadScaled><<<grid, kSweepThreads, 0, st>>>(nslots, d_order, d_rptr, d_rdep, d_ridx, d_rval, d_wd, wds, d_in, d_out, pout, d_ticket); break;
    default: sweep_rowwarp_kernel<kSweepBwd><<<grid, kSweepThreads, 0, st>>>(nslots, d_order, d_rptr, d_rdep, d_ridx, d_rval, d_wd, wds, d_in, d_out, pout, d_ticket); break;
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
