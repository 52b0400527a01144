// spmv.cu -- y = A x for the six storage formats of the hot path (CSR, ELL, DIA, JAD, BSR;
// CSC is served by the CSR kernel on a transposed device mirror, see host/lis_matrix_dev.c).
//
// Design (B200, HBM-bound, fp64, int32 indices):
//  * CSR  : "row-block / product-tile" kernel.  A CTA owns kRows consecutive rows, i.e. one
//           CONTIGUOUS slice of idx[]/val[].  Phase 1 streams that slice with 128-bit coalesced
//           loads (4 idx + 4 val per thread per step), gathers x through the read-only path and
//           parks the rounded products in shared memory.  Phase 2 lets thread r walk row r's
//           products in storage order.  The expensive part (HBM stream + gather) is perfectly
//           coalesced and load-balanced regardless of row length; the ordered sum is what makes
//           the result bit-identical to the reference loop (src/matvec/lis_matvec_csr.c:98-109).
//           Rows longer than a tile simply span several tiles with the accumulator kept in a
//           register.
//  * ELL/DIA/JAD : column-major sweeps, thread per row, every load coalesced.
//  * BSR  : thread per block row, template on (bnr,bnc) for the 4x4 table of the reference.
#include "common.cuh"
#include "../../../include/lis_b200_kernels.h"

namespace lisb {

constexpr int kCsrThreads = 256;          // threads per CTA == rows per CTA
constexpr int kCsrTile    = 2048;         // products per shared-memory tile (16 KB)

// Accumulate, for row `r` of this CTA's row block [r0, rend), the products of one CSR piece
// (ptr/idx/val) in storage order into `acc`.  All threads of the CTA must call it.
__device__ __forceinline__ double csr_block_accumulate(
        double acc, int r0, int rend, int r, bool row_ok,
        const int *__restrict__ ptr, const int *__restrict__ idx, const double *__restrict__ val,
        const double *__restrict__ x, double *prod /* kCsrTile doubles of smem */)
{
    const int tid = threadIdx.x;
    const int a0 = __ldg(ptr + r0);
    const int a1 = __ldg(ptr + rend);
    int ps = a1, pe = a1;
    if (row_ok) { ps = __ldg(ptr + r); pe = __ldg(ptr + r + 1); }

    for (int w = a0 & ~3; w < a1; w += kCsrTile) {
        // ---- phase 1: products of window [w, w+kCsrTile) ∩ [a0, a1) ------------------------
#pragma unroll
        for (int k = 0; k < kCsrTile / (4 * kCsrThreads); ++k) {
            const int j = w + 4 * (tid + k * kCsrThreads);
            if (j < a1) {            // arrays are readable up to nnz rounded up to 4 entries
                const int4    c  = ld_stream4(reinterpret_cast<const int4 *>(idx + j));
                const double2 v0 = ld_stream2(reinterpret_cast<const double2 *>(val + j));
                const double2 v1 = ld_stream2(reinterpret_cast<const double2 *>(val + j + 2));
                // entries outside [a0,a1) belong to other CTAs (or are padding): their column
                // may be anything, so clamp the gather address instead of branching.
                const bool k0 = (j     >= a0) & (j     < a1);
                const bool k1 = (j + 1 >= a0) & (j + 1 < a1);
                const bool k2 = (j + 2 >= a0) & (j + 2 < a1);
                const bool k3 = (j + 3 >= a0) & (j + 3 < a1);
                const double x0 = __ldg(x + (k0 ? c.x : 0));
                const double x1 = __ldg(x + (k1 ? c.y : 0));
                const double x2 = __ldg(x + (k2 ? c.z : 0));
                const double x3 = __ldg(x + (k3 ? c.w : 0));
                double2 p0, p1;
                p0.x = mul(v0.x, x0); p0.y = mul(v0.y, x1);
                p1.x = mul(v1.x, x2); p1.y = mul(v1.y, x3);
                double2 *dst = reinterpret_cast<double2 *>(prod + (j - w));
                dst[0] = p0; dst[1] = p1;
            }
        }
        __syncthreads();
        // ---- phase 2: ordered row sums ---------------------------------------------------
        const int wend = w + kCsrTile;
        const int s = ps > w ? ps : w;
        const int e = pe < wend ? pe : wend;
        for (int j = s; j < e; ++j) acc = add(acc, prod[j - w]);
        __syncthreads();
    }
    return acc;
}

__global__ void __launch_bounds__(kCsrThreads)
csr_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ idx,
           const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
    __shared__ __align__(16) double prod[kCsrTile];
    const int r0 = blockIdx.x * kCsrThreads;
    const int rend = min(r0 + kCsrThreads, n);
    const int r = r0 + threadIdx.x;
    const bool ok = r < n;
    double acc = csr_block_accumulate(0.0, r0, rend, r, ok, ptr, idx, val, x, prod);
    if (ok) y[r] = acc;
}

// split order: t = D[i]*x[i]; t += L row; t += U row   (src/matvec/lis_matvec_csr.c:69-86)
__global__ void __launch_bounds__(kCsrThreads)
csr_split_kernel(int n, const double *__restrict__ diag,
                 const int *__restrict__ lptr, const int *__restrict__ lidx, const double *__restrict__ lval,
                 const int *__restrict__ uptr, const int *__restrict__ uidx, const double *__restrict__ uval,
                 const double *__restrict__ x, double *__restrict__ y)
{
    __shared__ __align__(16) double prod[kCsrTile];
    const int r0 = blockIdx.x * kCsrThreads;
    const int rend = min(r0 + kCsrThreads, n);
    const int r = r0 + threadIdx.x;
    const bool ok = r < n;
    double acc = ok ? mul(diag[r], x[r]) : 0.0;
    acc = csr_block_accumulate(acc, r0, rend, r, ok, lptr, lidx, lval, x, prod);
    acc = csr_block_accumulate(acc, r0, rend, r, ok, uptr, uidx, uval, x, prod);
    if (ok) y[r] = acc;
}

// y = A x and, in the same pass, sum_i x[i]*y[i]  (q = A p ; <p,q> of CG)
__global__ void __launch_bounds__(kCsrThreads)
csr_dot_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ idx,
               const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
               double *partial, unsigned int *counter, double *result)
{
    __shared__ __align__(16) double prod[kCsrTile];
    __shared__ double red[32];
    const int r0 = blockIdx.x * kCsrThreads;
    const int rend = min(r0 + kCsrThreads, n);
    const int r = r0 + threadIdx.x;
    const bool ok = r < n;
    double acc = csr_block_accumulate(0.0, r0, rend, r, ok, ptr, idx, val, x, prod);
    double d = 0.0;
    if (ok) { y[r] = acc; d = mul(x[r], acc); }
    double mine[1] = { block_reduce<false, kCsrThreads>(d, red) };
    grid_finish<false, kCsrThreads, 1>(mine, partial, counter, result, red);
}

// ---- CSR, short-row matrices: TMA-staged row blocks ------------------------------------------
// For matrices whose rows are short (stencils, FEM: a block of kRows rows owns at most kTile
// stored entries -- the host checks that), the idx/val/ptr slices of a row block are
// brought into shared memory by the TMA bulk-copy engine (no registers, no L1 traffic) by a
// dedicated producer warp running kStages blocks ahead, and consumer thread r walks row r out
// of shared memory in storage order, gathering x through the read-only path.  Lanes of a warp
// own consecutive rows, so for banded matrices the k-th gather of the warp is coalesced --
// this is what the product-tile kernel above cannot offer (its lanes own consecutive ENTRIES,
// whose columns jump between the diagonals; measured 93 % L1 utilisation at 0.84 of HBM peak).
// Persistent: gridDim.x CTAs stride over the row blocks.
template <int kRows>
struct CsrTmaSmem {
    static constexpr int kPtrInts = kRows + 4;
    static __host__ __device__ size_t stage_bytes(int tile) { return (size_t)tile * 12 + kPtrInts * 4; }
};

// kHalo: the row-partitioned product with the halo exchange INSIDE the kernel (one process per GPU,
// peers' memory mapped through CUDA IPC, NVLink loads/stores; host/lis_comm.c sets the table up):
//   1. all CTAs copy this rank's exported x entries straight into the neighbours' inboxes (P2P
//      stores), fence, and the last CTA to finish raises this rank's flag in every neighbour;
//   2. the row blocks that read no halo entry -- [int_lo, int_hi), found by the host -- run first;
//   3. before its first other block a CTA waits for the neighbours' flags of this epoch, and those
//      blocks read columns >= n from the inbox instead of x[n..).
// No pack kernel, no NCCL group, no unpack copy, no second stream: the transfer overlaps the
// interior rows by construction.  Inboxes are double-buffered by epoch parity: a neighbour can only
// be one product ahead (it needs this rank's push of the current epoch to finish its own).
#ifndef LISB_EMU
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
// a plain (L1-cached, coherent at kernel boundaries and after an acquire) global load the compiler may schedule freely
__device__ __forceinline__ double ld_global(const double *p) {
    double v;
    asm("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

__device__ __forceinline__ int halo_block_of(int v, int int_lo, int int_hi)
{
    const int ni = int_hi - int_lo;                  // interior blocks first, then the others in row order
    if (v < ni) return int_lo + v;
    v -= ni;
    return v < int_lo ? v : v + ni;
}

// one row out of a staged block: entries [j, e) of sidx/sval in storage order; up to kGather gathers of x in flight per
// thread; tail entries are clamped to the row's last entry and masked out of the sum.  kInbox: columns >= n are read from
// the inbox (a pointer select and ONE kind of load: two predicated loads of different kinds made the compiler order the
// gathers in two groups with the arithmetic in between; profiles/r02_bench_2gpu_vmm.log).  Cached loads of the inbox are
// safe: nothing of it can sit in L1 (invalidated at kernel start and again by the flag acquire, never read before that).
template <bool kInbox>
__device__ __forceinline__ double csr_row_walk(const int *__restrict__ sidx, const double *__restrict__ sval, int j, int e,
                                               const double *__restrict__ x, const double *inbox, int n)
{
    constexpr int kGather = 8;
    double acc = 0.0;
    for (; j < e; j += kGather) {
        int c[kGather];
        double v[kGather], xv[kGather];
#pragma unroll
        for (int k = 0; k < kGather; ++k) {
            const int jj = min(j + k, e - 1);
            c[k] = sidx[jj];
            v[k] = sval[jj];
        }
#pragma unroll
        for (int k = 0; k < kGather; ++k) xv[k] = kInbox ? ld_global(c[k] < n ? x + c[k] : inbox + c[k]) : __ldg(x + c[k]);
#pragma unroll
        for (int k = 0; k < kGather; ++k)
            if (j + k < e) acc = add(acc, mul(v[k], xv[k]));
    }
    return acc;
}

template <int kRows, int kStages, bool kDot, bool kHalo>
__global__ void __launch_bounds__(kRows + 32, 1152 / (kRows + 32))
csr_tma_kernel(int n, int nblocks, int tile, const int *__restrict__ ptr, const int *__restrict__ idx,
               const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
               const double *__restrict__ dotx /* kDot: row r pairs with dotx[r] (x itself, or x + first row of a row range) */,
               double *partial, unsigned int *counter, double *result,
               const lisb200_p2p *__restrict__ pd, unsigned long long epoch, int int_lo, int int_hi)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages];
    __shared__ double red[32];
    const int tid = threadIdx.x;
    const size_t stage_bytes = CsrTmaSmem<kRows>::stage_bytes(tile);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kRows / 32); }
        mbar_fence_init();
    }
    const int par = kHalo ? (int)(epoch & 1ull) : 0;
    if (kHalo) {
        // ---- push: exported entries into the neighbours' inboxes of this epoch's parity
        const int tot = pd->n_export;
        for (int q = blockIdx.x * (kRows + 32) + tid; q < tot; q += gridDim.x * (kRows + 32)) {
            int s = 0;
            while (q >= pd->exp_start[s + 1]) ++s;
            pd->peer_inbox[s][(long long)par * pd->peer_stride[s] + (q - pd->exp_start[s])] = x[__ldg(pd->export_index + q)];
        }
        __threadfence_system();
    }
    __syncthreads();
    if (kHalo && tid == 0) {
        const unsigned int arrived = atomicAdd(pd->push_count, 1u);
        if (arrived == gridDim.x - 1) {
            *pd->push_count = 0u;                    // every CTA has arrived: ready for the next product
            __threadfence_system();
            for (int s = 0; s < pd->n_nbr; ++s) st_release_sys(pd->peer_flag[s] + par * LISB200_P2P_MAX, epoch);
        }
    }
    const double *inbox = kHalo ? pd->inbox + (long long)par * pd->inbox_stride - n : nullptr;   // column c >= n -> inbox[c]
    const int n_first = kHalo ? int_hi - int_lo : 0;

    double dsum = 0.0;
    if (tid >= kRows) {
        // ---------------- producer warp: one lane drives the TMA ----------------
        if (tid == kRows) {
            int it = 0;
            int vb = blockIdx.x;                     // position in the visiting order; b: the row block
            // slice bounds of the block after this one are fetched one iteration ahead, so the
            // producer never sits on a DRAM round trip between "stage free" and "TMA issued"
            int a0 = 0, a1 = 0;
            if (vb < nblocks) {
                const int b = kHalo ? halo_block_of(vb, int_lo, int_hi) : vb;
                a0 = __ldg(ptr + b * kRows); a1 = __ldg(ptr + min(b * kRows + kRows, n));
            }
            for (; vb < nblocks; vb += gridDim.x, ++it) {
                const int b = kHalo ? halo_block_of(vb, int_lo, int_hi) : vb;
                const int s = it % kStages;
                const int r0 = b * kRows;
                const int rend = min(r0 + kRows, n);
                int na0 = 0, na1 = 0;
                if (vb + (int)gridDim.x < nblocks) {
                    const int bn = kHalo ? halo_block_of(vb + gridDim.x, int_lo, int_hi) : vb + gridDim.x;
                    na0 = __ldg(ptr + bn * kRows); na1 = __ldg(ptr + min(bn * kRows + kRows, n));
                }
                if (it >= kStages) mbar_wait(&empty_bar[s], ((it / kStages) - 1) & 1);
                unsigned char *st = smem_raw + (size_t)s * stage_bytes;
                double *sval = reinterpret_cast<double *>(st);
                int *sidx = reinterpret_cast<int *>(st + (size_t)tile * 8);
                int *sptr = sidx + tile;
                const int w = a0 & ~3;
                const uint32_t cnt = (uint32_t)((a1 - w + 3) & ~3);
                const uint32_t nptr = (uint32_t)((rend - r0 + 1 + 3) & ~3);
                mbar_expect_tx(&full_bar[s], cnt * 12u + nptr * 4u);
                tma_load_1d(sptr, ptr + r0, nptr * 4u, &full_bar[s]);
                if (cnt) {
                    tma_load_1d(sidx, idx + w, cnt * 4u, &full_bar[s]);
                    tma_load_1d(sval, val + w, cnt * 8u, &full_bar[s]);
                }
                a0 = na0; a1 = na1;
            }
        }
    } else {
        // ---------------- consumers: thread r walks row r ----------------
        int it = 0;
        bool have_halo = false;
        for (int vb = blockIdx.x; vb < nblocks; vb += gridDim.x, ++it) {
            const int b = kHalo ? halo_block_of(vb, int_lo, int_hi) : vb;
            if (kHalo && !have_halo && vb >= n_first) {
                // first block that may read halo entries: the neighbours' pushes of this epoch must have landed
                have_halo = true;
                const unsigned long long t0 = global_timer_ns();
                for (int q = 0; q < pd->n_nbr; ++q) {
                    const unsigned long long *f = pd->my_flag + par * LISB200_P2P_MAX + pd->nbr_rank[q];
                    while (ld_acquire_sys(f) != epoch) {
                        if (global_timer_ns() - t0 > 20000000000ull) { *pd->error = 1; break; }   // 20 s: a neighbour never pushed
                    }
                }
            }
            const int s = it % kStages;
            unsigned char *st = smem_raw + (size_t)s * stage_bytes;
            const double *sval = reinterpret_cast<const double *>(st);
            const int *sidx = reinterpret_cast<const int *>(st + (size_t)tile * 8);
            const int *sptr = sidx + tile;
            const int r = b * kRows + tid;
            mbar_wait(&full_bar[s], (it / kStages) & 1);
            if (r < n) {
                const int w = sptr[0] & ~3;
                int j = sptr[tid] - w;
                const int e = sptr[tid + 1] - w;
                // interior blocks (all but the two boundary planes of a slab) take the plain read-only gather; only
                // the blocks behind the flag wait pay for the two-source gather
                const double acc = (kHalo && vb >= n_first) ? csr_row_walk<true>(sidx, sval, j, e, x, inbox, n)
                                                            : csr_row_walk<false>(sidx, sval, j, e, x, inbox, n);
                y[r] = acc;
                if (kDot) dsum = add(dsum, mul(__ldg(dotx + r), acc));
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]);
        }
    }
    if (kDot) {
        // per-CTA partial of <x,y>: rows of this CTA in block order per thread, then the fixed tree
        __syncthreads();
        double v = warp_reduce<false>(dsum);
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        double mine[1] = {0.0};
        if (tid < 32) {
            constexpr int nw = (kRows + 32) / 32;
            double t = tid < nw ? red[tid] : 0.0;
            t = warp_reduce<false>(t);
            mine[0] = t;
        }
        __syncthreads();
        grid_finish<false, kRows + 32, 1>(mine, partial, counter, result, red);
    }
}

// ---- ELL -------------------------------------------------------------------------------
// y[i] = 0; for j<maxnzr: y[i] += value[j*ld+i]*x[index[j*ld+i]]  (lis_matvec_ell.c:110-128)
// Column-major sweep, every idx/val load coalesced.  kPair: a thread owns rows 2t and 2t+1 and
// moves them with one 64-bit index load and one 128-bit value load per slot (ld even).
template <bool kPair>
__global__ void __launch_bounds__(256)
ell_kernel(int n, int maxnzr, int ld, const int *__restrict__ idx, const double *__restrict__ val,
           const double *__restrict__ x, double *__restrict__ y)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (kPair) {
        const int i = 2 * t;
        if (i + 1 < n) {
            double t0 = 0.0, t1 = 0.0;
            int j = 0;
            for (; j + 2 <= maxnzr; j += 2) {        // 2 slots x 2 rows in flight
                const size_t o = (size_t)j * ld + i;
                const int2 ca = *reinterpret_cast<const int2 *>(idx + o);
                const int2 cb = *reinterpret_cast<const int2 *>(idx + o + ld);
                const double2 va = ld_stream2(reinterpret_cast<const double2 *>(val + o));
                const double2 vb = ld_stream2(reinterpret_cast<const double2 *>(val + o + ld));
                const double xa0 = __ldg(x + ca.x), xa1 = __ldg(x + ca.y), xb0 = __ldg(x + cb.x), xb1 = __ldg(x + cb.y);
                t0 = add(t0, mul(va.x, xa0)); t1 = add(t1, mul(va.y, xa1));
                t0 = add(t0, mul(vb.x, xb0)); t1 = add(t1, mul(vb.y, xb1));
            }
            for (; j < maxnzr; ++j) {
                const size_t o = (size_t)j * ld + i;
                const int2 c = *reinterpret_cast<const int2 *>(idx + o);
                const double2 v = ld_stream2(reinterpret_cast<const double2 *>(val + o));
                t0 = add(t0, mul(v.x, __ldg(x + c.x))); t1 = add(t1, mul(v.y, __ldg(x + c.y)));
            }
            *reinterpret_cast<double2 *>(y + i) = make_double2(t0, t1);
            return;
        }
        if (i >= n) return;
        double tt = 0.0;                               // odd n: the last row alone
        for (int j = 0; j < maxnzr; ++j) {
            const size_t o = (size_t)j * ld + i;
            tt = add(tt, mul(ld_stream(val + o), __ldg(x + ld_stream(idx + o))));
        }
        y[i] = tt;
        return;
    }
    const int i = t;
    if (i >= n) return;
    double tt = 0.0;
    int j = 0;
    for (; j + 4 <= maxnzr; j += 4) {       // 4 independent idx/val streams in flight
        const size_t o = (size_t)j * ld + i;
        const int c0 = ld_stream(idx + o), c1 = ld_stream(idx + o + ld);
        const int c2 = ld_stream(idx + o + 2 * (size_t)ld), c3 = ld_stream(idx + o + 3 * (size_t)ld);
        const double v0 = ld_stream(val + o), v1 = ld_stream(val + o + ld);
        const double v2 = ld_stream(val + o + 2 * (size_t)ld), v3 = ld_stream(val + o + 3 * (size_t)ld);
        const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
        tt = add(tt, mul(v0, x0)); tt = add(tt, mul(v1, x1));
        tt = add(tt, mul(v2, x2)); tt = add(tt, mul(v3, x3));
    }
    for (; j < maxnzr; ++j) {
        const size_t o = (size_t)j * ld + i;
        tt = add(tt, mul(ld_stream(val + o), __ldg(x + ld_stream(idx + o))));
    }
    y[i] = tt;
}

// ---- DIA -------------------------------------------------------------------------------
// y[i] = 0; for each diagonal j with offset off[j]: rows max(0,-off) <= i < min(n, xlen-off)
//   y[i] += value[j*ld+i]*x[i+off]                                (lis_matvec_dia.c:150-172)
// kPair: two adjacent rows per thread, one 128-bit value load per diagonal (ld even).
constexpr int kDiaMaxOff = 64;
template <bool kPair>
__global__ void __launch_bounds__(256)
dia_kernel(int n, int xlen, int nnd, int ld, const int *__restrict__ off,
           const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
    __shared__ int soff[kDiaMaxOff];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = kPair ? 2 * t : t;
    double t0 = 0.0, t1 = 0.0;
    for (int j0 = 0; j0 < nnd; j0 += kDiaMaxOff) {
        const int nj = min(kDiaMaxOff, nnd - j0);
        __syncthreads();
        if (threadIdx.x < nj) soff[threadIdx.x] = off[j0 + threadIdx.x];
        __syncthreads();
        if (kPair && i + 1 < n) {
#pragma unroll 4
            for (int j = 0; j < nj; ++j) {
                const int c = i + soff[j];
                const bool in0 = c >= 0 && c < xlen, in1 = c + 1 >= 0 && c + 1 < xlen;
                if (in0 | in1) {
                    const double2 v = ld_stream2(reinterpret_cast<const double2 *>(val + (size_t)(j0 + j) * ld + i));
                    if (in0) t0 = add(t0, mul(v.x, __ldg(x + c)));
                    if (in1) t1 = add(t1, mul(v.y, __ldg(x + c + 1)));
                }
            }
        } else if (i < n) {
#pragma unroll 4
            for (int j = 0; j < nj; ++j) {
                const int c = i + soff[j];
                if (c >= 0 && c < xlen)
                    t0 = add(t0, mul(ld_stream(val + (size_t)(j0 + j) * ld + i), __ldg(x + c)));
            }
        }
    }
    if (kPair && i + 1 < n) *reinterpret_cast<double2 *>(y + i) = make_double2(t0, t1);
    else if (i < n) y[i] = t0;
}

// ---- JAD -------------------------------------------------------------------------------
// w[i] = sum_j value[jptr[j]+i]*x[index[jptr[j]+i]] for all j with i < jptr[j+1]-jptr[j];
// y[perm[i]] = w[i]                                                (lis_matvec_jad.c:171-196)
// A row sits in the first `cnt` jagged diagonals (their lengths never increase), so the trip count
// is known up front: four diagonals' index/value loads and x gathers are in flight before the first
// add (round 1 walked one diagonal at a time behind a data-dependent break: 0.72 of the HBM peak).
__global__ void __launch_bounds__(256)
jad_kernel(int n, int maxnzr, const int *__restrict__ jptr, const int *__restrict__ perm,
           const int *__restrict__ idx, const double *__restrict__ val,
           const double *__restrict__ x, double *__restrict__ y)
{
    extern __shared__ int sjp[];            // maxnzr+1 entries (or 0 when it does not fit)
    const bool cached = (maxnzr + 1) * (int)sizeof(int) <= 16384;
    if (cached) {
        for (int j = threadIdx.x; j <= maxnzr; j += blockDim.x) sjp[j] = jptr[j];
        __syncthreads();
    }
    const int *jp = cached ? sjp : jptr;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int out = __ldg(perm + i);
    // cnt = number of diagonals longer than i (binary search over the non-increasing lengths)
    int lo = 0, hi = maxnzr;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (jp[mid + 1] - jp[mid] > i) lo = mid + 1; else hi = mid;
    }
    const int cnt = lo;
    double t = 0.0;
    int j = 0;
    for (; j + 4 <= cnt; j += 4) {
        const size_t o0 = (size_t)jp[j] + i, o1 = (size_t)jp[j + 1] + i, o2 = (size_t)jp[j + 2] + i, o3 = (size_t)jp[j + 3] + i;
        const int c0 = ld_stream(idx + o0), c1 = ld_stream(idx + o1), c2 = ld_stream(idx + o2), c3 = ld_stream(idx + o3);
        const double v0 = ld_stream(val + o0), v1 = ld_stream(val + o1), v2 = ld_stream(val + o2), v3 = ld_stream(val + o3);
        const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
        t = add(t, mul(v0, x0)); t = add(t, mul(v1, x1)); t = add(t, mul(v2, x2)); t = add(t, mul(v3, x3));
    }
    if (j < cnt) {                           // 1..3 left: same loads, the missing ones clamped to the last valid diagonal
        const int r = cnt - j;
        const size_t o0 = (size_t)jp[j] + i, o1 = (size_t)jp[j + (r > 1 ? 1 : 0)] + i, o2 = (size_t)jp[j + (r > 2 ? 2 : 0)] + i;
        const int c0 = ld_stream(idx + o0), c1 = ld_stream(idx + o1), c2 = ld_stream(idx + o2);
        const double v0 = ld_stream(val + o0), v1 = ld_stream(val + o1), v2 = ld_stream(val + o2);
        const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2);
        t = add(t, mul(v0, x0));
        if (r > 1) t = add(t, mul(v1, x1));
        if (r > 2) t = add(t, mul(v2, x2));
    }
    y[out] = t;
}

// ---- BSR -------------------------------------------------------------------------------
// per block row bi: t[0..bnr) = 0; for each block bc (storage order): for j<bnc, for i<bnr:
//   t[i] += value[bc*bs + j*bnr + i] * x[bindex[bc]*bnc + j]       (lis_matvec_bsr.c:134-146)
// The unrolled RxC kernels of the reference accumulate each t[i] in exactly this order.
// Block-row tile kernel for the 4x4 table of block shapes (2x2 is the reference's default).
// Round 1 ran a thread per block row: every thread walked its own 36-byte-strided run of blocks and
// DRAM moved 3.4x the algorithmic bytes (profiles/r02_ncu_bsr_v1.txt, 0.25 of the HBM peak at 512^3).
// Here a CTA owns kBsrRows consecutive block rows, i.e. ONE contiguous slice of value[]: phase 1
// streams that slice with coalesced loads (element e of the slice belongs to block e / (R*C),
// column (e % (R*C)) / R), multiplies by the matching x entry and parks the rounded product in
// shared memory; phase 2 lets thread r add the products of block row r in the reference's order --
// block by block, inside a block column by column, t[i] += a[j*R+i]*x[j] (lis_matvec_bsr.c:134-146,
// 338-343) -- so the bits are the reference's.
constexpr int kBsrRows = 256;             // block rows per CTA == threads per CTA
constexpr int kBsrTileDoubles = 4096;     // products per shared-memory window (32 KB)

// Shared-memory layout of a window: component-major -- product (block b, element c of the block)
// at prod[c * WB + b] -- so that in phase 2 the lanes of a warp (consecutive block rows, i.e. block
// offsets a row length apart) read 8-byte words a row length apart: conflict-free for the odd and
// 2-way for the even row lengths of stencils, where the block-major layout of the first version was
// 8-way (32-byte blocks 7 blocks apart; profiles/r02_ncu_bsr_v2.txt).
template <int R, int C>
__global__ void __launch_bounds__(kBsrRows, 4)
bsr_tile_kernel(int n, int ncols, int nr, const int *__restrict__ bptr, const int *__restrict__ bidx,
                const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
    constexpr int BS = R * C;
    constexpr bool kVec = (BS % 2) == 0;                       // a slice then starts 16-byte aligned
    constexpr int WB = kBsrTileDoubles / BS;                   // blocks per window
    constexpr int W = WB * BS;                                 // elements per window
    constexpr int kPer = kVec ? 2 : 1;                         // elements per thread per step
    constexpr int kSteps = (W + kPer * kBsrRows - 1) / (kPer * kBsrRows);
    constexpr int kUnroll = 4;                                 // steps whose loads are in flight together
    __shared__ __align__(16) double prod[W];
    const int tid = threadIdx.x;
    const bool aligned32 = ((reinterpret_cast<uintptr_t>(val) & 31) | (reinterpret_cast<uintptr_t>(x) & 15)) == 0;
    const int br0 = blockIdx.x * kBsrRows;
    const int brend = min(br0 + kBsrRows, nr);
    const int bi = br0 + tid;
    const bool ok = bi < nr;
    const long long b0 = __ldg(bptr + br0), b1 = __ldg(bptr + brend);       // block range of the CTA
    long long ps = b1, pe = b1;                                            // block range of this thread's block row
    if (ok) { ps = __ldg(bptr + bi); pe = __ldg(bptr + bi + 1); }
    double t[R];
#pragma unroll
    for (int i = 0; i < R; ++i) t[i] = 0.0;
    for (long long wb = b0; wb < b1; wb += WB) {
        const long long wbend = wb + WB < b1 ? wb + WB : b1;
        const long long e0 = wb * BS;
        const int cnt = (int)(wbend - wb) * BS;                // elements in this window
        // ---- phase 1: products of the window's elements, kUnroll steps of loads in flight together
        if (R == 2 && C == 2 && aligned32) {
            // the reference's default shape: a thread takes a whole block -- ONE 256-bit load of its four values
            // (32 lanes: 1 KB contiguous), one index load, one 128-bit load of its x pair -- half the load
            // instructions and L1 requests of the two-elements-per-thread path below (which ran at 84 % of the
            // L1 throughput, profiles/r02_ncu_bsr_jad_v2.txt)
            constexpr int kBlkSteps = (WB + kBsrRows - 1) / kBsrRows;            // 4
            const int nblk = (int)(wbend - wb);
            double4 a[kBlkSteps];
            int bc[kBlkSteps];
#pragma unroll
            for (int u = 0; u < kBlkSteps; ++u) {
                const int bl = tid + u * kBsrRows;
                bc[u] = -1;
                a[u] = make_double4(0.0, 0.0, 0.0, 0.0);
                if (bl < nblk) {
                    a[u] = ld_stream4d(val + e0 + 4 * (long long)bl);
                    bc[u] = __ldg(bidx + wb + bl) * 2;
                }
            }
            double2 xx[kBlkSteps];
#pragma unroll
            for (int u = 0; u < kBlkSteps; ++u) {
                xx[u] = make_double2(0.0, 0.0);
                if (bc[u] >= 0) {
                    // a padded last block column holds structural zeros; x behind them does not exist
                    if (bc[u] + 1 < ncols) xx[u] = __ldg(reinterpret_cast<const double2 *>(x + bc[u]));
                    else xx[u].x = __ldg(x + bc[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < kBlkSteps; ++u)
                if (bc[u] >= 0) {
                    const int bl = tid + u * kBsrRows;
                    prod[0 * WB + bl] = mul(a[u].x, xx[u].x);
                    prod[1 * WB + bl] = mul(a[u].y, xx[u].x);
                    prod[2 * WB + bl] = mul(a[u].z, xx[u].y);
                    prod[3 * WB + bl] = mul(a[u].w, xx[u].y);
                }
        } else
        for (int s0 = 0; s0 < kSteps; s0 += kUnroll) {
            double a0[kUnroll], a1[kUnroll];
            int bcol[kUnroll], off[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                off[u] = kPer * (tid + (s0 + u) * kBsrRows);
                const bool in = s0 + u < kSteps && off[u] < cnt;
                a0[u] = 0.0; a1[u] = 0.0; bcol[u] = 0;
                if (in) {
                    if (kVec) {
                        const double2 a = ld_stream2(reinterpret_cast<const double2 *>(val + e0 + off[u]));
                        a0[u] = a.x; a1[u] = a.y;
                    } else {
                        a0[u] = ld_stream(val + e0 + off[u]);
                    }
                    bcol[u] = __ldg(bidx + wb + off[u] / BS) * C;
                } else off[u] = -1;
            }
            double x0[kUnroll], x1[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                x0[u] = 0.0; x1[u] = 0.0;
                if (off[u] >= 0) {
                    const int rem = off[u] % BS;
                    // a padded last block column holds structural zeros; x behind them does not exist
                    const int c0 = bcol[u] + rem / R;
                    x0[u] = c0 < ncols ? __ldg(x + c0) : 0.0;
                    if (kVec) {
                        if (R % 2 == 0) x1[u] = x0[u];                 // rem is even: both elements sit in the same block column
                        else { const int c1 = bcol[u] + (rem + 1) / R; x1[u] = c1 < ncols ? __ldg(x + c1) : 0.0; }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (off[u] >= 0) {
                    const int bl = off[u] / BS, rem = off[u] % BS;
                    prod[rem * WB + bl] = mul(a0[u], x0[u]);
                    if (kVec) prod[(rem + 1) * WB + bl] = mul(a1[u], x1[u]);      // rem even, BS even: same block
                }
        }
        __syncthreads();
        // ---- phase 2: ordered sums of this thread's block row inside the window
        {
            const long long s = ps > wb ? ps : wb, e = pe < wbend ? pe : wbend;
            for (long long kb = s; kb < e; ++kb) {
                const double *p = prod + (kb - wb);
#pragma unroll
                for (int j = 0; j < C; ++j)
#pragma unroll
                    for (int i = 0; i < R; ++i) t[i] = add(t[i], p[(j * R + i) * WB]);
            }
        }
        __syncthreads();
    }
    if (ok) {
#pragma unroll
        for (int i = 0; i < R; ++i)
            if (bi * R + i < n) y[bi * R + i] = t[i];
    }
}

// generic block size (bnr or bnc > 4): one thread per scalar row, same per-row order
__global__ void __launch_bounds__(256)
bsr_generic_kernel(int n, int ncols, int nr, int bnr, int bnc, const int *__restrict__ bptr,
                   const int *__restrict__ bidx, const double *__restrict__ val,
                   const double *__restrict__ x, double *__restrict__ y)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int bi = row / bnr, i = row - bi * bnr;
    const int bs = bnr * bnc;
    double t = 0.0;
    const int s = __ldg(bptr + bi), e = __ldg(bptr + bi + 1);
    for (int bc = s; bc < e; ++bc) {
        const int bj = __ldg(bidx + bc) * bnc;
        const double *v = val + (size_t)bc * bs + i;
        for (int j = 0; j < bnc; ++j) t = add(t, mul(__ldg(v + (size_t)j * bnr), bj + j < ncols ? __ldg(x + bj + j) : 0.0));
    }
    y[row] = t;
}

template <int R>
static int launch_bsr_c(int n, int ncols, int nr, int bnc, const int *bptr, const int *bidx, const double *val,
                        const double *x, double *y, cudaStream_t st)
{
    const int grid = (nr + kBsrRows - 1) / kBsrRows;
    switch (bnc) {
    case 1: bsr_tile_kernel<R, 1><<<grid, kBsrRows, 0, st>>>(n, ncols, nr, bptr, bidx, val, x, y); break;
    case 2: bsr_tile_kernel<R, 2><<<grid, kBsrRows, 0, st>>>(n, ncols, nr, bptr, bidx, val, x, y); break;
    case 3: bsr_tile_kernel<R, 3><<<grid, kBsrRows, 0, st>>>(n, ncols, nr, bptr, bidx, val, x, y); break;
    case 4: bsr_tile_kernel<R, 4><<<grid, kBsrRows, 0, st>>>(n, ncols, nr, bptr, bidx, val, x, y); break;
    default: return -1;
    }
    return 0;
}

}  // namespace lisb

using namespace lisb;

// ---- TMA row-block variant: plan + launch ---------------------------------------------------
namespace lisb {
static int g_sm_count = 0;
static int sm_count_spmv() {
    if (g_sm_count <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sm_count <= 0)
            g_sm_count = 148;
    }
    return g_sm_count;
}

struct HaloArgs { const lisb200_p2p *pd; unsigned long long epoch; int int_lo, int_hi; };

template <int kRows, int kStages, bool kDot, bool kHalo>
static int launch_csr_tma_s(int n, int tile, const int *ptr, const int *idx, const double *val, const double *x, double *y,
                            const double *dotx, double *partial, unsigned int *counter, double *result, HaloArgs h, cudaStream_t st)
{
    const size_t smem = kStages * CsrTmaSmem<kRows>::stage_bytes(tile);
    auto kern = csr_tma_kernel<kRows, kStages, kDot, kHalo>;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    int per_sm = (int)((size_t)(222 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    const int max_by_threads = 1152 / (kRows + 32);      // register budget of __launch_bounds__
    if (per_sm > max_by_threads) per_sm = max_by_threads;
    const int nblocks = (n + kRows - 1) / kRows;
    int grid = sm_count_spmv() * per_sm;
    if (grid > nblocks) grid = nblocks;
    kern<<<grid, kRows + 32, smem, st>>>(n, nblocks, tile, ptr, idx, val, x, y, dotx, partial, counter, result,
                                         h.pd, h.epoch, h.int_lo / kRows, h.int_hi / kRows);
    LISB_CHECK_LAUNCH();
    return 0;
}

template <int kRows, bool kDot, bool kHalo = false>
static int launch_csr_tma(int n, int tile, int stages, const int *ptr, const int *idx, const double *val, const double *x, double *y,
                          const double *dotx, double *partial, unsigned int *counter, double *result, cudaStream_t st,
                          HaloArgs h = HaloArgs{nullptr, 0ull, 0, 0})
{
    switch (stages) {
    case 3: return launch_csr_tma_s<kRows, 3, kDot, kHalo>(n, tile, ptr, idx, val, x, y, dotx, partial, counter, result, h, st);
    case 4: return launch_csr_tma_s<kRows, 4, kDot, kHalo>(n, tile, ptr, idx, val, x, y, dotx, partial, counter, result, h, st);
    case 2: return launch_csr_tma_s<kRows, 2, kDot, kHalo>(n, tile, ptr, idx, val, x, y, dotx, partial, counter, result, h, st);
    default: break;
    }
    if (kHalo) return (int)cudaErrorInvalidValue;        // the plan only yields depths 2..4
    switch (stages) {
    case 6: return launch_csr_tma_s<kRows, 6, kDot, false>(n, tile, ptr, idx, val, x, y, dotx, partial, counter, result, h, st);
    case 8: return launch_csr_tma_s<kRows, 8, kDot, false>(n, tile, ptr, idx, val, x, y, dotx, partial, counter, result, h, st);
    default: return (int)cudaErrorInvalidValue;
    }
}
}  // namespace lisb

/* row-partitioned product with the halo exchange inside the kernel; see include/lis_b200_kernels.h */
extern "C" int lisb200_spmv_csr_tma_p2p(int n, int rows_per_block, int tile, int stages, const int *d_ptr, const int *d_idx,
                                        const double *d_val, const double *d_x, double *d_y, int with_dot, double *d_partial,
                                        unsigned int *d_counter, double *d_result, const lisb200_p2p *d_table,
                                        unsigned long long epoch, int interior_lo, int interior_hi, void *stream)
{
    if (n <= 0 || d_table == nullptr || epoch == 0) return (int)cudaErrorInvalidValue;
    if (interior_lo % rows_per_block || interior_hi % rows_per_block || interior_lo > interior_hi || interior_hi > n) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    const HaloArgs h{d_table, epoch, interior_lo, interior_hi};
    if (with_dot) {
        switch (rows_per_block) {
        case 256: return launch_csr_tma<256, true, true>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, d_x, d_partial, d_counter, d_result, st, h);
        case 128: return launch_csr_tma<128, true, true>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, d_x, d_partial, d_counter, d_result, st, h);
        case 64:  return launch_csr_tma<64, true, true>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, d_x, d_partial, d_counter, d_result, st, h);
        default:  return (int)cudaErrorInvalidValue;
        }
    }
    switch (rows_per_block) {
    case 256: return launch_csr_tma<256, false, true>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, nullptr, nullptr, nullptr, nullptr, st, h);
    case 128: return launch_csr_tma<128, false, true>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, nullptr, nullptr, nullptr, nullptr, st, h);
    case 64:  return launch_csr_tma<64, false, true>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, nullptr, nullptr, nullptr, nullptr, st, h);
    default:  return (int)cudaErrorInvalidValue;
    }
}

// rows_per_block in {256,128,64}; tile = entries staged per row block (multiple of 4);
// stages in {2,3,4,6,8} with stages * (12*tile + 4*(rows+4)) <= ~224 KB
extern "C" int lisb200_spmv_csr_tma(int n, int rows_per_block, int tile, int stages, const int *d_ptr, const int *d_idx,
                                    const double *d_val, const double *d_x, double *d_y, void *stream)
{
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (rows_per_block) {
    case 256: return launch_csr_tma<256, false>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, nullptr, nullptr, nullptr, nullptr, st);
    case 128: return launch_csr_tma<128, false>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, nullptr, nullptr, nullptr, nullptr, st);
    case 64:  return launch_csr_tma<64, false>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, nullptr, nullptr, nullptr, nullptr, st);
    default:  return (int)cudaErrorInvalidValue;
    }
}

extern "C" int lisb200_spmv_csr_tma_dot_rows(int n, int rows_per_block, int tile, int stages, const int *d_ptr, const int *d_idx,
                                             const double *d_val, const double *d_x, double *d_y, const double *d_dotx,
                                             double *d_partial, unsigned int *d_counter, double *d_result, void *stream);

extern "C" int lisb200_spmv_csr_tma_dot(int n, int rows_per_block, int tile, int stages, const int *d_ptr, const int *d_idx,
                                        const double *d_val, const double *d_x, double *d_y, double *d_partial,
                                        unsigned int *d_counter, double *d_result, void *stream)
{
    return lisb200_spmv_csr_tma_dot_rows(n, rows_per_block, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, d_x, d_partial, d_counter, d_result, stream);
}

/* the same on a row range: d_ptr / d_y / d_dotx point at the range's first row (d_dotx = x + first row), d_x at the
 * whole vector; the range's share of <x,y> goes to *d_result (empty range: 0) */
extern "C" int lisb200_spmv_csr_tma_dot_rows(int n, int rows_per_block, int tile, int stages, const int *d_ptr, const int *d_idx,
                                             const double *d_val, const double *d_x, double *d_y, const double *d_dotx,
                                             double *d_partial, unsigned int *d_counter, double *d_result, void *stream)
{
    if (n <= 0) return (int)cudaMemsetAsync(d_result, 0, sizeof(double), (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    switch (rows_per_block) {
    case 256: return launch_csr_tma<256, true>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, d_dotx, d_partial, d_counter, d_result, st);
    case 128: return launch_csr_tma<128, true>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, d_dotx, d_partial, d_counter, d_result, st);
    case 64:  return launch_csr_tma<64, true>(n, tile, stages, d_ptr, d_idx, d_val, d_x, d_y, d_dotx, d_partial, d_counter, d_result, st);
    default:  return (int)cudaErrorInvalidValue;
    }
}

// Plan for the TMA kernel from the HOST row pointers: rows per block R, tile (entries staged per
// block) and pipeline depth.  What the sweeps say (profiles/r01_sweep_csr_tma.txt):
//   * ~512 consumer threads per SM keep enough x gathers in flight (7-pt: 512 consumers/4 stages
//     899 GFLOP/s vs 1024 consumers/2 stages 831; 27-pt: 320 consumers 1013 vs 128 consumers 542);
//   * beyond that, shared memory is better spent on pipeline depth (>= 3 stages hide the drain
//     bubble of a stage that all warps of the CTA must release before it is refilled).
// So: depth = what fits next to 512 consumers' worth of staged rows (2..4), then the R that
// yields the most resident consumers (ties: larger R).  Returns 0 and fills the plan, or 1 when
// the rows are too long/ragged for a thread-per-row walk (use lisb200_spmv_csr).
extern "C" int lisb200_spmv_csr_tma_plan(int n, const int *h_ptr, int *rows_per_block, int *tile, int *stages)
{
    if (n <= 0) return 1;
    const int cand[3] = {256, 128, 64};
    const long long smem_sm = 222 * 1024;
    long long best_consumers = 0;
    for (int c = 0; c < 3; ++c) {
        const int R = cand[c];
        long long worst = 0;
        for (long long r0 = 0; r0 < n; r0 += R) {
            const long long rend = r0 + R < n ? r0 + R : n;
            const long long w = h_ptr[r0] & ~3;
            const long long cnt = ((long long)h_ptr[rend] - w + 3) & ~3LL;
            if (cnt > worst) worst = cnt;
        }
        if (worst > 64LL * R) continue;                             // very ragged: one thread per row would crawl
        long long t = (worst + 255) & ~255LL;
        if (t < 256) t = 256;
        const long long stage = 12 * t + 4 * (R + 4);
        long long st = smem_sm / (512 * stage / R);                 // depth affordable with 512 consumers resident
        if (st > 4) st = 4;
        if (st < 2) st = 2;
        long long ctas = smem_sm / (st * stage + 1024);
        const long long by_threads = 1152 / (R + 32);          // launch bound of the kernel
        if (ctas > by_threads) ctas = by_threads;
        if (ctas < 1) continue;
        const long long consumers = ctas * R;
        if (consumers > best_consumers) {
            best_consumers = consumers;
            *rows_per_block = R; *tile = (int)t; *stages = (int)st;
        }
    }
    return best_consumers > 0 ? 0 : 1;
}

extern "C" int lisb200_spmv_csr(int n, const int *d_ptr, const int *d_idx, const double *d_val,
                                const double *d_x, double *d_y, void *stream)
{
    if (n <= 0) return 0;
    const int grid = (n + kCsrThreads - 1) / kCsrThreads;
    csr_kernel<<<grid, kCsrThreads, 0, (cudaStream_t)stream>>>(n, d_ptr, d_idx, d_val, d_x, d_y);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_spmv_csr_split(int n, const double *d_diag,
                                      const int *d_lptr, const int *d_lidx, const double *d_lval,
                                      const int *d_uptr, const int *d_uidx, const double *d_uval,
                                      const double *d_x, double *d_y, void *stream)
{
    if (n <= 0) return 0;
    const int grid = (n + kCsrThreads - 1) / kCsrThreads;
    csr_split_kernel<<<grid, kCsrThreads, 0, (cudaStream_t)stream>>>(
        n, d_diag, d_lptr, d_lidx, d_lval, d_uptr, d_uidx, d_uval, d_x, d_y);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_spmv_csr_dot_slots(int n) { return n > 0 ? (n + kCsrThreads - 1) / kCsrThreads : 1; }

extern "C" int lisb200_spmv_csr_dot(int n, const int *d_ptr, const int *d_idx, const double *d_val,
                                    const double *d_x, double *d_y, double *d_partial,
                                    unsigned int *d_counter, double *d_result, void *stream)
{
    if (n <= 0) return 0;
    const int grid = (n + kCsrThreads - 1) / kCsrThreads;
    csr_dot_kernel<<<grid, kCsrThreads, 0, (cudaStream_t)stream>>>(
        n, d_ptr, d_idx, d_val, d_x, d_y, d_partial, d_counter, d_result);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_spmv_ell(int n, int maxnzr, int ld, const int *d_idx, const double *d_val,
                                const double *d_x, double *d_y, void *stream)
{
    if (n <= 0) return 0;
    const bool pair = (ld % 2 == 0) && (((uintptr_t)d_idx & 7) == 0) && (((uintptr_t)d_val & 15) == 0) && (((uintptr_t)d_y & 15) == 0);
    if (pair) ell_kernel<true><<<((n + 1) / 2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, maxnzr, ld, d_idx, d_val, d_x, d_y);
    else ell_kernel<false><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, maxnzr, ld, d_idx, d_val, d_x, d_y);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_spmv_dia(int n, int xlen, int nnd, int ld, const int *d_off,
                                const double *d_val, const double *d_x, double *d_y, void *stream)
{
    if (n <= 0) return 0;
    const bool pair = (ld % 2 == 0) && (((uintptr_t)d_val & 15) == 0) && (((uintptr_t)d_y & 15) == 0);
    if (pair) dia_kernel<true><<<((n + 1) / 2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, xlen, nnd, ld, d_off, d_val, d_x, d_y);
    else dia_kernel<false><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, xlen, nnd, ld, d_off, d_val, d_x, d_y);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_spmv_jad(int n, int maxnzr, const int *d_jptr, const int *d_perm,
                                const int *d_idx, const double *d_val,
                                const double *d_x, double *d_y, void *stream)
{
    if (n <= 0) return 0;
    size_t sm = (size_t)(maxnzr + 1) * sizeof(int);
    if (sm > 16384) sm = 0;
    jad_kernel<<<(n + 255) / 256, 256, sm, (cudaStream_t)stream>>>(n, maxnzr, d_jptr, d_perm, d_idx, d_val, d_x, d_y);
    LISB_CHECK_LAUNCH();
    return 0;
}

/* ncols: entries of x (n for a square matrix, n + halo columns for a row-partitioned one) */
extern "C" int lisb200_spmv_bsr_cols(int n, int ncols, int nr, int bnr, int bnc, const int *d_bptr,
                                     const int *d_bidx, const double *d_val,
                                     const double *d_x, double *d_y, void *stream)
{
    if (n <= 0 || nr <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = -1;
    if (bnc >= 1 && bnc <= 4 && ((uintptr_t)d_val & 15) == 0) {
        switch (bnr) {
        case 1: rc = launch_bsr_c<1>(n, ncols, nr, bnc, d_bptr, d_bidx, d_val, d_x, d_y, st); break;
        case 2: rc = launch_bsr_c<2>(n, ncols, nr, bnc, d_bptr, d_bidx, d_val, d_x, d_y, st); break;
        case 3: rc = launch_bsr_c<3>(n, ncols, nr, bnc, d_bptr, d_bidx, d_val, d_x, d_y, st); break;
        case 4: rc = launch_bsr_c<4>(n, ncols, nr, bnc, d_bptr, d_bidx, d_val, d_x, d_y, st); break;
        default: break;
        }
    }
    if (rc != 0)
        bsr_generic_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, ncols, nr, bnr, bnc, d_bptr, d_bidx, d_val, d_x, d_y);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_spmv_bsr(int n, int nr, int bnr, int bnc, const int *d_bptr,
                                const int *d_bidx, const double *d_val,
                                const double *d_x, double *d_y, void *stream)
{
    return lisb200_spmv_bsr_cols(n, n, nr, bnr, bnc, d_bptr, d_bidx, d_val, d_x, d_y, stream);
}
