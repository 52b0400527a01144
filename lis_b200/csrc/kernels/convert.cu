// convert.cu -- storage-format conversion on the device: CSR -> ELL / DIA / JAD / BSR.
//
// The reference converts on the host, through CSR (lis_matrix_convert, src/matrix/lis_matrix_ops.c:
// 127-322; builders src/matrix/lis_matrix_ell.c:957-1070, lis_matrix_dia.c:1190-1305,
// lis_matrix_jad.c:1590-1770, lis_matrix_bsr.c:350-545).  At 512^3 that host pass, not the solver,
// is the wall-clock cost of `-storage ell`.  These kernels produce the same arrays -- the SERIAL
// layouts, entry for entry, which host/lis_convert.c also produces and the SpMV kernels read -- from
// a CSR mirror that is already in HBM.  Integer/byte work, HBM-bound: every kernel is a thread-per-row
// (or per jagged position / block row) sweep whose writes are coalesced in the column-major target;
// the CSR side is read through L1 (adjacent threads own adjacent rows, so a warp's reads cover a
// contiguous slice).  Small tables that the public struct needs on the host anyway (DIA offsets, JAD
// pointers, BSR block pointers) are prefix-summed there between two launches.
#include "common.cuh"
#include "../../../include/lis_b200_kernels.h"

namespace lisb {

constexpr int kCvThreads = 256;

__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const int w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- longest row (ELL slots, JAD diagonals) -------------------------------------------------
__global__ void __launch_bounds__(kCvThreads)
row_len_max_kernel(int n, const int *__restrict__ ptr, int *out_max)
{
    __shared__ int wmax[kCvThreads / 32];
    int m = 0;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int len = ptr[i + 1] - ptr[i];
        m = len > m ? len : m;
    }
    m = warp_max_int(m);
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = threadIdx.x < kCvThreads / 32 ? wmax[threadIdx.x] : 0;
        t = warp_max_int(t);
        if (threadIdx.x == 0) atomicMax(out_max, t);
    }
}

// ---- ELL: value[j*ld+i], index[j*ld+i]; unused slots hold (0.0, i)   lis_matrix_ell.c:1035-1052
__global__ void __launch_bounds__(kCvThreads)
csr2ell_kernel(int n, int maxnzr, int ld, const int *__restrict__ ptr, const int *__restrict__ idx,
               const double *__restrict__ val, int *__restrict__ eidx, double *__restrict__ eval)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = ptr[i], len = ptr[i + 1] - s;
    for (int j = 0; j < maxnzr; ++j) {
        const size_t o = (size_t)j * ld + i;
        if (j < len) { eidx[o] = idx[s + j]; eval[o] = val[s + j]; }
        else { eidx[o] = i; eval[o] = 0.0; }
    }
}

// ---- DIA ---------------------------------------------------------------------------------------
// Rows must be sorted by column (the host sorts Ain first, like lis_matrix_dia.c:1217).
// 1. flags[col - row + n] = 1 for every stored entry; 2. per 1024-flag segment: how many offsets
// occur; (host: prefix sum over the segments -> nnd, bases) 3. ordered compaction of the offsets;
// 4. fill: row i walks the ascending offsets and its own ascending entries together and writes
// EVERY diagonal slot of its row (value or 0.0): one coalesced store per diagonal, no memset, no search.
constexpr int kDiaSeg = 1024;

__global__ void __launch_bounds__(kCvThreads)
dia_mark_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ idx, unsigned char *flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int e = ptr[i + 1];
    for (int j = ptr[i]; j < e; ++j) flags[(size_t)((long long)idx[j] - i + n)] = 1;
}

__global__ void __launch_bounds__(kCvThreads)
dia_count_kernel(long long span, const unsigned char *__restrict__ flags, int *__restrict__ seg_count)
{
    __shared__ int wsum[kCvThreads / 32];
    const long long base = (long long)blockIdx.x * kDiaSeg;
    int c = 0;
#pragma unroll
    for (int q = 0; q < kDiaSeg / kCvThreads; ++q) {
        const long long k = base + threadIdx.x * (kDiaSeg / kCvThreads) + q;
        if (k < span && flags[k]) ++c;
    }
    c = warp_sum_int(c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = threadIdx.x < kCvThreads / 32 ? wsum[threadIdx.x] : 0;
        t = warp_sum_int(t);
        if (threadIdx.x == 0) seg_count[blockIdx.x] = t;
    }
}

// seg_base[s] = number of offsets in the segments before s (exclusive prefix sum, from the host)
__global__ void __launch_bounds__(kCvThreads)
dia_compact_kernel(long long span, int n, const unsigned char *__restrict__ flags,
                   const int *__restrict__ seg_base, const int *__restrict__ seg_count, int *__restrict__ off_out)
{
    __shared__ int tcount[kCvThreads];
    if (seg_count[blockIdx.x] == 0) return;
    constexpr int per = kDiaSeg / kCvThreads;
    const long long base = (long long)blockIdx.x * kDiaSeg + threadIdx.x * per;
    int c = 0;
#pragma unroll
    for (int q = 0; q < per; ++q) if (base + q < span && flags[base + q]) ++c;
    tcount[threadIdx.x] = c;
    __syncthreads();
    int before = 0;
    for (int t = 0; t < (int)threadIdx.x; ++t) before += tcount[t];     // 256 x 255/2 adds per non-empty segment: few segments are
    int o = seg_base[blockIdx.x] + before;
#pragma unroll
    for (int q = 0; q < per; ++q)
        if (base + q < span && flags[base + q]) off_out[o++] = (int)(base + q - (long long)n);
}

__global__ void __launch_bounds__(kCvThreads)
csr2dia_fill_kernel(int n, int nnd, int ld, const int *__restrict__ ptr, const int *__restrict__ idx,
                    const double *__restrict__ val, const int *__restrict__ off, double *__restrict__ dval)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = ptr[i];
    const int e = ptr[i + 1];
    for (int k = 0; k < nnd; ++k) {
        const int o = __ldg(off + k);
        double v = 0.0;
        while (p < e && idx[p] - i == o) { v = val[p]; ++p; }      // duplicates: the last one stays, as on the host
        dval[(size_t)k * ld + i] = v;
    }
}

// ---- JAD ---------------------------------------------------------------------------------------
// perm lists the rows by descending length, rows of equal length in ascending order (a stable
// counting sort, host/lis_convert.c csr2jad).  bin = maxnzr - len (<= 255).  A CTA owns 4096
// consecutive rows: kernel 1 counts its rows per bin; the host turns the [cta][bin] table into
// start positions (bin-major, then CTA order); kernel 2 ranks every row among the rows of its bin
// that precede it in the CTA (tiles of 256 rows, O(tile^2) compares in shared memory) and writes
// perm; kernel 3 lays the entries out: position p of jagged diagonal j = j-th entry of row perm[p].
constexpr int kJadBins = 256;
constexpr int kJadTile = kCvThreads;
constexpr int kJadRowsPerCta = 4096;

__global__ void __launch_bounds__(kCvThreads)
jad_hist_kernel(int n, int maxnzr, const int *__restrict__ ptr, int *__restrict__ cta_bin)
{
    __shared__ int bins[kJadBins];
    bins[threadIdx.x] = 0;
    __syncthreads();
    const int r0 = blockIdx.x * kJadRowsPerCta;
    for (int t = 0; t < kJadRowsPerCta; t += kJadTile) {
        const int i = r0 + t + threadIdx.x;
        if (i < n) atomicAdd(&bins[maxnzr - (ptr[i + 1] - ptr[i])], 1);
    }
    __syncthreads();
    cta_bin[(size_t)blockIdx.x * kJadBins + threadIdx.x] = bins[threadIdx.x];
}

__global__ void __launch_bounds__(kCvThreads)
jad_perm_kernel(int n, int maxnzr, const int *__restrict__ ptr, const int *__restrict__ cta_base, int *__restrict__ perm)
{
    __shared__ int run[kJadBins];
    __shared__ int keys[kJadTile];
    run[threadIdx.x] = cta_base[(size_t)blockIdx.x * kJadBins + threadIdx.x];
    const int r0 = blockIdx.x * kJadRowsPerCta;
    for (int t = 0; t < kJadRowsPerCta && r0 + t < n; t += kJadTile) {
        const int i = r0 + t + threadIdx.x;
        const int key = i < n ? maxnzr - (ptr[i + 1] - ptr[i]) : -1;
        __syncthreads();                         // run[] of the previous tile is final, keys[] free
        keys[threadIdx.x] = key;
        __syncthreads();
        int p = -1;
        if (key >= 0) {
            int rank = 0;
            for (int s = 0; s < (int)threadIdx.x; ++s) rank += (keys[s] == key);
            p = run[key] + rank;
        }
        __syncthreads();                         // everybody has read run[] before it moves on
        if (key >= 0) { perm[p] = i; atomicAdd(&run[key], 1); }
    }
}

__global__ void __launch_bounds__(kCvThreads)
csr2jad_fill_kernel(int n, int maxnzr, const int *__restrict__ ptr, const int *__restrict__ idx,
                    const double *__restrict__ val, const int *__restrict__ jptr, const int *__restrict__ perm,
                    int *__restrict__ jidx, double *__restrict__ jval)
{
    __shared__ int sjp[kJadBins + 1];
    for (int j = threadIdx.x; j <= maxnzr; j += blockDim.x) sjp[j] = jptr[j];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int row = perm[p];
    const int s = ptr[row], len = ptr[row + 1] - s;
    for (int j = 0; j < len; ++j) {
        const size_t o = (size_t)sjp[j] + p;
        jidx[o] = idx[s + j];
        jval[o] = val[s + j];
    }
}

// ---- BSR ---------------------------------------------------------------------------------------
// Block row bi = scalar rows [bi*bnr, bi*bnr+bnr).  Its blocks are the distinct idx/bnc values met
// while scanning those rows in order (first-seen order, lis_matrix_bsr.c:411-470); inside a block
// value[j*bnr + ii], unset entries 0.0.  A thread owns a block row and keeps the block columns it
// has met in a small local list (linear search: block rows of the matrices BSR is used for hold a
// handful of blocks).  More than kBsrMaxBlocks distinct blocks in one block row raise `overflow`
// and the host converts instead.
constexpr int kBsrMaxBlocks = 64;

__global__ void __launch_bounds__(128)
bsr_count_kernel(int n, int nr, int bnr, int bnc, const int *__restrict__ ptr, const int *__restrict__ idx,
                 int *__restrict__ count, int *overflow)
{
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= nr) return;
    int seen[kBsrMaxBlocks];
    int cnt = 0;
    bool over = false;
    for (int ii = 0; ii < bnr; ++ii) {
        const int r = bi * bnr + ii;
        if (r >= n) break;
        const int e = ptr[r + 1];
        for (int k = ptr[r]; k < e; ++k) {
            const int bj = idx[k] / bnc;
            int q = 0;
            while (q < cnt && seen[q] != bj) ++q;
            if (q == cnt) {
                if (cnt < kBsrMaxBlocks) seen[cnt++] = bj;
                else over = true;
            }
        }
    }
    count[bi] = cnt;
    if (over) atomicMax(overflow, 1);
}

__global__ void __launch_bounds__(128)
csr2bsr_fill_kernel(int n, int nr, int bnr, int bnc, const int *__restrict__ ptr, const int *__restrict__ idx,
                    const double *__restrict__ val, const int *__restrict__ bptr,
                    int *__restrict__ bidx, double *__restrict__ bval)
{
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= nr) return;
    int seen[kBsrMaxBlocks];
    int cnt = 0;
    const int b0 = bptr[bi];
    const int bs = bnr * bnc;
    for (int ii = 0; ii < bnr; ++ii) {
        const int r = bi * bnr + ii;
        if (r >= n) break;
        const int e = ptr[r + 1];
        for (int k = ptr[r]; k < e; ++k) {
            const int c = idx[k];
            const int bj = c / bnc, j = c - bj * bnc;
            int q = 0;
            while (q < cnt && seen[q] != bj) ++q;
            if (q >= kBsrMaxBlocks) continue;                     // cannot happen: such matrices convert on the host
            double *blk = bval + (size_t)(b0 + q) * bs;
            if (q == cnt) {
                seen[cnt++] = bj;
                bidx[b0 + q] = bj;
                for (int z = 0; z < bs; ++z) blk[z] = 0.0;
            }
            blk[j * bnr + ii] = val[k];
        }
    }
}

static int cv_sm_count() {
    static int sms = 0;
    if (sms <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

// *unsorted = 1 if some row's columns are not STRICTLY ascending.  A repeated column also sends the matrix through the host
// sort (lis_matrix_sort_csr, src/matrix/lis_matrix_csr.c:1486): which of the copies ends last is a property of that sort,
// and the DIA conversion keeps the last one.
__global__ void __launch_bounds__(kCvThreads)
csr_unsorted_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ idx, int *unsorted)
{
    const int stride = gridDim.x * blockDim.x;
    int bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        for (int j = ptr[i] + 1; j < ptr[i + 1]; ++j) bad |= idx[j - 1] >= idx[j];
    if (bad) *unsorted = 1;
}

}  // namespace lisb

using namespace lisb;

/* *d_out = 1 when some row of the CSR matrix has its columns out of ascending order, else 0 */
extern "C" int lisb200_csr_rows_unsorted(int n, const int *d_ptr, const int *d_idx, int *d_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    if (n <= 0) return 0;
    long long grid = ((long long)n + kCvThreads - 1) / kCvThreads;
    const long long cap = (long long)cv_sm_count() * 16;
    if (grid > cap) grid = cap;
    csr_unsorted_kernel<<<(int)grid, kCvThreads, 0, st>>>(n, d_ptr, d_idx, d_out);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_csr_max_row_len(int n, const int *d_ptr, int *d_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    if (n <= 0) return 0;
    long long grid = ((long long)n + kCvThreads - 1) / kCvThreads;
    const long long cap = (long long)cv_sm_count() * 8;
    if (grid > cap) grid = cap;
    row_len_max_kernel<<<(int)grid, kCvThreads, 0, st>>>(n, d_ptr, d_out);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_csr2ell(int n, int maxnzr, int ld, const int *d_ptr, const int *d_idx, const double *d_val,
                               int *d_eidx, double *d_eval, void *stream)
{
    if (n <= 0 || maxnzr <= 0) return 0;
    csr2ell_kernel<<<(n + kCvThreads - 1) / kCvThreads, kCvThreads, 0, (cudaStream_t)stream>>>(
        n, maxnzr, ld, d_ptr, d_idx, d_val, d_eidx, d_eval);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_dia_segments(int n, int np) { return (int)(((long long)n + np + kDiaSeg - 1) / kDiaSeg); }

extern "C" int lisb200_csr2dia_mark(int n, int np, const int *d_ptr, const int *d_idx,
                                    unsigned char *d_flags, int *d_seg_count, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const long long span = (long long)n + np;
    if (n <= 0) return 0;
    cudaError_t e = cudaMemsetAsync(d_flags, 0, (size_t)span, st);
    if (e != cudaSuccess) return (int)e;
    dia_mark_kernel<<<(n + kCvThreads - 1) / kCvThreads, kCvThreads, 0, st>>>(n, d_ptr, d_idx, d_flags);
    LISB_CHECK_LAUNCH();
    dia_count_kernel<<<lisb200_dia_segments(n, np), kCvThreads, 0, st>>>(span, d_flags, d_seg_count);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_csr2dia_fill(int n, int np, int nnd, int ld, const int *d_ptr, const int *d_idx, const double *d_val,
                                    const unsigned char *d_flags, const int *d_seg_base, const int *d_seg_count,
                                    int *d_off, double *d_dval, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0 || nnd <= 0) return 0;
    dia_compact_kernel<<<lisb200_dia_segments(n, np), kCvThreads, 0, st>>>((long long)n + np, n, d_flags, d_seg_base, d_seg_count, d_off);
    LISB_CHECK_LAUNCH();
    csr2dia_fill_kernel<<<(n + kCvThreads - 1) / kCvThreads, kCvThreads, 0, st>>>(n, nnd, ld, d_ptr, d_idx, d_val, d_off, d_dval);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_jad_ctas(int n) { return n > 0 ? (n + kJadRowsPerCta - 1) / kJadRowsPerCta : 0; }
extern "C" int lisb200_jad_bins(void) { return kJadBins; }

extern "C" int lisb200_csr2jad_hist(int n, int maxnzr, const int *d_ptr, int *d_cta_bin, void *stream)
{
    if (n <= 0) return 0;
    if (maxnzr >= kJadBins) return (int)cudaErrorInvalidValue;
    jad_hist_kernel<<<lisb200_jad_ctas(n), kCvThreads, 0, (cudaStream_t)stream>>>(n, maxnzr, d_ptr, d_cta_bin);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_csr2jad_fill(int n, int maxnzr, const int *d_ptr, const int *d_idx, const double *d_val,
                                    const int *d_cta_base, const int *d_jptr, int *d_perm, int *d_jidx, double *d_jval,
                                    void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return 0;
    if (maxnzr >= kJadBins) return (int)cudaErrorInvalidValue;
    jad_perm_kernel<<<lisb200_jad_ctas(n), kCvThreads, 0, st>>>(n, maxnzr, d_ptr, d_cta_base, d_perm);
    LISB_CHECK_LAUNCH();
    csr2jad_fill_kernel<<<(n + kCvThreads - 1) / kCvThreads, kCvThreads, 0, st>>>(n, maxnzr, d_ptr, d_idx, d_val, d_jptr, d_perm, d_jidx, d_jval);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_bsr_max_blocks(void) { return kBsrMaxBlocks; }

extern "C" int lisb200_csr2bsr_count(int n, int nr, int bnr, int bnc, const int *d_ptr, const int *d_idx,
                                     int *d_count, int *d_overflow, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_overflow, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    if (n <= 0 || nr <= 0) return 0;
    bsr_count_kernel<<<(nr + 127) / 128, 128, 0, st>>>(n, nr, bnr, bnc, d_ptr, d_idx, d_count, d_overflow);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_csr2bsr_fill(int n, int nr, int bnr, int bnc, const int *d_ptr, const int *d_idx, const double *d_val,
                                    const int *d_bptr, int *d_bidx, double *d_bval, void *stream)
{
    if (n <= 0 || nr <= 0) return 0;
    csr2bsr_fill_kernel<<<(nr + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, nr, bnr, bnc, d_ptr, d_idx, d_val, d_bptr, d_bidx, d_bval);
    LISB_CHECK_LAUNCH();
    return 0;
}
