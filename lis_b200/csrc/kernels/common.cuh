// common.cuh -- shared device helpers for the lis_b200 sm_100a kernels.
//
// Arithmetic discipline: every fp64 multiply/add that must match the reference's CPU path is
// spelled with the round-to-nearest intrinsics (__dmul_rn/__dadd_rn/...), which the compiler
// never contracts into FMA.  The translation units are additionally built with -fmad=false.
#pragma once
#ifdef LISB_EMU            // tests/cudaemu: kernel sources compiled for the host emulator (test infrastructure)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#define LISB_CHECK_LAUNCH() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return (int)e__; } while (0)

namespace lisb {

constexpr int kWarp = 32;

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

#ifndef LISB_EMU           // the inline-PTX helpers; cuda_emu.h restates them for the host emulator
// streaming (read-once) loads: bypass L1 allocation so the gathered x keeps the L1
__device__ __forceinline__ double ld_stream(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_stream2(const double2 *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
// 256-bit streaming load (sm_100: LDG.E.256): four doubles, address 32-byte aligned
__device__ __forceinline__ double4 ld_stream4d(const double *p) {
    double4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream4(const int4 *p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS: UBLKCP) -----------------------------------
// One elected thread arms the barrier with the byte count and issues the copies; the bytes land
// in shared memory without passing through registers or L1; consumers spin on the phase parity.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif  // !LISB_EMU

// ---- deterministic block reduction (sum or max) ------------------------------------------
// Fixed shape: xor-shuffle tree inside each warp, then warp 0 combines the per-warp values
// with the same tree.  The result is valid in thread 0.
template <bool kMax>
__device__ __forceinline__ double combine(double a, double b) {
    if (kMax) return a > b ? a : b;
    return __dadd_rn(a, b);
}
template <bool kMax>
__device__ __forceinline__ double warp_reduce(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = combine<kMax>(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <bool kMax, int kThreads>
__device__ __forceinline__ double block_reduce(double v, double *smem /* >= 32 doubles */) {
    constexpr int nw = kThreads / kWarp;
    v = warp_reduce<kMax>(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();                       // smem may still be in use by a previous reduction
    if (lane == 0) smem[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? smem[lane] : (kMax ? 0.0 : 0.0);
        t = warp_reduce<kMax>(t);
        v = t;
    }
    return v;
}

// ---- grid-level finish: every CTA stores one partial, the last CTA to arrive folds them ---
// The fold order depends only on (nparts, kThreads), never on which CTA happens to be last:
// thread t sums partial[t], partial[t+kThreads], ... sequentially, then block_reduce.
template <bool kMax, int kThreads, int kOut>
__device__ __forceinline__ void grid_finish(const double (&mine)[kOut], double *partial,
                                            unsigned int *counter, double *result,
                                            double *smem, bool sqrt_first = false) {
    __shared__ bool is_last;
    const unsigned int nparts = gridDim.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kOut; ++k) partial[(size_t)k * nparts + blockIdx.x] = mine[k];
        __threadfence();
        unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == nparts - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int k = 0; k < kOut; ++k) {
        double t = 0.0;
        const volatile double *p = partial + (size_t)k * nparts;
        for (unsigned int i = threadIdx.x; i < nparts; i += kThreads) t = combine<kMax>(t, p[i]);
        t = block_reduce<kMax, kThreads>(t, smem);
        if (threadIdx.x == 0) result[k] = t;
    }
    if (threadIdx.x == 0) {
        *counter = 0u;                      // ready for the next reduction on this stream
        __threadfence_system();             // result may live in mapped host memory
    }
}

}  // namespace lisb
