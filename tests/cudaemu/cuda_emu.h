/*
 * cuda_emu.h -- TEST INFRASTRUCTURE ONLY.  Lets the product's CUDA kernel sources
 * (lis_b200/csrc/kernels/*.cu) be compiled by g++ and executed on the host, one CUDA thread
 * per fiber, so that their indexing, tiling, barrier and pipeline logic can be checked against
 * the oracle on a machine without a GPU (tests/test_emu_kernels.py).
 *
 * It is NOT a CPU fallback: nothing under lis_b200/ includes or links it, the product library
 * still returns LIS_ERR_DEVICE without a GPU, and no number measured through it is ever
 * reported.  What it checks: every index expression, alignment assumption (128-bit loads, TMA
 * bulk copies: 16-byte addresses and sizes), out-of-bounds access (device allocations end at a
 * guard page), shared-memory staging, mbarrier phase logic, the last-CTA reduction fold and the
 * dependency polling of the one-launch triangular sweeps.  What it cannot check: the memory
 * model, the PTX itself, performance.
 *
 * Execution model: CTAs run one after another in launch order; the threads of a CTA are
 * ucontext fibers scheduled round-robin.  A fiber runs until it finishes or reaches
 * __syncthreads / __syncwarp / a warp shuffle / a spin-wait (mbarrier wait, dependency poll).
 * tests/cudaemu/cu2cpp.py rewrites `kernel<<<g, b, s, st>>>(args)` into emu::launch(...) and
 * `extern __shared__ T name[]` into a pointer to the launch's dynamic shared memory; the kernel
 * bodies are compiled unchanged (common.cuh swaps its inline-PTX helpers for the ones below
 * under LISB_EMU).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <functional>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __launch_bounds__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)

namespace emu {

struct ThreadCtx {
    uint3 tid, bid;
    dim3 bdim, gdim;
};
extern ThreadCtx *g_cur;            /* the CUDA thread whose fiber is running */

void launch(const char *name, dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()> &body);
void *dyn_smem();
void sync_block();
void sync_warp();
void spin_yield();                  /* inside a spin-wait loop */
uint64_t shfl_exchange(uint64_t mine, int src_lane);   /* value of lane src_lane (own value if out of range) */
[[noreturn]] void fail(const char *what);

static inline void check_aligned(const void *p, size_t a, const char *what) {
    if (((uintptr_t)p) & (a - 1)) fail(what);
}

}  // namespace emu

#define threadIdx (emu::g_cur->tid)
#define blockIdx  (emu::g_cur->bid)
#define blockDim  (emu::g_cur->bdim)
#define gridDim   (emu::g_cur->gdim)

/* C++ convenience overload cuda_runtime.h only provides under nvcc */
template <class T> static inline cudaError_t cudaFuncSetAttribute(T *entry, enum cudaFuncAttribute attr, int value) {
    return ::cudaFuncSetAttribute(reinterpret_cast<const void *>(entry), attr, value);
}

/* ---- device intrinsics (the translation units are built with -ffp-contract=off) ---- */
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)mask; emu::sync_warp(); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline unsigned int atomicAdd(unsigned int *p, unsigned int v) { unsigned int o = *p; *p = o + v; return o; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline double atomicAdd(double *p, double v) { double o = *p; *p = o + v; return o; }
static inline int atomicMax(int *p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline unsigned int atomicMax(unsigned int *p, unsigned int v) { unsigned int o = *p; if (v > o) *p = v; return o; }
static inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline unsigned int atomicOr(unsigned int *p, unsigned int v) { unsigned int o = *p; *p = o | v; return o; }
static inline int atomicCAS(int *p, int cmp, int v) { int o = *p; if (o == cmp) *p = v; return o; }
static inline unsigned int atomicExch(unsigned int *p, unsigned int v) { unsigned int o = *p; *p = v; return o; }

template <class T> static inline T emu_shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes");
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    bits = emu::shfl_exchange(bits, src);
    T out;
    memcpy(&out, &bits, sizeof(T));
    return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) { return emu_shfl(v, (int)((threadIdx.x & 31) ^ lane_mask)); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta) { return emu_shfl(v, (int)((threadIdx.x & 31) + delta)); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned delta) { return emu_shfl(v, (int)(threadIdx.x & 31) - (int)delta); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl(v, src & 31); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (emu_shfl<int>(pred ? 1 : 0, l) ? 1u : 0u) << l;
    return r;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }

/* ---- the inline-PTX helpers of common.cuh / sweep.cu, restated for the host ---- */
namespace lisb {

__forceinline__ double ld_stream(const double *p) { return *p; }
__forceinline__ int ld_stream(const int *p) { return *p; }
__forceinline__ double4 ld_stream4d(const double *p) { emu::check_aligned(p, 32, "ld_stream4d: address not 32-byte aligned"); double4 v; v.x = p[0]; v.y = p[1]; v.z = p[2]; v.w = p[3]; return v; }
__forceinline__ double2 ld_stream2(const double2 *p) { emu::check_aligned(p, 16, "ld_stream2: address not 16-byte aligned"); return *p; }
__forceinline__ int4 ld_stream4(const int4 *p) { emu::check_aligned(p, 16, "ld_stream4: address not 16-byte aligned"); return *p; }
__forceinline__ int2 ld_stream2(const int2 *p) { emu::check_aligned(p, 8, "ld_stream2(int2): address not 8-byte aligned"); return *p; }

/* mbarrier state packed into the 64-bit word the kernels declare:
 *   [0]      phase parity of the CURRENT (incomplete) phase
 *   [1..20]  pending arrivals    [21..40] arrival count the barrier was initialised with
 *   [41..63] pending transaction bytes (biased: expect_tx adds, complete_tx subtracts) */
struct EmuMbar { uint64_t phase : 1, pending : 20, count : 20; int64_t tx : 23; };
static_assert(sizeof(EmuMbar) == 8, "EmuMbar must overlay a uint64_t");
__forceinline__ EmuMbar *emu_bar(uint64_t *b) { return reinterpret_cast<EmuMbar *>(b); }
__forceinline__ void emu_bar_try_complete(EmuMbar *m) {
    if (m->pending == 0 && m->tx == 0) { m->phase ^= 1; m->pending = m->count; }
}
__forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { EmuMbar *m = emu_bar(bar); m->phase = 0; m->pending = count; m->count = count; m->tx = 0; }
__forceinline__ void mbar_fence_init() {}
__forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    EmuMbar *m = emu_bar(bar);
    if (bytes >= (1u << 20)) emu::fail("mbarrier expect_tx: byte count exceeds the 2^20-1 transaction limit");
    if (m->pending == 0) emu::fail("mbarrier arrive.expect_tx: more arrivals than the barrier was initialised for");
    m->tx += bytes; m->pending -= 1;
    emu_bar_try_complete(m);
}
__forceinline__ void mbar_arrive(uint64_t *bar) {
    EmuMbar *m = emu_bar(bar);
    if (m->pending == 0) emu::fail("mbarrier arrive: more arrivals than the barrier was initialised for");
    m->pending -= 1;
    emu_bar_try_complete(m);
}
__forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    /* try_wait.parity succeeds once the phase with that parity has completed */
    while (emu_bar(bar)->phase == (parity & 1u)) emu::spin_yield();
}
__forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    emu::check_aligned(dst_smem, 16, "cp.async.bulk: shared-memory address not 16-byte aligned");
    emu::check_aligned(src_gmem, 16, "cp.async.bulk: global address not 16-byte aligned");
    if (bytes == 0 || (bytes & 15u)) emu::fail("cp.async.bulk: size must be a non-zero multiple of 16 bytes");
    memcpy(dst_smem, src_gmem, bytes);
    EmuMbar *m = emu_bar(bar);
    m->tx -= bytes;
    emu_bar_try_complete(m);
}
/* dependency polling of the one-launch sweeps (sweep.cu) */
__forceinline__ unsigned long long ld_poll(const double *p) {
    emu::spin_yield();
    unsigned long long v; memcpy(&v, p, 8); return v;
}
__forceinline__ void st_publish(double *p, double v) { *p = v; }
/* flags of the in-kernel halo exchange (spmv.cu, kHalo) */
__forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) { emu::spin_yield(); return *(const volatile unsigned long long *)p; }
__forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) { *(volatile unsigned long long *)p = v; }
__forceinline__ unsigned long long global_timer_ns() { return 0ull; }
__forceinline__ double ld_global(const double *p) { return *p; }

}  // namespace lisb
