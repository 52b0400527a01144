// blas1.cu -- lis_vector_* BLAS-1 kernels (src/vector/lis_vector_opv.c, lis_vector_ops.c of
// the reference) plus the fused CG vector updates.  All HBM-bound streams:
//  * elementwise ops: 128-bit loads/stores, grid sized to cover the SMs several times over,
//    grid-stride loop; mul and add rounded separately => bit-identical to the CPU loops;
//  * reductions: per-thread strided partial sums (4 independent accumulators), warp-shuffle +
//    shared-memory block tree, then the LAST CTA folds the per-CTA partials in a fixed order
//    and writes the scalar where the host can see it (mapped pinned memory) -- one launch per
//    reduction, deterministic for a given n.
#include "common.cuh"
#include "../../../include/lis_b200_kernels.h"

namespace lisb {

constexpr int kEwThreads  = 256;
constexpr int kRedThreads = 256;
constexpr int kRedMaxGrid = 148 * 8;       // upper bound on reduction CTAs (fits kReduceSlots)
constexpr int kReduceSlots = 4 * kRedMaxGrid;

static int g_sms = 0;
static int sm_count() {
    if (g_sms <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sms <= 0)
            g_sms = 148;
    }
    return g_sms;
}

static inline int ew_grid(int n) {
    long long need = ((long long)n + 2 * kEwThreads - 1) / (2 * kEwThreads);   // 2 elems / thread / step
    long long cap = (long long)sm_count() * 16;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

static inline bool aligned16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

// ---- generic elementwise driver ------------------------------------------------------------
// Op::apply(i-th element) is expressed on double2 lanes; F describes loads/stores.
template <class F>
__global__ void __launch_bounds__(kEwThreads) ew_kernel(int n, F f, bool vec)
{
    const int stride = gridDim.x * blockDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const int n2 = n >> 1;
#pragma unroll 4
        for (int i = t; i < n2; i += stride) f.vec2(i);
        if (t == 0 && (n & 1)) f.one(n - 1);
    } else {
        for (int i = t; i < n; i += stride) f.one(i);
    }
}

struct CopyF {
    const double *x; double *y;
    __device__ void one(int i) const { y[i] = x[i]; }
    __device__ void vec2(int i) const { reinterpret_cast<double2 *>(y)[i] = reinterpret_cast<const double2 *>(x)[i]; }
};
struct AxpyF {      // y += alpha*x
    double a; const double *x; double *y;
    __device__ void one(int i) const { y[i] = add(y[i], mul(a, x[i])); }
    __device__ void vec2(int i) const {
        double2 xv = reinterpret_cast<const double2 *>(x)[i];
        double2 yv = reinterpret_cast<double2 *>(y)[i];
        yv.x = add(yv.x, mul(a, xv.x)); yv.y = add(yv.y, mul(a, xv.y));
        reinterpret_cast<double2 *>(y)[i] = yv;
    }
};
struct AxpyDevF {   // y += (scale * *alpha_dev) * x : alpha comes from an earlier reduction on the same stream
    const double *alpha_dev; double scale; const double *x; double *y;
    __device__ void one(int i) const { const double a = mul(scale, *alpha_dev); y[i] = add(y[i], mul(a, x[i])); }
    __device__ void vec2(int i) const {
        const double a = mul(scale, *alpha_dev);
        double2 xv = reinterpret_cast<const double2 *>(x)[i];
        double2 yv = reinterpret_cast<double2 *>(y)[i];
        yv.x = add(yv.x, mul(a, xv.x)); yv.y = add(yv.y, mul(a, xv.y));
        reinterpret_cast<double2 *>(y)[i] = yv;
    }
};
struct XpayF {      // y = x + alpha*y
    double a; const double *x; double *y;
    __device__ void one(int i) const { y[i] = add(x[i], mul(a, y[i])); }
    __device__ void vec2(int i) const {
        double2 xv = reinterpret_cast<const double2 *>(x)[i];
        double2 yv = reinterpret_cast<double2 *>(y)[i];
        yv.x = add(xv.x, mul(a, yv.x)); yv.y = add(xv.y, mul(a, yv.y));
        reinterpret_cast<double2 *>(y)[i] = yv;
    }
};
struct AxpyzF {     // z = alpha*x + y
    double a; const double *x; const double *y; double *z;
    __device__ void one(int i) const { z[i] = add(mul(a, x[i]), y[i]); }
    __device__ void vec2(int i) const {
        double2 xv = reinterpret_cast<const double2 *>(x)[i];
        double2 yv = reinterpret_cast<const double2 *>(y)[i];
        double2 zv; zv.x = add(mul(a, xv.x), yv.x); zv.y = add(mul(a, xv.y), yv.y);
        reinterpret_cast<double2 *>(z)[i] = zv;
    }
};
struct ScaleF {     // x = alpha*x
    double a; double *x;
    __device__ void one(int i) const { x[i] = mul(a, x[i]); }
    __device__ void vec2(int i) const {
        double2 v = reinterpret_cast<double2 *>(x)[i];
        v.x = mul(a, v.x); v.y = mul(a, v.y);
        reinterpret_cast<double2 *>(x)[i] = v;
    }
};
struct PmulF {      // z = x*y
    const double *x; const double *y; double *z;
    __device__ void one(int i) const { z[i] = mul(x[i], y[i]); }
    __device__ void vec2(int i) const {
        double2 xv = reinterpret_cast<const double2 *>(x)[i];
        double2 yv = reinterpret_cast<const double2 *>(y)[i];
        double2 zv; zv.x = mul(xv.x, yv.x); zv.y = mul(xv.y, yv.y);
        reinterpret_cast<double2 *>(z)[i] = zv;
    }
};
struct PdivF {      // z = x/y
    const double *x; const double *y; double *z;
    __device__ void one(int i) const { z[i] = __ddiv_rn(x[i], y[i]); }
    __device__ void vec2(int i) const {
        double2 xv = reinterpret_cast<const double2 *>(x)[i];
        double2 yv = reinterpret_cast<const double2 *>(y)[i];
        double2 zv; zv.x = __ddiv_rn(xv.x, yv.x); zv.y = __ddiv_rn(xv.y, yv.y);
        reinterpret_cast<double2 *>(z)[i] = zv;
    }
};
struct SetF {
    double a; double *x;
    __device__ void one(int i) const { x[i] = a; }
    __device__ void vec2(int i) const { reinterpret_cast<double2 *>(x)[i] = make_double2(a, a); }
};
struct AbsF {
    double *x;
    __device__ void one(int i) const { x[i] = fabs(x[i]); }
    __device__ void vec2(int i) const {
        double2 v = reinterpret_cast<double2 *>(x)[i];
        v.x = fabs(v.x); v.y = fabs(v.y);
        reinterpret_cast<double2 *>(x)[i] = v;
    }
};
struct RecipF {     // x = 1.0/x
    double *x;
    __device__ void one(int i) const { x[i] = __ddiv_rn(1.0, x[i]); }
    __device__ void vec2(int i) const {
        double2 v = reinterpret_cast<double2 *>(x)[i];
        v.x = __ddiv_rn(1.0, v.x); v.y = __ddiv_rn(1.0, v.y);
        reinterpret_cast<double2 *>(x)[i] = v;
    }
};
struct ShiftF {     // x = x - sigma
    double s; double *x;
    __device__ void one(int i) const { x[i] = sub(x[i], s); }
    __device__ void vec2(int i) const {
        double2 v = reinterpret_cast<double2 *>(x)[i];
        v.x = sub(v.x, s); v.y = sub(v.y, s);
        reinterpret_cast<double2 *>(x)[i] = v;
    }
};
struct SwapF {
    double *x; double *y;
    __device__ void one(int i) const { double t = y[i]; y[i] = x[i]; x[i] = t; }
    __device__ void vec2(int i) const {
        double2 a = reinterpret_cast<double2 *>(x)[i];
        double2 b = reinterpret_cast<double2 *>(y)[i];
        reinterpret_cast<double2 *>(x)[i] = b; reinterpret_cast<double2 *>(y)[i] = a;
    }
};
struct BicgstabPF {     // p = r + beta*(p + nomega*v): axpy(-omega,v,p) then xpay(r,beta,p), one pass
    double nomega, beta; const double *v; const double *r; double *p;
    __device__ void one(int i) const { p[i] = add(r[i], mul(beta, add(p[i], mul(nomega, v[i])))); }
    __device__ void vec2(int i) const {
        const double2 vv = reinterpret_cast<const double2 *>(v)[i];
        const double2 rv = reinterpret_cast<const double2 *>(r)[i];
        double2 pv = reinterpret_cast<double2 *>(p)[i];
        pv.x = add(rv.x, mul(beta, add(pv.x, mul(nomega, vv.x))));
        pv.y = add(rv.y, mul(beta, add(pv.y, mul(nomega, vv.y))));
        reinterpret_cast<double2 *>(p)[i] = pv;
    }
};
struct ScatterAddF {        // y[idx[i]] += src[i]; idx holds no value twice (one neighbour's export list)
    const int *idx; const double *src; double *y;
    __device__ void one(int i) const { const int k = idx[i]; y[k] = add(y[k], src[i]); }
    __device__ void vec2(int i) const { one(2 * i); one(2 * i + 1); }
};
struct GatherF {
    const int *idx; const double *x; double *out;
    __device__ void one(int i) const { out[i] = x[idx[i]]; }
    __device__ void vec2(int i) const { one(2 * i); one(2 * i + 1); }
};

template <class F>
static int launch_ew(int n, const F &f, bool vec, void *stream)
{
    if (n <= 0) return 0;
    ew_kernel<F><<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(n, f, vec);
    LISB_CHECK_LAUNCH();
    return 0;
}

// ---- reductions ----------------------------------------------------------------------------
// kind: 0 dot, 1 sum x*x, 2 sum |x|, 3 max |x|, 4 sum x
template <int kKind>
__device__ __forceinline__ double red_term(double x, double y) {
    if (kKind == 0) return mul(x, y);
    if (kKind == 1) return mul(x, x);
    if (kKind == 2) return fabs(x);
    if (kKind == 3) return fabs(x);
    return x;
}

template <int kKind>
__global__ void __launch_bounds__(kRedThreads)
reduce_kernel(int n, const double *__restrict__ x, const double *__restrict__ y, bool vec,
              double *partial, unsigned int *counter, double *result)
{
    constexpr bool kMax = (kKind == 3);
    __shared__ double red[32];
    const int stride = gridDim.x * blockDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double a0 = 0.0, a1 = 0.0;
    if (vec) {
        const int n2 = n >> 1;
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        const double2 *y2 = reinterpret_cast<const double2 *>(y);
#pragma unroll 4
        for (int i = t; i < n2; i += stride) {
            const double2 xv = x2[i];
            double2 yv = xv;
            if (kKind == 0) yv = y2[i];
            a0 = combine<kMax>(a0, red_term<kKind>(xv.x, yv.x));
            a1 = combine<kMax>(a1, red_term<kKind>(xv.y, yv.y));
        }
        if (t == 0 && (n & 1)) a0 = combine<kMax>(a0, red_term<kKind>(x[n - 1], kKind == 0 ? y[n - 1] : 0.0));
    } else {
        for (int i = t; i < n; i += stride)
            a0 = combine<kMax>(a0, red_term<kKind>(x[i], kKind == 0 ? y[i] : 0.0));
    }
    double mine[1] = { block_reduce<kMax, kRedThreads>(combine<kMax>(a0, a1), red) };
    grid_finish<kMax, kRedThreads, 1>(mine, partial, counter, result, red);
}

// r[0] = <a,b>, r[1] = <a,a>: one pass over a; element -> thread map, accumulators and tree of
// reduce_kernel<0> / <1>, so both scalars carry the bits of the two separate reductions
__global__ void __launch_bounds__(kRedThreads)
dot2_kernel(int n, const double *__restrict__ a, const double *__restrict__ b, bool vec,
            double *partial, unsigned int *counter, double *result)
{
    __shared__ double red[32];
    const int stride = gridDim.x * blockDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
    if (vec) {
        const int n2 = n >> 1;
#pragma unroll 4
        for (int i = t; i < n2; i += stride) {
            const double2 av = reinterpret_cast<const double2 *>(a)[i];
            const double2 bv = reinterpret_cast<const double2 *>(b)[i];
            s0 = add(s0, mul(av.x, bv.x)); s1 = add(s1, mul(av.y, bv.y));
            q0 = add(q0, mul(av.x, av.x)); q1 = add(q1, mul(av.y, av.y));
        }
        if (t == 0 && (n & 1)) { s0 = add(s0, mul(a[n - 1], b[n - 1])); q0 = add(q0, mul(a[n - 1], a[n - 1])); }
    } else {
        for (int i = t; i < n; i += stride) {
            const double av = a[i], bv = b[i];
            s0 = add(s0, mul(av, bv));
            q0 = add(q0, mul(av, av));
        }
    }
    double mine[2];
    mine[0] = block_reduce<false, kRedThreads>(add(s0, s1), red);
    mine[1] = block_reduce<false, kRedThreads>(add(q0, q1), red);
    grid_finish<false, kRedThreads, 2>(mine, partial, counter, result, red);
}

// x += alpha*p ; r += (-alpha)*q ; rr = sum r*r
__global__ void __launch_bounds__(kRedThreads)
cg_update_kernel(int n, double alpha, const double *__restrict__ p, const double *__restrict__ q,
                 double *__restrict__ x, double *__restrict__ r, bool vec,
                 double *partial, unsigned int *counter, double *result)
{
    __shared__ double red[32];
    const int stride = gridDim.x * blockDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const double na = -alpha;
    double a0 = 0.0, a1 = 0.0;
    if (vec) {
        const int n2 = n >> 1;
#pragma unroll 2
        for (int i = t; i < n2; i += stride) {
            const double2 pv = reinterpret_cast<const double2 *>(p)[i];
            const double2 qv = reinterpret_cast<const double2 *>(q)[i];
            double2 xv = reinterpret_cast<double2 *>(x)[i];
            double2 rv = reinterpret_cast<double2 *>(r)[i];
            xv.x = add(xv.x, mul(alpha, pv.x)); xv.y = add(xv.y, mul(alpha, pv.y));
            rv.x = add(rv.x, mul(na, qv.x));    rv.y = add(rv.y, mul(na, qv.y));
            reinterpret_cast<double2 *>(x)[i] = xv;
            reinterpret_cast<double2 *>(r)[i] = rv;
            a0 = add(a0, mul(rv.x, rv.x)); a1 = add(a1, mul(rv.y, rv.y));
        }
        if (t == 0 && (n & 1)) {
            const int i = n - 1;
            x[i] = add(x[i], mul(alpha, p[i]));
            const double rv = add(r[i], mul(na, q[i]));
            r[i] = rv; a0 = add(a0, mul(rv, rv));
        }
    } else {
        for (int i = t; i < n; i += stride) {
            x[i] = add(x[i], mul(alpha, p[i]));
            const double rv = add(r[i], mul(na, q[i]));
            r[i] = rv; a0 = add(a0, mul(rv, rv));
        }
    }
    double mine[1] = { block_reduce<false, kRedThreads>(add(a0, a1), red) };
    grid_finish<false, kRedThreads, 1>(mine, partial, counter, result, red);
}

// cg_update_kernel and the jacobi_dot_kernel of the NEXT iteration in one pass over r:
//   x += alpha*p ; r += (-alpha)*q ; z = r*dinv ; result[0] = sum r*r ; result[1] = <r,z>
// (src/solver/lis_solver_cg.c:205-211, then :171-177 of the following iteration; lis_psolve_jacobi is
// a multiply, src/precon/lis_precon_jacobi.c:88-147).  Separately the two launches move 48 + 24 B per
// element, here 64, and the host waits once instead of twice.  Element -> thread map, accumulator
// pairs and trees are those of the two kernels: x, r, z and both scalars carry the same bits.
__global__ void __launch_bounds__(kRedThreads)
cg_update_jacobi_kernel(int n, double alpha, const double *__restrict__ p, const double *__restrict__ q,
                        double *__restrict__ x, double *__restrict__ r, const double *__restrict__ dinv,
                        double *__restrict__ z, bool vec, double *partial, unsigned int *counter, double *result)
{
    __shared__ double red[32];
    const int stride = gridDim.x * blockDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const double na = -alpha;
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    if (vec) {
        const int n2 = n >> 1;
#pragma unroll 2
        for (int i = t; i < n2; i += stride) {
            const double2 pv = reinterpret_cast<const double2 *>(p)[i];
            const double2 qv = reinterpret_cast<const double2 *>(q)[i];
            const double2 dv = reinterpret_cast<const double2 *>(dinv)[i];
            double2 xv = reinterpret_cast<double2 *>(x)[i];
            double2 rv = reinterpret_cast<double2 *>(r)[i];
            xv.x = add(xv.x, mul(alpha, pv.x)); xv.y = add(xv.y, mul(alpha, pv.y));
            rv.x = add(rv.x, mul(na, qv.x));    rv.y = add(rv.y, mul(na, qv.y));
            double2 zv; zv.x = mul(rv.x, dv.x); zv.y = mul(rv.y, dv.y);
            reinterpret_cast<double2 *>(x)[i] = xv;
            reinterpret_cast<double2 *>(r)[i] = rv;
            reinterpret_cast<double2 *>(z)[i] = zv;
            a0 = add(a0, mul(rv.x, rv.x)); a1 = add(a1, mul(rv.y, rv.y));
            b0 = add(b0, mul(rv.x, zv.x)); b1 = add(b1, mul(rv.y, zv.y));
        }
        if (t == 0 && (n & 1)) {
            const int i = n - 1;
            x[i] = add(x[i], mul(alpha, p[i]));
            const double rv = add(r[i], mul(na, q[i]));
            const double zv = mul(rv, dinv[i]);
            r[i] = rv; z[i] = zv;
            a0 = add(a0, mul(rv, rv)); b0 = add(b0, mul(rv, zv));
        }
    } else {
        for (int i = t; i < n; i += stride) {
            x[i] = add(x[i], mul(alpha, p[i]));
            const double rv = add(r[i], mul(na, q[i]));
            const double zv = mul(rv, dinv[i]);
            r[i] = rv; z[i] = zv;
            a0 = add(a0, mul(rv, rv)); b0 = add(b0, mul(rv, zv));
        }
    }
    double mine[2];
    mine[0] = block_reduce<false, kRedThreads>(add(a0, a1), red);
    mine[1] = block_reduce<false, kRedThreads>(add(b0, b1), red);
    grid_finish<false, kRedThreads, 2>(mine, partial, counter, result, red);
}

// One link of GMRES' modified Gram-Schmidt chain (src/solver/lis_solver_gmres.c:225-236) in one pass:
//   w += (scale * *alpha) * v        alpha = the previous link's <w,v>, still on the device
//   kNorm ? sum w*w : <w,u>          the next link's coefficient / the norm that ends the chain
// Unfused this is axpy (read v,w; write w) + dot (read w,u): 40 B per element; here 32 (24 for the
// norm).  Element -> thread map, accumulators and tree are those of reduce_kernel<0/1>, the update
// is AxpyDevF's: w and the scalar carry the same bits as the two separate launches.
template <bool kNorm>
__global__ void __launch_bounds__(kRedThreads)
mgs_step_kernel(int n, const double *__restrict__ alpha, double scale, const double *__restrict__ v,
                double *__restrict__ w, const double *__restrict__ u, bool vec,
                double *partial, unsigned int *counter, double *result)
{
    __shared__ double red[32];
    const int stride = gridDim.x * blockDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const double a = alpha ? mul(scale, *alpha) : scale;     // no device coefficient: `scale` is the coefficient (plain axpy)
    double a0 = 0.0, a1 = 0.0;
    if (vec) {
        const int n2 = n >> 1;
#pragma unroll 4
        for (int i = t; i < n2; i += stride) {
            const double2 vv = reinterpret_cast<const double2 *>(v)[i];
            double2 wv = reinterpret_cast<double2 *>(w)[i];
            wv.x = add(wv.x, mul(a, vv.x)); wv.y = add(wv.y, mul(a, vv.y));
            reinterpret_cast<double2 *>(w)[i] = wv;
            double2 uv = wv;
            if (!kNorm) uv = reinterpret_cast<const double2 *>(u)[i];
            a0 = add(a0, mul(wv.x, uv.x)); a1 = add(a1, mul(wv.y, uv.y));
        }
        if (t == 0 && (n & 1)) {
            const int i = n - 1;
            const double wv = add(w[i], mul(a, v[i]));
            w[i] = wv; a0 = add(a0, mul(wv, kNorm ? wv : u[i]));
        }
    } else {
        for (int i = t; i < n; i += stride) {
            const double wv = add(w[i], mul(a, v[i]));
            w[i] = wv; a0 = add(a0, mul(wv, kNorm ? wv : u[i]));
        }
    }
    double mine[1] = { block_reduce<false, kRedThreads>(add(a0, a1), red) };
    grid_finish<false, kRedThreads, 1>(mine, partial, counter, result, red);
}

// BiCGSTAB's closing updates (src/solver/lis_solver_bicgstab.c:272-279) in one pass:
//   x += alpha*phat ; x += omega*shat ; r += (-omega)*t ; rr = sum r*r
// 56 B per element instead of 80 for three axpys and a norm; bits of x, r and rr unchanged
// (the two x updates stay two rounded steps; thread map and tree of reduce_kernel<1>).
__global__ void __launch_bounds__(kRedThreads)
bicgstab_update_kernel(int n, double alpha, double omega, const double *__restrict__ phat, const double *__restrict__ shat,
                       const double *__restrict__ tv, double *__restrict__ x, double *__restrict__ r, bool vec,
                       double *partial, unsigned int *counter, double *result)
{
    __shared__ double red[32];
    const int stride = gridDim.x * blockDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const double no = -omega;
    double a0 = 0.0, a1 = 0.0;
    if (vec) {
        const int n2 = n >> 1;
#pragma unroll 2
        for (int i = t; i < n2; i += stride) {
            const double2 pv = reinterpret_cast<const double2 *>(phat)[i];
            const double2 sv = reinterpret_cast<const double2 *>(shat)[i];
            const double2 tt = reinterpret_cast<const double2 *>(tv)[i];
            double2 xv = reinterpret_cast<double2 *>(x)[i];
            double2 rv = reinterpret_cast<double2 *>(r)[i];
            xv.x = add(add(xv.x, mul(alpha, pv.x)), mul(omega, sv.x));
            xv.y = add(add(xv.y, mul(alpha, pv.y)), mul(omega, sv.y));
            rv.x = add(rv.x, mul(no, tt.x)); rv.y = add(rv.y, mul(no, tt.y));
            reinterpret_cast<double2 *>(x)[i] = xv;
            reinterpret_cast<double2 *>(r)[i] = rv;
            a0 = add(a0, mul(rv.x, rv.x)); a1 = add(a1, mul(rv.y, rv.y));
        }
        if (t == 0 && (n & 1)) {
            const int i = n - 1;
            x[i] = add(add(x[i], mul(alpha, phat[i])), mul(omega, shat[i]));
            const double rv = add(r[i], mul(no, tv[i]));
            r[i] = rv; a0 = add(a0, mul(rv, rv));
        }
    } else {
        for (int i = t; i < n; i += stride) {
            x[i] = add(add(x[i], mul(alpha, phat[i])), mul(omega, shat[i]));
            const double rv = add(r[i], mul(no, tv[i]));
            r[i] = rv; a0 = add(a0, mul(rv, rv));
        }
    }
    double mine[1] = { block_reduce<false, kRedThreads>(add(a0, a1), red) };
    grid_finish<false, kRedThreads, 1>(mine, partial, counter, result, red);
}

// z = r .* dinv ; rho = <r,z>
__global__ void __launch_bounds__(kRedThreads)
jacobi_dot_kernel(int n, const double *__restrict__ r, const double *__restrict__ dinv,
                  double *__restrict__ z, bool vec,
                  double *partial, unsigned int *counter, double *result)
{
    __shared__ double red[32];
    const int stride = gridDim.x * blockDim.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double a0 = 0.0, a1 = 0.0;
    if (vec) {
        const int n2 = n >> 1;
#pragma unroll 4
        for (int i = t; i < n2; i += stride) {
            const double2 rv = reinterpret_cast<const double2 *>(r)[i];
            const double2 dv = reinterpret_cast<const double2 *>(dinv)[i];
            double2 zv; zv.x = mul(rv.x, dv.x); zv.y = mul(rv.y, dv.y);
            reinterpret_cast<double2 *>(z)[i] = zv;
            a0 = add(a0, mul(rv.x, zv.x)); a1 = add(a1, mul(rv.y, zv.y));
        }
        if (t == 0 && (n & 1)) {
            const int i = n - 1;
            const double zv = mul(r[i], dinv[i]);
            z[i] = zv; a0 = add(a0, mul(r[i], zv));
        }
    } else {
        for (int i = t; i < n; i += stride) {
            const double zv = mul(r[i], dinv[i]);
            z[i] = zv; a0 = add(a0, mul(r[i], zv));
        }
    }
    double mine[1] = { block_reduce<false, kRedThreads>(add(a0, a1), red) };
    grid_finish<false, kRedThreads, 1>(mine, partial, counter, result, red);
}

static inline int red_grid(int n) {
    long long need = ((long long)n + 8 * kRedThreads - 1) / (8 * kRedThreads);
    long long cap = (long long)sm_count() * 8;
    if (cap > kRedMaxGrid) cap = kRedMaxGrid;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// d[i] = first entry of row i whose column is i, else 0
__global__ void __launch_bounds__(256)
csr_diag_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ idx,
                const double *__restrict__ val, double *__restrict__ d)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = 0.0;
    const int e = ptr[i + 1];
    for (int j = ptr[i]; j < e; ++j)
        if (idx[j] == i) { v = val[j]; break; }
    d[i] = v;
}

// first stored entry of row i whose column is i: value -= sigma   (A <- A - sigma I,
// src/matrix/lis_matrix_csr.c:565-603); rows without a stored diagonal are left alone
__global__ void __launch_bounds__(256)
csr_shift_diag_kernel(int n, const int *__restrict__ ptr, const int *__restrict__ idx, double *__restrict__ val, double sigma)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int e = ptr[i + 1];
    for (int j = ptr[i]; j < e; ++j)
        if (idx[j] == i) { val[j] = sub(val[j], sigma); break; }
}

}  // namespace lisb

using namespace lisb;

extern "C" int lisb200_sm_count(void) { return sm_count(); }
extern "C" const char *lisb200_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }
extern "C" int lisb200_reduce_slots(void) { return kReduceSlots; }

extern "C" int lisb200_copy(int n, const double *x, double *y, void *s)
{ return launch_ew(n, CopyF{x, y}, aligned16(x) && aligned16(y), s); }
extern "C" int lisb200_axpy(int n, double a, const double *x, double *y, void *s)
{ return launch_ew(n, AxpyF{a, x, y}, aligned16(x) && aligned16(y), s); }
extern "C" int lisb200_axpy_dev(int n, const double *d_alpha, double scale, const double *x, double *y, void *s)
{ return launch_ew(n, AxpyDevF{d_alpha, scale, x, y}, aligned16(x) && aligned16(y), s); }
extern "C" int lisb200_xpay(int n, const double *x, double a, double *y, void *s)
{ return launch_ew(n, XpayF{a, x, y}, aligned16(x) && aligned16(y), s); }
extern "C" int lisb200_axpyz(int n, double a, const double *x, const double *y, double *z, void *s)
{ return launch_ew(n, AxpyzF{a, x, y, z}, aligned16(x) && aligned16(y) && aligned16(z), s); }
extern "C" int lisb200_scale(int n, double a, double *x, void *s)
{ return launch_ew(n, ScaleF{a, x}, aligned16(x), s); }
extern "C" int lisb200_pmul(int n, const double *x, const double *y, double *z, void *s)
{ return launch_ew(n, PmulF{x, y, z}, aligned16(x) && aligned16(y) && aligned16(z), s); }
extern "C" int lisb200_pdiv(int n, const double *x, const double *y, double *z, void *s)
{ return launch_ew(n, PdivF{x, y, z}, aligned16(x) && aligned16(y) && aligned16(z), s); }
extern "C" int lisb200_set_all(int n, double a, double *x, void *s)
{ return launch_ew(n, SetF{a, x}, aligned16(x), s); }
extern "C" int lisb200_abs(int n, double *x, void *s)
{ return launch_ew(n, AbsF{x}, aligned16(x), s); }
extern "C" int lisb200_reciprocal(int n, double *x, void *s)
{ return launch_ew(n, RecipF{x}, aligned16(x), s); }
extern "C" int lisb200_shift(int n, double sigma, double *x, void *s)
{ return launch_ew(n, ShiftF{sigma, x}, aligned16(x), s); }
extern "C" int lisb200_swap(int n, double *x, double *y, void *s)
{ return launch_ew(n, SwapF{x, y}, aligned16(x) && aligned16(y), s); }
extern "C" int lisb200_gather(int count, const int *idx, const double *x, double *out, void *s)
{ return launch_ew(count, GatherF{idx, x, out}, false, s); }
extern "C" int lisb200_scatter_add(int count, const int *idx, const double *src, double *y, void *s)
{ return launch_ew(count, ScatterAddF{idx, src, y}, false, s); }

extern "C" int lisb200_reduce(int kind, int n, const double *x, const double *y,
                              double *partial, unsigned int *counter, double *result, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) {       // empty vector: the reference loops do not execute, the sum stays 0.0
        cudaError_t e = cudaMemsetAsync(result, 0, sizeof(double), st);
        return (int)e;
    }
    const int grid = red_grid(n);
    const bool vec = aligned16(x) && (kind != 0 || aligned16(y));
    switch (kind) {
    case 0: reduce_kernel<0><<<grid, kRedThreads, 0, st>>>(n, x, y, vec, partial, counter, result); break;
    case 1: reduce_kernel<1><<<grid, kRedThreads, 0, st>>>(n, x, x, vec, partial, counter, result); break;
    case 2: reduce_kernel<2><<<grid, kRedThreads, 0, st>>>(n, x, x, vec, partial, counter, result); break;
    case 3: reduce_kernel<3><<<grid, kRedThreads, 0, st>>>(n, x, x, vec, partial, counter, result); break;
    case 4: reduce_kernel<4><<<grid, kRedThreads, 0, st>>>(n, x, x, vec, partial, counter, result); break;
    default: return (int)cudaErrorInvalidValue;
    }
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_dot2(int n, const double *a, const double *b, double *partial,
                            unsigned int *counter, double *result2, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return (int)cudaMemsetAsync(result2, 0, 2 * sizeof(double), st);
    dot2_kernel<<<red_grid(n), kRedThreads, 0, st>>>(n, a, b, aligned16(a) && aligned16(b), partial, counter, result2);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_cg_update(int n, double alpha, const double *p, const double *q,
                                 double *x, double *r, double *partial, unsigned int *counter,
                                 double *rr, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return (int)cudaMemsetAsync(rr, 0, sizeof(double), st);
    const bool vec = aligned16(p) && aligned16(q) && aligned16(x) && aligned16(r);
    cg_update_kernel<<<red_grid(n), kRedThreads, 0, st>>>(n, alpha, p, q, x, r, vec, partial, counter, rr);
    LISB_CHECK_LAUNCH();
    return 0;
}

/* returns cudaErrorInvalidValue (nothing launched) when the pointers' 16-byte alignment is mixed: the two
 * separate kernels would then pick different accumulator layouts and the bits would differ */
extern "C" int lisb200_cg_update_jacobi(int n, double alpha, const double *p, const double *q, double *x, double *r,
                                        const double *dinv, double *z, double *partial, unsigned int *counter,
                                        double *rr_rho, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return (int)cudaMemsetAsync(rr_rho, 0, 2 * sizeof(double), st);
    const bool v1 = aligned16(p) && aligned16(q) && aligned16(x) && aligned16(r);      /* cg_update_kernel's choice */
    const bool v2 = aligned16(r) && aligned16(dinv) && aligned16(z);                   /* jacobi_dot_kernel's */
    if (v1 != v2) return (int)cudaErrorInvalidValue;
    cg_update_jacobi_kernel<<<red_grid(n), kRedThreads, 0, st>>>(n, alpha, p, q, x, r, dinv, z, v1, partial, counter, rr_rho);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_jacobi_dot(int n, const double *r, const double *dinv, double *z,
                                  double *partial, unsigned int *counter, double *rho, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return (int)cudaMemsetAsync(rho, 0, sizeof(double), st);
    const bool vec = aligned16(r) && aligned16(dinv) && aligned16(z);
    jacobi_dot_kernel<<<red_grid(n), kRedThreads, 0, st>>>(n, r, dinv, z, vec, partial, counter, rho);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_mgs_step(int norm, int n, const double *d_alpha, double scale, const double *v, double *w, const double *u,
                                double *partial, unsigned int *counter, double *result, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return (int)cudaMemsetAsync(result, 0, sizeof(double), st);
    /* the packed path only where the separate axpy and reduction would both take theirs */
    const bool vec = aligned16(v) && aligned16(w) && (norm || aligned16(u));
    if (!vec && (aligned16(w) && (norm || aligned16(u)))) return (int)cudaErrorInvalidValue;   /* mixed alignment: launch separately */
    if (norm) mgs_step_kernel<true><<<red_grid(n), kRedThreads, 0, st>>>(n, d_alpha, scale, v, w, w, vec, partial, counter, result);
    else mgs_step_kernel<false><<<red_grid(n), kRedThreads, 0, st>>>(n, d_alpha, scale, v, w, u, vec, partial, counter, result);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_bicgstab_p(int n, double omega, double beta, const double *v, const double *r, double *p, void *s)
{ return launch_ew(n, BicgstabPF{-omega, beta, v, r, p}, aligned16(v) && aligned16(r) && aligned16(p), s); }

extern "C" int lisb200_bicgstab_update(int n, double alpha, double omega, const double *phat, const double *shat, const double *t,
                                       double *x, double *r, double *partial, unsigned int *counter, double *rr, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return (int)cudaMemsetAsync(rr, 0, sizeof(double), st);
    const bool vec = aligned16(phat) && aligned16(shat) && aligned16(t) && aligned16(x) && aligned16(r);
    bicgstab_update_kernel<<<red_grid(n), kRedThreads, 0, st>>>(n, alpha, omega, phat, shat, t, x, r, vec, partial, counter, rr);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_csr_shift_diagonal(int n, const int *ptr, const int *idx, double *val, double sigma, void *stream)
{
    if (n <= 0) return 0;
    csr_shift_diag_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, ptr, idx, val, sigma);
    LISB_CHECK_LAUNCH();
    return 0;
}

extern "C" int lisb200_csr_get_diagonal(int n, const int *ptr, const int *idx, const double *val,
                                        double *d, void *stream)
{
    if (n <= 0) return 0;
    csr_diag_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, ptr, idx, val, d);
    LISB_CHECK_LAUNCH();
    return 0;
}
