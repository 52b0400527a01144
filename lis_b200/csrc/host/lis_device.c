/*
 * lis_device.c -- device runtime: stream, scratch, mapped scalars, managed vector storage.
 * See lis_device.h.  Plain C over the CUDA runtime C API.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "lis_device.h"
#include "lis_b200_kernels.h"

typedef struct {
    int probed, available, device;
    cudaStream_t stream;
    int busy;
    double *partial; size_t partial_slots;
    unsigned int *counter;
    double *h_scalar, *d_scalar;
    double *dev_scalars, *h_fetch;     /* device-resident reduction results + pinned landing zone */
    /* copy pipeline (lisd_pipe_*): one stream per PCIe direction + per-chunk events */
    cudaStream_t s_in, s_out;
    cudaEvent_t *ev; int nev;
    cudaEvent_t ev_fork, ev_join;
    double *stage_x, *stage_y; size_t stage_xn, stage_yn;      /* plain device staging of the pipelined product */
    cudaStream_t s_aux; cudaEvent_t ev_aux_fork, ev_aux_join;  /* second compute stream (interior rows during the halo exchange) */
} lisd_ctx_t;

static lisd_ctx_t g_ctx;
static void pool_clear(void);

static void lisd_probe(void)
{
    if (g_ctx.probed) return;
    g_ctx.probed = 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        g_ctx.available = 0;
        return;
    }
    int dev = 0;
    const char *lr = getenv("LOCAL_RANK");
    if (lr && *lr) dev = atoi(lr) % ndev;
    if (cudaSetDevice(dev) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaHostAlloc((void **)&g_ctx.h_scalar, LISD_NSCALARS * sizeof(double), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaHostGetDevicePointer((void **)&g_ctx.d_scalar, g_ctx.h_scalar, 0) != cudaSuccess) { cudaGetLastError(); return; }
    memset(g_ctx.h_scalar, 0, LISD_NSCALARS * sizeof(double));
    if (cudaMalloc((void **)&g_ctx.dev_scalars, LISD_NSCALARS * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaHostAlloc((void **)&g_ctx.h_fetch, LISD_NSCALARS * sizeof(double), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaMalloc((void **)&g_ctx.counter, 64) != cudaSuccess) { cudaGetLastError(); return; }
    cudaMemset(g_ctx.counter, 0, 64);
    g_ctx.device = dev;
    g_ctx.available = 1;
}

int lisd_available(void) { lisd_probe(); return g_ctx.available; }
int lisd_device_id(void) { lisd_probe(); return g_ctx.device; }

LIS_INT lisd_require(const char *what)
{
    if (lisd_available()) return LIS_SUCCESS;
    LIS_SETERR1(LIS_ERR_DEVICE, "%s needs a CUDA device (sm_100a); lis_b200 has no CPU compute path\n", what);
    return LIS_ERR_DEVICE;
}

void *lisd_stream(void) { lisd_probe(); return (void *)g_ctx.stream; }
void *lis_b200_stream(void) { return lisd_stream(); }
void lisd_mark_busy(void) { g_ctx.busy = 1; }

LIS_INT lisd_check(int rc, const char *what)
{
    if (rc == 0) return LIS_SUCCESS;
    if (rc == (int)cudaErrorMemoryAllocation) {
        LIS_SETERR1(LIS_ERR_OUT_OF_MEMORY, "%s: out of device memory\n", what);
        return LIS_ERR_OUT_OF_MEMORY;
    }
    LIS_SETERR2(LIS_ERR_DEVICE, "%s: CUDA error: %s\n", what, lisb200_error_string(rc));
    return LIS_ERR_DEVICE;
}

LIS_INT lisd_sync(void)
{
    if (!g_ctx.available || !g_ctx.busy) return LIS_SUCCESS;
    g_ctx.busy = 0;
    LIS_INT err = lisd_check((int)cudaStreamSynchronize(g_ctx.stream), "stream synchronize");
    if (!err && lisd_p2p_error()) {
        LIS_SETERR(LIS_ERR_DEVICE, "row-partitioned product: a neighbour's halo never arrived (ranks out of step?)\n");
        return LIS_ERR_DEVICE;
    }
    return err;
}

void lisd_shutdown(void)
{
    if (!g_ctx.available) return;
    cudaStreamSynchronize(g_ctx.stream);
    pool_clear();
    if (g_ctx.partial) cudaFree(g_ctx.partial);
    if (g_ctx.counter) cudaFree(g_ctx.counter);
    if (g_ctx.h_scalar) cudaFreeHost(g_ctx.h_scalar);
    if (g_ctx.dev_scalars) cudaFree(g_ctx.dev_scalars);
    if (g_ctx.h_fetch) cudaFreeHost(g_ctx.h_fetch);
    if (g_ctx.s_aux) {
        cudaStreamSynchronize(g_ctx.s_aux);
        cudaEventDestroy(g_ctx.ev_aux_fork); cudaEventDestroy(g_ctx.ev_aux_join); cudaStreamDestroy(g_ctx.s_aux);
        g_ctx.s_aux = NULL; g_ctx.ev_aux_fork = NULL; g_ctx.ev_aux_join = NULL;
    }
    if (g_ctx.stage_x) cudaFree(g_ctx.stage_x);
    if (g_ctx.stage_y) cudaFree(g_ctx.stage_y);
    g_ctx.stage_x = g_ctx.stage_y = NULL; g_ctx.stage_xn = g_ctx.stage_yn = 0;
    for (int i = 0; i < g_ctx.nev; i++) cudaEventDestroy(g_ctx.ev[i]);
    free(g_ctx.ev);
    g_ctx.ev = NULL; g_ctx.nev = 0;
    if (g_ctx.s_in) {
        cudaEventDestroy(g_ctx.ev_fork); cudaEventDestroy(g_ctx.ev_join); cudaStreamDestroy(g_ctx.s_in); cudaStreamDestroy(g_ctx.s_out);
        g_ctx.s_in = g_ctx.s_out = NULL;
    }
    cudaStreamDestroy(g_ctx.stream);
    memset(&g_ctx, 0, sizeof(g_ctx));
}

/* ---- memory ----------------------------------------------------------------------------
 * Vector storage is managed memory (the host may touch v->value between API calls, exactly as
 * with the reference).  Two things keep it at HBM speed: a fresh block is populated on the
 * device with one bulk prefetch instead of GPU page faults, and destroyed vectors go to a
 * small size-keyed pool, because lis_solve creates and destroys its work vectors on every
 * call (src/solver/lis_solver.c:828,909) and cudaMallocManaged/cudaFree are device-wide
 * synchronising calls costing milliseconds each. */
#define LISD_POOL_SLOTS 64
static struct { void *p; size_t bytes; } g_pool[LISD_POOL_SLOTS];
static size_t g_pool_bytes = 0;
static const size_t g_pool_cap = (size_t)48 << 30;      /* at most 48 GB parked */

static void pool_clear(void)
{
    for (int i = 0; i < LISD_POOL_SLOTS; i++)
        if (g_pool[i].p) { cudaFree(g_pool[i].p); g_pool[i].p = NULL; g_pool[i].bytes = 0; }
    g_pool_bytes = 0;
}

LIS_INT lisd_alloc_vector(size_t count, LIS_SCALAR **value, LIS_INT *managed)
{
    size_t bytes = (count > 0 ? count : 1) * sizeof(LIS_SCALAR);
    bytes = (bytes + 255) & ~(size_t)255;
    *value = NULL;
    if (lisd_available()) {
        void *p = NULL;
        for (int i = 0; i < LISD_POOL_SLOTS && !p; i++)
            if (g_pool[i].p && g_pool[i].bytes == bytes) { p = g_pool[i].p; g_pool[i].p = NULL; g_pool_bytes -= bytes; }
        static int force_device = -1;
        if (force_device < 0) { const char *fv = getenv("LIS_B200_VECTORS"); force_device = (fv && strcmp(fv, "device") == 0) ? 1 : 0; }
        if (p == NULL && force_device) {
            /* LIS_B200_VECTORS=device: plain device memory (experiments; v->value is then reachable through the API only) */
            if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); LIS_SETERR_MEM(bytes); return LIS_ERR_OUT_OF_MEMORY; }
            *value = (LIS_SCALAR *)p; *managed = 2;
            return LIS_SUCCESS;
        }
        if (p == NULL) {
            cudaError_t e = cudaMallocManaged(&p, bytes, cudaMemAttachGlobal);
            if (e != cudaSuccess) {
                cudaGetLastError();
                pool_clear();                                  /* give parked blocks back and retry once */
                e = cudaMallocManaged(&p, bytes, cudaMemAttachGlobal);
            }
            if (e != cudaSuccess) {
                /* managed memory refused (seen with 8 processes on one node once NCCL has set up its
                 * collectives): fall back to plain device memory.  Such a vector is reachable from
                 * the host through the API only (get/set_value(s), scatter/gather), not by
                 * dereferencing v->value */
                static int warned = 0;
                size_t fr = 0, tot = 0;
                const char *why = lisb200_error_string((int)e);
                cudaGetLastError();
                e = cudaMalloc(&p, bytes);
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    cudaMemGetInfo(&fr, &tot);
                    LIS_SETERR3(LIS_ERR_OUT_OF_MEMORY, "vector allocation of %D MiB failed: %s; device free %D MiB\n",
                                (LIS_INT)(bytes >> 20), lisb200_error_string((int)e), (LIS_INT)(fr >> 20));
                    return LIS_ERR_OUT_OF_MEMORY;
                }
                if (!warned) { warned = 1; fprintf(stderr, "lis_b200: cudaMallocManaged failed (%s); using device-only vector storage\n", why); }
                *value = (LIS_SCALAR *)p;
                *managed = 2;
                return LIS_SUCCESS;
            }
        }
        /* populate / bring the pages to HBM in one go (advisory: faults still work if it fails) */
        if (cudaMemPrefetchAsync(p, bytes, g_ctx.device, g_ctx.stream) != cudaSuccess) cudaGetLastError();
        g_ctx.busy = 1;
        *value = (LIS_SCALAR *)p;
        *managed = 1;
    } else {
        /* no device: keep the data-structure API usable (I/O, conversion, tests of the host
         * logic); every compute entry point still fails with LIS_ERR_DEVICE */
        *value = (LIS_SCALAR *)malloc(bytes);
        if (*value == NULL) { LIS_SETERR_MEM(bytes); return LIS_ERR_OUT_OF_MEMORY; }
        *managed = 0;
    }
    return LIS_SUCCESS;
}

void lisd_free_vector_bytes(LIS_SCALAR *value, LIS_INT managed, size_t count)
{
    if (value == NULL) return;
    if (!managed) { free(value); return; }
    if (managed == 2) { lisd_sync(); cudaFree(value); return; }
    size_t bytes = (count > 0 ? count : 1) * sizeof(LIS_SCALAR);
    bytes = (bytes + 255) & ~(size_t)255;
    if (g_ctx.available && g_pool_bytes + bytes <= g_pool_cap) {
        for (int i = 0; i < LISD_POOL_SLOTS; i++)
            if (g_pool[i].p == NULL) {
                /* work queued on the stream may still use the block; stream order protects the
                 * next owner, who only touches it through the same stream */
                g_pool[i].p = value; g_pool[i].bytes = bytes; g_pool_bytes += bytes;
                return;
            }
    }
    lisd_sync();
    cudaFree(value);
}

void lisd_free_vector(LIS_SCALAR *value, LIS_INT managed)
{
    if (value == NULL) return;
    if (managed) { lisd_sync(); cudaFree(value); }
    else free(value);
}

/* ---- arrays that are the public host arrays of a matrix AND its device mirror ------------------
 * lis_matrix_convert on the device writes its result into managed memory: Aout->index / Aout->value
 * point at it (the host may read it; pages migrate on demand), the mirror points at the same bytes,
 * nothing is downloaded or uploaded.  lis_free() recognises such blocks through lisd_shared_release. */
#define LISD_SHARED_MAX 1024
static void *g_shared[LISD_SHARED_MAX];
static int g_nshared = 0;

void *lisd_shared_alloc(size_t bytes)
{
    if (!lisd_available() || g_nshared >= LISD_SHARED_MAX) return NULL;
    void *p = NULL;
    if (bytes == 0) bytes = 16;
    if (cudaMallocManaged(&p, bytes, cudaMemAttachGlobal) != cudaSuccess) { cudaGetLastError(); return NULL; }
    if (cudaMemPrefetchAsync(p, bytes, g_ctx.device, g_ctx.stream) != cudaSuccess) cudaGetLastError();
    g_ctx.busy = 1;
    g_shared[g_nshared++] = p;
    return p;
}

int lisd_is_shared(const void *p)
{
    for (int i = 0; i < g_nshared; i++) if (g_shared[i] == p) return 1;
    return 0;
}

int lisd_shared_release(void *p)
{
    if (p == NULL) return 0;
    for (int i = 0; i < g_nshared; i++)
        if (g_shared[i] == p) {
            g_shared[i] = g_shared[--g_nshared];
            if (g_ctx.available) { lisd_sync(); cudaFree(p); }
            return 1;
        }
    return 0;
}

LIS_INT lisd_malloc(void **p, size_t bytes)
{
    *p = NULL;
    LIS_INT err = lisd_require("device allocation");
    if (err) return err;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); LIS_SETERR_MEM(bytes); return LIS_ERR_OUT_OF_MEMORY; }
    return LIS_SUCCESS;
}

void lisd_free(void *p) { if (p) { lisd_sync(); cudaFree(p); } }

LIS_INT lisd_upload(void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return LIS_SUCCESS;
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_ctx.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx.stream);
    return lisd_check((int)e, "host to device copy");
}

LIS_INT lisd_download(void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return LIS_SUCCESS;
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_ctx.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g_ctx.stream);
    g_ctx.busy = 0;
    return lisd_check((int)e, "device to host copy");
}

LIS_INT lisd_memset(void *dst, int byte, size_t bytes)
{
    if (bytes == 0) return LIS_SUCCESS;
    g_ctx.busy = 1;
    return lisd_check((int)cudaMemsetAsync(dst, byte, bytes, g_ctx.stream), "memset");
}

/* ---- copy pipeline ---------------------------------------------------------------------------
 * Host -> device on one stream, kernels on the main stream, device -> host on a third, chained
 * by events per chunk, so that both PCIe directions and the SMs work at the same time
 * (lis_b200_matvec_host).  Chunk c uses events 2c (its input has landed) and 2c+1 (its kernel
 * has finished). */
LIS_INT lisd_pipe_begin(int nchunks)
{
    cudaError_t e = cudaSuccess;
    if (g_ctx.s_in == NULL) {
        e = cudaStreamCreateWithFlags(&g_ctx.s_in, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g_ctx.s_out, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_ctx.ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_ctx.ev_join, cudaEventDisableTiming);
        if (e != cudaSuccess) { g_ctx.s_in = NULL; return lisd_check((int)e, "copy pipeline setup"); }
    }
    if (2 * nchunks > g_ctx.nev) {
        cudaEvent_t *nv = (cudaEvent_t *)realloc(g_ctx.ev, sizeof(cudaEvent_t) * (size_t)(2 * nchunks));
        if (nv == NULL) { LIS_SETERR_MEM(2 * nchunks * sizeof(cudaEvent_t)); return LIS_OUT_OF_MEMORY; }
        g_ctx.ev = nv;
        while (g_ctx.nev < 2 * nchunks) {
            e = cudaEventCreateWithFlags(&g_ctx.ev[g_ctx.nev], cudaEventDisableTiming);
            if (e != cudaSuccess) return lisd_check((int)e, "copy pipeline setup");
            g_ctx.nev++;
        }
    }
    /* the incoming copies overwrite a vector that work already queued on the main stream may
     * still read: the copy stream starts behind it */
    e = cudaEventRecord(g_ctx.ev_fork, g_ctx.stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_ctx.s_in, g_ctx.ev_fork, 0);
    return lisd_check((int)e, "copy pipeline");
}

/* Staging vectors in plain device memory (cudaMalloc): copies between pinned host memory and such
 * memory are asynchronous by specification, which vector storage (managed memory) does not
 * promise.  Cached; grown on demand. */
LIS_INT lisd_pipe_staging(size_t xcount, size_t ycount, double **xs, double **ys)
{
    if (xcount > g_ctx.stage_xn) {
        lisd_sync();
        if (g_ctx.stage_x) cudaFree(g_ctx.stage_x);
        g_ctx.stage_x = NULL; g_ctx.stage_xn = 0;
        if (cudaMalloc((void **)&g_ctx.stage_x, (xcount + 8) * sizeof(double)) != cudaSuccess) { cudaGetLastError(); LIS_SETERR_MEM(xcount * sizeof(double)); return LIS_ERR_OUT_OF_MEMORY; }
        g_ctx.stage_xn = xcount;
    }
    if (ycount > g_ctx.stage_yn) {
        lisd_sync();
        if (g_ctx.stage_y) cudaFree(g_ctx.stage_y);
        g_ctx.stage_y = NULL; g_ctx.stage_yn = 0;
        if (cudaMalloc((void **)&g_ctx.stage_y, (ycount + 8) * sizeof(double)) != cudaSuccess) { cudaGetLastError(); LIS_SETERR_MEM(ycount * sizeof(double)); return LIS_ERR_OUT_OF_MEMORY; }
        g_ctx.stage_yn = ycount;
    }
    *xs = g_ctx.stage_x; *ys = g_ctx.stage_y;
    return LIS_SUCCESS;
}

LIS_INT lisd_d2d(void *dst, const void *src, size_t bytes)
{
    if (bytes == 0) return LIS_SUCCESS;
    g_ctx.busy = 1;
    return lisd_check((int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, g_ctx.stream), "device copy");
}

LIS_INT lisd_pipe_h2d(int c, void *dst, const void *src, size_t bytes)
{
    cudaError_t e = bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_ctx.s_in) : cudaSuccess;
    if (e == cudaSuccess) e = cudaEventRecord(g_ctx.ev[2 * c], g_ctx.s_in);
    return lisd_check((int)e, "host to device copy");
}

LIS_INT lisd_pipe_wait_in(int c)
{
    return lisd_check((int)cudaStreamWaitEvent(g_ctx.stream, g_ctx.ev[2 * c], 0), "copy pipeline");
}

LIS_INT lisd_pipe_d2h(int c, void *dst, const void *src, size_t bytes)
{
    cudaError_t e = cudaEventRecord(g_ctx.ev[2 * c + 1], g_ctx.stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_ctx.s_out, g_ctx.ev[2 * c + 1], 0);
    if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_ctx.s_out);
    return lisd_check((int)e, "device to host copy");
}

/* the main stream continues behind the last outgoing copy; one lisd_sync() then covers all three */
LIS_INT lisd_pipe_end(void)
{
    cudaError_t e = cudaEventRecord(g_ctx.ev_join, g_ctx.s_out);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_ctx.stream, g_ctx.ev_join, 0);
    g_ctx.busy = 1;
    return lisd_check((int)e, "copy pipeline");
}

/* ---- second compute stream -----------------------------------------------------------------
 * lisd_aux_fork: work enqueued on the returned stream from now on starts behind everything the main
 * stream holds at this point; lisd_aux_join: the main stream continues behind it. */
LIS_INT lisd_aux_fork(void **stream)
{
    cudaError_t e = cudaSuccess;
    if (g_ctx.s_aux == NULL) {
        e = cudaStreamCreateWithFlags(&g_ctx.s_aux, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_ctx.ev_aux_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_ctx.ev_aux_join, cudaEventDisableTiming);
        if (e != cudaSuccess) { g_ctx.s_aux = NULL; return lisd_check((int)e, "second stream setup"); }
    }
    e = cudaEventRecord(g_ctx.ev_aux_fork, g_ctx.stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_ctx.s_aux, g_ctx.ev_aux_fork, 0);
    *stream = (void *)g_ctx.s_aux;
    return lisd_check((int)e, "second stream");
}

LIS_INT lisd_aux_join(void)
{
    cudaError_t e = cudaEventRecord(g_ctx.ev_aux_join, g_ctx.s_aux);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(g_ctx.stream, g_ctx.ev_aux_join, 0);
    g_ctx.busy = 1;
    return lisd_check((int)e, "second stream");
}

/* ---- vector residency ---------------------------------------------------------------------
 * value[] is managed memory.  Kernels run at full HBM speed once the pages are on the device;
 * we prefetch explicitly when the host may have touched the vector since the last kernel. */
LIS_INT lisd_vec_device(LIS_VECTOR v)
{
    if (!v->b200_managed) {
        LIS_SETERR(LIS_ERR_DEVICE, "vector storage is not device accessible\n");
        return LIS_ERR_DEVICE;
    }
    if (v->b200_resident || v->b200_managed == 2) return LIS_SUCCESS;
    size_t bytes = v->b200_capacity * sizeof(LIS_SCALAR);
    cudaError_t e = cudaMemPrefetchAsync(v->value, bytes, g_ctx.device, g_ctx.stream);
    if (e != cudaSuccess) cudaGetLastError();      /* advisory only: page faults still work */
    v->b200_resident = 1;
    g_ctx.busy = 1;
    return LIS_SUCCESS;
}

void lisd_vec_host(LIS_VECTOR v)
{
    lisd_sync();
    if (v->b200_managed == 2) return;          /* device-only storage: callers go through upload/download */
    if (v->b200_managed && v->b200_resident && g_ctx.available) {
        /* bring the whole vector back in one bulk migration instead of page faults */
        size_t bytes = v->b200_capacity * sizeof(LIS_SCALAR);
        if (cudaMemPrefetchAsync(v->value, bytes, cudaCpuDeviceId, g_ctx.stream) == cudaSuccess)
            cudaStreamSynchronize(g_ctx.stream);
        else
            cudaGetLastError();
    }
    v->b200_resident = 0;
}

/* Host view of a vector's values for the one-off host passes (I/O, the diagonal of a non-CSR matrix, print).
 * Managed / host storage: v->value itself, migrated to the host first.  Device-only storage (b200_managed == 2, the
 * fallback when managed memory is refused): a heap copy -- filled from the device when `load` is set -- that
 * lisd_vec_host_done writes back when `store` is set, and frees.  NULL (with the error set) when that copy fails. */
LIS_SCALAR *lisd_vec_host_view(LIS_VECTOR v, int load)
{
    if (v->b200_managed != 2) { lisd_vec_host(v); return v->value; }
    const size_t bytes = (size_t)(v->b200_capacity > 0 ? v->b200_capacity : 1) * sizeof(LIS_SCALAR);
    LIS_SCALAR *h = (LIS_SCALAR *)malloc(bytes);
    if (h == NULL) { LIS_SETERR_MEM(bytes); return NULL; }
    if (lisd_sync() != LIS_SUCCESS) { free(h); return NULL; }
    if (load && lisd_download(h, v->value, bytes) != LIS_SUCCESS) { free(h); return NULL; }
    return h;
}

LIS_INT lisd_vec_host_done(LIS_VECTOR v, LIS_SCALAR *view, int store)
{
    if (view == NULL || view == v->value) return LIS_SUCCESS;
    LIS_INT err = LIS_SUCCESS;
    if (store) err = lisd_upload(v->value, view, (size_t)(v->b200_capacity > 0 ? v->b200_capacity : 1) * sizeof(LIS_SCALAR));
    free(view);
    return err;
}

/* ---- reductions ------------------------------------------------------------------------ */
double *lisd_partial(size_t slots)
{
    if (slots < (size_t)lisb200_reduce_slots()) slots = (size_t)lisb200_reduce_slots();
    if (slots > g_ctx.partial_slots) {
        lisd_sync();
        if (g_ctx.partial) cudaFree(g_ctx.partial);
        g_ctx.partial = NULL; g_ctx.partial_slots = 0;
        if (cudaMalloc((void **)&g_ctx.partial, slots * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return NULL;
        }
        g_ctx.partial_slots = slots;
    }
    return g_ctx.partial;
}

unsigned int *lisd_counter(void) { return g_ctx.counter; }
double *lisd_dev_scalar(int slot) { return g_ctx.dev_scalars + slot; }

/* queue a copy of device scalar slots [first, first+count) to pinned memory; valid after the
 * next lisd_sync (or any later host-synchronous call); read with lisd_fetched() */
LIS_INT lisd_dev_scalars_fetch(int first, int count)
{
    if (count <= 0) return LIS_SUCCESS;
    g_ctx.busy = 1;
    return lisd_check((int)cudaMemcpyAsync(g_ctx.h_fetch + first, g_ctx.dev_scalars + first, sizeof(double) * (size_t)count,
                                           cudaMemcpyDeviceToHost, g_ctx.stream), "scalar read-back");
}
double lisd_fetched(int slot) { return ((volatile double *)g_ctx.h_fetch)[slot]; }
/* where a reduction kernel writes its scalar(s): mapped pinned host memory on one rank (the
 * host reads it right after the stream sync), a device buffer feeding ncclAllGather otherwise */
double *lisd_scalar_dev(int slot)
{
    if (lisd_reduce_uses_nccl()) return lisd_reduce_dev_buffer() + slot;
    return g_ctx.d_scalar + slot;
}
double lisd_scalar_get(int slot) { return ((volatile double *)g_ctx.h_scalar)[slot]; }
