/*
 * lis_comm.c -- process group for the row-partitioned (multi-GPU) path: one process per GPU.
 *
 * Replaces the MPI layer of the reference: lis_matrix_g2l (src/matrix/lis_matrix_mpi.c:55-420),
 * lis_commtable_create (:594-826), lis_send_recv (:834-954) and the one-scalar MPI_Allreduce of
 * every reduction (src/vector/lis_vector_ops.c:119,263).  Same semantics: contiguous 1-D row
 * partition (LIS_GET_ISIE or caller-given local sizes), columns relabelled local-then-halo
 * with halo slots ordered by global index, neighbour lists, halo values received straight
 * into x[n .. np), reductions combined in RANK ORDER so every rank gets the same bits.
 *
 * Two planes:
 *   control (host, setup + small scalars): a POSIX shared-memory segment shared by the ranks
 *     of one node; generation-counted allgather.  Works without a GPU (CPU tests).
 *   data (device): NCCL over NVLink -- grouped ncclSend/ncclRecv of the packed halo on the
 *     library stream, ncclAllGather of the per-rank reduction partials.  libnccl is dlopen'ed
 *     (no link-time dependency; inside a PyTorch process its bundled NCCL is reused).
 *
 * Bootstrap: lis_initialize reads RANK / WORLD_SIZE / LOCAL_RANK (and MASTER_PORT for the
 * segment name) from the environment, or the host program calls lis_b200_comm_attach(rank,
 * nranks, token) with a job-unique token.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <fcntl.h>
#include <dlfcn.h>
#include <sched.h>
#include <time.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <cuda_runtime_api.h>
#include "lis_device.h"
#include "lis_host.h"
#include "lis_b200_kernels.h"

#define LISC_MAXR 16
#define LISC_SLOT (1u << 20)                 /* bytes per rank per exchange round */
#define LISC_MAGIC 0x4c49534232303001ull
#define LISC_TIMEOUT_S 180.0

typedef struct {
    volatile uint64_t magic;
    volatile uint64_t token;
    volatile uint64_t epoch;
    volatile int32_t nranks;
    volatile uint64_t seq_in[LISC_MAXR * 8];     /* one cache line per rank */
    volatile uint64_t seq_out[LISC_MAXR * 8];
    unsigned char slot[LISC_MAXR][LISC_SLOT];
} lisc_shm_t;

/* ---- NCCL through dlopen ---- */
typedef struct { char internal[128]; } lisc_nccl_id;
typedef void *lisc_nccl_comm;
typedef int (*fn_ncclGetUniqueId)(lisc_nccl_id *);
typedef int (*fn_ncclCommInitRank)(lisc_nccl_comm *, int, lisc_nccl_id, int);
typedef int (*fn_ncclCommDestroy)(lisc_nccl_comm);
typedef int (*fn_ncclAllGather)(const void *, void *, size_t, int, lisc_nccl_comm, cudaStream_t);
typedef int (*fn_ncclSend)(const void *, size_t, int, int, lisc_nccl_comm, cudaStream_t);
typedef int (*fn_ncclRecv)(void *, size_t, int, int, lisc_nccl_comm, cudaStream_t);
typedef int (*fn_ncclGroup)(void);
typedef const char *(*fn_ncclGetErrorString)(int);
#define LISC_NCCL_DOUBLE 8                    /* ncclFloat64 */

static struct {
    int rank, nranks, attached;
    char name[96];
    lisc_shm_t *shm;
    uint64_t gen;
    /* data plane */
    void *dl;
    lisc_nccl_comm comm;
    int nccl_ok;
    fn_ncclGetUniqueId GetUniqueId; fn_ncclCommInitRank CommInitRank; fn_ncclCommDestroy CommDestroy;
    fn_ncclAllGather AllGather; fn_ncclSend Send; fn_ncclRecv Recv; fn_ncclGroup GroupStart, GroupEnd;
    fn_ncclGetErrorString GetErrorString;
    double *d_red, *d_all, *h_all;            /* reduction staging: device partials, gathered, pinned */
} g = { .rank = 0, .nranks = 1 };

int lisd_rank(void) { return g.rank; }
int lisd_nranks(void) { return g.nranks; }

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int spin_until(volatile uint64_t *p, uint64_t want)
{
    const double t0 = now_s();
    for (unsigned it = 0; *p < want; it++) {
        if ((it & 1023) == 1023) {
            sched_yield();
            if (now_s() - t0 > LISC_TIMEOUT_S) return -1;
        }
    }
    __sync_synchronize();
    return 0;
}

/* one exchange round: rank k contributes lens[k] (<= LISC_SLOT) bytes, which land at
 * out + offs[k] on every rank.  Generation counters: seq_in[k] = "k has published round g",
 * seq_out[k] = "k has finished copying round g" (nobody overwrites a slot before that). */
static LIS_INT shm_round(const void *mine, void *out, const size_t *lens, const size_t *offs)
{
    lisc_shm_t *s = g.shm;
    const uint64_t gen = ++g.gen;
    if (lens[g.rank] > LISC_SLOT) { LIS_SETERR(LIS_ERR_ILL_ARG, "control-plane message too large\n"); return LIS_ERR_ILL_ARG; }
    for (int k = 0; k < g.nranks; k++)
        if (spin_until(&s->seq_out[k * 8], gen - 1)) goto timeout;
    if (lens[g.rank]) memcpy(s->slot[g.rank], mine, lens[g.rank]);
    __sync_synchronize();
    s->seq_in[g.rank * 8] = gen;
    for (int k = 0; k < g.nranks; k++)
        if (spin_until(&s->seq_in[k * 8], gen)) goto timeout;
    for (int k = 0; k < g.nranks; k++)
        if (lens[k]) memcpy((char *)out + offs[k], s->slot[k], lens[k]);
    __sync_synchronize();
    s->seq_out[g.rank * 8] = gen;
    return LIS_SUCCESS;
timeout:
    LIS_SETERR(LIS_ERR_DEVICE, "process group: a rank did not arrive within the timeout\n");
    return LIS_ERR_DEVICE;
}

/* every rank contributes `len` bytes; out receives nranks blocks of `len` */
static LIS_INT shm_allgather(const void *mine, size_t len, void *out)
{
    size_t lens[LISC_MAXR], offs[LISC_MAXR];
    for (int k = 0; k < g.nranks; k++) { lens[k] = len; offs[k] = (size_t)k * len; }
    return shm_round(mine, out, lens, offs);
}

/* variable-length allgather: lens[k] bytes from rank k land at out + offs[k]; any size */
static LIS_INT shm_allgatherv(const void *mine, void *out, const size_t *lens, const size_t *offs)
{
    size_t maxlen = 0;
    for (int k = 0; k < g.nranks; k++) if (lens[k] > maxlen) maxlen = lens[k];
    for (size_t done = 0; done < maxlen; done += LISC_SLOT) {
        size_t cl[LISC_MAXR], co[LISC_MAXR];
        for (int k = 0; k < g.nranks; k++) {
            cl[k] = lens[k] > done ? (lens[k] - done < LISC_SLOT ? lens[k] - done : LISC_SLOT) : 0;
            co[k] = offs[k] + done;
        }
        LIS_INT err = shm_round((const char *)mine + (cl[g.rank] ? done : 0), out, cl, co);
        if (err) return err;
    }
    return LIS_SUCCESS;
}

static void nccl_load(void)
{
    if (g.dl) return;
    g.dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!g.dl) g.dl = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!g.dl) return;
#define LOAD(field, sym) g.field = (fn_##sym)dlsym(g.dl, #sym)
    LOAD(GetUniqueId, ncclGetUniqueId); LOAD(CommInitRank, ncclCommInitRank); LOAD(CommDestroy, ncclCommDestroy);
    LOAD(AllGather, ncclAllGather); LOAD(Send, ncclSend); LOAD(Recv, ncclRecv);
    g.GroupStart = (fn_ncclGroup)dlsym(g.dl, "ncclGroupStart"); g.GroupEnd = (fn_ncclGroup)dlsym(g.dl, "ncclGroupEnd");
    LOAD(GetErrorString, ncclGetErrorString);
#undef LOAD
}

static LIS_INT nccl_check(int rc, const char *what)
{
    if (rc == 0) return LIS_SUCCESS;
    LIS_SETERR2(LIS_ERR_DEVICE, "%s: NCCL error: %s\n", what, g.GetErrorString ? g.GetErrorString(rc) : "?");
    return LIS_ERR_DEVICE;
}

static LIS_INT attach(int rank, int nranks, uint64_t token, int token_is_unique)
{
    if (g.attached) return LIS_SUCCESS;
    if (nranks <= 1) { g.rank = 0; g.nranks = 1; return LIS_SUCCESS; }
    if (nranks > LISC_MAXR || rank < 0 || rank >= nranks) {
        LIS_SETERR2(LIS_ERR_ILL_ARG, "process group: rank %D of %D is not supported\n", rank, nranks);
        return LIS_ERR_ILL_ARG;
    }
    snprintf(g.name, sizeof(g.name), "/lisb200_%u_%016llx", (unsigned)getuid(), (unsigned long long)token);
    const uint64_t start = (uint64_t)time(NULL);
    if (rank == 0) {
        shm_unlink(g.name);                                  /* leftovers of a crashed job */
        const int fd = shm_open(g.name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)sizeof(lisc_shm_t)) != 0) {
            LIS_SETERR1(LIS_ERR_DEVICE, "process group: cannot create shared segment %s\n", g.name);
            return LIS_ERR_DEVICE;
        }
        g.shm = (lisc_shm_t *)mmap(NULL, sizeof(lisc_shm_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (g.shm == MAP_FAILED) { g.shm = NULL; LIS_SETERR(LIS_ERR_DEVICE, "process group: mmap failed\n"); return LIS_ERR_DEVICE; }
        g.shm->token = token; g.shm->nranks = nranks; g.shm->epoch = start;
        __sync_synchronize();
        g.shm->magic = LISC_MAGIC;
    } else {
        /* open, map and validate; a segment that is not (yet) the one rank 0 of THIS job made --
         * missing, half initialised, or left behind by a crashed job that used the same
         * environment-derived name -- is dropped and looked up again */
        const double t0 = now_s();
        for (;;) {
            const int fd = shm_open(g.name, O_RDWR, 0600);
            if (fd >= 0) {
                struct stat st;
                lisc_shm_t *m = NULL;
                if (fstat(fd, &st) == 0 && (size_t)st.st_size >= sizeof(lisc_shm_t))
                    m = (lisc_shm_t *)mmap(NULL, sizeof(lisc_shm_t), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
                close(fd);
                if (m && m != MAP_FAILED) {
                    const double t1 = now_s();
                    while (m->magic != LISC_MAGIC && now_s() - t1 < 0.5) usleep(500);
                    if (m->magic == LISC_MAGIC && m->token == token && m->nranks == nranks &&
                        (token_is_unique || m->epoch + 600 >= start)) { g.shm = m; break; }
                    munmap(m, sizeof(lisc_shm_t));
                }
            }
            if (now_s() - t0 > LISC_TIMEOUT_S) {
                LIS_SETERR1(LIS_ERR_DEVICE, "process group: segment %s never became ready\n", g.name);
                return LIS_ERR_DEVICE;
            }
            usleep(2000);
        }
    }
    g.rank = rank; g.nranks = nranks; g.gen = 0; g.attached = 1;
    /* first exchange doubles as a barrier: afterwards the name is no longer needed */
    int all[LISC_MAXR];
    LIS_INT err = lisd_allgather_int(&rank, 1, all);
    if (err) return err;
    if (rank == 0) shm_unlink(g.name);

    /* data plane */
    if (lisd_available()) {
        /* point-to-point halo traffic and 8-byte gathers gain nothing from NVLS multicast; its setup
         * has been seen to break later cudaMallocManaged calls with 8 processes per node */
        setenv("NCCL_NVLS_ENABLE", "0", 0);
        /* LIS_B200_TRANSPORT=host: do not touch NCCL, stage the halo through host memory */
        const char *tr = getenv("LIS_B200_TRANSPORT");
        if (!(tr && strcmp(tr, "host") == 0)) nccl_load();
        lisc_nccl_id id;
        memset(&id, 0, sizeof(id));
        int have = g.dl && g.GetUniqueId && g.CommInitRank && g.AllGather && g.Send && g.Recv && g.GroupStart && g.GroupEnd;
        if (have && rank == 0) have = g.GetUniqueId(&id) == 0;
        lisc_nccl_id ids[LISC_MAXR];
        err = shm_allgather(&id, sizeof(id), ids);
        if (err) return err;
        int haves[LISC_MAXR];
        err = lisd_allgather_int(&have, 1, haves);
        if (err) return err;
        for (int k = 0; k < nranks; k++) have = have && haves[k];
        if (have) {
            err = nccl_check(g.CommInitRank(&g.comm, nranks, ids[0], rank), "ncclCommInitRank");
            if (err) return err;
            if (cudaMalloc((void **)&g.d_red, 8 * sizeof(double)) != cudaSuccess ||
                cudaMalloc((void **)&g.d_all, 8 * LISC_MAXR * sizeof(double)) != cudaSuccess ||
                cudaHostAlloc((void **)&g.h_all, 8 * LISC_MAXR * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                LIS_SETERR_MEM(8 * LISC_MAXR * sizeof(double));
                return LIS_ERR_OUT_OF_MEMORY;
            }
            g.nccl_ok = 1;
        } else if (rank == 0) {
            fprintf(stderr, "lis_b200: NCCL not available, halo exchange is staged through host memory\n");
        }
    }
    return LIS_SUCCESS;
}

LIS_INT lis_b200_comm_attach(LIS_INT rank, LIS_INT nranks, unsigned long long token)
{
    return attach((int)rank, (int)nranks, (uint64_t)token, 1);
}

LIS_INT lisd_comm_init(void)
{
    if (g.attached) return LIS_SUCCESS;
    const char *ws = getenv("WORLD_SIZE"), *rk = getenv("RANK");
    if (!ws || !rk || atoi(ws) <= 1) return LIS_SUCCESS;
    const char *port = getenv("MASTER_PORT"), *job = getenv("LIS_B200_JOB");
    uint64_t token = 1469598103934665603ull;
    for (const char *p = job ? job : (port ? port : "0"); *p; p++) token = (token ^ (unsigned char)*p) * 1099511628211ull;
    return attach(atoi(rk), atoi(ws), token, job != NULL);
}

void lisd_comm_finalize(void)
{
    if (!g.attached) return;
    if (g.nccl_ok) {
        cudaStreamSynchronize((cudaStream_t)lisd_stream());
        if (g.CommDestroy) g.CommDestroy(g.comm);
        cudaFree(g.d_red); cudaFree(g.d_all); cudaFreeHost(g.h_all);
        g.d_red = g.d_all = g.h_all = NULL;
        g.nccl_ok = 0;
    }
    if (g.shm) { munmap(g.shm, sizeof(lisc_shm_t)); g.shm = NULL; }
    g.attached = 0; g.rank = 0; g.nranks = 1;
}

/* ------------------------------------------------------------------ host collectives */
LIS_INT lisd_allgather_int(const int *mine, int count, int *all)
{
    if (g.nranks == 1) { memcpy(all, mine, sizeof(int) * (size_t)count); return LIS_SUCCESS; }
    return shm_allgather(mine, sizeof(int) * (size_t)count, all);
}

/* sum / max of `count` host scalars over the ranks, combined in rank order on every rank */
static LIS_INT allreduce_host(double *vals, int count, int is_max)
{
    if (g.nranks == 1) return LIS_SUCCESS;
    double all[LISC_MAXR * 8];
    if (count > 8) { LIS_SETERR(LIS_ERR_ILL_ARG, "too many scalars in one reduction\n"); return LIS_ERR_ILL_ARG; }
    LIS_INT err = shm_allgather(vals, sizeof(double) * (size_t)count, all);
    if (err) return err;
    for (int c = 0; c < count; c++) {
        double t = all[c];
        for (int k = 1; k < g.nranks; k++) {
            const double v = all[k * count + c];
            if (is_max) t = v > t ? v : t; else t = t + v;
        }
        vals[c] = t;
    }
    return LIS_SUCCESS;
}
LIS_INT lisd_allreduce_sum(double *vals, int count) { return allreduce_host(vals, count, 0); }
LIS_INT lisd_allreduce_max(double *vals, int count) { return allreduce_host(vals, count, 1); }
LIS_INT lis_b200_allreduce_sum(double *vals, LIS_INT count) { return allreduce_host(vals, (int)count, 0); }

/* device path of a reduction: the kernel left `count` partial scalars in lisd_scalar_dev();
 * NCCL all-gathers them over NVLink, one pinned copy brings the nranks x count table to the
 * host, the host folds it in rank order (same bits on every rank) */
/* Default: the scalars go through the host control plane (mapped scalar -> shm exchange, ~2 us,
 * no NCCL collective involved); LIS_B200_REDUCE=nccl selects the ncclAllGather route. */
static int g_reduce_mode = -1;
int lisd_reduce_uses_nccl(void)
{
    if (g_reduce_mode < 0) { const char *e = getenv("LIS_B200_REDUCE"); g_reduce_mode = (e && strcmp(e, "nccl") == 0) ? 1 : 0; }
    return g_reduce_mode == 1 && g.nranks > 1 && g.nccl_ok;
}
/* 0: host control plane (default), 1: ncclAllGather.  Call on every rank at the same point, with
 * no reduction in flight.  Returns the previous mode. */
LIS_INT lis_b200_set_reduce(LIS_INT nccl)
{
    lisd_sync();
    const int old = g_reduce_mode == 1;
    g_reduce_mode = nccl ? 1 : 0;
    return old;
}
double *lisd_reduce_dev_buffer(void) { return g.d_red; }

LIS_INT lisd_reduce_nccl_finish(double *vals, int count, int is_max)
{
    cudaStream_t st = (cudaStream_t)lisd_stream();
    LIS_INT err = nccl_check(g.AllGather(g.d_red, g.d_all, (size_t)count, LISC_NCCL_DOUBLE, g.comm, st), "ncclAllGather(reduction)");
    if (err) return err;
    cudaError_t e = cudaMemcpyAsync(g.h_all, g.d_all, sizeof(double) * (size_t)count * (size_t)g.nranks, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    err = lisd_check((int)e, "reduction read-back");
    if (err) return err;
    for (int c = 0; c < count; c++) {
        double t = g.h_all[c];
        for (int k = 1; k < g.nranks; k++) {
            const double v = g.h_all[k * count + c];
            if (is_max) t = v > t ? v : t; else t = t + v;
        }
        vals[c] = t;
    }
    return LIS_SUCCESS;
}

/* value[] holds this rank's slice at ranges[rank]; afterwards every rank holds all of it */
LIS_INT lisd_allgatherv_host(double *value, const LIS_INT *ranges, int nprocs)
{
    if (g.nranks == 1) return LIS_SUCCESS;
    size_t lens[LISC_MAXR], offs[LISC_MAXR];
    for (int k = 0; k < nprocs; k++) { lens[k] = sizeof(double) * (size_t)(ranges[k + 1] - ranges[k]); offs[k] = sizeof(double) * (size_t)ranges[k]; }
    double *mine = (double *)malloc(lens[g.rank] ? lens[g.rank] : 8);
    if (!mine) { LIS_SETERR_MEM(lens[g.rank]); return LIS_OUT_OF_MEMORY; }
    memcpy(mine, value + ranges[g.rank], lens[g.rank]);
    LIS_INT err = shm_allgatherv(mine, value, lens, offs);
    free(mine);
    return err;
}

/* ------------------------------------------------------------------ global -> local numbering */
static int cmp_int(const void *a, const void *b)
{
    const int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

static int owner_of(const LIS_INT *ranges, int nprocs, int gcol)
{
    int lo = 0, hi = nprocs;                  /* ranges[lo] <= gcol < ranges[hi] */
    while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (ranges[mid] <= gcol) lo = mid; else hi = mid; }
    return lo;
}

/* CSR with global columns -> local columns: owned columns become c - is, the others get halo
 * slots n, n+1, ... in ascending global order (lis_matrix_g2l_csr, :222-320) */
LIS_INT lisd_matrix_g2l(LIS_MATRIX A)
{
    if (A->nprocs == 1 || A->l2g_map != NULL) return LIS_SUCCESS;
    if (A->matrix_type != LIS_MATRIX_CSR) {
        LIS_SETERR(LIS_ERR_NOT_IMPLEMENTED, "row-partitioned matrices must be handed over in CSR (convert afterwards)\n");
        return LIS_ERR_NOT_IMPLEMENTED;
    }
    const LIS_INT n = A->n, is = A->is, ie = A->ie, nnz = A->ptr[n];
    LIS_INT nh = 0;
    for (LIS_INT j = 0; j < nnz; j++) if (A->index[j] < is || A->index[j] >= ie) nh++;
    int *halo = (int *)malloc(sizeof(int) * (size_t)(nh > 0 ? nh : 1));
    if (!halo) { LIS_SETERR_MEM(nh); return LIS_OUT_OF_MEMORY; }
    nh = 0;
    for (LIS_INT j = 0; j < nnz; j++) if (A->index[j] < is || A->index[j] >= ie) halo[nh++] = A->index[j];
    qsort(halo, (size_t)nh, sizeof(int), cmp_int);
    LIS_INT nu = 0;
    for (LIS_INT k = 0; k < nh; k++) if (k == 0 || halo[k] != halo[k - 1]) halo[nu++] = halo[k];
    A->l2g_map = (LIS_INT *)lis_malloc(sizeof(LIS_INT) * (size_t)(nu > 0 ? nu : 1), "lis_matrix_g2l::l2g_map");
    if (!A->l2g_map) { free(halo); LIS_SETERR_MEM(nu); return LIS_OUT_OF_MEMORY; }
    memcpy(A->l2g_map, halo, sizeof(int) * (size_t)nu);
    free(halo);
    for (LIS_INT j = 0; j < nnz; j++) {
        const LIS_INT c = A->index[j];
        if (c >= is && c < ie) A->index[j] = c - is;
        else {
            const int *hit = (const int *)bsearch(&c, A->l2g_map, (size_t)nu, sizeof(int), cmp_int);
            A->index[j] = n + (LIS_INT)(hit - A->l2g_map);
        }
    }
    A->np = n + nu;
    A->is_sorted = LIS_FALSE;
    lisd_matrix_drop(A);
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ communication table */
struct LIS_COMMTABLE_STRUCT {
    int nranks, rank;
    int n;                                    /* owned rows: halo values land at x[n ...] */
    int n_import, n_export;
    int import_ptr[LISC_MAXR + 1];            /* halo slots owned by rank k: [import_ptr[k], import_ptr[k+1]) */
    int export_ptr[LISC_MAXR + 1];            /* entries of export_index wanted by rank k */
    int *export_index;                        /* local row numbers, host */
    int *d_export_index;                      /* device copy */
    double *d_ws;                             /* packed send buffer, device */
    double *d_wr;                             /* receive buffer, device (plain cudaMalloc: NCCL never sees managed memory) */
    int *peer_export_ptr;                     /* [k*(nranks+1) + j]: export_ptr of rank k (host-staged transport) */
    int *peer_n_export;                       /* total exported entries of rank k */
    double *h_stage;                          /* host staging for the transport without NCCL */
    size_t h_stage_len;
    int neibpetot;                            /* number of ranks exchanged with (information) */
    /* in-kernel halo exchange over peer memory (p2p_prepare): 0 not looked at yet, 1 ready, -1 not available */
    int p2p;
    unsigned long long p2p_epoch;             /* products done through the table */
    double *p2p_inbox;                        /* [2][stride] doubles + [2][LISB200_P2P_MAX] flags: one exportable block (host/lis_peer.c) */
    size_t p2p_bytes; int p2p_fd; unsigned long long p2p_handle;
    void *p2p_peer_base[LISC_MAXR];           /* neighbours' inboxes mapped into this process */
    size_t p2p_peer_bytes[LISC_MAXR];
    unsigned int *p2p_push_count;
    lisb200_p2p *d_p2p;                       /* the kernel's table, device */
};

static void p2p_release(LIS_COMMTABLE t)
{
    for (int k = 0; k < LISC_MAXR; k++)
        if (t->p2p_peer_base[k]) { lisd_peer_unmap(t->p2p_peer_base[k], t->p2p_peer_bytes[k]); t->p2p_peer_base[k] = NULL; }
    if (t->p2p_inbox) lisd_peer_free(t->p2p_inbox, t->p2p_bytes, t->p2p_fd, t->p2p_handle);
    if (t->p2p_push_count) cudaFree(t->p2p_push_count);
    if (t->d_p2p) cudaFree(t->d_p2p);
    t->p2p_inbox = NULL; t->p2p_push_count = NULL; t->d_p2p = NULL; t->p2p_fd = -1;
    cudaGetLastError();
}

void lisd_commtable_destroy(LIS_COMMTABLE t)
{
    if (t == NULL) return;
    free(t->export_index);
    lisd_free(t->d_export_index);
    lisd_free(t->d_ws);
    lisd_free(t->d_wr);
    free(t->peer_export_ptr); free(t->peer_n_export); free(t->h_stage);
    if (t->p2p == 1) { lisd_sync(); p2p_release(t); }
    free(t);
}

static LIS_INT commtable_to_device(LIS_COMMTABLE t)
{
    LIS_INT err = LIS_SUCCESS;
    if (!lisd_available()) return LIS_SUCCESS;
    if (t->n_export) {
        err = lisd_malloc((void **)&t->d_export_index, sizeof(int) * (size_t)t->n_export);
        if (!err) err = lisd_upload(t->d_export_index, t->export_index, sizeof(int) * (size_t)t->n_export);
        if (!err) err = lisd_malloc((void **)&t->d_ws, sizeof(double) * (size_t)t->n_export);
    }
    if (!err && t->n_import) err = lisd_malloc((void **)&t->d_wr, sizeof(double) * (size_t)t->n_import);
    return err;
}

LIS_INT lisd_commtable_create(LIS_MATRIX A)
{
    if (A->nprocs == 1 || A->commtable) return LIS_SUCCESS;
    const int np_ = A->nprocs, me = A->my_rank;
    const LIS_INT nh = A->np - A->n;
    LIS_COMMTABLE t = (LIS_COMMTABLE)calloc(1, sizeof(struct LIS_COMMTABLE_STRUCT));
    if (!t) { LIS_SETERR_MEM(sizeof(struct LIS_COMMTABLE_STRUCT)); return LIS_OUT_OF_MEMORY; }
    t->nranks = np_; t->rank = me; t->n = A->n; t->n_import = nh;
    /* import side: my halo list is sorted by global index, i.e. grouped by owner */
    for (int k = 0; k <= np_; k++) t->import_ptr[k] = 0;
    for (LIS_INT h = 0; h < nh; h++) t->import_ptr[owner_of(A->ranges, np_, A->l2g_map[h]) + 1]++;
    for (int k = 0; k < np_; k++) t->import_ptr[k + 1] += t->import_ptr[k];
    /* export side: everybody publishes its halo list; I pick what falls into my rows */
    int counts[LISC_MAXR], mine = (int)nh;
    LIS_INT err = lisd_allgather_int(&mine, 1, counts);
    if (err) { free(t); return err; }
    size_t lens[LISC_MAXR], offs[LISC_MAXR], total = 0;
    for (int k = 0; k < np_; k++) { lens[k] = sizeof(int) * (size_t)counts[k]; offs[k] = total; total += lens[k]; }
    int *all = (int *)malloc(total ? total : 4);
    if (!all) { free(t); LIS_SETERR_MEM(total); return LIS_OUT_OF_MEMORY; }
    err = shm_allgatherv(A->l2g_map, all, lens, offs);
    if (err) { free(all); free(t); return err; }
    int nexp = 0;
    for (int k = 0; k < np_; k++) {
        const int *lst = (const int *)((const char *)all + offs[k]);
        t->export_ptr[k] = nexp;
        if (k == me) continue;
        for (int h = 0; h < counts[k]; h++) if (lst[h] >= A->is && lst[h] < A->ie) nexp++;
    }
    t->export_ptr[np_] = nexp;
    t->n_export = nexp;
    t->export_index = (int *)malloc(sizeof(int) * (size_t)(nexp > 0 ? nexp : 1));
    if (!t->export_index) { free(all); free(t); LIS_SETERR_MEM(nexp); return LIS_OUT_OF_MEMORY; }
    nexp = 0;
    for (int k = 0; k < np_; k++) {
        const int *lst = (const int *)((const char *)all + offs[k]);
        if (k == me) continue;
        for (int h = 0; h < counts[k]; h++) if (lst[h] >= A->is && lst[h] < A->ie) t->export_index[nexp++] = lst[h] - A->is;
    }
    free(all);
    for (int k = 0; k < np_; k++)
        if (k != me && (t->import_ptr[k + 1] > t->import_ptr[k] || t->export_ptr[k + 1] > t->export_ptr[k])) t->neibpetot++;
    /* every rank's export table: lets the host-staged transport find "what rank k packed for me" */
    t->peer_export_ptr = (int *)malloc(sizeof(int) * (size_t)np_ * (size_t)(np_ + 1));
    t->peer_n_export = (int *)malloc(sizeof(int) * (size_t)np_);
    if (!t->peer_export_ptr || !t->peer_n_export) { lisd_commtable_destroy(t); LIS_SETERR_MEM(np_ * np_); return LIS_OUT_OF_MEMORY; }
    err = lisd_allgather_int(t->export_ptr, np_ + 1, t->peer_export_ptr);
    if (err) { lisd_commtable_destroy(t); return err; }
    for (int k = 0; k < np_; k++) t->peer_n_export[k] = t->peer_export_ptr[k * (np_ + 1) + np_];
    err = commtable_to_device(t);
    if (err) { lisd_commtable_destroy(t); return err; }
    A->commtable = t;
    return LIS_SUCCESS;
}

LIS_INT lisd_commtable_duplicate(LIS_MATRIX Ain, LIS_MATRIX Aout)
{
    const LIS_COMMTABLE s = Ain->commtable;
    if (s == NULL) return LIS_SUCCESS;
    LIS_COMMTABLE t = (LIS_COMMTABLE)malloc(sizeof(struct LIS_COMMTABLE_STRUCT));
    if (!t) { LIS_SETERR_MEM(sizeof(struct LIS_COMMTABLE_STRUCT)); return LIS_OUT_OF_MEMORY; }
    memcpy(t, s, sizeof(*t));
    t->d_export_index = NULL; t->d_ws = NULL; t->d_wr = NULL;
    t->h_stage = NULL; t->h_stage_len = 0;
    t->p2p = 0; t->p2p_epoch = 0; t->p2p_inbox = NULL; t->p2p_push_count = NULL; t->d_p2p = NULL; t->p2p_fd = -1;
    memset(t->p2p_peer_base, 0, sizeof(t->p2p_peer_base));
    t->peer_export_ptr = (int *)malloc(sizeof(int) * (size_t)s->nranks * (size_t)(s->nranks + 1));
    t->peer_n_export = (int *)malloc(sizeof(int) * (size_t)s->nranks);
    if (!t->peer_export_ptr || !t->peer_n_export) { free(t->peer_export_ptr); free(t->peer_n_export); free(t); LIS_SETERR_MEM(s->nranks); return LIS_OUT_OF_MEMORY; }
    memcpy(t->peer_export_ptr, s->peer_export_ptr, sizeof(int) * (size_t)s->nranks * (size_t)(s->nranks + 1));
    memcpy(t->peer_n_export, s->peer_n_export, sizeof(int) * (size_t)s->nranks);
    t->export_index = (int *)malloc(sizeof(int) * (size_t)(s->n_export > 0 ? s->n_export : 1));
    if (!t->export_index) { free(t); LIS_SETERR_MEM(s->n_export); return LIS_OUT_OF_MEMORY; }
    memcpy(t->export_index, s->export_index, sizeof(int) * (size_t)s->n_export);
    LIS_INT err = commtable_to_device(t);
    if (err) { lisd_commtable_destroy(t); return err; }
    Aout->commtable = t;
    return LIS_SUCCESS;
}

/* sizes and index lists, for tests and diagnostics: out = {n_import, n_export, neighbours} */
LIS_INT lis_b200_commtable_info(LIS_MATRIX A, LIS_INT *out, LIS_INT *import_ptr, LIS_INT *export_ptr, LIS_INT *export_index,
                                LIS_INT *l2g_map, LIS_INT cap)
{
    const LIS_COMMTABLE t = A->commtable;
    for (LIS_INT k = 0; l2g_map && A->l2g_map && k < A->np - A->n && k < cap; k++) l2g_map[k] = A->l2g_map[k];
    if (t == NULL) { out[0] = out[1] = out[2] = 0; return LIS_SUCCESS; }
    out[0] = t->n_import; out[1] = t->n_export; out[2] = t->neibpetot;
    for (int k = 0; k <= t->nranks; k++) { if (import_ptr) import_ptr[k] = t->import_ptr[k]; if (export_ptr) export_ptr[k] = t->export_ptr[k]; }
    for (int k = 0; export_index && k < t->n_export && k < cap; k++) export_index[k] = t->export_index[k];
    return LIS_SUCCESS;
}

/* ------------------------------------------------------------------ halo exchange
 * pack ws[i] = x[export_index[i]] (one gather kernel), then one NCCL group of sends/receives;
 * received values land in a device buffer and are copied into x[n ...].  Asynchronous on the stream. */
/* transport without NCCL: packed entries -> pinned-less host staging -> control-plane allgatherv ->
 * pick the segments addressed to this rank -> halo part of x.  Host-synchronous; slower than the
 * NVLink path but independent of libnccl (and what the CPU-side multi-rank tests exercise). */
static LIS_INT halo_exchange_staged(LIS_COMMTABLE t, LIS_INT n, double *x)
{
    const int np_ = t->nranks, me = t->rank;
    cudaStream_t st = (cudaStream_t)lisd_stream();
    size_t lens[LISC_MAXR], offs[LISC_MAXR], total = 0;
    for (int k = 0; k < np_; k++) { lens[k] = sizeof(double) * (size_t)t->peer_n_export[k]; offs[k] = total; total += lens[k]; }
    const size_t need = total + sizeof(double) * ((size_t)t->n_export + (size_t)t->n_import + 2);
    if (need > t->h_stage_len) {
        free(t->h_stage);
        t->h_stage = (double *)malloc(need);
        t->h_stage_len = t->h_stage ? need : 0;
        if (!t->h_stage) { LIS_SETERR_MEM(need); return LIS_OUT_OF_MEMORY; }
    }
    double *all = t->h_stage, *mine = (double *)((char *)t->h_stage + total), *wr = mine + t->n_export + 1;
    LIS_INT err;
    if (t->n_export) {
        lisd_mark_busy();
        err = lisd_check(lisb200_gather(t->n_export, t->d_export_index, x, t->d_ws, st), "halo pack");
        if (!err) err = lisd_download(mine, t->d_ws, sizeof(double) * (size_t)t->n_export);
        if (err) return err;
    } else {
        err = lisd_sync();
        if (err) return err;
    }
    err = shm_allgatherv(mine, all, lens, offs);
    if (err) return err;
    for (int k = 0; k < np_; k++) {
        if (k == me) continue;
        const int ni = t->import_ptr[k + 1] - t->import_ptr[k];
        if (ni == 0) continue;
        const double *src = (const double *)((const char *)all + offs[k]) + t->peer_export_ptr[k * (np_ + 1) + me];
        memcpy(wr + t->import_ptr[k], src, sizeof(double) * (size_t)ni);
    }
    if (t->n_import) return lisd_upload(x + n, wr, sizeof(double) * (size_t)t->n_import);
    return LIS_SUCCESS;
}

static LIS_INT halo_exchange_raw(LIS_COMMTABLE t, LIS_INT n, double *x)
{
    if (t == NULL || g.nranks == 1) return LIS_SUCCESS;
    if (!g.nccl_ok) return halo_exchange_staged(t, n, x);
    cudaStream_t st = (cudaStream_t)lisd_stream();
    LIS_INT err;
    if (t->n_export) {
        lisd_mark_busy();
        err = lisd_check(lisb200_gather(t->n_export, t->d_export_index, x, t->d_ws, st), "halo pack");
        if (err) return err;
    }
    err = nccl_check(g.GroupStart(), "ncclGroupStart");
    if (err) return err;
    for (int k = 0; k < t->nranks; k++) {
        if (k == t->rank) continue;
        const int ne = t->export_ptr[k + 1] - t->export_ptr[k], ni = t->import_ptr[k + 1] - t->import_ptr[k];
        if (ne) { err = nccl_check(g.Send(t->d_ws + t->export_ptr[k], (size_t)ne, LISC_NCCL_DOUBLE, k, g.comm, st), "ncclSend"); if (err) { g.GroupEnd(); return err; } }
        if (ni) { err = nccl_check(g.Recv(t->d_wr + t->import_ptr[k], (size_t)ni, LISC_NCCL_DOUBLE, k, g.comm, st), "ncclRecv"); if (err) { g.GroupEnd(); return err; } }
    }
    lisd_mark_busy();
    err = nccl_check(g.GroupEnd(), "ncclGroupEnd");
    if (err) return err;
    /* unpack: the halo is contiguous in x (like the reference's copy of wr into x[n+pad..], :946-951) */
    if (t->n_import)
        err = lisd_check((int)cudaMemcpyAsync(x + n, t->d_wr, sizeof(double) * (size_t)t->n_import, cudaMemcpyDeviceToDevice, st), "halo unpack");
    return err;
}

/* ------------------------------------------------------------------ reverse halo reduction (lis_reduce,
 * src/matrix/lis_matrix_mpi.c:958-996): after y = A_loc^T x the entries y[n .. np) belong to rows other ranks
 * own.  Each goes back to its owner, which adds what it receives to y[export_index[..]], neighbour after
 * neighbour in rank order.  The segment rank k returns to me is its import segment for owner me: as long as my
 * export segment for k and in the same order. */
static int peer_import_offset(LIS_COMMTABLE t, int k, int owner)          /* import_ptr[owner] of rank k, from the export tables */
{
    const int np_ = t->nranks;
    int off = 0;
    for (int j = 0; j < owner; j++) off += t->peer_export_ptr[j * (np_ + 1) + k + 1] - t->peer_export_ptr[j * (np_ + 1) + k];
    return off;
}

static LIS_INT halo_reduce_apply(LIS_COMMTABLE t, double *y)
{
    cudaStream_t st = (cudaStream_t)lisd_stream();
    for (int k = 0; k < t->nranks; k++) {
        const int ne = t->export_ptr[k + 1] - t->export_ptr[k];
        if (k == t->rank || ne == 0) continue;
        lisd_mark_busy();
        LIS_INT err = lisd_check(lisb200_scatter_add(ne, t->d_export_index + t->export_ptr[k], t->d_ws + t->export_ptr[k], y, st), "halo reduce");
        if (err) return err;
    }
    return LIS_SUCCESS;
}

static LIS_INT halo_reduce_staged(LIS_COMMTABLE t, LIS_INT n, double *y)
{
    const int np_ = t->nranks, me = t->rank;
    size_t lens[LISC_MAXR], offs[LISC_MAXR], total = 0;
    for (int k = 0; k < np_; k++) { lens[k] = sizeof(double) * (size_t)peer_import_offset(t, k, np_); offs[k] = total; total += lens[k]; }
    const size_t need = total + sizeof(double) * ((size_t)t->n_export + (size_t)t->n_import + 2);
    if (need > t->h_stage_len) {
        free(t->h_stage);
        t->h_stage = (double *)malloc(need);
        t->h_stage_len = t->h_stage ? need : 0;
        if (!t->h_stage) { LIS_SETERR_MEM(need); return LIS_OUT_OF_MEMORY; }
    }
    double *all = t->h_stage, *mine = (double *)((char *)t->h_stage + total), *back = mine + t->n_import + 1;
    LIS_INT err = t->n_import ? lisd_download(mine, y + n, sizeof(double) * (size_t)t->n_import) : lisd_sync();
    if (err) return err;
    err = shm_allgatherv(mine, all, lens, offs);
    if (err) return err;
    for (int k = 0; k < np_; k++) {
        const int ne = t->export_ptr[k + 1] - t->export_ptr[k];
        if (k == me || ne == 0) continue;
        memcpy(back + t->export_ptr[k], (const double *)((const char *)all + offs[k]) + peer_import_offset(t, k, me), sizeof(double) * (size_t)ne);
    }
    if (t->n_export) { err = lisd_upload(t->d_ws, back, sizeof(double) * (size_t)t->n_export); if (err) return err; }
    return halo_reduce_apply(t, y);
}

LIS_INT lisd_halo_reduce_raw(LIS_MATRIX A, double *d_y)
{
    LIS_COMMTABLE t = A->commtable;
    const LIS_INT n = A->n;
    if (t == NULL || g.nranks == 1) return LIS_SUCCESS;
    if (!g.nccl_ok) return halo_reduce_staged(t, n, d_y);
    cudaStream_t st = (cudaStream_t)lisd_stream();
    LIS_INT err = LIS_SUCCESS;
    /* the halo part of y is contiguous and already grouped by owner; it goes through the plain device buffer
     * (vector storage is managed memory, which NCCL never sees) */
    if (t->n_import) {
        lisd_mark_busy();
        err = lisd_check((int)cudaMemcpyAsync(t->d_wr, d_y + n, sizeof(double) * (size_t)t->n_import, cudaMemcpyDeviceToDevice, st), "halo reduce pack");
        if (err) return err;
    }
    err = nccl_check(g.GroupStart(), "ncclGroupStart");
    if (err) return err;
    for (int k = 0; k < t->nranks; k++) {
        if (k == t->rank) continue;
        const int ne = t->export_ptr[k + 1] - t->export_ptr[k], ni = t->import_ptr[k + 1] - t->import_ptr[k];
        if (ni) { err = nccl_check(g.Send(t->d_wr + t->import_ptr[k], (size_t)ni, LISC_NCCL_DOUBLE, k, g.comm, st), "ncclSend"); if (err) { g.GroupEnd(); return err; } }
        if (ne) { err = nccl_check(g.Recv(t->d_ws + t->export_ptr[k], (size_t)ne, LISC_NCCL_DOUBLE, k, g.comm, st), "ncclRecv"); if (err) { g.GroupEnd(); return err; } }
    }
    lisd_mark_busy();
    err = nccl_check(g.GroupEnd(), "ncclGroupEnd");
    if (err) return err;
    return halo_reduce_apply(t, d_y);
}

/* public seam of the reference (src/matrix/lis_matrix_mpi.c:958): x has np entries */
LIS_INT lis_reduce(LIS_COMMTABLE commtable, LIS_SCALAR x[])
{
    if (commtable == NULL || g.nranks == 1) return LIS_SUCCESS;
    struct LIS_MATRIX_STRUCT fake;
    memset(&fake, 0, sizeof(fake));
    fake.commtable = commtable; fake.n = commtable->n;
    LIS_INT err = lisd_halo_reduce_raw(&fake, x);
    if (err) return err;
    return lisd_sync();
}

/* ------------------------------------------------------------------ halo exchange inside the SpMV kernel
 * One process per GPU; where every rank's GPU can map its neighbours' memory (all GPUs visible to every process,
 * NVLink peers, the CUDA virtual-memory API) the row-partitioned CSR product needs no pack kernel, no NCCL group and no
 * unpack copy: kernels/spmv.cu (csr_tma_kernel<.., kHalo>) stores the exported x entries straight into the
 * neighbours' inboxes, raises a flag there, runs the rows that read no halo entry, waits for the neighbours'
 * flags and reads the halo columns from its own inbox.  This file sets up what the kernel needs, once per
 * communication table: inbox + flags (one cuMemCreate block exported as a POSIX file descriptor, which travels to
 * the neighbours over a unix socket -- host/lis_peer.c), the neighbours' blocks imported and mapped, and the
 * table of addresses in device memory.  Anything missing -- a GPU hidden by CUDA_VISIBLE_DEVICES, no peer
 * access, an unsymmetric neighbour relation, LIS_B200_P2P=0 -- leaves the NCCL exchange in place. */
static struct { int probed, ok, enabled; int *h_error, *d_error; int peer_dev[LISC_MAXR]; } gp = { .enabled = -1 };

/* History (profiles/r02_session11.sh, 2 B200s): the first version mapped the inboxes with cudaIpcOpenMemHandle, which needs
 * cudaDeviceEnablePeerAccess -- and once a process has made that call EVERY kernel of it that reads cudaMalloc memory slows
 * down: the 512^3 CSR product took 2.49 ms instead of 2.10 ms whatever exchange was used.  The inbox is now one
 * virtual-memory-API allocation (host/lis_peer.c: cuMemCreate + POSIX-fd handle, imported and mapped by the neighbours):
 * nothing else becomes peer-accessible, as with NCCL's own buffers.  LIS_B200_P2P=0 / lis_b200_set_p2p(0) select the NCCL
 * exchange. */
static int p2p_enabled(void)
{
    if (gp.enabled < 0) { const char *e = getenv("LIS_B200_P2P"); gp.enabled = (e && e[0] == '0') ? 0 : 1; }
    return gp.enabled;
}

/* 1 (default): use the in-kernel exchange where it is available, 0: NCCL send/recv.  Returns the old setting.
 * Call on every rank alike. */
LIS_INT lis_b200_set_p2p(LIS_INT on) { const int old = p2p_enabled(); gp.enabled = on ? 1 : 0; return old; }

static int all_agree(int mine)
{
    int all[LISC_MAXR];
    if (lisd_allgather_int(&mine, 1, all)) return 0;
    for (int k = 0; k < g.nranks; k++) if (!all[k]) return 0;
    return 1;
}

static void p2p_probe(void)              /* collective */
{
    gp.probed = 1; gp.ok = 0;
    int want = g.nccl_ok && lisd_peer_available();
    char bus[32], all[LISC_MAXR][32];
    memset(bus, 0, sizeof(bus));
    if (want && cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), lisd_device_id()) != cudaSuccess) { cudaGetLastError(); want = 0; }
    if (shm_allgather(bus, sizeof(bus), all)) return;
    int ok = want;
    for (int k = 0; k < g.nranks && ok; k++) {
        int dev = -1;
        gp.peer_dev[k] = -1;
        if (k == g.rank) continue;
        if (all[k][0] == 0 || cudaDeviceGetByPCIBusId(&dev, all[k]) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }   /* hidden from this rank */
        if (dev == lisd_device_id() || !lisd_peer_can_access(dev)) { ok = 0; break; }
        gp.peer_dev[k] = dev;
    }
    if (ok && (cudaHostAlloc((void **)&gp.h_error, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
               cudaHostGetDevicePointer((void **)&gp.d_error, gp.h_error, 0) != cudaSuccess)) { cudaGetLastError(); ok = 0; }
    if (ok) *gp.h_error = 0;
    gp.ok = all_agree(ok);
    if (g.rank == 0 && getenv("LIS_B200_VERBOSE"))
        fprintf(stderr, "lis_b200: in-kernel halo exchange over peer memory %s\n", gp.ok ? "available" : "not available (NCCL send/recv)");
}

int lisd_p2p_error(void) { return gp.h_error && *(volatile int *)gp.h_error; }

typedef struct { long long stride; unsigned long long bytes; int ok; } p2p_msg;

/* collective: every rank comes here at its first CSR product on the table; local_ok: this rank's matrix takes the
 * TMA row-block kernel (the plan depends on the local row lengths, so the ranks must agree before any of them
 * switches to the in-kernel exchange) */
static void p2p_prepare(LIS_COMMTABLE t, int local_ok)
{
    const int me = t->rank, np_ = t->nranks;
    static unsigned serial = 0;
    char job[128];
    int sock = -1, got[LISC_MAXR];
    t->p2p = -1;
    if (!gp.probed) p2p_probe();
    int ok = gp.ok && local_ok, nn = 0;
    for (int k = 0; k < np_; k++) {
        const int ne = t->export_ptr[k + 1] - t->export_ptr[k], ni = t->import_ptr[k + 1] - t->import_ptr[k];
        got[k] = 0;
        if (k == me) continue;
        if ((ne > 0) != (ni > 0)) ok = 0;                     /* the double-buffer argument needs mutual neighbours */
        if (ne > 0) nn++;
    }
    if (nn == 0 || nn > LISB200_P2P_MAX) ok = 0;
    /* my inbox: one exportable block; a socket to receive the neighbours' descriptors on */
    snprintf(job, sizeof(job), "%.80s-%u", g.name[0] ? g.name + 1 : "job", ++serial);
    const long long stride = ((long long)t->n_import + 15) & ~15LL;
    const size_t bytes = sizeof(double) * (size_t)(2 * stride) + sizeof(unsigned long long) * 2 * LISB200_P2P_MAX + 64;
    void *inbox = NULL;
    if (ok) {
        if (!lisd_peer_alloc(bytes, &inbox, &t->p2p_bytes, &t->p2p_fd, &t->p2p_handle)) ok = 0;
        else {
            t->p2p_inbox = (double *)inbox;
            if (cudaMemset(inbox, 0, t->p2p_bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); ok = 0; }
        }
    }
    if (ok) { sock = lisd_fd_socket(job, me); if (sock < 0) ok = 0; }
    p2p_msg mine, all[LISC_MAXR];
    memset(&mine, 0, sizeof(mine));
    mine.stride = stride; mine.bytes = (unsigned long long)t->p2p_bytes; mine.ok = ok;
    if (shm_allgather(&mine, sizeof(mine), all)) ok = 0;      /* also the barrier: every socket is bound */
    for (int k = 0; k < np_; k++) if (!all[k].ok) ok = 0;
    if (ok) {
        for (int k = 0; k < np_ && ok; k++)
            if (k != me && t->export_ptr[k + 1] > t->export_ptr[k] && !lisd_fd_send(sock, job, k, me, t->p2p_fd)) ok = 0;
        for (int r = 0; r < nn && ok; r++) {
            int from = -1, fd = -1;
            if (!lisd_fd_recv(sock, &from, &fd, 60000) || from < 0 || from >= np_ || from == me || got[from]) { ok = 0; break; }
            got[from] = 1;
            if (!lisd_peer_import(fd, (size_t)all[from].bytes, &t->p2p_peer_base[from])) { t->p2p_peer_base[from] = NULL; ok = 0; }
            else t->p2p_peer_bytes[from] = (size_t)all[from].bytes;
            close(fd);
        }
    }
    lisb200_p2p tb;
    memset(&tb, 0, sizeof(tb));
    if (ok) {
        int s = 0;
        for (int k = 0; k < np_; k++) {
            if (k == me || t->export_ptr[k + 1] == t->export_ptr[k]) continue;
            if (!t->p2p_peer_base[k]) { ok = 0; break; }
            tb.exp_start[s] = t->export_ptr[k];
            tb.nbr_rank[s] = k;
            tb.peer_inbox[s] = (double *)t->p2p_peer_base[k] + peer_import_offset(t, k, me);
            tb.peer_stride[s] = all[k].stride;
            tb.peer_flag[s] = (unsigned long long *)((double *)t->p2p_peer_base[k] + 2 * all[k].stride) + me;
            s++;
        }
        tb.exp_start[s] = t->n_export;
        tb.n_nbr = s; tb.n_export = t->n_export; tb.export_index = t->d_export_index;
        tb.inbox = t->p2p_inbox; tb.inbox_stride = stride;
        tb.my_flag = (const unsigned long long *)(t->p2p_inbox + 2 * stride);
        tb.error = gp.d_error;
    }
    if (ok) {
        if (cudaMalloc((void **)&t->p2p_push_count, 64) != cudaSuccess || cudaMemset(t->p2p_push_count, 0, 64) != cudaSuccess ||
            cudaMalloc((void **)&t->d_p2p, sizeof(tb)) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        else {
            tb.push_count = t->p2p_push_count;
            if (cudaMemcpy(t->d_p2p, &tb, sizeof(tb), cudaMemcpyHostToDevice) != cudaSuccess) { cudaGetLastError(); ok = 0; }
        }
    }
    const int agreed = all_agree(ok);                         /* also: nobody closes a socket somebody still sends to */
    if (sock >= 0) close(sock);
    if (agreed) { t->p2p = 1; t->p2p_epoch = 0; return; }
    p2p_release(t);
}

/* undo p2p_prepare for A's communication table (collective; the next product with the exchange enabled sets it up again) */
LIS_INT lis_b200_p2p_release(LIS_MATRIX A)
{
    LIS_COMMTABLE t = A ? A->commtable : NULL;
    if (t == NULL) return LIS_SUCCESS;
    LIS_INT err = lisd_sync();
    if (t->p2p == 1) p2p_release(t);
    t->p2p = 0;
    all_agree(1);                                   /* nobody unmaps while a neighbour's kernel may still push */
    return err;
}

static unsigned long long g_p2p_products = 0;
/* products of this process that exchanged their halo inside the kernel (diagnostics, bench.py) */
unsigned long long lis_b200_p2p_products(void) { return g_p2p_products; }

/* the table for the fused product on A (device pointer) and the epoch of this product, or NULL: use lisd_halo_exchange */
const lisb200_p2p *lisd_p2p_begin(LIS_MATRIX A, int local_ok, unsigned long long *epoch)
{
    LIS_COMMTABLE t = A->commtable;
    if (t == NULL || g.nranks == 1 || !p2p_enabled()) return NULL;
    if (t->p2p == 0) p2p_prepare(t, local_ok);
    if (t->p2p != 1) return NULL;
    *epoch = ++t->p2p_epoch;
    g_p2p_products++;
    return t->d_p2p;
}

LIS_INT lisd_halo_exchange(LIS_MATRIX A, LIS_VECTOR x) { return halo_exchange_raw(A->commtable, A->n, x->value); }
LIS_INT lisd_halo_exchange_raw(LIS_MATRIX A, double *d_x) { return halo_exchange_raw(A->commtable, A->n, d_x); }

/* public seam of the reference (src/matrix/lis_matrix_mpi.c:834): x has np entries, the
 * halo is received into x[n .. np).  Host-synchronous like every public entry point. */
LIS_INT lis_send_recv(LIS_COMMTABLE commtable, LIS_SCALAR x[])
{
    if (commtable == NULL || g.nranks == 1) return LIS_SUCCESS;
    LIS_INT err = halo_exchange_raw(commtable, commtable->n, x);
    if (err) return err;
    return lisd_sync();
}
