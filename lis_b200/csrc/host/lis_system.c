/*
 * lis_system.c -- library lifetime, argument list, tracked allocation, error/printf, timer,
 * sorting helpers and the 1-D row partition.  Host C restatement of the slice of the
 * reference's src/system/ that the hot path and its drivers need:
 *   lis_init.c:122-232 (initialize/args), :401-472 (ranges), lis_memory.c:100-327,
 *   lis_error.c:118-205, lis_time.c:62-115, lis_sort.c.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <time.h>
#include <sys/time.h>
#include "lis_device.h"
#include "lis_host.h"

/* ---------------------------------------------------------------- tracked allocation
 * The reference keeps a linked list of its own allocations so that handles can be validated
 * (lis_is_malloc) and so that lis_free() can tell its blocks from caller-malloc'ed arrays it
 * adopted.  Here: an open-addressing pointer set (O(1) instead of an O(#allocations) walk). */
static void **g_set = NULL;
static size_t g_set_cap = 0, g_set_used = 0, g_set_tomb = 0;
#define TOMB ((void *)(size_t)1)

static size_t ptr_hash(const void *p, size_t cap)
{
    size_t h = (size_t)p;
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33;
    return h & (cap - 1);
}

static void set_insert_raw(void **tab, size_t cap, void *p)
{
    size_t i = ptr_hash(p, cap);
    while (tab[i] != NULL && tab[i] != TOMB) i = (i + 1) & (cap - 1);
    tab[i] = p;
}

static void set_grow(void)
{
    size_t ncap = g_set_cap ? g_set_cap * 2 : 1024;
    void **nt = (void **)calloc(ncap, sizeof(void *));
    if (!nt) return;
    for (size_t i = 0; i < g_set_cap; i++)
        if (g_set[i] != NULL && g_set[i] != TOMB) set_insert_raw(nt, ncap, g_set[i]);
    free(g_set);
    g_set = nt; g_set_cap = ncap; g_set_tomb = 0;
}

static void set_add(void *p)
{
    if ((g_set_used + g_set_tomb + 1) * 2 > g_set_cap) set_grow();
    if (!g_set) return;
    set_insert_raw(g_set, g_set_cap, p);
    g_set_used++;
}

static int set_find(const void *p, size_t *slot)
{
    if (!g_set || p == NULL) return 0;
    size_t i = ptr_hash(p, g_set_cap);
    while (g_set[i] != NULL) {
        if (g_set[i] == p) { if (slot) *slot = i; return 1; }
        i = (i + 1) & (g_set_cap - 1);
    }
    return 0;
}

void *lis_malloc(size_t size, char *tag)
{
    (void)tag;
    void *p = malloc(size ? size : 1);
    if (p) set_add(p);
    return p;
}

void *lis_calloc(size_t size, char *tag)
{
    (void)tag;
    void *p = calloc(size ? size : 1, 1);
    if (p) set_add(p);
    return p;
}

void *lis_realloc(void *p, size_t size)
{
    size_t slot;
    const int tracked = set_find(p, &slot);
    void *q = realloc(p, size ? size : 1);
    if (q == NULL) return NULL;
    if (tracked && q != p) { g_set[slot] = TOMB; g_set_used--; g_set_tomb++; set_add(q); }
    else if (!tracked) set_add(q);
    return q;
}

/* blocks that were adopted from the caller (plain malloc) fall through to free(), exactly
 * like src/system/lis_memory.c:198-219 */
void lis_free(void *p)
{
    size_t slot;
    if (p == NULL) return;
    if (lisd_shared_release(p)) return;           /* a converted matrix's arrays living in managed memory */
    if (set_find(p, &slot)) { g_set[slot] = TOMB; g_set_used--; g_set_tomb++; }
    free(p);
}

void lis_free2(LIS_INT n, ...)
{
    va_list ap;
    va_start(ap, n);
    for (LIS_INT i = 0; i < n; i++) {
        void *p = va_arg(ap, void *);
        if (p) lis_free(p);
    }
    va_end(ap);
}

LIS_INT lis_is_malloc(void *p) { return set_find(p, NULL) ? LIS_TRUE : LIS_FALSE; }

/* ---------------------------------------------------------------- errors and printing */
static const char *lis_code_name(LIS_INT code)
{
    switch (code) {
    case LIS_ERR_ILL_ARG: return "LIS_ERR_ILL_ARG";
    case LIS_BREAKDOWN: return "LIS_BREAKDOWN";
    case LIS_ERR_OUT_OF_MEMORY: return "LIS_ERR_OUT_OF_MEMORY";
    case LIS_MAXITER: return "LIS_MAXITER";
    case LIS_ERR_NOT_IMPLEMENTED: return "LIS_ERR_NOT_IMPLEMENTED";
    case LIS_ERR_FILE_IO: return "LIS_ERR_FILE_IO";
    case LIS_ERR_DEVICE: return "LIS_ERR_DEVICE";
    default: return "LIS_ERROR";
    }
}

static void expand_D(const char *in, char *out, size_t cap)
{
    size_t o = 0;
    for (size_t i = 0; in[i] && o + 2 < cap; i++) {
        if (in[i] == '%' && in[i + 1] == 'D') { out[o++] = '%'; out[o++] = 'd'; i++; }
        else out[o++] = in[i];
    }
    out[o] = 0;
}

LIS_INT lis_error(const char *file, const char *func, const LIS_INT line, const LIS_INT code, const char *mess, ...)
{
    char fmt[2048];
    va_list ap;
    expand_D(mess, fmt, sizeof(fmt));
    fprintf(stderr, "%s(%d) : %s : error %s : ", file, (int)line, func, lis_code_name(code));
    va_start(ap, mess);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    return code;
}

LIS_INT lis_printf(LIS_Comm comm, const char *mess, ...)
{
    (void)comm;
    if (lisd_rank() != 0) return LIS_SUCCESS;
    char fmt[2048];
    va_list ap;
    expand_D(mess, fmt, sizeof(fmt));
    va_start(ap, mess);
    vprintf(fmt, ap);
    va_end(ap);
    return LIS_SUCCESS;
}

/* call-depth trace behind the reference's LIS_DEBUG_FUNC_IN / _OUT macros (src/system/lis_error.c:67-95) */
LIS_INT lis_debug_trace_func(LIS_INT flag, char *func)
{
    static int depth = 0;
    if (flag) { lis_printf(LIS_COMM_WORLD, "%*s : %s\n", depth + 3, "IN ", func); depth++; }
    else { depth--; lis_printf(LIS_COMM_WORLD, "%*s : %s\n", depth + 3, "OUT", func); }
    return LIS_SUCCESS;
}

/* the reference's "the application owns MPI_Init/Finalize" switch (src/system/lis_init.c:99): nothing to own here */
void lis_do_not_handle_mpi(void) {}

void CHKERR(LIS_INT err)
{
    if (err) {
        lis_finalize();
        exit((int)err);
    }
}

/* ---------------------------------------------------------------- time */
double lis_wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1.0e-9 * (double)ts.tv_nsec;
}

void lis_date(char *date)
{
    time_t t = time(NULL);
    struct tm *tmv = localtime(&t);
    sprintf(date, "%04d/%02d/%02d %02d:%02d:%02d", tmv->tm_year + 1900, tmv->tm_mon + 1, tmv->tm_mday,
            tmv->tm_hour, tmv->tm_min, tmv->tm_sec);
}

/* ---------------------------------------------------------------- command-line options
 * lis_initialize keeps "-name value" pairs from argv; lis_solver_set_optionC replays them
 * into a solver (reference: src/system/lis_init.c:248-366, src/solver/lis_solver.c:1095). */
static lis_arg_t *g_args = NULL;
static int g_nargs = 0;
static int g_initialized = 0;

static int g_num_threads = 1;

const lis_arg_t *lis_host_args(int *count) { *count = g_nargs; return g_args; }
int lis_host_num_threads(void) { return g_num_threads; }

/* ---------------------------------------------------------------- host worker threads for one-off set-up passes
 * (sweep schedule, conversions): fn(lo, hi, ctx) over disjoint chunks of [0, count).  Not the emulated OpenMP thread
 * count above (that one decides block partitions and so results); this one only spreads work whose result does not
 * depend on it.  LIS_B200_HOST_THREADS overrides (1 = inline); small ranges run inline. */
#include <pthread.h>
#include <unistd.h>
typedef struct { void (*fn)(size_t, size_t, void *); void *ctx; size_t lo, hi; } host_chunk_t;
static void *host_chunk_run(void *p) { host_chunk_t *c = (host_chunk_t *)p; c->fn(c->lo, c->hi, c->ctx); return NULL; }

int lis_host_worker_count(void)
{
    const char *e = getenv("LIS_B200_HOST_THREADS");
    long t = e && e[0] >= '1' && e[0] <= '9' ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (t < 1) t = 1;
    if (t > 32) t = 32;
    return (int)t;
}

void lis_host_parallel_for(size_t count, size_t grain, void (*fn)(size_t lo, size_t hi, void *ctx), void *ctx)
{
    int nt = lis_host_worker_count();
    if (grain < 1) grain = 1;
    if ((size_t)nt > count / grain) nt = (int)(count / grain);
    if (nt <= 1) { if (count) fn(0, count, ctx); return; }
    pthread_t tid[32];
    host_chunk_t ch[32];
    int started[32];
    for (int k = 0; k < nt; k++) {
        ch[k].fn = fn; ch[k].ctx = ctx;
        ch[k].lo = count * (size_t)k / (size_t)nt;
        ch[k].hi = count * (size_t)(k + 1) / (size_t)nt;
        started[k] = k + 1 < nt && pthread_create(&tid[k], NULL, host_chunk_run, &ch[k]) == 0;
    }
    for (int k = 0; k < nt; k++) if (!started[k]) host_chunk_run(&ch[k]);       /* the last chunk, and any that got no thread */
    for (int k = 0; k < nt; k++) if (started[k]) pthread_join(tid[k], NULL);
}
void lis_host_set_num_threads(int n) { g_num_threads = n > 0 ? n : 1; }
LIS_INT lis_b200_set_num_threads(LIS_INT nthreads)
{
    const LIS_INT old = g_num_threads;
    lis_host_set_num_threads((int)nthreads);
    return old;
}

static void args_clear(void)
{
    for (int i = 0; i < g_nargs; i++) { free(g_args[i].name); free(g_args[i].value); }
    free(g_args);
    g_args = NULL; g_nargs = 0;
}

static char *lower_dup(const char *s)
{
    size_t n = strlen(s);
    char *r = (char *)malloc(n + 1);
    for (size_t i = 0; i <= n; i++) r[i] = (char)tolower((unsigned char)s[i]);
    return r;
}

LIS_INT lis_initialize(int *argc, char **argv[])
{
    args_clear();
    if (argc && argv && *argv) {
        g_args = (lis_arg_t *)calloc((size_t)(*argc > 0 ? *argc : 1), sizeof(lis_arg_t));
        for (int i = 1; i < *argc; i++) {
            const char *a = (*argv)[i];
            if (a && a[0] == '-' && i + 1 < *argc && !(a[1] >= '0' && a[1] <= '9') && a[1] != '.') {
                g_args[g_nargs].name = lower_dup(a + 1);
                g_args[g_nargs].value = lower_dup((*argv)[i + 1]);
                g_nargs++;
                i++;
            }
        }
    }
    /* -omp_num_threads N (src/system/lis_init.c:163-186): here the emulated thread count that
     * fixes the SSOR block partition; LIS_B200_NUM_THREADS in the environment is the default */
    {
        const char *e = getenv("LIS_B200_NUM_THREADS");
        if (e && atoi(e) > 0) lis_host_set_num_threads(atoi(e));
        for (int i = 0; i < g_nargs; i++)
            if (strcmp(g_args[i].name, "omp_num_threads") == 0 && atoi(g_args[i].value) > 0)
                lis_host_set_num_threads(atoi(g_args[i].value));
    }
    LIS_INT err = lisd_comm_init();
    if (err) return err;
    g_initialized = 1;
    return LIS_SUCCESS;
}

LIS_INT lis_finalize(void)
{
    lis_precon_register_free();
    lisd_sync();
    lisd_comm_finalize();
    lisd_shutdown();
    args_clear();
    g_initialized = 0;
    return LIS_SUCCESS;
}

/* ---------------------------------------------------------------- 1-D row partition */
LIS_INT lis_ranges_create(LIS_Comm comm, LIS_INT *local_n, LIS_INT *global_n, LIS_INT **ranges,
                          LIS_INT *is, LIS_INT *ie, LIS_INT *nprocs, LIS_INT *my_rank)
{
    (void)comm;
    const int np = lisd_nranks(), me = lisd_rank();
    *nprocs = np; *my_rank = me;
    if (np == 1) {
        if (*local_n == 0) *local_n = *global_n; else *global_n = *local_n;
        *is = 0; *ie = *local_n;
        *ranges = NULL;
        return LIS_SUCCESS;
    }
    LIS_INT *tr = (LIS_INT *)lis_malloc((size_t)(np + 1) * sizeof(LIS_INT), "lis_ranges_create::ranges");
    if (tr == NULL) { LIS_SETERR_MEM((np + 1) * sizeof(LIS_INT)); return LIS_ERR_OUT_OF_MEMORY; }
    int *all = (int *)malloc(sizeof(int) * (size_t)np);
    int mine = (int)*local_n;
    LIS_INT err = lisd_allgather_int(&mine, 1, all);
    if (err) { free(all); lis_free(tr); return err; }
    long long total = 0;
    for (int k = 0; k < np; k++) total += all[k];
    tr[0] = 0;
    if (total == 0) {                    /* nobody gave a local size: split global_n */
        for (int k = 0; k < np; k++) {
            LIS_INT s, e;
            LIS_GET_ISIE(k, np, *global_n, s, e);
            tr[k + 1] = e;
            if (k == me) { *is = s; *ie = e; }
        }
        *local_n = *ie - *is;
    } else {
        for (int k = 0; k < np; k++) tr[k + 1] = tr[k] + all[k];
        *global_n = tr[np];
        *is = tr[me]; *ie = tr[me + 1];
    }
    free(all);
    *ranges = tr;
    return LIS_SUCCESS;
}

/* ---------------------------------------------------------------- sorting helpers
 * Same contracts as src/system/lis_sort.c (inclusive index range [is, ie]); implemented as
 * introspection-free heap sorts so worst-case inputs cannot go quadratic. */
#define LIS_SIFT(ROOT, LIM, SWAP2, CMP)                                                     \
    {                                                                                       \
        LIS_INT root_ = (ROOT);                                                             \
        for (;;) {                                                                          \
            LIS_INT child_ = 2 * root_ + 1;                                                 \
            if (child_ >= (LIM)) break;                                                     \
            if (child_ + 1 < (LIM) && CMP(a[child_], a[child_ + 1])) child_++;              \
            if (!CMP(a[root_], a[child_])) break;                                           \
            { LIS_INT t_ = a[root_]; a[root_] = a[child_]; a[child_] = t_; SWAP2(root_, child_) } \
            root_ = child_;                                                                 \
        }                                                                                   \
    }
#define DEFINE_HEAPSORT(NAME, DECL2, SWAP2, CMP)                                            \
    void NAME(LIS_INT is, LIS_INT ie, LIS_INT *i1 DECL2)                                    \
    {                                                                                       \
        const LIS_INT n = ie - is + 1;                                                      \
        if (n < 2) return;                                                                  \
        LIS_INT *a = i1 + is;                                                               \
        for (LIS_INT start = n / 2 - 1; start >= 0; start--) LIS_SIFT(start, n, SWAP2, CMP) \
        for (LIS_INT end = n - 1; end > 0; end--) {                                         \
            { LIS_INT t_ = a[0]; a[0] = a[end]; a[end] = t_; SWAP2(0, end) }                \
            LIS_SIFT(0, end, SWAP2, CMP)                                                    \
        }                                                                                   \
    }

#define NO_DECL
#define NO_SWAP(i, j)
#define LESS(x, y) ((x) < (y))
#define GREATER(x, y) ((x) > (y))
DEFINE_HEAPSORT(lis_sort_i, NO_DECL, NO_SWAP, LESS)

#define D_DECL , LIS_SCALAR *d1
#define D_SWAP(i, j) { LIS_SCALAR *d_ = d1 + is; LIS_SCALAR s_ = d_[i]; d_[i] = d_[j]; d_[j] = s_; }

#define I_DECL , LIS_INT *i2
#define I_SWAP(i, j) { LIS_INT *b_ = i2 + is; LIS_INT s_ = b_[i]; b_[i] = b_[j]; b_[j] = s_; }
DEFINE_HEAPSORT(lis_sort_ii, I_DECL, I_SWAP, LESS)
DEFINE_HEAPSORT(lis_sortr_ii, I_DECL, I_SWAP, GREATER)

/* Keys ascending, the satellite carried along.  Rows of sparse matrices mostly arrive strictly ascending: one pass
 * and out.  Otherwise the partition scheme is the reference's (src/system/lis_sort.c:90-118: the middle element is
 * parked at the end and is the pivot value, a two-sided scan swaps out-of-place pairs, both sides are sorted the same
 * way), because among EQUAL keys the final order is a property of the scheme: a row that stores a column twice must end
 * with the same copy last as there -- CSR -> DIA / VBR keep that one.  Explicit stack, smaller side first. */
void lis_sort_id(LIS_INT is, LIS_INT ie, LIS_INT *i1, LIS_SCALAR *d1)
{
    if (ie <= is) return;
    LIS_INT k = is + 1;
    while (k <= ie && i1[k - 1] < i1[k]) k++;
    if (k > ie) return;
    LIS_INT lo_stack[96], hi_stack[96];
    int top = 0;
    lo_stack[0] = is; hi_stack[0] = ie; top = 1;
    while (top > 0) {
        LIS_INT lo = lo_stack[--top], hi = hi_stack[top];
        while (lo < hi) {
            const LIS_INT mid = (lo + hi) / 2;
            const LIS_INT pivot = i1[mid];
            { const LIS_INT t = i1[mid]; i1[mid] = i1[hi]; i1[hi] = t; }
            { const LIS_SCALAR t = d1[mid]; d1[mid] = d1[hi]; d1[hi] = t; }
            LIS_INT i = lo, j = hi;
            while (i <= j) {
                while (i1[i] < pivot) i++;
                while (i1[j] > pivot) j--;
                if (i <= j) {
                    const LIS_INT t = i1[i]; i1[i] = i1[j]; i1[j] = t;
                    const LIS_SCALAR u = d1[i]; d1[i] = d1[j]; d1[j] = u;
                    i++; j--;
                }
            }
            /* [lo, j] and [i, hi] remain; the larger one waits on the stack */
            if (j - lo < hi - i) {
                if (i < hi) { lo_stack[top] = i; hi_stack[top] = hi; top++; }
                hi = j;
            } else {
                if (lo < j) { lo_stack[top] = lo; hi_stack[top] = j; top++; }
                lo = i;
            }
        }
    }
}
