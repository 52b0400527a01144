/*
 * emu_runtime.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).  The fiber scheduler that runs
 * one CTA at a time, plus host-memory stand-ins for the handful of CUDA runtime calls the
 * product's host code and kernel launchers make.  "Device" allocations end flush against a
 * PROT_NONE guard page, so a kernel that reads or writes past the end of a buffer (beyond the
 * 16-byte granule) dies with SIGSEGV instead of passing silently.
 */
#include "cuda_emu.h"
#include <stdio.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
#include <map>
#include <set>
#include <string>
#include <vector>

namespace emu {

ThreadCtx *g_cur = nullptr;

enum State { kRunnable, kAtBlockBarrier, kAtWarpBarrier, kDone };

/* Context switch: on x86-64 a dozen instructions (callee-saved registers + stack pointer);
 * swapcontext() elsewhere -- it makes a sigprocmask system call per switch, which dominated the
 * run time of the emulated solver tests. */
#if defined(__x86_64__)
#define EMU_ASM_SWITCH 1
extern "C" void emu_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");
#endif

struct Fiber {
    ucontext_t ctx;
    void *sp = nullptr;
    char *stack = nullptr;
    ThreadCtx tc;
    State state = kDone;
    uint64_t shfl = 0;
};

static constexpr size_t kStackBytes = 256 * 1024;
static std::vector<Fiber *> g_fibers;          /* pool, grown on demand */
static ucontext_t g_sched;
static void *g_sched_sp = nullptr;
static Fiber *g_running = nullptr;
static const std::function<void()> *g_body = nullptr;
static std::vector<unsigned char> g_dyn_smem;
static int g_nthreads = 0;
static unsigned long long g_spins = 0;

[[noreturn]] void fail(const char *what)
{
    fprintf(stderr, "cuda_emu: %s", what);
    if (g_cur) fprintf(stderr, " (block %u, thread %u)", g_cur->bid.x, g_cur->tid.x);
    fprintf(stderr, "\n");
    abort();
}

static inline void switch_to_sched(Fiber *f)
{
#ifdef EMU_ASM_SWITCH
    emu_switch(&f->sp, g_sched_sp);
#else
    swapcontext(&f->ctx, &g_sched);
#endif
}
static inline void switch_to_fiber(Fiber *f)
{
#ifdef EMU_ASM_SWITCH
    emu_switch(&g_sched_sp, f->sp);
#else
    swapcontext(&g_sched, &f->ctx);
#endif
}

static void fiber_entry()
{
    (*g_body)();
    g_running->state = kDone;
    g_spins = 0;
    switch_to_sched(g_running);
    fail("a finished fiber was resumed");
}

static void to_scheduler()
{
    Fiber *f = g_running;
    switch_to_sched(f);
    g_cur = &f->tc;
}

void *dyn_smem() { return g_dyn_smem.data(); }

void sync_block() { g_running->state = kAtBlockBarrier; to_scheduler(); }
void sync_warp() { g_running->state = kAtWarpBarrier; to_scheduler(); }

void spin_yield()
{
    if (++g_spins > 200000000ull) fail("deadlock: threads keep spinning and nothing completes");
    to_scheduler();
}

uint64_t shfl_exchange(uint64_t mine, int src_lane)
{
    Fiber *f = g_running;
    f->shfl = mine;
    sync_warp();                                   /* everyone has published */
    const int tid = (int)f->tc.tid.x;
    const int base = tid & ~31;
    uint64_t v = mine;
    if (src_lane >= 0 && src_lane < 32 && base + src_lane < g_nthreads) v = g_fibers[base + src_lane]->shfl;
    sync_warp();                                   /* everyone has read before the next publish */
    return v;
}

static void run_block(dim3 grid, dim3 block, unsigned bx)
{
    const int nt = (int)(block.x * block.y * block.z);
    g_nthreads = nt;
    while ((int)g_fibers.size() < nt) {
        Fiber *f = new Fiber;
        f->stack = (char *)mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (f->stack == MAP_FAILED) fail("fiber stack allocation failed");
        g_fibers.push_back(f);
    }
    for (int t = 0; t < nt; ++t) {
        Fiber *f = g_fibers[t];
        f->tc.tid = make_uint3((unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y));
        f->tc.bid = make_uint3(bx, 0, 0);
        f->tc.bdim = block; f->tc.gdim = grid;
        f->state = kRunnable;
#ifdef EMU_ASM_SWITCH
        {
            /* six callee-saved registers, the entry address `ret` jumps to, one pad slot so that
             * the entry sees the stack alignment of a called function (rsp = 16k + 8) */
            void **top = (void **)(((uintptr_t)f->stack + kStackBytes) & ~(uintptr_t)15);
            void **sp = top - 8;
            for (int k = 0; k < 6; ++k) sp[k] = nullptr;
            sp[6] = (void *)fiber_entry;
            sp[7] = nullptr;
            f->sp = sp;
        }
#else
        getcontext(&f->ctx);
        f->ctx.uc_stack.ss_sp = f->stack;
        f->ctx.uc_stack.ss_size = kStackBytes;
        f->ctx.uc_link = &g_sched;
        makecontext(&f->ctx, fiber_entry, 0);
#endif
    }
    for (;;) {
        int live = 0;
        for (int t = 0; t < nt; ++t) {
            Fiber *f = g_fibers[t];
            if (f->state != kRunnable) continue;
            g_running = f; g_cur = &f->tc;
            switch_to_fiber(f);
        }
        /* release barriers whose participants have all arrived (exited threads do not count) */
        int at_block = 0;
        for (int t = 0; t < nt; ++t) {
            const State s = g_fibers[t]->state;
            if (s != kDone) ++live;
            if (s == kAtBlockBarrier) ++at_block;
        }
        if (live == 0) break;
        bool released = false;
        if (at_block == live) {
            for (int t = 0; t < nt; ++t) if (g_fibers[t]->state == kAtBlockBarrier) g_fibers[t]->state = kRunnable;
            released = true;
        }
        for (int w = 0; w < nt; w += 32) {
            int wl = 0, ww = 0;
            for (int t = w; t < w + 32 && t < nt; ++t) {
                const State s = g_fibers[t]->state;
                if (s != kDone) ++wl;
                if (s == kAtWarpBarrier) ++ww;
            }
            if (wl > 0 && ww == wl) {
                for (int t = w; t < w + 32 && t < nt; ++t) if (g_fibers[t]->state == kAtWarpBarrier) g_fibers[t]->state = kRunnable;
                released = true;
            }
        }
        if (released) { g_spins = 0; continue; }
        bool any_runnable = false;
        for (int t = 0; t < nt; ++t) if (g_fibers[t]->state == kRunnable) any_runnable = true;
        if (!any_runnable) fail("deadlock: every live thread of the CTA waits at a barrier that cannot complete (divergent __syncthreads / __syncwarp?)");
    }
    g_cur = nullptr; g_running = nullptr;
}

static std::map<std::string, long> g_launches;      /* kernel expression as written at the launch site -> count */

/* Plain device memory (cudaMalloc) is not addressable from the host on the real machine.  Here it is host memory, so a
 * host-side dereference would silently work; to make it fault like the real thing the pages of every cudaMalloc block
 * are PROT_NONE except while a kernel runs or the runtime itself copies / fills (DeviceAccess).  Managed and pinned
 * blocks stay open.  LISB_EMU_PROTECT=0 switches the protection off. */
struct DeviceAccess { DeviceAccess(); ~DeviceAccess(); };

void launch(const char *name, dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()> &body)
{
    DeviceAccess open_device;
    g_launches[name] += 1;
    if (g_running) fail("nested kernel launch");
    if (grid.y != 1 || grid.z != 1) fail("only 1-D grids are emulated");
    const size_t nt = (size_t)block.x * block.y * block.z;
    if (nt == 0 || nt > 1024) fail("invalid block size");
    if (dyn_smem_bytes > 227 * 1024) fail("dynamic shared memory exceeds 227 KB");
    g_dyn_smem.assign(dyn_smem_bytes + 128, 0xA5);
    g_body = &body;
    g_spins = 0;
    for (unsigned b = 0; b < grid.x; ++b) run_block(grid, block, b);
    g_body = nullptr;
}

/* ---- guarded "device" memory ---------------------------------------------------------------- */
struct Alloc { void *base; size_t len; };
static std::map<void *, Alloc> g_allocs;

static void *guarded_alloc(size_t bytes)
{
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t user = ((bytes ? bytes : 1) + 15) & ~(size_t)15;
    const size_t body = (user + page - 1) / page * page;
    char *base = (char *)mmap(nullptr, body + page, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) return nullptr;
    mprotect(base + body, page, PROT_NONE);
    char *p = base + body - user;
    memset(base, 0xCD, body);                      /* uninitialised device memory is not zero */
    g_allocs[p] = Alloc{base, body + page};
    return p;
}

static std::set<void *> g_device_only;             /* cudaMalloc blocks (keys of g_allocs) */
static int g_device_open = 0;
static bool protect_enabled()
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("LISB_EMU_PROTECT"); on = !(e && e[0] == '0'); }
    return on != 0;
}
/* with memory protection keys (x86 PKU) every device block carries one key and opening / closing the device is a
 * register write; without them each block is mprotect'ed */
static int g_pkey = -2;                           /* -2: not tried, -1: unavailable */
static void device_pages(void *p, int prot)
{
    const Alloc &a = g_allocs[p];
    mprotect(a.base, a.len - (size_t)sysconf(_SC_PAGESIZE), prot);
}
static void device_set(int prot)
{
    if (g_pkey >= 0) pkey_set(g_pkey, prot == PROT_NONE ? PKEY_DISABLE_ACCESS : 0);
    else for (void *p : g_device_only) device_pages(p, prot);
}
DeviceAccess::DeviceAccess()
{
    if (g_device_open++ == 0 && protect_enabled()) device_set(PROT_READ | PROT_WRITE);
}
DeviceAccess::~DeviceAccess()
{
    if (--g_device_open == 0 && protect_enabled()) device_set(PROT_NONE);
}
/* a fault inside a closed device block: say so (with the host call chain) before dying with the SIGSEGV the tests expect */
static void on_segv(int sig, siginfo_t *si, void *)
{
    const char *addr = (const char *)si->si_addr;
    for (void *p : g_device_only) {
        const Alloc &a = g_allocs[p];
        if (addr >= (const char *)a.base && addr < (const char *)a.base + a.len) {
            static const char msg[] = "cuda_emu: the host touched plain device memory (cudaMalloc) outside a kernel or runtime copy\n";
            if (write(2, msg, sizeof(msg) - 1) < 0) {}
            void *bt[32];
            backtrace_symbols_fd(bt, backtrace(bt, 32), 2);
            break;
        }
    }
    signal(sig, SIG_DFL);
    raise(sig);
}
static void *device_alloc(size_t bytes)
{
    static bool hooked = false;
    if (!hooked && protect_enabled()) {
        hooked = true;
        struct sigaction sa;
        memset(&sa, 0, sizeof(sa));
        sa.sa_sigaction = on_segv;
        sa.sa_flags = SA_SIGINFO | SA_NODEFER;
        sigaction(SIGSEGV, &sa, nullptr);
    }
    if (g_pkey == -2 && protect_enabled()) {
        const char *e = getenv("LISB_EMU_PKEY");           /* LISB_EMU_PKEY=0: take the mprotect path */
        g_pkey = (e && e[0] == '0') ? -1 : pkey_alloc(0, g_device_open == 0 ? PKEY_DISABLE_ACCESS : 0);
    }
    void *p = guarded_alloc(bytes);
    if (p && protect_enabled()) {
        g_device_only.insert(p);
        if (g_pkey >= 0) {
            const Alloc &a = g_allocs[p];
            if (pkey_mprotect(a.base, a.len - (size_t)sysconf(_SC_PAGESIZE), PROT_READ | PROT_WRITE, g_pkey) != 0) { fprintf(stderr, "cuda_emu: pkey_mprotect failed\n"); abort(); }
        } else if (g_device_open == 0) device_pages(p, PROT_NONE);
    }
    return p;
}

static void guarded_free(void *p)
{
    if (!p) return;
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) { fprintf(stderr, "cuda_emu: cudaFree of an unknown pointer %p\n", p); abort(); }
    g_device_only.erase(p);
    munmap(it->second.base, it->second.len);
    g_allocs.erase(it);
}

}  // namespace emu

/* launch log for the tests: how many launches had `substr` in the kernel expression; NULL resets */
extern "C" long emu_launch_count(const char *substr)
{
    if (substr == nullptr) { emu::g_launches.clear(); return 0; }
    long n = 0;
    for (auto &kv : emu::g_launches) if (kv.first.find(substr) != std::string::npos) n += kv.second;
    return n;
}

/* ---- CUDA runtime stand-ins (host memory; streams and events are ordered by program order) ---- */
extern "C" {

/* SM count the launchers size their grids with: 148 (B200) by default, so that reduction trees
 * and persistent grids are the ones the real device gets; LISB_EMU_SMS=2 makes small inputs
 * exercise the persistent loops (several row blocks per CTA) */
static int emu_sms()
{
    const char *e = getenv("LISB_EMU_SMS");
    const int v = e ? atoi(e) : 148;
    return v > 0 ? v : 148;
}

cudaError_t cudaGetDeviceCount(int *c) { *c = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "emulated device error"; }
cudaError_t cudaDeviceGetAttribute(int *v, enum cudaDeviceAttr a, int)
{
    *v = a == cudaDevAttrMultiProcessorCount ? emu_sms() : 0;
    return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void *, enum cudaFuncAttribute, int) { return cudaSuccess; }
/* Streams.  The first stream created (the library's main stream) and the kernels run eagerly, in
 * program order.  Every OTHER stream is lazy: its copies are queued and carried out as late as
 * the CUDA ordering rules allow -- when somebody waits on an event recorded behind them, or
 * synchronises the stream.  A kernel that reads a buffer without waiting for the event of the
 * copy that fills it therefore sees stale data here, and a kernel that overwrites a buffer whose
 * outgoing copy is still queued corrupts that copy: missing cross-stream dependencies fail the
 * tests instead of being hidden by a synchronous emulation. */
struct EmuOp { void *dst; const void *src; size_t n; cudaEvent_t marker; };
static cudaStream_t g_main_stream = nullptr;
static std::map<cudaStream_t, std::vector<EmuOp>> g_lazy;     /* pending operations per lazy stream */
static std::map<cudaEvent_t, cudaStream_t> g_event_on;        /* event -> lazy stream it is pending on */

/* Live handles.  Real CUDA has undefined behaviour (in practice: a crash) on a stream or event that
 * was already destroyed; here every use or second destroy of a dead handle aborts the test. */
static std::set<cudaStream_t> g_live_streams;
static std::set<cudaEvent_t> g_live_events;
static void emu_need_stream(cudaStream_t s, const char *what)
{
    if (s != nullptr && !g_live_streams.count(s)) { fprintf(stderr, "cuda_emu: %s on a stream that was destroyed or never created (%p)\n", what, (void *)s); abort(); }
}
static void emu_need_event(cudaEvent_t e, const char *what)
{
    if (!g_live_events.count(e)) { fprintf(stderr, "cuda_emu: %s on an event that was destroyed or never created (%p)\n", what, (void *)e); abort(); }
}

static void emu_flush(cudaStream_t s, cudaEvent_t upto)
{
    auto it = g_lazy.find(s);
    if (it == g_lazy.end()) return;
    std::vector<EmuOp> &q = it->second;
    size_t k = 0;
    for (; k < q.size(); ++k) {
        if (q[k].marker) {
            g_event_on.erase(q[k].marker);
            if (q[k].marker == upto) { ++k; break; }
        } else {
            emu::DeviceAccess open_device;
            memmove(q[k].dst, q[k].src, q[k].n);
        }
    }
    q.erase(q.begin(), q.begin() + (long)k);
}
static void emu_flush_all() { for (auto &kv : g_lazy) emu_flush(kv.first, nullptr); }
static bool emu_is_lazy(cudaStream_t s) { return s != nullptr && s != g_main_stream; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int)
{
    static int k = 0;
    *s = (cudaStream_t)(uintptr_t)(0x100 + 16 * ++k);
    if (g_main_stream == nullptr) g_main_stream = *s;
    g_live_streams.insert(*s);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) { emu_need_stream(s, "cudaStreamSynchronize"); emu_flush(s, nullptr); return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { emu_flush_all(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s)
{
    if (s == nullptr) { fprintf(stderr, "cuda_emu: cudaStreamDestroy(NULL)\n"); abort(); }
    emu_need_stream(s, "cudaStreamDestroy");
    emu_flush(s, nullptr); g_lazy.erase(s); g_live_streams.erase(s);
    if (s == g_main_stream) g_main_stream = nullptr;
    return cudaSuccess;
}
static cudaError_t emu_wait_event(cudaEvent_t e)
{
    auto it = g_event_on.find(e);
    if (it != g_event_on.end()) emu_flush(it->second, e);
    return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned int)
{
    emu_need_stream(s, "cudaStreamWaitEvent"); emu_need_event(e, "cudaStreamWaitEvent");
    auto it = g_event_on.find(e);
    if (it != g_event_on.end()) emu_flush(it->second, e);
    return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned int) { static int k = 0; *e = (cudaEvent_t)(uintptr_t)(0x100000 + 16 * ++k); g_live_events.insert(*e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s)
{
    emu_need_stream(s, "cudaEventRecord"); emu_need_event(e, "cudaEventRecord");
    auto it = g_event_on.find(e);
    if (it != g_event_on.end()) emu_flush(it->second, e);            /* re-recording: the old one is done with */
    if (emu_is_lazy(s) && !g_lazy[s].empty()) { g_lazy[s].push_back(EmuOp{nullptr, nullptr, 0, e}); g_event_on[e] = s; }
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) { emu_need_event(e, "cudaEventSynchronize"); return emu_wait_event(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { emu_need_event(e, "cudaEventDestroy"); emu_wait_event(e); g_live_events.erase(e); return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) { *p = emu::device_alloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaMallocManaged(void **p, size_t n, unsigned int) { *p = emu::guarded_alloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { emu_flush_all(); emu::guarded_free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned int) { *p = emu::guarded_alloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { emu_flush_all(); emu::guarded_free(p); return cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned int) { *d = h; return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, enum cudaMemcpyKind, cudaStream_t st)
{
    emu_need_stream(st, "cudaMemcpyAsync");
    if (emu_is_lazy(st)) g_lazy[st].push_back(EmuOp{d, s, n, nullptr});
    else { emu::DeviceAccess open_device; memmove(d, s, n); }
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, enum cudaMemcpyKind) { emu_flush_all(); emu::DeviceAccess open_device; memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st) { emu_need_stream(st, "cudaMemsetAsync"); emu::DeviceAccess open_device; memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { emu::DeviceAccess open_device; memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemPrefetchAsync(const void *, size_t, int, cudaStream_t st) { emu_need_stream(st, "cudaMemPrefetchAsync"); return cudaSuccess; }
cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)1 << 40; return cudaSuccess; }
/* peer memory / IPC: not available here, so the in-kernel halo exchange stays off and the staged transport runs */
cudaError_t cudaDeviceGetPCIBusId(char *b, int len, int dev) { (void)dev; if (len > 0) b[0] = 0; return cudaErrorNotSupported; }
cudaError_t cudaDeviceGetByPCIBusId(int *dev, const char *b) { (void)b; *dev = -1; return cudaErrorNotSupported; }
cudaError_t cudaDeviceCanAccessPeer(int *can, int a, int b) { (void)a; (void)b; *can = 0; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int dev, unsigned int f) { (void)dev; (void)f; return cudaErrorNotSupported; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { (void)h; (void)p; return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned int f) { (void)h; (void)f; *p = 0; return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *p) { (void)p; return cudaErrorNotSupported; }


}  // extern "C"
