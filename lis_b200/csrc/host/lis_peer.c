/*
 * lis_peer.c -- device memory that the neighbours' GPUs can store into, and nothing else.
 *
 * The halo exchange inside the SpMV kernel (kernels/spmv.cu, host/lis_comm.c p2p_prepare) needs each rank's inbox
 * mapped into its neighbours' address spaces.  The legacy route -- cudaIpcGetMemHandle / cudaIpcOpenMemHandle -- needs
 * cudaDeviceEnablePeerAccess, and that call was measured to slow EVERY kernel of the process that reads cudaMalloc
 * memory by ~20 % on B200 (512^3 CSR product: 2.49 ms instead of 2.10 ms, profiles/r02_session11.sh).  The virtual-memory
 * API maps exactly one allocation: cuMemCreate with a POSIX-fd shareable handle here, cuMemImportFromShareableHandle +
 * cuMemMap + cuMemSetAccess for the local device there; no peer access is enabled for anything else (this is also how
 * NCCL maps its buffers).  The file descriptors travel between the processes of the node as SCM_RIGHTS messages over
 * abstract-namespace unix datagram sockets.
 *
 * libcuda is reached through dlopen (the library links the runtime only); everything here reports failure instead of
 * aborting, and the caller falls back to the NCCL exchange.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <unistd.h>
#include <dlfcn.h>
#include <errno.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <cuda.h>
#include <cuda_runtime_api.h>
#include "lis_device.h"

typedef CUresult (*fn_cuDeviceGet)(CUdevice *, int);
typedef CUresult (*fn_cuDeviceGetAttribute)(int *, CUdevice_attribute, CUdevice);
typedef CUresult (*fn_cuDeviceCanAccessPeer)(int *, CUdevice, CUdevice);
typedef CUresult (*fn_cuMemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags);
typedef CUresult (*fn_cuMemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
typedef CUresult (*fn_cuMemRelease)(CUmemGenericAllocationHandle);
typedef CUresult (*fn_cuMemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
typedef CUresult (*fn_cuMemAddressFree)(CUdeviceptr, size_t);
typedef CUresult (*fn_cuMemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
typedef CUresult (*fn_cuMemUnmap)(CUdeviceptr, size_t);
typedef CUresult (*fn_cuMemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
typedef CUresult (*fn_cuMemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
typedef CUresult (*fn_cuMemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType);

static struct {
    int tried, ok;
    void *dl;
    fn_cuDeviceGet DeviceGet; fn_cuDeviceGetAttribute DeviceGetAttribute; fn_cuDeviceCanAccessPeer DeviceCanAccessPeer;
    fn_cuMemGetAllocationGranularity GetGranularity; fn_cuMemCreate Create; fn_cuMemRelease Release;
    fn_cuMemAddressReserve Reserve; fn_cuMemAddressFree AddressFree; fn_cuMemMap Map; fn_cuMemUnmap Unmap; fn_cuMemSetAccess SetAccess;
    fn_cuMemExportToShareableHandle Export; fn_cuMemImportFromShareableHandle Import;
} P;

static int peer_load(void)
{
    if (P.tried) return P.ok;
    P.tried = 1;
    P.dl = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!P.dl) return 0;
#define SYM(field, name) do { P.field = (fn_##name)dlsym(P.dl, #name); if (!P.field) return 0; } while (0)
    SYM(DeviceGet, cuDeviceGet); SYM(DeviceGetAttribute, cuDeviceGetAttribute); SYM(DeviceCanAccessPeer, cuDeviceCanAccessPeer);
    SYM(GetGranularity, cuMemGetAllocationGranularity); SYM(Create, cuMemCreate); SYM(Release, cuMemRelease);
    SYM(Reserve, cuMemAddressReserve); SYM(AddressFree, cuMemAddressFree); SYM(Map, cuMemMap); SYM(Unmap, cuMemUnmap);
    SYM(SetAccess, cuMemSetAccess); SYM(Export, cuMemExportToShareableHandle); SYM(Import, cuMemImportFromShareableHandle);
#undef SYM
    CUdevice dev;
    int fd_ok = 0;
    if (P.DeviceGet(&dev, lisd_device_id()) != CUDA_SUCCESS) return 0;
    if (P.DeviceGetAttribute(&fd_ok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev) != CUDA_SUCCESS || !fd_ok) return 0;
    P.ok = 1;
    return 1;
}

int lisd_peer_available(void) { return lisd_available() && peer_load(); }

/* can my device load/store memory of the device with runtime ordinal `peer_dev`? */
int lisd_peer_can_access(int peer_dev)
{
    CUdevice a, b;
    int can = 0;
    if (!lisd_peer_available()) return 0;
    if (P.DeviceGet(&a, lisd_device_id()) != CUDA_SUCCESS || P.DeviceGet(&b, peer_dev) != CUDA_SUCCESS) return 0;
    if (P.DeviceCanAccessPeer(&can, a, b) != CUDA_SUCCESS) return 0;
    return can;
}

static void alloc_prop(CUmemAllocationProp *prop, int dev)
{
    memset(prop, 0, sizeof(*prop));
    prop->type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop->location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop->location.id = dev;
    prop->requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
}

static int map_for_me(CUmemGenericAllocationHandle h, size_t size, void **ptr)
{
    CUdeviceptr va = 0;
    CUmemAccessDesc acc;
    if (P.Reserve(&va, size, 0, 0, 0) != CUDA_SUCCESS) return 0;
    if (P.Map(va, size, 0, h, 0) != CUDA_SUCCESS) { P.AddressFree(va, size); return 0; }
    memset(&acc, 0, sizeof(acc));
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = lisd_device_id();
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (P.SetAccess(va, size, &acc, 1) != CUDA_SUCCESS) { P.Unmap(va, size); P.AddressFree(va, size); return 0; }
    *ptr = (void *)va;
    return 1;
}

/* a block of at least `bytes` on my device, exportable: *ptr (mapped here), *size (rounded up), *fd (to hand to neighbours),
 * *handle (for lisd_peer_free) */
int lisd_peer_alloc(size_t bytes, void **ptr, size_t *size, int *fd, unsigned long long *handle)
{
    CUmemAllocationProp prop;
    CUmemGenericAllocationHandle h;
    size_t gran = 0;
    int f = -1;
    if (!lisd_peer_available()) return 0;
    alloc_prop(&prop, lisd_device_id());
    if (P.GetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return 0;
    const size_t sz = (bytes + gran - 1) / gran * gran;
    if (P.Create(&h, sz, &prop, 0) != CUDA_SUCCESS) return 0;
    if (!map_for_me(h, sz, ptr)) { P.Release(h); return 0; }
    if (P.Export(&f, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) {
        P.Unmap((CUdeviceptr)*ptr, sz); P.AddressFree((CUdeviceptr)*ptr, sz); P.Release(h);
        return 0;
    }
    *size = sz; *fd = f; *handle = (unsigned long long)h;
    return 1;
}

/* map a neighbour's block (its exported fd, received over the socket) on my device */
int lisd_peer_import(int fd, size_t size, void **ptr)
{
    CUmemGenericAllocationHandle h;
    if (!lisd_peer_available()) return 0;
    if (P.Import(&h, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) return 0;
    const int ok = map_for_me(h, size, ptr);
    P.Release(h);                         /* the mapping keeps the memory alive */
    return ok;
}

void lisd_peer_unmap(void *ptr, size_t size)
{
    if (!ptr || !P.ok) return;
    P.Unmap((CUdeviceptr)ptr, size);
    P.AddressFree((CUdeviceptr)ptr, size);
}

void lisd_peer_free(void *ptr, size_t size, int fd, unsigned long long handle)
{
    if (!P.ok) return;
    lisd_peer_unmap(ptr, size);
    if (fd >= 0) close(fd);
    P.Release((CUmemGenericAllocationHandle)handle);
}

/* ---- file descriptors between the processes of the node: SCM_RIGHTS over abstract unix datagram sockets ---- */
static void sock_name(struct sockaddr_un *sa, socklen_t *len, const char *job, int rank)
{
    memset(sa, 0, sizeof(*sa));
    sa->sun_family = AF_UNIX;
    const int n = snprintf(sa->sun_path + 1, sizeof(sa->sun_path) - 2, "lisb200-%s-%d", job, rank);      /* sun_path[0] = 0: abstract */
    *len = (socklen_t)(offsetof(struct sockaddr_un, sun_path) + 1 + (size_t)n);
}

/* this rank's receiving socket (bind before anybody sends: the caller puts a barrier behind it); -1 on failure */
int lisd_fd_socket(const char *job, int rank)
{
    struct sockaddr_un sa;
    socklen_t len;
    const int s = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
    if (s < 0) return -1;
    sock_name(&sa, &len, job, rank);
    if (bind(s, (struct sockaddr *)&sa, len) != 0) { close(s); return -1; }
    return s;
}

int lisd_fd_send(int sock, const char *job, int to_rank, int my_rank, int fd)
{
    struct sockaddr_un sa;
    socklen_t len;
    struct msghdr msg;
    struct iovec iov;
    union { struct cmsghdr h; char buf[CMSG_SPACE(sizeof(int))]; } ctl;
    int payload = my_rank;
    sock_name(&sa, &len, job, to_rank);
    memset(&msg, 0, sizeof(msg)); memset(&ctl, 0, sizeof(ctl));
    iov.iov_base = &payload; iov.iov_len = sizeof(payload);
    msg.msg_name = &sa; msg.msg_namelen = len; msg.msg_iov = &iov; msg.msg_iovlen = 1;
    msg.msg_control = ctl.buf; msg.msg_controllen = sizeof(ctl.buf);
    struct cmsghdr *c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd, sizeof(int));
    return sendmsg(sock, &msg, 0) == (ssize_t)sizeof(payload);
}

/* one message: *from_rank and the descriptor it carried; 0 on timeout / error */
int lisd_fd_recv(int sock, int *from_rank, int *fd, int timeout_ms)
{
    struct pollfd pf = { .fd = sock, .events = POLLIN };
    struct msghdr msg;
    struct iovec iov;
    union { struct cmsghdr h; char buf[CMSG_SPACE(sizeof(int))]; } ctl;
    int payload = -1;
    if (poll(&pf, 1, timeout_ms) <= 0) return 0;
    memset(&msg, 0, sizeof(msg)); memset(&ctl, 0, sizeof(ctl));
    iov.iov_base = &payload; iov.iov_len = sizeof(payload);
    msg.msg_iov = &iov; msg.msg_iovlen = 1; msg.msg_control = ctl.buf; msg.msg_controllen = sizeof(ctl.buf);
    if (recvmsg(sock, &msg, MSG_CMSG_CLOEXEC) != (ssize_t)sizeof(payload)) return 0;
    struct cmsghdr *c = CMSG_FIRSTHDR(&msg);
    if (!c || c->cmsg_level != SOL_SOCKET || c->cmsg_type != SCM_RIGHTS) return 0;
    memcpy(fd, CMSG_DATA(c), sizeof(int));
    *from_rank = payload;
    return 1;
}
