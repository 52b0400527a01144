#!/usr/bin/env python
"""managed_probe.py -- why does cudaMallocManaged fail once 8 processes share a node (round 1: "unknown error", vector
storage fell back to device-only memory)?  One process per GPU under torchrun, ALL GPUs visible to every process (no
CUDA_VISIBLE_DEVICES narrowing), NCCL communicator set up the way bench.py / the library do it, then a series of managed
allocations with the error code of each.  Variants through the environment (NCCL_NVLS_ENABLE, NCCL_P2P_DISABLE,
NCCL_CUMEM_ENABLE, PROBE_NARROW=1 to narrow like bench.py does).  Not a test; prints one line per rank."""
import ctypes as C
import os
import sys

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
if os.environ.get("PROBE_NARROW") == "1":
    os.environ["CUDA_VISIBLE_DEVICES"] = str(local)
    local = 0
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

torch.cuda.set_device(local)
dev = torch.device("cuda", local)
rt = C.CDLL("libcudart.so.12")
rt.cudaMallocManaged.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
rt.cudaGetErrorString.restype = C.c_char_p
rt.cudaMemPrefetchAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]


def managed(tag, count=6, gb=1):
    res = []
    ptrs = []
    for k in range(count):
        p = C.c_void_p()
        e = rt.cudaMallocManaged(C.byref(p), gb << 30, 1)
        if e == 0:
            e2 = rt.cudaMemPrefetchAsync(p, gb << 30, local, None)
            rt.cudaDeviceSynchronize()
            res.append("ok" if e2 == 0 else f"prefetch:{rt.cudaGetErrorString(e2).decode()}")
            ptrs.append(p)
        else:
            res.append(rt.cudaGetErrorString(e).decode())
            rt.cudaGetLastError()
    for p in ptrs:
        rt.cudaFree(p)
    print(f"[rank {rank}] {tag}: {res}", flush=True)


managed("before NCCL")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    t = torch.ones(1 << 20, device=dev)
    dist.all_reduce(t)
    torch.cuda.synchronize()
    managed("after torch NCCL all_reduce")
    if os.environ.get("PROBE_PEER") == "1":
        rt.cudaDeviceCanAccessPeer.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int]
        errs = []
        for peer in range(torch.cuda.device_count()):
            if peer != local:
                errs.append(rt.cudaDeviceEnablePeerAccess(peer, 0))
        rt.cudaGetLastError()
        managed(f"after cudaDeviceEnablePeerAccess to {len(errs)} peers (rc {sorted(set(errs))})")
    # the library's own communicator + halo exchange path
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import lis_b200
    lib = lis_b200.load_library()
    lib.lis_b200_comm_attach.argtypes = [C.c_int, C.c_int, C.c_ulonglong]
    tok = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == 0:
        tok[0] = int.from_bytes(os.urandom(7), "little")
    dist.broadcast(tok, 0)
    os.environ["LOCAL_RANK"] = str(local)
    shim = lis_b200.load_shim()
    rc = lib.lis_b200_comm_attach(rank, world, int(tok.item()))
    managed(f"after lis_b200_comm_attach rc={rc}")
    dist.barrier()
    dist.destroy_process_group()
if rank == 0:
    try:
        print("max_map_count", open("/proc/sys/vm/max_map_count").read().strip(), "ulimit -l", os.popen("ulimit -l").read().strip(), flush=True)
    except Exception:
        pass
