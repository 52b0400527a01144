"""Experiment: rows-per-block x pipeline depth of the TMA CSR kernel.
Run on the GPU box: python profiles/sweep_csr_tma.py [grid] [7|27]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import lis_b200  # noqa: E402

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
K = lis_b200.load_kernels()
stencil = int(sys.argv[2]) if len(sys.argv) > 2 else 7
if stencil == 7:
    ptr, idx, val = bench.poisson7_device(torch, grid, grid, grid, 0, grid, dev)
else:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import harness
    p_, i_, v_ = harness.poisson3d_27pt(grid, grid, grid)
    ptr, idx, val = torch.from_numpy(p_).to(dev), torch.from_numpy(i_).to(dev), torch.from_numpy(v_).to(dev)
n, nnz = ptr.numel() - 1, idx.numel()
x = torch.rand(n, device=dev, dtype=torch.float64); y = torch.zeros_like(x)
idx = torch.cat([idx, torch.zeros(8, device=dev, dtype=torch.int32)])
val = torch.cat([val, torch.zeros(8, device=dev, dtype=torch.float64)])
ptr = torch.cat([ptr, torch.zeros(4, device=dev, dtype=torch.int32)])
stream = torch.cuda.Stream(device=dev); sp = C.c_void_p(stream.cuda_stream)
bytes_csr = 12.0 * nnz + 20.0 * n
ref = None
import numpy as np
hp = ptr.cpu().numpy()
configs = []
for rows in (256, 128, 64):
    w = int(max(((hp[min(r0 + rows, n)] - (hp[r0] & ~3) + 3) & ~3) for r0 in range(0, n, rows * max(1, n // (rows * 4096)))))
    tile = (max(w, hp[min(rows, n)]) + 255) & ~255
    configs.append((rows, int(tile)))
print("configs", configs, flush=True)


def tile_kernel():
    rc = K.lisb200_spmv_csr(n, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(), x.data_ptr(), y.data_ptr(), sp)
    assert rc == 0


s0 = bench.time_launches(torch, stream, tile_kernel, 20, 3) / 20
ref = y.clone()
print(f"product-tile kernel: {2 * nnz / s0 / 1e9:8.1f} GFLOP/s {bytes_csr / s0 / 1e9:8.1f} GB/s", flush=True)
for rows, tile in configs:
    for stages in (2, 3, 4, 6, 8):
        if stages * (12 * tile + 4 * (rows + 4)) > 222 * 1024:
            continue

        def f():
            rc = K.lisb200_spmv_csr_tma(n, rows, tile, stages, ptr.data_ptr(), idx.data_ptr(), val.data_ptr(), x.data_ptr(), y.data_ptr(), sp)
            assert rc == 0, rc
        try:
            s = bench.time_launches(torch, stream, f, 20, 3) / 20
        except Exception as e:
            print(rows, tile, stages, "failed", e); continue
        ok = True
        if ref is None:
            ref = y.clone()
        else:
            ok = torch.equal(ref.view(torch.int64), y.view(torch.int64))
        print(f"rows={rows:4d} tile={tile:5d} stages={stages}: {2 * nnz / s / 1e9:8.1f} GFLOP/s {bytes_csr / s / 1e9:8.1f} GB/s same_bits={ok}", flush=True)
