#!/usr/bin/env python
"""run_configs.py -- BASELINE.json configs 3, 4 and 5 through lis_solve, one process per GPU (not a test).

    [torchrun --nproc-per-node N ...] python profiles/run_configs.py CONFIG [--size S] [--threads T] [--impl lis_b200|reference]
                                                                      [--opts "..."] [--out FILE]

CONFIG  cg7   config 3: test/test3.c's 7-pt Poisson system, S^3 rows PER RANK (slabs of an (S*N) x S x S box), -i cg -p jacobi
        su    config 4: seeded unsymmetric banded matrix (tests/shim/lis_shim.c shim_banded_rows; S rows in TOTAL, 70 entries per
              row, |i-j| <= 1e5, diagonal = 0.17 * sum|off-diagonals|), rows partitioned by LIS_GET_ISIE, -i bicgstab -p ssor;
              --threads T = SSOR blocks per rank (the reference's OpenMP thread count per process)
        gm27  config 5: 27-pt stencil of test/spmvtest3b.c, S x S x S rows PER RANK ... (S*N) x S x S box, -i gmres -restart 30 -p jacobi
b = A*1, x0 = 0, -tol 1e-12.  --impl reference runs the compiled reference (OpenMP build, one process, all cores unless --threads).
Rank 0 prints one JSON line (times are the maximum over the ranks)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1 and "LOCAL_RANK" in os.environ and "CUDA_VISIBLE_DEVICES" not in os.environ:
    os.environ["CUDA_VISIBLE_DEVICES"] = os.environ["LOCAL_RANK"]      # one process, one visible GPU
os.environ.setdefault("NCCL_NVLS_ENABLE", "0")

import numpy as np  # noqa: E402
import lis_b200  # noqa: E402

libc = C.CDLL("libc.so.6"); libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]


def isie(k, nprocs, n):
    """LIS_GET_ISIE (include/lis.h:1067-1078)"""
    q, r = divmod(n, nprocs)
    is_ = k * q + min(k, r)
    return is_, is_ + q + (1 if k < r else 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["cg7", "su", "gm27"])
    ap.add_argument("--size", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--impl", default="lis_b200", choices=["lis_b200", "reference"])
    ap.add_argument("--opts", default="")
    ap.add_argument("--dom", type=float, default=0.17)
    ap.add_argument("--slab", type=int, default=0, help="gm27/cg7: planes per rank of an (slab*N) x size x size box instead of size^3 per rank")
    ap.add_argument("--out", default="")
    ap.add_argument("--maxiter", type=int, default=0, help="cap the iterations (a rate measurement, status 4)")
    a = ap.parse_args()
    ref = a.impl == "reference"
    if ref:
        assert world == 1, "the reference has no multi-process build here (no MPI)"
        cores = a.threads or len(os.sched_getaffinity(0))
        shim = lis_b200.Shim(os.path.join(ROOT, "oracle", "_ref", "libref_shim_omp.so"), f"-omp_num_threads {cores}")
        shim.set_threads(cores)
    else:
        shim = lis_b200.load_shim()
        if a.threads:
            shim.set_threads(a.threads)
    L = shim.lib
    t0 = time.time()
    if a.config == "cg7":
        g = a.size or 256
        sl = a.slab or g
        L.shim_poisson7.restype = C.c_longlong; L.shim_poisson7.argtypes = [C.c_int] * 6 + [C.c_void_p] * 3
        n = sl * g * g
        nnz = L.shim_poisson7(sl * world, g, g, rank * sl, (rank + 1) * sl, 0, None, None, None)
        pp, pi, pv = libc.malloc(4 * (n + 1)), libc.malloc(4 * nnz), libc.malloc(8 * nnz)
        L.shim_poisson7(sl * world, g, g, rank * sl, (rank + 1) * sl, 0, pp, pi, pv)
        opts = "-i cg -p jacobi -tol 1e-12 -maxiter 20000 " + a.opts
        what = f"test3.c 7-pt Poisson {sl * world}x{g}x{g}"
        flops_it = lambda nnz_g, n_g: 2.0 * nnz_g + 13.0 * n_g
    elif a.config == "gm27":
        g = a.size or 128
        sl = a.slab or g
        L.shim_poisson27.restype = C.c_longlong; L.shim_poisson27.argtypes = [C.c_int] * 5 + [C.c_void_p] * 3
        n = sl * g * g
        nnz = L.shim_poisson27(sl * world, g, g, rank * sl, (rank + 1) * sl, None, None, None)
        pp, pi, pv = libc.malloc(4 * (n + 1)), libc.malloc(4 * nnz), libc.malloc(8 * nnz)
        L.shim_poisson27(sl * world, g, g, rank * sl, (rank + 1) * sl, pp, pi, pv)
        opts = "-i gmres -restart 30 -p jacobi -tol 1e-12 -maxiter 20000 " + a.opts
        what = f"spmvtest3b.c 27-pt stencil {sl * world}x{g}x{g}"
        flops_it = lambda nnz_g, n_g: 2.0 * nnz_g + 65.0 * n_g          # (2m+5) n with m = 30
    else:
        gn = a.size or 1000000
        i0, i1 = isie(rank, world, gn)
        n = i1 - i0
        nnz = 70 * n
        L.shim_banded_rows.restype = C.c_longlong
        L.shim_banded_rows.argtypes = [C.c_int] * 5 + [C.c_double, C.c_double, C.c_ulonglong] + [C.c_void_p] * 3
        pp, pi, pv = libc.malloc(4 * (n + 1)), libc.malloc(4 * nnz), libc.malloc(8 * nnz)
        L.shim_banded_rows(gn, i0, i1, 70, 100000, a.dom, 0.0, 7, pp, pi, pv)
        opts = "-i bicgstab -p ssor -tol 1e-12 -maxiter 5000 " + a.opts
        what = f"banded unsymmetric n={gn}, 70 entries/row, |i-j|<=1e5, diag = {a.dom}*sum|offdiag|"
        flops_it = lambda nnz_g, n_g: 4.0 * nnz_g + 2.0 * 2.0 * nnz_g + 26.0 * n_g   # 2 products + 2 SSOR applies (~2 nnz each)
    if a.maxiter:
        opts += f" -maxiter {a.maxiter}"
    gen_s = time.time() - t0
    t0 = time.time()
    if ref:
        L.shim_mv_open.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        h = L.shim_mv_open(1, n, pp, pi, pv, 0, 0, 1)
    else:
        L.shim_mv_open_dist.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        h = L.shim_mv_open_dist(1, n, pp, pi, pv, 1)
    assert h >= 0, h
    open_s = time.time() - t0
    oi = np.zeros(4, np.int32); od = np.zeros(6, np.float64); rh = np.zeros(32768)
    L.shim_mv_solve_ones.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.shim_mv_solve_ones(h, opts.encode(), oi.ctypes.data, od.ctypes.data, rh.ctypes.data, len(rh))
    assert rc == 0 or oi[1] == 4, (rc, oi)      # a -maxiter cap (status 4) is a valid rate measurement
    mine = np.array([od[2], od[3], od[4], od[5], float(nnz), float(n)])
    if world > 1:
        lib = lis_b200.load_library()
        lib.lis_b200_allreduce_sum.argtypes = [C.c_void_p, C.c_int]
        # max over ranks of the times / the error, sum of the sizes: through sums of one-hot slices
        tab = np.zeros((world, 6)); tab[rank] = mine
        flat = tab.reshape(-1).copy()
        for s in range(0, flat.size, 8):
            chunk = np.ascontiguousarray(flat[s:s + 8])
            assert lib.lis_b200_allreduce_sum(chunk.ctypes.data, len(chunk)) == 0
            flat[s:s + 8] = chunk
        tab = flat.reshape(world, 6)
    else:
        tab = mine.reshape(1, 6)
    shim.lib.shim_end()
    if rank != 0:
        return
    itime, ptime, wall, xerr = tab[:, 0].max(), tab[:, 1].max(), tab[:, 2].max(), tab[:, 3].max()
    nnz_g, n_g = tab[:, 4].sum(), tab[:, 5].sum()
    it = int(oi[0])
    out = {"config": a.config, "impl": a.impl, "n_ranks": world, "what": what, "options": opts.strip(), "n": int(n_g), "nnz": int(nnz_g),
           "threads_or_blocks_per_rank": (cores if ref else (a.threads or 1)),
           "iters": it, "status": int(oi[1]), "relres": float(od[0]), "max_abs_x_minus_1": float(xerr),
           "iter_time_s": float(itime), "precon_time_s": float(ptime), "solve_wall_s": float(wall),
           "ms_per_iter": float(itime) / max(it, 1) * 1e3, "iters_per_s": it / float(itime) if itime > 0 else None,
           "gflops": flops_it(nnz_g, n_g) * it / float(itime) / 1e9 if itime > 0 else None,
           "generate_s": gen_s, "assemble_s": open_s, "rhistory_head": [float(v) for v in rh[:3]], "rhistory_tail": [float(v) for v in rh[max(0, int(oi[3]) - 2):int(oi[3])]]}
    line = json.dumps(out)
    print(line, flush=True)
    if a.out:
        with open(a.out, "a") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
