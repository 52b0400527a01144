#!/usr/bin/env python
"""make_cg_fullsize.py -- TEST INFRASTRUCTURE.  Runs the compiled REFERENCE (oracle/_ref, OpenMP
build, all cores of this machine) on BASELINE.json config 3 -- test/test3.c's system: 7-point Poisson
on an N^3 grid, rows in test3.c order, b = A*1, x0 = 0, `-i cg -p jacobi -tol 1e-12` -- and stores
the iteration count and the full-precision residual history as tests/golden/cg_poisson_<N>.npz.
The GPU box has no reference tree and cannot afford ~10 minutes of CPU solve inside bench.py; the
bench and the -m gpu tests compare the lis_b200 run with this file.

usage: python tests/golden/make_cg_fullsize.py N [threads]
"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import lis_b200  # noqa: E402


def main():
    N = int(sys.argv[1])
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else len(os.sched_getaffinity(0))
    shim = lis_b200.Shim(os.path.join(ROOT, "oracle", "_ref", "libref_shim_omp.so" if threads > 1 else "libref_shim_serial.so"),
                         f"-omp_num_threads {threads}" if threads > 1 else "")
    L = shim.lib
    libc = C.CDLL("libc.so.6"); libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]
    L.shim_poisson7.restype = C.c_longlong
    L.shim_poisson7.argtypes = [C.c_int] * 6 + [C.c_void_p] * 3
    n = N ** 3
    nnz = L.shim_poisson7(N, N, N, 0, N, 0, None, None, None)
    p_ptr, p_idx, p_val = libc.malloc(4 * (n + 1)), libc.malloc(4 * nnz), libc.malloc(8 * nnz)
    assert L.shim_poisson7(N, N, N, 0, N, 0, p_ptr, p_idx, p_val) == nnz
    L.shim_mv_open.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    h = L.shim_mv_open(1, n, p_ptr, p_idx, p_val, 0, 0, 1)
    assert h >= 0, h
    # b = A*1: 6 minus the number of neighbours, exact in any summation order
    g = np.arange(N)
    edge = ((g > 0).astype(np.float64) + (g < N - 1))
    b = (6.0 - (edge[:, None, None] + edge[None, :, None] + edge[None, None, :])).reshape(-1)
    x = np.zeros(n)
    rh = np.zeros(8192)
    oi = np.zeros(4, np.int32); od = np.zeros(5, np.float64)
    L.shim_mv_solve_b.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    t0 = time.time()
    err = L.shim_mv_solve_b(h, b"-i cg -p jacobi -tol 1e-12 -maxiter 6000", b.ctypes.data, x.ctypes.data, oi.ctypes.data, od.ctypes.data,
                            rh.ctypes.data, len(rh))
    wall = time.time() - t0
    assert err == 0 and oi[1] == 0, (err, oi)
    it = int(oi[0])
    print(f"reference CG+Jacobi {N}^3, {threads} threads: {it} iterations, resid {od[0]:.6e}, max|x-1| {np.abs(x - 1).max():.3e}, {wall:.1f}s wall")
    out = os.path.join(HERE, f"cg_poisson_{N}.npz" if threads > 1 else f"cg_poisson_{N}_serial.npz")
    np.savez_compressed(out, grid=N, threads=threads, iters=it, resid=od[0], rhistory=rh[:int(oi[3])], xerr=np.abs(x - 1).max(), wall_s=wall,
                        options="-i cg -p jacobi -tol 1e-12")
    print("wrote", out)


if __name__ == "__main__":
    main()
