"""Solver-level measurements on the GPU box (not a test): CG+Jacobi on the 7-pt cube to
convergence, BiCGSTAB+SSOR on a synthetic unsymmetric banded matrix, GMRES(30)+Jacobi on the 27-pt
cube -- lis_b200 next to the reference's OpenMP build on the host cores, same driver code (shim).
usage: python profiles/run_solvers.py [cg_grid] [su_rows] [gm_grid] [--noref]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness as H  # noqa: E402
import lis_b200  # noqa: E402


def su_matrix(n, per_row=70, band=100000, seed=7):
    """diagonal + (per_row-1) off-diagonals within |i-j| <= band, values
    uniform(-1,1), diagonal = 1 + sum|off-diag| (strictly dominant), storage order random."""
    rng = np.random.default_rng(seed)
    k = per_row - 1
    band = min(band, n - 1)
    off = rng.integers(1, band + 1, size=(n, k), dtype=np.int64) * rng.choice(np.array([-1, 1]), size=(n, k))
    rows = np.arange(n, dtype=np.int64)[:, None]
    cols = rows + off
    cols = np.where(cols < 0, rows - off, cols)           # reflect at the ends
    cols = np.where(cols >= n, rows - off, cols)
    cols = np.clip(cols, 0, n - 1)
    cols = np.where(cols == rows, (rows + 1) % n, cols)
    vals = rng.uniform(-1, 1, size=(n, k))
    diag = 1.0 + np.abs(vals).sum(1)
    pos = rng.integers(0, per_row, n)                      # where the diagonal sits inside the row
    idx = np.empty((n, per_row), np.int32); val = np.empty((n, per_row), np.float64)
    m = np.ones((n, per_row), bool); m[np.arange(n), pos] = False
    idx[m] = cols.reshape(-1); val[m] = vals.reshape(-1)
    idx[~m] = np.arange(n); val[~m] = diag
    ptr = (np.arange(n + 1, dtype=np.int64) * per_row).astype(np.int32)
    return ptr, idx.reshape(-1), val.reshape(-1)


def run(shim, name, ptr, idx, val, b, opts, **kw):
    t0 = time.time()
    r = shim.solve(ptr, idx, val, b, opts, rh_cap=40000, **kw)
    wall = time.time() - t0
    it = r["iter"]
    print(f"  {name:10s} {opts:42s} iter={it:5d} status={r['status']} resid={r['resid']:.3e} solver_time={r['itime'] + r['ptime']:.3f}s "
          f"({(r['itime'] + r['ptime']) / max(it, 1) * 1e3:.3f} ms/it) wall={wall:.1f}s", flush=True)
    return r


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    cg_grid = int(args[0]) if len(args) > 0 else 256
    su_rows = int(args[1]) if len(args) > 1 else 1000000
    gm_grid = int(args[2]) if len(args) > 2 else 128
    noref = "--noref" in sys.argv
    H.ensure_built()
    g = lis_b200.load_shim()
    ref = None if noref else H.ref_shim("omp")
    T = ref.max_threads() if ref else 1
    print(f"reference OpenMP threads: {T}")

    if cg_grid:
        print(f"CG + Jacobi, 7-pt {cg_grid}^3 (test/test3.c system, b = A*1)")
        ptr, idx, val = H.poisson3d_7pt(cg_grid, cg_grid, cg_grid)
        n = len(ptr) - 1
        b, _ = g.spmv("csr", ptr, idx, val, np.ones(n))
        rg = run(g, "lis_b200", ptr, idx, val, b, "-i cg -p jacobi -maxiter 5000")
        if ref:
            rr = run(ref, f"ref omp{T}", ptr, idx, val, b, "-i cg -p jacobi -maxiter 5000")
            k = min(len(rg["rhistory"]), len(rr["rhistory"]))
            rel = np.abs(rg["rhistory"][:k] - rr["rhistory"][:k]) / rr["rhistory"][:k]
            print(f"  iterations {rg['iter']} vs {rr['iter']}; history gap: first quarter {rel[:k // 4].max():.2e}, "
                  f"first half {rel[:k // 2].max():.2e}, all {rel.max():.2e}; |x-1|max {np.abs(rg['x'] - 1).max():.2e}")
        for blocks in (1, T):
            g.set_threads(blocks)
            rs = run(g, f"b200 T={blocks}", ptr, idx, val, b, "-i cg -p ssor -maxiter 5000")
        g.set_threads(1)
        if ref:
            ref.set_threads(T)
            rr = run(ref, f"ref omp{T}", ptr, idx, val, b, "-i cg -p ssor -maxiter 5000")
            print(f"  CG+SSOR iterations (blocks={T}) {rs['iter']} vs {rr['iter']}")
    if su_rows:
        print(f"BiCGSTAB + SSOR, unsymmetric banded, n={su_rows}, 70 nnz/row")
        ptr, idx, val = su_matrix(su_rows)
        n = su_rows
        b = H.rand_vec(n, 3)
        for blocks in (1, T):
            g.set_threads(blocks)
            rg = run(g, f"b200 T={blocks}", ptr, idx, val, b, "-i bicgstab -p ssor")
        g.set_threads(1)
        run(g, "b200", ptr, idx, val, b, "-i bicgstab -p jacobi")
        if ref:
            ref.set_threads(T)
            rr = run(ref, f"ref omp{T}", ptr, idx, val, b, "-i bicgstab -p ssor")
            print(f"  iterations (blocks={T}) {rg['iter']} vs {rr['iter']}")
    if gm_grid:
        print(f"GMRES(30) + Jacobi, 27-pt {gm_grid}^3")
        ptr, idx, val = H.poisson3d_27pt(gm_grid, gm_grid, gm_grid)
        n = len(ptr) - 1
        b, _ = g.spmv("csr", ptr, idx, val, np.ones(n))
        rg = run(g, "lis_b200", ptr, idx, val, b, "-i gmres -restart 30 -p jacobi -maxiter 3000")
        if ref:
            rr = run(ref, f"ref omp{T}", ptr, idx, val, b, "-i gmres -restart 30 -p jacobi -maxiter 3000")
            print(f"  iterations {rg['iter']} vs {rr['iter']}")


if __name__ == "__main__":
    main()
