"""Runs eigensolver cases through one build of tests/shim/lis_shim.c (path = argv[1]) with the library
initialised with argv[2] (the inner linear solver's options travel through lis_initialize, like on a
driver's command line) and stores what came back in argv[3] (.npz).  Cases: argv[4:] = "name|options"."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402
import lis_b200  # noqa: E402


def matrices():
    yield "p7", H.poisson3d_7pt(6, 5, 4)
    yield "p1d", H.poisson1d(40)
    ptr, idx, val = H.random_csr(60, 4, 7, sorted_rows=True)
    # symmetrise: A + A^T (dense detour is fine at this size), diagonally dominant
    import scipy.sparse as sp
    a = sp.csr_matrix((val, idx, ptr), shape=(60, 60))
    s = (a + a.T).tocsr(); s.sort_indices()
    yield "symrand", (s.indptr.astype(np.int32), s.indices.astype(np.int32), s.data.astype(np.float64))


def main():
    path, init_args, out = sys.argv[1], sys.argv[2], sys.argv[3]
    shim = lis_b200.Shim(path, init_args)
    L = shim.lib
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C"); f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    L.shim_esolve.argtypes = [C.c_int, C.c_int, i32p, i32p, f64p, f64p, C.c_char_p, i32p, f64p, f64p, C.c_int, f64p, f64p, i32p, C.c_int]
    res = {}
    for case in sys.argv[4:]:
        name, opts = case.split("|")
        for mname, (ptr, idx, val) in matrices():
            if os.environ.get("ESOLVE_MATRICES") and mname not in os.environ["ESOLVE_MATRICES"].split(","):
                continue
            n = len(ptr) - 1
            x = H.rand_vec(n, 3)
            oi = np.zeros(8, np.int32); od = np.zeros(4); rh = np.zeros(20000)
            ev = np.zeros(16); er = np.zeros(16); ei = np.zeros(16, np.int32)
            rc = L.shim_esolve(1, n, ptr, idx, val, x, opts.encode(), oi, od, rh, len(rh), ev, er, ei, 16)
            key = f"{name}_{mname}"
            res[key + "_rc"] = np.array([rc, oi[0], oi[1], oi[2], oi[4]])
            res[key + "_d"] = od[:2].copy()
            res[key + "_rh"] = rh[:oi[3]].copy()
            res[key + "_x"] = x.copy()
            res[key + "_ev"] = ev[:max(int(oi[4]), 0)].copy(); res[key + "_er"] = er[:max(int(oi[4]), 0)].copy()
            res[key + "_ei"] = ei[:max(int(oi[4]), 0)].copy()
    np.savez(out, **res)


if __name__ == "__main__":
    main()
