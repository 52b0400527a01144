"""ctypes bindings: the product library, its raw kernel ABI, and the shared test shim."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(REPO_ROOT, "lis_b200", "_lib")

# LIS_MATRIX_* storage-format codes (include/lis.h)
FMT = {"csr": 1, "csc": 2, "msr": 3, "dia": 4, "ell": 5, "jad": 6, "bsr": 7, "bsc": 8, "vbr": 9, "coo": 10, "dns": 11}

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(verbose: bool = False) -> None:
    """Compile the CUDA kernels (sm_100a) + host C into lis_b200/_lib (nvcc cross-compiles
    without a GPU)."""
    r = subprocess.run(["make", "-C", REPO_ROOT, "-j8", "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building liblis_b200.so failed")


def _require(path: str) -> str:
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `make` (or __graft_entry__.build()); "
                           "lis_b200 has no fallback path without its CUDA library")
    return path


def load_library() -> C.CDLL:
    """The product: liblis_b200.so (lis.h API + lisb200_* kernel ABI)."""
    return C.CDLL(_require(os.path.join(LIB_DIR, "liblis_b200.so")))


def load_kernels() -> C.CDLL:
    """Same library, with argtypes set for the raw kernel C-ABI of include/lis_b200_kernels.h
    that the benchmarks drive directly (device pointers as integers)."""
    lib = load_library()
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    lib.lisb200_sm_count.restype = ci
    lib.lisb200_error_string.restype = C.c_char_p
    lib.lisb200_error_string.argtypes = [ci]
    lib.lisb200_reduce_slots.restype = ci
    sigs = {
        "lisb200_spmv_csr": [ci, vp, vp, vp, vp, vp, vp],
        "lisb200_spmv_csr_tma": [ci, ci, ci, ci, vp, vp, vp, vp, vp, vp],
        "lisb200_spmv_csr_tma_dot": [ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "lisb200_spmv_csr_split": [ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "lisb200_spmv_csr_dot": [ci, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "lisb200_spmv_ell": [ci, ci, ci, vp, vp, vp, vp, vp],
        "lisb200_spmv_dia": [ci, ci, ci, ci, vp, vp, vp, vp, vp],
        "lisb200_spmv_jad": [ci, ci, vp, vp, vp, vp, vp, vp, vp],
        "lisb200_spmv_bsr": [ci, ci, ci, ci, vp, vp, vp, vp, vp, vp],
        "lisb200_copy": [ci, vp, vp, vp],
        "lisb200_axpy": [ci, cd, vp, vp, vp],
        "lisb200_xpay": [ci, vp, cd, vp, vp],
        "lisb200_axpyz": [ci, cd, vp, vp, vp, vp],
        "lisb200_scale": [ci, cd, vp, vp],
        "lisb200_pmul": [ci, vp, vp, vp, vp],
        "lisb200_set_all": [ci, cd, vp, vp],
        "lisb200_reduce": [ci, ci, vp, vp, vp, vp, vp, vp],
        "lisb200_dot2": [ci, vp, vp, vp, vp, vp, vp],
        "lisb200_cg_update": [ci, cd, vp, vp, vp, vp, vp, vp, vp, vp],
        "lisb200_jacobi_dot": [ci, vp, vp, vp, vp, vp, vp, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ci
    return lib


def device_available() -> bool:
    """True when the product library finds a usable CUDA device."""
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


class Shim:
    """numpy-facing wrapper of one build of tests/shim/lis_shim.c.

    ``Shim(path)`` works for the lis_b200 build and for the reference builds alike: the C
    entry points are identical, only the library underneath differs."""

    def __init__(self, path: str, init_args: str = ""):
        self.path = path
        self.lib = C.CDLL(_require(path))
        L = self.lib
        L.shim_begin.argtypes = [C.c_char_p]
        L.shim_spmv.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int, C.c_int, C.c_int, C.c_int,
                                _f64p, _f64p, C.c_int, C.POINTER(C.c_double)]
        L.shim_convert_open.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int, C.c_int, C.c_int]
        L.shim_convert_dims.argtypes = [C.c_int, _i32p]
        L.shim_convert_copy.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.shim_vec_op.argtypes = [C.c_int, C.c_int, C.c_double, _f64p, _f64p, _f64p, _f64p, C.POINTER(C.c_double)]
        L.shim_get_diagonal.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, C.c_int, C.c_int, _f64p]
        L.shim_psolve.argtypes = [C.c_int, _i32p, _i32p, _f64p, C.c_char_p, _f64p, _f64p]
        L.shim_psolveh.argtypes = [C.c_int, _i32p, _i32p, _f64p, C.c_char_p, _f64p, _f64p]
        L.shim_solve.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_char_p,
                                 _i32p, _f64p, _f64p, C.c_int]
        err = L.shim_begin(init_args.encode())
        if err:
            raise RuntimeError(f"lis_initialize failed ({err}) in {path}")
        self.is_b200 = bool(L.shim_is_b200())

    # -- helpers
    @staticmethod
    def _csr(ptr, idx, val):
        return (np.ascontiguousarray(ptr, np.int32), np.ascontiguousarray(idx, np.int32),
                np.ascontiguousarray(val, np.float64))

    def set_threads(self, n: int) -> None:
        self.lib.shim_set_threads(int(n))

    def max_threads(self) -> int:
        return int(self.lib.shim_max_threads())

    def spmv(self, fmt, ptr, idx, val, x, *, bnr=0, bnc=0, sort_rows=False, split=False, iters=0):
        """y = A x through lis_matrix_set_csr -> lis_matrix_convert(fmt) -> lis_matvec.
        Returns (y, seconds spent in `iters` timed lis_matvec calls)."""
        ptr, idx, val = self._csr(ptr, idx, val)
        n = len(ptr) - 1
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(max(n, 1), np.float64)
        sec = C.c_double(0.0)
        code = FMT[fmt] if isinstance(fmt, str) else int(fmt)
        err = self.lib.shim_spmv(code, n, ptr, idx, val, bnr, bnc, int(sort_rows), int(split),
                                 x if n else np.zeros(1), y, iters, C.byref(sec))
        if err:
            raise RuntimeError(f"shim_spmv({fmt}) failed with Lis error {err}")
        return y[:n], sec.value

    def convert(self, fmt, ptr, idx, val, *, bnr=0, bnc=0, sort_rows=False):
        """Arrays of the converted storage format, as the library lays them out."""
        ptr, idx, val = self._csr(ptr, idx, val)
        n = len(ptr) - 1
        code = FMT[fmt] if isinstance(fmt, str) else int(fmt)
        h = self.lib.shim_convert_open(code, n, ptr, idx, val, bnr, bnc, int(sort_rows))
        if h < 0:
            raise RuntimeError(f"shim_convert_open({fmt}) failed ({h})")
        return self._grab_handle(h, fmt)

    def input_mm(self, path, fmt="csr"):
        """lis_input on a Matrix Market file -> (matrix arrays dict, b or None, x or None)."""
        self.lib.shim_input_open.argtypes = [C.c_char_p, C.c_int, _i32p, _f64p, C.c_int]
        has = np.zeros(2, np.int32)
        cap = 1 << 22
        bx = np.zeros(cap, np.float64)
        h = self.lib.shim_input_open(path.encode(), FMT[fmt], has, bx, cap)
        if h < 0:
            raise RuntimeError(f"lis_input({path}) failed ({h})")
        out = self._grab_handle(h, fmt)
        n = out["n"]
        return out, (bx[:n].copy() if has[0] else None), (bx[n:2 * n].copy() if has[1] else None)

    def assemble(self, n, rows, cols, vals, flags=None, fmt="csr"):
        """lis_matrix_set_value entry by entry + lis_matrix_assemble -> arrays dict."""
        self.lib.shim_assemble_open.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _f64p, _i32p, C.c_int]
        rows = np.ascontiguousarray(rows, np.int32); cols = np.ascontiguousarray(cols, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        flags = np.ascontiguousarray(flags if flags is not None else np.zeros(len(rows)), np.int32)
        h = self.lib.shim_assemble_open(n, len(rows), rows, cols, vals, flags, FMT[fmt])
        if h < 0:
            raise RuntimeError(f"assembly failed ({h})")
        return self._grab_handle(h, fmt)

    def sort_id(self, keys, vals):
        self.lib.shim_sort_id.argtypes = [C.c_int, _i32p, _f64p]
        k = np.ascontiguousarray(keys, np.int32).copy(); v = np.ascontiguousarray(vals, np.float64).copy()
        self.lib.shim_sort_id(len(k), k, v)
        return k, v

    def parse_options(self, text):
        self.lib.shim_parse_options.argtypes = [C.c_char_p, _i32p, _f64p]
        o = np.zeros(8, np.int32); d = np.zeros(2, np.float64)
        err = self.lib.shim_parse_options(text.encode(), o, d)
        keys = ["solver", "precon", "maxiter", "restart", "storage", "output", "conv_cond", "initx_zeros"]
        return int(err), dict(zip(keys, map(int, o))), dict(tol=float(d[0]), ssor_omega=float(d[1]))

    def _grab_handle(self, h, fmt):
        n_unused = 0
        dims = np.zeros(11, np.int32)
        self.lib.shim_convert_dims(h, dims)
        d = dict(zip(["n", "nnz", "maxnzr", "nnd", "nr", "bnr", "bnc", "bnnz", "type", "nc", "ndz"], map(int, dims)))
        n = d["n"]

        def grab(which, count, dtype):
            out = np.empty(max(count, 1), dtype)
            rc = self.lib.shim_convert_copy(h, which, out.ctypes.data, count)
            if rc:
                raise RuntimeError(f"shim_convert_copy({which}) failed ({rc})")
            return out[:count]

        out = dict(d)
        if fmt in ("csr", "csc"):
            out["ptr"] = grab(0, n + 1, np.int32)
            nnz = int(out["ptr"][-1])
            out["index"] = grab(1, nnz, np.int32)
            out["value"] = grab(2, nnz, np.float64)
        elif fmt == "ell":
            out["index"] = grab(1, n * d["maxnzr"], np.int32)
            out["value"] = grab(2, n * d["maxnzr"], np.float64)
        elif fmt == "dia":
            out["index"] = grab(1, d["nnd"], np.int32)
            out["value"] = grab(2, n * d["nnd"], np.float64)
        elif fmt == "jad":
            out["ptr"] = grab(0, d["maxnzr"] + 1, np.int32)
            out["row"] = grab(3, n, np.int32)
            nnz = int(out["ptr"][-1])
            out["index"] = grab(1, nnz, np.int32)
            out["value"] = grab(2, nnz, np.float64)
        elif fmt == "bsr":
            out["bptr"] = grab(4, d["nr"] + 1, np.int32)
            out["bindex"] = grab(5, d["bnnz"], np.int32)
            out["value"] = grab(2, d["bnnz"] * d["bnr"] * d["bnc"], np.float64)
        elif fmt == "msr":
            out["index"] = grab(1, d["nnz"] + d["ndz"] + 1, np.int32)
            out["value"] = grab(2, d["nnz"] + d["ndz"] + 1, np.float64)
        elif fmt == "coo":
            out["row"] = grab(3, d["nnz"], np.int32)
            out["col"] = grab(6, d["nnz"], np.int32)
            out["value"] = grab(2, d["nnz"], np.float64)
        elif fmt == "bsc":
            out["bptr"] = grab(4, d["nc"] + 1, np.int32)
            out["bindex"] = grab(5, d["bnnz"], np.int32)
            out["value"] = grab(2, d["bnnz"] * d["bnr"] * d["bnc"], np.float64)
        elif fmt == "vbr":
            out["row"] = grab(3, d["nr"] + 1, np.int32)
            out["col"] = grab(6, d["nc"] + 1, np.int32)
            out["ptr"] = grab(0, d["bnnz"] + 1, np.int32)
            out["bptr"] = grab(4, d["nr"] + 1, np.int32)
            out["bindex"] = grab(5, d["bnnz"], np.int32)
            out["value"] = grab(2, d["nnz"], np.float64)
        elif fmt == "dns":
            out["value"] = grab(2, n * n, np.float64)
        self.lib.shim_convert_close(h)
        return out

    _OPS = {"axpy": 0, "xpay": 1, "axpyz": 2, "scale": 3, "copy": 4, "set_all": 5, "pmul": 6, "pdiv": 7,
            "reciprocal": 8, "abs": 9, "shift": 10, "swap": 11, "dot": 20, "nrm2": 21, "nrm1": 22, "nrmi": 23,
            "sum": 24}

    def vec_op(self, op, x, y=None, alpha=0.0):
        """One lis_vector_* call on fresh vectors; returns (out_a, out_b, scalar)."""
        x = np.ascontiguousarray(x, np.float64)
        n = len(x)
        y = np.ascontiguousarray(y if y is not None else np.zeros(n), np.float64)
        oa, ob = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        s = C.c_double(0.0)
        pad = np.zeros(1)
        err = self.lib.shim_vec_op(self._OPS[op], n, float(alpha), x if n else pad, y if n else pad, oa, ob, C.byref(s))
        if err:
            raise RuntimeError(f"shim_vec_op({op}) failed with Lis error {err}")
        return oa[:n], ob[:n], s.value

    def vec_mismatch(self, op) -> int:
        return int(self.lib.shim_vec_mismatch(self._OPS[op]))

    def get_diagonal(self, fmt, ptr, idx, val, *, bnr=0, bnc=0):
        ptr, idx, val = self._csr(ptr, idx, val)
        n = len(ptr) - 1
        d = np.empty(n, np.float64)
        err = self.lib.shim_get_diagonal(FMT[fmt], n, ptr, idx, val, bnr, bnc, d)
        if err:
            raise RuntimeError(f"shim_get_diagonal failed with Lis error {err}")
        return d

    def psolve(self, ptr, idx, val, b, options, transposed=False):
        ptr, idx, val = self._csr(ptr, idx, val)
        n = len(ptr) - 1
        x = np.empty(n, np.float64)
        fn = self.lib.shim_psolveh if transposed else self.lib.shim_psolve
        err = fn(n, ptr, idx, val, options.encode(), np.ascontiguousarray(b, np.float64), x)
        if err:
            raise RuntimeError(f"shim_psolve failed with Lis error {err}")
        return x

    def solve(self, ptr, idx, val, b, options, *, fmt="csr", x0=None, rh_cap=20000):
        """lis_solve; returns dict(x, iter, status, err, resid, rhistory, time, itime, ptime)."""
        ptr, idx, val = self._csr(ptr, idx, val)
        n = len(ptr) - 1
        x = np.ascontiguousarray(x0 if x0 is not None else np.zeros(n), np.float64).copy()
        oi = np.zeros(4, np.int32)
        od = np.zeros(4, np.float64)
        rh = np.zeros(rh_cap, np.float64)
        err = self.lib.shim_solve(FMT[fmt], n, ptr, idx, val, np.ascontiguousarray(b, np.float64), x,
                                  options.encode(), oi, od, rh, rh_cap)
        return dict(x=x, iter=int(oi[0]), status=int(oi[1]), err=int(err), resid=float(od[0]),
                    rhistory=rh[:int(oi[3])].copy(), time=float(od[1]), itime=float(od[2]), ptime=float(od[3]))


def load_shim(init_args: str = "") -> Shim:
    """The shim built against lis_b200 (the product side of a parity test)."""
    return Shim(os.path.join(LIB_DIR, "liblis_b200_shim.so"), init_args)
