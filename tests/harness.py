"""Test harness: matrix generators, the CPU oracle (oracle/lis_oracle.c) and the compiled
reference (oracle/_ref) as numpy-facing objects.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liblis_oracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def ensure_built() -> None:
    import lis_b200
    if not os.path.exists(os.path.join(lis_b200.LIB_DIR, "liblis_b200_shim.so")):
        lis_b200.build()
    need_ref = os.path.isdir("/root/reference/src") and not os.path.exists(os.path.join(REF_DIR, "libref_shim_omp.so"))
    if not os.path.exists(ORACLE_SO) or need_ref:
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j4", "all"], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("oracle build failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])


HOSTCHECK_DIR = os.path.join(ROOT, "tests", "hostcheck", "_build")


def ensure_hostcheck() -> str:
    """Build (if needed) the mock-device library: the product's host C code + a CUDA/kernel mock on
    the oracle (tests/hostcheck).  Test infrastructure; never loaded by the product."""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostcheck")], capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("hostcheck build failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return HOSTCHECK_DIR


def hostcheck_shim():
    import lis_b200
    if "hostcheck" not in _ref_cache:
        _ref_cache["hostcheck"] = lis_b200.Shim(os.path.join(ensure_hostcheck(), "liblis_hostcheck_shim.so"))
    return _ref_cache["hostcheck"]


_ref_cache: dict = {}


def ref_shim(kind: str):
    """The REAL reference (compiled from /root/reference by oracle/Makefile) behind the shared
    shim; None when oracle/_ref does not exist."""
    import lis_b200
    if kind not in _ref_cache:
        path = os.path.join(REF_DIR, f"libref_shim_{kind}.so")
        _ref_cache[kind] = lis_b200.Shim(path) if os.path.exists(path) else None
    return _ref_cache[kind]


# ------------------------------------------------------------------ matrices
def poisson1d(n):
    """test/spmvtest1.c:139-146: rows (i-1, i+1, i) with values (-1, -1, 2)."""
    ptr = [0]
    idx, val = [], []
    for i in range(n):
        if i > 0:
            idx.append(i - 1); val.append(-1.0)
        if i < n - 1:
            idx.append(i + 1); val.append(-1.0)
        idx.append(i); val.append(2.0)
        ptr.append(len(idx))
    return np.array(ptr, np.int32), np.array(idx, np.int32), np.array(val, np.float64)


def poisson3d_7pt(l, m, n, diag=6.0, sort=False):
    """test/spmvtest3.c:142-157 and test/test3.c:116-126: order -mn,+mn,-n,+n,-1,+1,diag.
    Vectorised: builds the 7 candidate slots per row and masks the missing ones."""
    nn = l * m * n
    ii = np.arange(nn, dtype=np.int64)
    i = ii // (m * n)
    j = (ii - i * m * n) // n
    k = ii - i * m * n - j * n
    cols = np.stack([ii - m * n, ii + m * n, ii - n, ii + n, ii - 1, ii + 1, ii], axis=1)
    mask = np.stack([i > 0, i < l - 1, j > 0, j < m - 1, k > 0, k < n - 1, np.ones(nn, bool)], axis=1)
    vals = np.tile(np.array([-1.0] * 6 + [diag]), (nn, 1))
    if sort:
        order = np.argsort(np.where(mask, cols, np.iinfo(np.int64).max), axis=1, kind="stable")
        cols = np.take_along_axis(cols, order, 1)
        vals = np.take_along_axis(vals, order, 1)
        mask = np.take_along_axis(mask, order, 1)
    ptr = np.zeros(nn + 1, np.int64)
    np.cumsum(mask.sum(1), out=ptr[1:])
    return ptr.astype(np.int32), cols[mask].astype(np.int32), vals[mask]


def poisson3d_27pt(l, m, n):
    """test/spmvtest3b.c:148-163: 27-point stencil, 26 on the diagonal and -1 elsewhere, in the
    driver's loop order (sz, sy, sx from -1 to 1)."""
    nn = l * m * n
    ii = np.arange(nn, dtype=np.int64)
    i = ii // (m * n)
    j = (ii - i * m * n) // n
    k = ii - i * m * n - j * n
    cols, mask, vals = [], [], []
    for sz in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sx in (-1, 0, 1):
                ok = (i + sz >= 0) & (i + sz < l) & (j + sy >= 0) & (j + sy < m) & (k + sx >= 0) & (k + sx < n)
                cols.append(ii + sz * m * n + sy * n + sx)
                mask.append(ok)
                vals.append(np.full(nn, 26.0 if (sz == 0 and sy == 0 and sx == 0) else -1.0))
    cols = np.stack(cols, 1); mask = np.stack(mask, 1); vals = np.stack(vals, 1)
    ptr = np.zeros(nn + 1, np.int64)
    np.cumsum(mask.sum(1), out=ptr[1:])
    return ptr.astype(np.int32), cols[mask].astype(np.int32), vals[mask]


def random_csr(n, avg_nnz, seed, *, band=None, diag_dominant=True, sorted_rows=False, empty_rows=False,
               values="uniform"):
    """Ragged random matrix: per-row length varies from 0/1 to ~2*avg_nnz (one long row of
    ~40*avg_nnz), unsorted storage order unless sorted_rows, optional band |i-j| <= band,
    diagonal stored at a random position.  diag_dominant => strictly dominant (solvable)."""
    rng = np.random.default_rng(seed)
    ptr = [0]
    idx, val = [], []
    for i in range(n):
        ln = int(rng.integers(0 if empty_rows else 1, 2 * avg_nnz + 1))
        if i == n // 3:
            ln = min(n - 1, 40 * avg_nnz)
        lo, hi = (0, n) if band is None else (max(0, i - band), min(n, i + band + 1))
        ln = min(ln, hi - lo - 1)
        cand = rng.choice(np.arange(lo, hi), size=min(hi - lo, ln + 1), replace=False)
        cols = [int(c) for c in cand if c != i][:ln]
        v = rng.uniform(-1, 1, len(cols)) if values == "uniform" else rng.standard_normal(len(cols)) * 10.0 ** rng.integers(-8, 8, len(cols))
        if diag_dominant or rng.random() < 0.7:
            d = 1.0 + float(np.abs(v).sum()) if diag_dominant else float(rng.uniform(-1, 1))
            pos = int(rng.integers(0, len(cols) + 1))
            cols.insert(pos, i)
            v = np.insert(v, pos, d)
        if empty_rows and rng.random() < 0.1:
            cols, v = [], np.zeros(0)
        if sorted_rows and len(cols):
            o = np.argsort(cols)
            cols = list(np.array(cols)[o]); v = np.asarray(v)[o]
        idx.extend(cols); val.extend(list(v))
        ptr.append(len(idx))
    return np.array(ptr, np.int32), np.array(idx, np.int32), np.array(val, np.float64)


def rand_vec(n, seed, kind="uniform"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.uniform(-1, 1, n)
    if kind == "wide":       # many magnitudes and signs: stresses rounding order
        return rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6, n)
    raise ValueError(kind)


def reference_envelope(oracle, solver, ptr, idx, val, b, *, threads=(1, 2, 3, 4, 6, 8), **kw):
    """The reference's OWN spread: the same solve with the reduction order of 1..8 OpenMP
    threads (preconditioner held fixed).  Returns the runs; a GPU result is acceptable where it
    is as close to the serial run as the reference is to itself."""
    blocks = kw.pop("ssor_blocks", 1)
    return [oracle.solve(solver, ptr, idx, val, b, nthreads=t, ssor_blocks=blocks, **kw) for t in threads]


def check_against_envelope(g, runs, what, slack=20.0, floor=1e-11):
    """iteration count inside the reference's own range (equal when the reference is
    thread-count invariant); residual history within `slack` x the reference's own
    thread-count spread (+ floor) on the common prefix."""
    its = [r["iter"] for r in runs]
    assert all(r["status"] == 0 for r in runs) and g["status"] == 0, f"{what}: status {g['status']}"
    assert min(its) <= g["iter"] <= max(its), f"{what}: {g['iter']} iterations, reference range {min(its)}..{max(its)} ({its})"
    base = runs[0]["rhistory"]
    k = min([len(g["rhistory"])] + [len(r["rhistory"]) for r in runs])
    assert k >= 2, what
    spread = np.zeros(k)
    for r in runs[1:]:
        spread = np.maximum(spread, np.abs(r["rhistory"][:k] - base[:k]) / np.abs(base[:k]))
    gap = np.abs(g["rhistory"][:k] - base[:k]) / np.abs(base[:k])
    bad = np.nonzero(gap > slack * spread + floor)[0]
    assert bad.size == 0, (f"{what}: history leaves the reference's own envelope at iteration {bad[0]}: "
                           f"gap {gap[bad[0]]:.3g}, reference spread {spread[bad[0]]:.3g}")
    return its, gap, spread


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a, np.float64); b = np.ascontiguousarray(b, np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = np.nonzero(bits(a) != bits(b))[0]
    assert bad.size == 0, f"{what}: {bad.size} of {a.size} entries differ bitwise; first at {bad[0]}: {a[bad[0]]!r} vs {b[bad[0]]!r}"


def exact_dot(x, y):
    """Correctly rounded-ish reference for reductions: long-double accumulation of exact
    products via math.fsum (exact for the sum of the rounded fp64 products' error-free split)."""
    import math
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    # error-free products via Dekker/Veltkamp split are overkill here: fsum of float128 prods
    p = x.astype(np.longdouble) * y.astype(np.longdouble)
    return float(math.fsum(p.astype(np.float64))) if p.dtype == np.float64 else float(np.sum(np.sort(p)))


# ------------------------------------------------------------------ oracle bindings
class OrcSolver(C.Structure):
    _fields_ = [("precon", C.c_int), ("ssor_omega", C.c_double), ("tol", C.c_double), ("maxiter", C.c_int),
                ("restart", C.c_int), ("nthreads", C.c_int), ("iter", C.c_int), ("retcode", C.c_int),
                ("resid", C.c_double), ("ssor_blocks", C.c_int)]


class Oracle:
    """oracle/lis_oracle.c: the CPU restatement of the reference's algorithm (plain arrays)."""

    PRECON = {"none": 0, "jacobi": 1, "ssor": 3}

    def __init__(self):
        self.lib = L = C.CDLL(ORACLE_SO)
        ci, cd = C.c_int, C.c_double
        L.orc_spmv_csr.argtypes = [ci, _i32p, _i32p, _f64p, _f64p, _f64p]
        L.orc_spmv_csr_split.argtypes = [ci, _f64p, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _f64p, _f64p]
        L.orc_spmv_ell.argtypes = [ci, ci, _i32p, _f64p, _f64p, _f64p]
        L.orc_spmv_dia.argtypes = [ci, ci, _i32p, _f64p, _f64p, _f64p, ci]
        L.orc_spmv_jad.argtypes = [ci, ci, _i32p, _i32p, _i32p, _f64p, _f64p, _f64p, ci]
        L.orc_spmv_bsr.argtypes = [ci, ci, ci, ci, _i32p, _i32p, _f64p, _f64p, _f64p]
        L.orc_spmv_csc.argtypes = [ci, _i32p, _i32p, _f64p, _f64p, _f64p]
        L.orc_sort_csr_rows.argtypes = [ci, _i32p, _i32p, _f64p]
        L.orc_csr2ell_maxnzr.argtypes = [ci, _i32p]
        L.orc_csr2ell.argtypes = [ci, _i32p, _i32p, _f64p, ci, _i32p, _f64p]
        L.orc_csr2dia_nnd.argtypes = [ci, _i32p, _i32p]
        L.orc_csr2dia.argtypes = [ci, _i32p, _i32p, _f64p, ci, _i32p, _f64p, ci]
        L.orc_csr2jad_maxnzr.argtypes = [ci, _i32p]
        L.orc_csr2jad.argtypes = [ci, _i32p, _i32p, _f64p, ci, _i32p, _i32p, _i32p, _f64p, ci]
        L.orc_csr2bsr_bnnz.argtypes = [ci, _i32p, _i32p, ci, ci, _i32p]
        L.orc_csr2bsr.argtypes = [ci, _i32p, _i32p, _f64p, ci, ci, _i32p, _i32p, _f64p]
        L.orc_csr2csc.argtypes = [ci, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
        for name in ("orc_dot",):
            getattr(L, name).argtypes = [ci, _f64p, _f64p, ci]; getattr(L, name).restype = cd
        for name in ("orc_nrm2", "orc_nrm1", "orc_sum"):
            getattr(L, name).argtypes = [ci, _f64p, ci]; getattr(L, name).restype = cd
        L.orc_nrmi.argtypes = [ci, _f64p]; L.orc_nrmi.restype = cd
        L.orc_axpy.argtypes = [ci, cd, _f64p, _f64p]
        L.orc_xpay.argtypes = [ci, _f64p, cd, _f64p]
        L.orc_axpyz.argtypes = [ci, cd, _f64p, _f64p, _f64p]
        L.orc_scale.argtypes = [ci, cd, _f64p]
        L.orc_pmul.argtypes = [ci, _f64p, _f64p, _f64p]
        L.orc_pdiv.argtypes = [ci, _f64p, _f64p, _f64p]
        L.orc_reciprocal.argtypes = [ci, _f64p]
        L.orc_shift.argtypes = [ci, cd, _f64p]
        L.orc_abs.argtypes = [ci, _f64p]
        L.orc_csr_get_diagonal.argtypes = [ci, _i32p, _i32p, _f64p, _f64p]
        L.orc_csr_split_count.argtypes = [ci, _i32p, _i32p, C.POINTER(ci), C.POINTER(ci)]
        L.orc_csr_split.argtypes = [ci, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _f64p]
        L.orc_ssor_sweep.argtypes = [ci, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, ci]
        for name in ("orc_cg", "orc_bicgstab", "orc_gmres"):
            getattr(L, name).argtypes = [ci, _i32p, _i32p, _f64p, _f64p, _f64p, C.POINTER(OrcSolver), _f64p]

    @staticmethod
    def _csr(ptr, idx, val):
        return (np.ascontiguousarray(ptr, np.int32), np.ascontiguousarray(idx, np.int32),
                np.ascontiguousarray(val, np.float64))

    # ---- format builders (the reference's layouts; nthreads=1 is the serial layout)
    def sort_rows(self, ptr, idx, val):
        ptr, idx, val = self._csr(ptr, idx, val)
        idx, val = idx.copy(), val.copy()
        self.lib.orc_sort_csr_rows(len(ptr) - 1, ptr, idx, val)
        return ptr, idx, val

    def to_ell(self, ptr, idx, val):
        ptr, idx, val = self._csr(ptr, idx, val); n = len(ptr) - 1
        m = self.lib.orc_csr2ell_maxnzr(n, ptr)
        ei = np.zeros(max(n * m, 1), np.int32); ev = np.zeros(max(n * m, 1), np.float64)
        self.lib.orc_csr2ell(n, ptr, idx, val, m, ei, ev)
        return dict(maxnzr=m, index=ei[:n * m], value=ev[:n * m])

    def to_dia(self, ptr, idx, val, nthreads=1):
        ptr, idx, val = self.sort_rows(ptr, idx, val); n = len(ptr) - 1
        nnd = self.lib.orc_csr2dia_nnd(n, ptr, idx)
        off = np.zeros(max(nnd, 1), np.int32); dv = np.zeros(max(n * nnd, 1), np.float64)
        self.lib.orc_csr2dia(n, ptr, idx, val, nnd, off, dv, nthreads)
        return dict(nnd=nnd, index=off[:nnd], value=dv[:n * nnd])

    def to_jad(self, ptr, idx, val, nthreads=1):
        ptr, idx, val = self._csr(ptr, idx, val); n = len(ptr) - 1
        m = self.lib.orc_csr2jad_maxnzr(n, ptr)
        nnz = int(ptr[-1])
        perm = np.zeros(max(n, 1), np.int32); jp = np.zeros(nthreads * (m + 1), np.int32)
        ji = np.zeros(max(nnz, 1), np.int32); jv = np.zeros(max(nnz, 1), np.float64)
        self.lib.orc_csr2jad(n, ptr, idx, val, m, perm, jp, ji, jv, nthreads)
        return dict(maxnzr=m, row=perm[:n], ptr=jp, index=ji[:nnz], value=jv[:nnz])

    def to_bsr(self, ptr, idx, val, bnr, bnc):
        ptr, idx, val = self._csr(ptr, idx, val); n = len(ptr) - 1
        nr = 1 + (n - 1) // bnr
        bptr = np.zeros(nr + 1, np.int32)
        bnnz = self.lib.orc_csr2bsr_bnnz(n, ptr, idx, bnr, bnc, bptr)
        bidx = np.zeros(max(bnnz, 1), np.int32); bv = np.zeros(max(bnnz * bnr * bnc, 1), np.float64)
        self.lib.orc_csr2bsr(n, ptr, idx, val, bnr, bnc, bptr, bidx, bv)
        return dict(nr=nr, bnr=bnr, bnc=bnc, bnnz=bnnz, bptr=bptr, bindex=bidx[:bnnz], value=bv[:bnnz * bnr * bnc])

    def to_csc(self, ptr, idx, val):
        ptr, idx, val = self._csr(ptr, idx, val); n = len(ptr) - 1
        nnz = int(ptr[-1])
        cp = np.zeros(n + 1, np.int32); ci_ = np.zeros(max(nnz, 1), np.int32); cv = np.zeros(max(nnz, 1), np.float64)
        self.lib.orc_csr2csc(n, ptr, idx, val, cp, ci_, cv)
        return dict(ptr=cp, index=ci_[:nnz], value=cv[:nnz])

    def split(self, ptr, idx, val):
        ptr, idx, val = self._csr(ptr, idx, val); n = len(ptr) - 1
        nl, nu = C.c_int(0), C.c_int(0)
        self.lib.orc_csr_split_count(n, ptr, idx, C.byref(nl), C.byref(nu))
        lp = np.zeros(n + 1, np.int32); up = np.zeros(n + 1, np.int32)
        li = np.zeros(max(nl.value, 1), np.int32); ui = np.zeros(max(nu.value, 1), np.int32)
        lv = np.zeros(max(nl.value, 1)); uv = np.zeros(max(nu.value, 1)); d = np.zeros(max(n, 1))
        self.lib.orc_csr_split(n, ptr, idx, val, lp, li, lv, up, ui, uv, d)
        return dict(lptr=lp, lidx=li, lval=lv, uptr=up, uidx=ui, uval=uv, diag=d[:n])

    # ---- SpMV in any format, starting from CSR like the drivers do
    def spmv(self, fmt, ptr, idx, val, x, *, bnr=2, bnc=2, sort_rows=False, split=False, nthreads=1):
        ptr, idx, val = self._csr(ptr, idx, val)
        if sort_rows:
            ptr, idx, val = self.sort_rows(ptr, idx, val)
        n = len(ptr) - 1
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros(max(n, 1))
        L = self.lib
        if n == 0:
            return y[:0]
        if fmt == "csr" and split:
            s = self.split(ptr, idx, val)
            L.orc_spmv_csr_split(n, s["diag"], s["lptr"], s["lidx"], s["lval"], s["uptr"], s["uidx"], s["uval"], x, y)
        elif fmt == "csr":
            L.orc_spmv_csr(n, ptr, idx, val, x, y)
        elif fmt == "ell":
            e = self.to_ell(ptr, idx, val)
            L.orc_spmv_ell(n, e["maxnzr"], np.ascontiguousarray(e["index"]), np.ascontiguousarray(e["value"]), x, y)
        elif fmt == "dia":
            d = self.to_dia(ptr, idx, val, nthreads)
            L.orc_spmv_dia(n, d["nnd"], np.ascontiguousarray(d["index"]), np.ascontiguousarray(d["value"]), x, y, nthreads)
        elif fmt == "jad":
            j = self.to_jad(ptr, idx, val, nthreads)
            L.orc_spmv_jad(n, j["maxnzr"], j["ptr"], np.ascontiguousarray(j["row"]), np.ascontiguousarray(j["index"]),
                           np.ascontiguousarray(j["value"]), x, y, nthreads)
        elif fmt == "bsr":
            b = self.to_bsr(ptr, idx, val, bnr, bnc)
            xx = np.zeros(b["nr"] * max(bnr, bnc) + bnc + n)      # room for the padded last block column
            xx[:n] = x
            L.orc_spmv_bsr(n, b["nr"], bnr, bnc, b["bptr"], np.ascontiguousarray(b["bindex"]),
                           np.ascontiguousarray(b["value"]), xx, y)
        elif fmt == "csc":
            c = self.to_csc(ptr, idx, val)
            L.orc_spmv_csc(n, c["ptr"], np.ascontiguousarray(c["index"]), np.ascontiguousarray(c["value"]), x, y)
        else:
            raise ValueError(fmt)
        return y[:n]

    # ---- BLAS-1
    def vec_op(self, op, x, y=None, alpha=0.0, nthreads=1):
        x = np.ascontiguousarray(x, np.float64).copy(); n = len(x)
        y = np.ascontiguousarray(y if y is not None else np.zeros(n), np.float64).copy()
        z = np.zeros(max(n, 1)); L = self.lib
        px, py = (x, y) if n else (np.zeros(1), np.zeros(1))
        if op == "axpy": L.orc_axpy(n, alpha, px, py); return y, None, 0.0
        if op == "xpay": L.orc_xpay(n, px, alpha, py); return y, None, 0.0
        if op == "axpyz": L.orc_axpyz(n, alpha, px, py, z); return z[:n], None, 0.0
        if op == "scale": L.orc_scale(n, alpha, px); return x, None, 0.0
        if op == "copy": return x.copy(), None, 0.0
        if op == "set_all": return np.full(n, alpha), None, 0.0
        if op == "pmul": L.orc_pmul(n, px, py, z); return z[:n], None, 0.0
        if op == "pdiv": L.orc_pdiv(n, px, py, z); return z[:n], None, 0.0
        if op == "reciprocal": L.orc_reciprocal(n, px); return x, None, 0.0
        if op == "abs": L.orc_abs(n, px); return x, None, 0.0
        if op == "shift": L.orc_shift(n, alpha, px); return x, None, 0.0
        if op == "swap": return y, x, 0.0
        if op == "dot": return None, None, L.orc_dot(n, px, py, nthreads)
        if op == "nrm2": return None, None, L.orc_nrm2(n, px, nthreads)
        if op == "nrm1": return None, None, L.orc_nrm1(n, px, nthreads)
        if op == "nrmi": return None, None, L.orc_nrmi(n, px)
        if op == "sum": return None, None, L.orc_sum(n, px, nthreads)
        raise ValueError(op)

    def get_diagonal(self, ptr, idx, val):
        ptr, idx, val = self._csr(ptr, idx, val); n = len(ptr) - 1
        d = np.zeros(max(n, 1))
        self.lib.orc_csr_get_diagonal(n, ptr, idx, val, d)
        return d[:n]

    def psolve(self, ptr, idx, val, b, precon, omega=1.0, nthreads=1):
        ptr, idx, val = self._csr(ptr, idx, val); n = len(ptr) - 1
        b = np.ascontiguousarray(b, np.float64)
        if precon == "none":
            return b.copy()
        if precon == "jacobi":
            d = self.get_diagonal(ptr, idx, val)
            self.lib.orc_reciprocal(n, d)
            z = np.zeros(n); self.lib.orc_pmul(n, b, d, z)
            return z
        s = self.split(ptr, idx, val)
        wd = 1.0 / (omega * s["diag"])
        x = np.zeros(n)
        self.lib.orc_ssor_sweep(n, s["lptr"], s["lidx"], s["lval"], s["uptr"], s["uidx"], s["uval"],
                                np.ascontiguousarray(wd), b, x, nthreads)
        return x

    def solve(self, solver, ptr, idx, val, b, *, precon="none", tol=1e-12, maxiter=1000, restart=40, omega=1.0,
              nthreads=1, ssor_blocks=0, x0=None):
        ptr, idx, val = self._csr(ptr, idx, val); n = len(ptr) - 1
        s = OrcSolver(self.PRECON[precon], omega, tol, maxiter, restart, nthreads, 0, 0, 0.0, ssor_blocks)
        x = np.ascontiguousarray(x0 if x0 is not None else np.zeros(n), np.float64).copy()
        rh = np.zeros(maxiter + 2)
        fn = {"cg": self.lib.orc_cg, "bicgstab": self.lib.orc_bicgstab, "gmres": self.lib.orc_gmres}[solver]
        fn(n, ptr, idx, val, np.ascontiguousarray(b, np.float64), x, C.byref(s), rh)
        ln = s.iter + 1 - (1 if s.retcode != 0 else 0)
        return dict(x=x, iter=s.iter, status=s.retcode, resid=s.resid, rhistory=rh[:max(ln, 0)].copy())
