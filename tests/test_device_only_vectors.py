"""Vectors in plain device memory (the fallback lis_b200 takes when cudaMallocManaged is refused, seen with 8 processes
on one node; LIS_B200_VECTORS=device forces it): v->value is then not host-addressable, so every host pass of the
library itself must go through a copy (lisd_vec_host_view / lisd_upload / lisd_download) and never dereference it.

Both test runtimes keep cudaMalloc blocks inaccessible to the host outside kernels and runtime copies (protection keys /
mprotect), so a stray dereference in the product's host code dies with SIGSEGV here instead of passing; the same script
with managed vectors gives the reference results the device-only run has to reproduce bit for bit."""
import json
import os
import signal
import subprocess
import sys

import pytest

import harness as H

EMU_DIR = os.path.join(H.ROOT, "tests", "cudaemu")

_SCRIPT = r"""
import json, os, sys, tempfile
import numpy as np
sys.path.insert(0, sys.argv[2]); sys.path.insert(0, os.path.join(sys.argv[2], "tests"))
import lis_b200, harness as H
sh = lis_b200.Shim(sys.argv[1])
ptr, idx, val = H.poisson3d_7pt(6, 5, 4)
n = len(ptr) - 1
rng = np.random.default_rng(5)
b = rng.standard_normal(n)
out = {}
for opts, fmt in (("-i cg -p jacobi", "csr"), ("-i bicgstab -p ssor", "csr"), ("-i gmres -restart 12 -p ilu", "csr"), ("-i bicg -p none", "ell"),
                  ("-i cg -p ssor -storage jad", "csr"), ("-i bicgstab -p is", "csr"), ("-i gmres -p jacobi -scale jacobi", "dia"),
                  ("-i cg -p none -scale symm_diag", "csc"), ("-i bicgstab -p hybrid", "csr"), ("-i gs -p jacobi", "csr"),
                  ("-i bicgstab -p ilut -adds true", "csr"), ("-i cg -p jacobi -storage bsr", "csr")):
    r = sh.solve(ptr, idx, val, b, opts + " -maxiter 400", fmt=fmt)
    out[opts + "/" + fmt] = [int(r["err"]), int(r["status"]), int(r["iter"]), r["x"].tobytes().hex()]
for fmt in ("csr", "csc", "ell", "dia", "jad", "bsr", "msr", "coo"):
    d = sh.get_diagonal(fmt, ptr, idx, val, bnr=2, bnc=2)
    out["diag/" + fmt] = np.asarray(d).tobytes().hex()
for fmt in ("csr", "ell", "jad"):
    y = sh.spmv(fmt, ptr, idx, val, b)
    out["spmv/" + fmt] = np.asarray(y[0] if isinstance(y, tuple) else y).tobytes().hex()
for op in ("axpy", "xpay", "dot", "nrm2", "pmul", "reciprocal"):
    r = sh.vec_op(op, b, None if op in ("nrm2", "reciprocal") else b[::-1].copy(), 0.75)
    out["vec/" + op] = [np.asarray(t).tobytes().hex() for t in (r if isinstance(r, tuple) else (r,))]
print("RESULT " + json.dumps(out))
"""


def _run(lib, env_extra):
    env = dict(os.environ)
    env.pop("LIS_B200_VECTORS", None)
    env.update(env_extra)
    return subprocess.run([sys.executable, "-c", _SCRIPT, lib, H.ROOT], capture_output=True, text=True, env=env, timeout=900)


def _result(r):
    assert r.returncode == 0, (r.returncode, r.stdout[-1500:], r.stderr[-3000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def _libs(which):
    if which == "mock":
        return os.path.join(H.ensure_hostcheck(), "liblis_hostcheck_shim.so")
    r = subprocess.run(["make", "-C", EMU_DIR, "-j8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return os.path.join(EMU_DIR, "_build", "liblis_emu_shim.so")


@pytest.mark.parametrize("which", ["mock", "emu"])
def test_device_only_vectors_same_results_and_no_host_dereference(built, which):
    lib = _libs(which)
    managed = _result(_run(lib, {}))
    device = _result(_run(lib, {"LIS_B200_VECTORS": "device"}))
    assert managed.keys() == device.keys()
    for k in managed:
        assert managed[k] == device[k], k
    assert all(v[0] == 0 for k, v in managed.items() if k.startswith("-i ")), {k: v[:3] for k, v in managed.items() if k.startswith("-i ")}


_NEG = r"""
import ctypes as C, sys
lib = C.CDLL(sys.argv[1])
p = C.c_void_p()
assert lib.cudaMalloc(C.byref(p), C.c_size_t(4096)) == 0
assert lib.cudaMemset(p, 0, C.c_size_t(4096)) == 0
print("allocated", flush=True)
print((C.c_double * 8).from_address(p.value)[0])
"""


def test_mock_device_memory_is_closed_to_the_host(built):
    """the mock device's own negative control (the emulator's is in test_emu_kernels.py): the host reading a cudaMalloc
    block faults; with MOCK_PROTECT=0 the same script passes"""
    lib = os.path.join(H.ensure_hostcheck(), "liblis_hostcheck.so")
    r = subprocess.run([sys.executable, "-c", _NEG, lib], capture_output=True, text=True)
    assert "allocated" in r.stdout and r.returncode == -signal.SIGSEGV, (r.returncode, r.stdout, r.stderr[-500:])
    assert "touched plain device memory" in r.stderr
    r = subprocess.run([sys.executable, "-c", _NEG, lib], capture_output=True, text=True, env=dict(os.environ, MOCK_PROTECT="0"))
    assert r.returncode == 0, (r.returncode, r.stderr[-500:])
