"""The eigensolvers (power, inverse, Rayleigh quotient, CG, CR, subspace, Lanczos)
against outputs of the compiled serial reference in tests/golden/esolve.npz (made by
`tests/golden/make_golden.py esolve`, one reference process per case).

* mock device (tests/hostcheck, sequential reductions): eigenvalue, iteration count, status, residual
  history, eigenvector and the per-mode arrays equal the reference's BIT FOR BIT -- the host control
  flow, the scalar arithmetic and the small dense helpers are the reference's;
* kernel emulator / GPU: the reduction tree differs, so eigenvalues agree to 1e-9 relative, iteration
  counts to a few steps (like the reference between its own thread counts)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import harness as H

sys.path.insert(0, os.path.join(H.ROOT, "tests", "golden"))
from make_golden import ESOLVE_CASES, ESOLVE_INITS  # noqa: E402

GOLD = os.path.join(H.GOLDEN, "esolve.npz")
WORKER = os.path.join(H.ROOT, "tests", "esolve_worker.py")


def run_worker(lib, init, cases, matrices=None):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "o.npz")
        env = dict(os.environ)
        if matrices:
            env["ESOLVE_MATRICES"] = ",".join(matrices)
        r = subprocess.run([sys.executable, WORKER, lib, init, path, *cases], capture_output=True, text=True, timeout=int(os.environ.get("ESOLVE_TIMEOUT", "600")), env=env)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
        z = np.load(path)
        return {k: z[k] for k in z.files}


def cases_for(iname):
    return [c for c in ESOLVE_CASES if iname == "default" or c.split("|")[0] in ("ii", "rqi", "cg", "cr", "li", "ai")]


@pytest.mark.parametrize("iname", list(ESOLVE_INITS))
def test_eigensolvers_bit_for_bit_on_mock_device(built, iname):
    d = H.ensure_hostcheck()
    got = run_worker(os.path.join(d, "liblis_hostcheck_shim.so"), ESOLVE_INITS[iname], cases_for(iname))
    gold = np.load(GOLD)
    keys = [k for k in gold.files if k.startswith(iname + "_")]
    assert len(keys) > 50
    for k in keys:
        a = gold[k]
        b = got[k[len(iname) + 1:]]
        if "lirv" in k and k.endswith("_d"):
            b = b[:1]
        assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8)), (k, a[:4], b[:4])


def check_close(gold, got, iname, matrices, cases=None):
    for case in cases or cases_for(iname):
        name = case.split("|")[0]
        for m in matrices:
            key = f"{name}_{m}"
            g_d, r_d = got[key + "_d"], gold[f"{iname}_{key}_d"]
            assert got[key + "_rc"][0] == 0 and got[key + "_rc"][3] == 0, (key, got[key + "_rc"])
            if name != "lirv" and (gold[f"{iname}_{key}_rc"][2] != 0 or r_d[1] > 1e-10):
                continue        # not converged in the reference either (power iteration from a symmetric start): rounding decides
            assert abs(g_d[0] - r_d[0]) <= 1e-9 * abs(r_d[0]), (key, g_d[0], r_d[0])
            if name == "lirv":
                assert np.allclose(got[key + "_ev"], gold[f"{iname}_{key}_ev"], rtol=1e-9, atol=1e-12), key
                continue
            r_rc = gold[f"{iname}_{key}_rc"]
            assert got[key + "_rc"][2] == r_rc[2], (key, "status", got[key + "_rc"], r_rc)
            it, rit = int(got[key + "_rc"][1]), int(r_rc[1])
            if name != "rqi":           # Rayleigh quotient iteration's early phase is rounding-sensitive (8 vs 17 steps seen)
                assert abs(it - rit) <= max(5, rit // 2), (key, it, rit)
            if r_rc[2] == 0:
                # eigenvector up to sign
                x, rx = got[key + "_x"], gold[f"{iname}_{key}_x"]
                assert min(np.abs(x - rx).max(), np.abs(x + rx).max()) < 1e-6, key
            if name in ("si", "sipi", "li", "ai", "aicr"):
                conv = gold[f"{iname}_{key}_er"] <= 1e-10          # modes the reference itself converged
                if name in ("si", "sipi"):
                    # Subspace iteration starts mode j+1 from the converged vector of mode j and removes that
                    # vector from it (lis_esolver_si.c:190-215): what is left is the rounding noise of one dot
                    # product.  With the sequential dot it is ~1e-17 and inverse iteration amplifies it into the
                    # next eigenvector; with the GPU's reduction tree <v,v> can come out as exactly 1 (seen on
                    # the 40-row 1-D case on the kernel emulator), the start vector is exactly 0 and the mode is
                    # NaN -- in the reference too, given that dot.  Mode 0 is always compared (above); higher
                    # modes where this run produced one.
                    e = got[key + "_er"]
                    conv = conv & np.isfinite(e) & (np.nan_to_num(e, nan=1.0) <= 1e-10)
                    assert conv[0], key
                assert np.allclose(got[key + "_ev"][conv], gold[f"{iname}_{key}_ev"][conv], rtol=1e-8, atol=1e-12), key


def test_eigensolvers_with_estorage_formats(built):
    """-estorage <fmt>: the shifted solvers (Rayleigh quotient, Lanczos / Arnoldi refinement) need
    lis_matrix_shift_diagonal in the chosen format.  Mock device against the compiled serial reference run here
    (one process per case -- the reference does not survive several eigensolver kinds in one process): same
    status and iteration counts, eigenvalue to 1e-9"""
    ref = os.path.join(H.ROOT, "oracle", "_ref", "libref_shim_serial.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ours = os.path.join(H.ensure_hostcheck(), "liblis_hostcheck_shim.so")
    cases = ["rqimsr|-e rqi -estorage msr", "libsr|-e li -ss 2 -estorage bsr", "aicoo|-e ai -ss 2 -estorage coo", "rqijad|-e rqi -estorage jad",
             "iidns|-e ii -estorage dns", "livbr|-e li -ss 2 -estorage vbr", "rqibsc|-e rqi -estorage bsc", "crell|-e cr -estorage ell"]
    for c in cases:
        name = c.split("|")[0]
        a = run_worker(ours, "", [c], matrices=("p7",))
        b = run_worker(ref, "", [c], matrices=("p7",))
        ra, rb = [int(v) for v in a[f"{name}_p7_rc"]], [int(v) for v in b[f"{name}_p7_rc"]]
        if "jad" in name:            # our JAD layout orders equal-length rows differently: products agree to rounding, RQI's count moves by one
            assert ra[0] == rb[0] and ra[2:] == rb[2:] and abs(ra[1] - rb[1]) <= 2, (c, ra, rb)
        else:
            assert ra == rb, (c, ra, rb)
        assert abs(a[f"{name}_p7_d"][0] - b[f"{name}_p7_d"][0]) <= 1e-9 * abs(b[f"{name}_p7_d"][0]), c


@pytest.mark.parametrize("iname", list(ESOLVE_INITS))
def test_eigensolvers_on_kernel_emulator(built, iname):
    d = os.path.join(H.ROOT, "tests", "cudaemu")
    r = subprocess.run(["make", "-C", d, "-j8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    light = [c for c in cases_for(iname) if c.split("|")[0] in ("rqi", "cg", "cr", "li", "lirv", "aicr")]     # seconds, not minutes
    got = run_worker(os.path.join(d, "_build", "liblis_emu_shim.so"), ESOLVE_INITS[iname], light, matrices=("p7",))
    check_close(np.load(GOLD), got, iname, ("p7",), light)


@pytest.mark.gpu
@pytest.mark.parametrize("iname", list(ESOLVE_INITS))
def test_eigensolvers_on_gpu(built, iname):
    import lis_b200
    got = run_worker(os.path.join(lis_b200.LIB_DIR, "liblis_b200_shim.so"), ESOLVE_INITS[iname], cases_for(iname), matrices=("p7", "p1d"))
    check_close(np.load(GOLD), got, iname, ("p7", "p1d"))
