"""Seeded random shapes through the kernel emulator (tests/cudaemu): sizes around every tile / warp /
chunk boundary, empty rows, duplicate columns, one to 148 emulated SMs.
 * device-side conversion CSR -> ELL/DIA/JAD/BSR == the host builders' arrays, entry for entry;
 * the overlapped host-buffer product == the oracle's bits for both CSR kernels and random chunkings.
LIS_B200_FUZZ=N multiplies the number of trials (the full runs were 120 / 60 trials)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import harness as H
import lis_b200

EMU_DIR = os.path.join(H.ROOT, "tests", "cudaemu")
SCALE = int(os.environ.get("LIS_B200_FUZZ", "1"))


@pytest.fixture(scope="module")
def emu(built):
    r = subprocess.run(["make", "-C", EMU_DIR, "-j8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return lis_b200.Shim(os.path.join(EMU_DIR, "_build", "liblis_emu_shim.so"))


def rand_matrix(rng, n, maxlen, dup, sort, empty=0.2):
    lens = rng.integers(0, maxlen + 1, n)
    lens[rng.random(n) < empty] = 0
    cols = []
    for ln in lens:
        c = rng.integers(0, n, ln) if dup else (rng.choice(n, min(ln, n), replace=False) if ln else np.zeros(0, int))
        cols.append(np.sort(c) if sort else c)
    ptr = np.zeros(n + 1, np.int32)
    ptr[1:] = np.cumsum([len(c) for c in cols])
    idx = (np.concatenate(cols) if ptr[-1] else np.zeros(0)).astype(np.int32)
    val = rng.standard_normal(ptr[-1]) * 10.0 ** rng.integers(-4, 4, ptr[-1])
    return ptr, idx, val


def test_device_conversion_random_shapes(emu, monkeypatch):
    rng = np.random.default_rng(2024)
    for trial in range(24 * SCALE):
        n = int(rng.choice([1, 2, 3, 5, 31, 32, 33, 64, 100, 255, 256, 257, 1000, 1025, 2049, 4097]))
        maxlen = int(rng.choice([0, 1, 2, 3, 8, 20, 70]))
        dup, sort = bool(rng.random() < 0.3), bool(rng.random() < 0.5)
        ptr, idx, val = rand_matrix(rng, n, min(maxlen, n), dup, sort)
        for fmt, blk in (("ell", (0, 0)), ("dia", (0, 0)), ("jad", (0, 0)), ("bsr", (2, 2)), ("bsr", (3, 2)), ("bsr", (1, 4))):
            if fmt == "dia" and n > 300 and maxlen > 3:
                continue                                   # n*nnd dense diagonals: not a DIA matrix
            monkeypatch.setenv("LIS_B200_CONVERT", "host")
            want = emu.convert(fmt, ptr, idx, val, bnr=blk[0], bnc=blk[1])
            monkeypatch.setenv("LIS_B200_CONVERT", "device")
            got = emu.convert(fmt, ptr, idx, val, bnr=blk[0], bnc=blk[1])
            for k in want:
                what = (trial, n, maxlen, dup, sort, fmt, blk, k)
                if isinstance(want[k], np.ndarray):
                    assert got[k].shape == want[k].shape and np.array_equal(got[k].view(np.uint8), want[k].view(np.uint8)), what
                else:
                    assert got[k] == want[k], what


def test_overlapped_host_product_random_shapes(emu, oracle, monkeypatch):
    rng = np.random.default_rng(7)
    L = emu.lib
    vp = C.c_void_p
    L.shim_mv_open.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int]
    L.shim_mv_step_e2e_pipelined.argtypes = [C.c_int, vp, vp]
    for trial in range(12 * SCALE):
        n = int(rng.choice([8, 100, 257, 1024, 1500, 4096, 5000, 9999]))
        kind = rng.choice(["band", "rand", "stencil"])
        if kind == "band":
            ptr, idx, val = H.random_csr(n, int(rng.integers(1, 9)), int(trial), band=int(rng.integers(1, max(2, n // 3))),
                                         sorted_rows=bool(rng.random() < 0.5))
        elif kind == "rand":
            ptr, idx, val = H.random_csr(n, int(rng.integers(1, 9)), int(trial), empty_rows=True, diag_dominant=False)
        else:
            a = max(2, int(round(n ** (1 / 3))))
            ptr, idx, val = H.poisson3d_7pt(a, a, a + 1)
            n = len(ptr) - 1
        for kernel in ("tma", "tile"):
            monkeypatch.setenv("LIS_B200_CSR_KERNEL", kernel)
            monkeypatch.setenv("LIS_B200_PIPE_CHUNKS", str(int(rng.integers(2, 12))))
            monkeypatch.setenv("LISB_EMU_SMS", str(int(rng.choice([1, 2, 148]))))
            h = L.shim_mv_open(1, n, ptr.ctypes.data, idx.ctypes.data, val.ctypes.data, 0, 0, 0)
            assert h >= 0
            try:
                for rep in range(2):
                    hx = H.rand_vec(n, 100 + trial + rep, "wide")
                    hy = np.full(n, np.nan)
                    assert L.shim_mv_step_e2e_pipelined(h, hx.ctypes.data, hy.ctypes.data) == 0
                    H.assert_bits_equal(hy, oracle.spmv("csr", ptr, idx, val, hx), f"trial {trial} n={n} {kind} {kernel}")
            finally:
                L.shim_mv_close(h)


def test_triangular_sweeps_random_shapes(emu, oracle, ref_serial, monkeypatch):
    """the one-launch (dependency-polling) and level-launched SSOR sweeps for 1..8 blocks against the
    oracle, ILU(0)/ILU(2) applies and the transposed SSOR / ILU sweeps against the compiled reference,
    on random (banded and unstructured) dependency patterns: bit-exact"""
    rng = np.random.default_rng(99)
    try:
        for trial in range(10 * SCALE):
            n = int(rng.choice([1, 2, 7, 31, 32, 33, 100, 129, 500, 1023, 2000]))
            nnzr = int(rng.choice([1, 2, 4, 9, 30]))
            band = None if rng.random() < 0.4 else int(rng.integers(1, max(2, n // 2)))
            ptr, idx, val = H.random_csr(n, min(nnzr, n), int(trial), band=band, sorted_rows=bool(rng.random() < 0.5))
            b = H.rand_vec(n, trial, "wide")
            for threads in (1, int(rng.integers(2, 9))):
                emu.set_threads(threads)
                monkeypatch.setenv("LISB_EMU_SMS", str(int(rng.choice([1, 2, 148]))))
                for mode in ("syncfree", "levels"):
                    monkeypatch.setenv("LIS_B200_SSOR", mode)
                    x = emu.psolve(ptr, idx, val, b, "-p ssor -ssor_omega 1.2")
                    H.assert_bits_equal(x, oracle.psolve(ptr, idx, val, b, "ssor", omega=1.2, nthreads=threads),
                                        f"ssor trial {trial} n={n} blocks={threads} {mode}")
                if threads == 1 and n >= 2:
                    for opts in ("-p ilu", "-p ilu -ilu_fill 2", "-p ssor"):
                        for tr in ((False, True) if "ilu" in opts else (True,)):
                            H.assert_bits_equal(emu.psolve(ptr, idx, val, b, opts, transposed=tr),
                                                ref_serial.psolve(ptr, idx, val, b, opts, transposed=tr),
                                                f"{opts} transposed={tr} trial {trial} n={n}")
    finally:
        emu.set_threads(1)
