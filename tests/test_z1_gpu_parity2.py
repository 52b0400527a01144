"""GPU parity tests, part 2 (-m gpu): what was added after round 1's GPU minutes were spent -- the
overlapped host-buffer product, device-side format conversion, the fused Gram-Schmidt and BiCGSTAB
steps.  Same bars as test_gpu_parity.py (bit-exact against the oracle / against the unfused sequence).
Kept in a file that sorts after the proven suite; tests/test_emu_kernels.py runs the very same
functions on the kernel emulator in the CPU suite."""
import numpy as np
import pytest

import harness as H
import lis_b200

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", ["tma", "tile"])
@pytest.mark.parametrize("case", ["banded", "stencil27", "full_reach", "ragged"])
def test_matvec_host_pipelined(b200, oracle, monkeypatch, kernel, case):
    """lis_b200_matvec_host (copy-in, product, copy-out overlapped chunk-wise on three streams) leaves
    host_y, x and y with the bits of lis_vector_scatter + lis_matvec + lis_vector_gather, and its
    chunk plan never lets a row start before the x entries it reads have landed"""
    import ctypes as C
    monkeypatch.setenv("LIS_B200_CSR_KERNEL", kernel)
    monkeypatch.setenv("LIS_B200_PIPE_CHUNKS", "7")
    if case == "banded":
        ptr, idx, val = H.random_csr(5000, 5, 12, band=300, sorted_rows=True)
    elif case == "stencil27":
        ptr, idx, val = H.poisson3d_27pt(15, 14, 13)
    elif case == "full_reach":
        ptr, idx, val = H.random_csr(3000, 6, 5, values="wide")
    else:
        ptr, idx, val = H.random_csr(2600, 4, 13, empty_rows=True, diag_dominant=False)
    n = len(ptr) - 1
    L = b200.lib
    vp = C.c_void_p
    L.shim_mv_open.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int]
    L.shim_mv_step_e2e.argtypes = [C.c_int, vp, vp]
    L.shim_mv_step_e2e_pipelined.argtypes = [C.c_int, vp, vp]
    L.shim_mv_host_plan.argtypes = [C.c_int, C.c_int, vp, vp]
    L.shim_mv_get_xy.argtypes = [C.c_int, vp, vp]
    h = L.shim_mv_open(1, n, ptr.ctypes.data, idx.ctypes.data, val.ctypes.data, 0, 0, 0)
    assert h >= 0
    try:
        for seed in (1, 2, 3):                      # x changes between calls: stale reads would show
            hx = H.rand_vec(n, seed, "wide")
            hy = np.full(n, np.nan)
            assert L.shim_mv_step_e2e_pipelined(h, hx.ctypes.data, hy.ctypes.data) == 0
            want = oracle.spmv("csr", ptr, idx, val, hx)
            H.assert_bits_equal(hy, want, f"pipelined {case}/{kernel} seed {seed}")
            xo = np.empty(n); yo = np.empty(n)
            assert L.shim_mv_get_xy(h, xo.ctypes.data, yo.ctypes.data) == 0
            H.assert_bits_equal(xo, hx, "x vector after the pipelined product")
            H.assert_bits_equal(yo, want, "y vector after the pipelined product")
            hy2 = np.full(n, np.nan)
            assert L.shim_mv_step_e2e(h, hx.ctypes.data, hy2.ctypes.data) == 0
            H.assert_bits_equal(hy2, want, "three separate calls")
        rows = np.zeros(65, np.int32); need = np.zeros(64, np.int32)
        nch = L.shim_mv_host_plan(h, 64, rows.ctypes.data, need.ctypes.data)
        assert nch >= 2, nch
        assert rows[0] == 0 and rows[nch] == n and np.all(np.diff(rows[:nch + 1]) > 0)
        for c in range(nch):                         # no row of chunk c reads beyond x chunk need[c]
            cols = idx[ptr[rows[c]]:ptr[rows[c + 1]]]
            if len(cols):
                assert cols.max() < rows[need[c] + 1], (c, cols.max(), rows[need[c] + 1])
        if case in ("banded", "stencil27"):
            assert np.all(need[:nch - 1] <= np.arange(nch - 1) + 1), need[:nch]   # band: next chunk at most
    finally:
        L.shim_mv_close(h)


def test_matvec_host_pipelined_default_chunks(b200, oracle):
    """the chunking the library picks by itself (>= 2^18 rows per chunk, boundaries on multiples of
    1024 rows) on a 1.3 M-row stencil: same bits as the oracle and as the three-call sequence"""
    import ctypes as C
    ptr, idx, val = H.poisson3d_7pt(128, 128, 80, sort=True)
    n = len(ptr) - 1
    L = b200.lib
    vp = C.c_void_p
    L.shim_mv_open.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int]
    L.shim_mv_step_e2e.argtypes = [C.c_int, vp, vp]
    L.shim_mv_step_e2e_pipelined.argtypes = [C.c_int, vp, vp]
    L.shim_mv_host_plan.argtypes = [C.c_int, C.c_int, vp, vp]
    h = L.shim_mv_open(1, n, ptr.ctypes.data, idx.ctypes.data, val.ctypes.data, 0, 0, 0)
    assert h >= 0
    try:
        for seed in (5, 6):
            hx = H.rand_vec(n, seed, "wide")
            hy = np.full(n, np.nan); hy2 = np.full(n, np.nan)
            assert L.shim_mv_step_e2e_pipelined(h, hx.ctypes.data, hy.ctypes.data) == 0
            assert L.shim_mv_step_e2e(h, hx.ctypes.data, hy2.ctypes.data) == 0
            H.assert_bits_equal(hy, hy2, "overlapped vs three calls")
            H.assert_bits_equal(hy, oracle.spmv("csr", ptr, idx, val, hx), "overlapped vs oracle")
        rows = np.zeros(65, np.int32); need = np.zeros(64, np.int32)
        nch = L.shim_mv_host_plan(h, 64, rows.ctypes.data, need.ctypes.data)
        assert nch == 5 and np.all(rows[1:nch] % 1024 == 0), (nch, rows[:nch + 1])
        assert list(need[:nch]) == [1, 2, 3, 4, 4]           # a 7-point row reaches one grid plane ahead
    finally:
        L.shim_mv_close(h)


def test_queued_products_same_bits_as_synchronous_calls(b200, oracle):
    """lis_b200_matvec_async: products enqueued back to back with one synchronisation at the end (what bench.py's
    device-timed multi-GPU leg does) leave the bits of lis_matvec in y"""
    import ctypes as C
    ptr, idx, val = H.poisson3d_7pt(20, 17, 15, sort=True)
    n = len(ptr) - 1
    L = b200.lib
    vp = C.c_void_p
    L.shim_mv_open.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int]
    L.shim_mv_set_x.argtypes = [C.c_int, vp]; L.shim_mv_get_xy.argtypes = [C.c_int, vp, vp]
    h = L.shim_mv_open(1, n, ptr.ctypes.data, idx.ctypes.data, val.ctypes.data, 0, 0, 0)
    assert h >= 0
    try:
        x = H.rand_vec(n, 21, "wide")
        assert L.shim_mv_set_x(h, x.ctypes.data) == 0
        assert L.shim_mv_matvec_queue(h, 5) == 0
        xo = np.zeros(n); y = np.zeros(n)
        assert L.shim_mv_get_xy(h, xo.ctypes.data, y.ctypes.data) == 0
        H.assert_bits_equal(y, oracle.spmv("csr", ptr, idx, val, x), "queued products")
        H.assert_bits_equal(xo, x, "x untouched")
    finally:
        L.shim_mv_close(h)


def _conv_cases():
    yield "poisson3d_7pt_sorted", H.poisson3d_7pt(17, 13, 11, sort=True)
    yield "poisson3d_7pt_unsorted", H.poisson3d_7pt(12, 9, 10)
    yield "poisson3d_27pt", H.poisson3d_27pt(9, 8, 7)
    yield "random_ragged", H.random_csr(2500, 7, 11, values="wide")
    yield "random_banded_sorted", H.random_csr(9001, 5, 12, band=40, sorted_rows=True)
    yield "random_empty_rows", H.random_csr(4200, 4, 13, empty_rows=True, diag_dominant=False)
    yield "single_row", (np.array([0, 1], np.int32), np.array([0], np.int32), np.array([3.5]))
    yield "poisson1d_5000", H.poisson1d(5000)


@pytest.mark.parametrize("fmt,blk", [("ell", (0, 0)), ("dia", (0, 0)), ("jad", (0, 0)), ("bsr", (2, 2)), ("bsr", (3, 2)),
                                     ("bsr", (1, 4)), ("bsr", (4, 4))])
def test_device_conversion_same_arrays_as_host(b200, oracle, monkeypatch, fmt, blk):
    """LIS_B200_CONVERT=device (kernels/convert.cu): lis_matrix_convert builds ELL / DIA / JAD / BSR in
    HBM; every public array must equal the host builder's (which is pinned to the reference's
    layouts in test_oracle_vs_reference.py), and the product on the converted matrix -- served by
    the mirror the conversion left on the device -- must carry the oracle's bits"""
    for name, (ptr, idx, val) in _conv_cases():
        if fmt == "dia" and name == "random_ragged":
            continue
        monkeypatch.setenv("LIS_B200_CONVERT", "host")
        want = b200.convert(fmt, ptr, idx, val, bnr=blk[0], bnc=blk[1])
        monkeypatch.setenv("LIS_B200_CONVERT", "device")
        got = b200.convert(fmt, ptr, idx, val, bnr=blk[0], bnc=blk[1])
        assert set(got) == set(want)
        for key in want:
            if isinstance(want[key], np.ndarray):
                assert got[key].dtype == want[key].dtype and got[key].shape == want[key].shape, (fmt, name, key)
                if want[key].dtype == np.float64:
                    H.assert_bits_equal(got[key], want[key], f"{fmt}/{name}/{key}")
                else:
                    assert np.array_equal(got[key], want[key]), (fmt, name, key)
            else:
                assert got[key] == want[key], (fmt, name, key, got[key], want[key])
        x = H.rand_vec(len(ptr) - 1, 3, "wide")
        y, _ = b200.spmv(fmt, ptr, idx, val, x, bnr=blk[0], bnc=blk[1])
        H.assert_bits_equal(y, oracle.spmv(fmt, ptr, idx, val, x, bnr=blk[0] or 2, bnc=blk[1] or 2), f"spmv after device {fmt}/{name}")


def test_device_conversion_falls_back_to_host_builder(b200, monkeypatch):
    """rows longer than 255 entries (JAD) and block rows with more than 64 blocks (BSR) are outside
    what the conversion kernels cover: the host builder takes over, same arrays"""
    rng = np.random.default_rng(8)
    n = 600
    lens = rng.integers(1, 6, n); lens[17] = 300; lens[400] = 256
    ptr = np.zeros(n + 1, np.int32); ptr[1:] = np.cumsum(lens)
    idx = np.concatenate([rng.choice(n, l, replace=False) for l in lens]).astype(np.int32)
    val = rng.standard_normal(ptr[-1])
    for fmt, blk in (("jad", (0, 0)), ("bsr", (2, 2))):
        monkeypatch.setenv("LIS_B200_CONVERT", "host")
        want = b200.convert(fmt, ptr, idx, val, bnr=blk[0], bnc=blk[1])
        monkeypatch.setenv("LIS_B200_CONVERT", "device")
        got = b200.convert(fmt, ptr, idx, val, bnr=blk[0], bnc=blk[1])
        for key in want:
            if isinstance(want[key], np.ndarray):
                assert np.array_equal(got[key], want[key]), (fmt, key)


@pytest.mark.parametrize("opts", ["-i gmres -restart 30 -p jacobi", "-i gmres -restart 7 -p none", "-i fgmres -restart 20 -p ssor"])
def test_gram_schmidt_fused_chain_same_bits(b200, oracle, monkeypatch, opts):
    """GMRES / FGMRES orthogonalisation: axpy fused with the following dot / norm (mgs_step_kernel),
    axpy and dot as separate launches chained on the device, and the reference's call-for-call
    sequence with a host wait per dot must give the same residual history and solution, bit for bit"""
    for name, (ptr, idx, val) in (("p7", H.poisson3d_7pt(12, 11, 10)), ("unsym", H.random_csr(1501, 6, 3)), ("odd", H.poisson1d(333))):
        b = oracle.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
        runs = {}
        for mode, env in (("fused", {}), ("chain", {"LIS_B200_MGS": "chain"}), ("waits", {"LIS_B200_FUSE": "0"})):
            for key in ("LIS_B200_MGS", "LIS_B200_FUSE"):
                monkeypatch.delenv(key, raising=False)
            for key, v in env.items():
                monkeypatch.setenv(key, v)
            runs[mode] = b200.solve(ptr, idx, val, b, opts + " -maxiter 70")
        for mode in ("chain", "waits"):
            assert runs[mode]["iter"] == runs["fused"]["iter"] and runs[mode]["status"] == runs["fused"]["status"], (name, opts, mode)
            H.assert_bits_equal(runs[mode]["rhistory"], runs["fused"]["rhistory"], f"{name} {opts} rhistory fused vs {mode}")
            H.assert_bits_equal(runs[mode]["x"], runs["fused"]["x"], f"{name} {opts} x fused vs {mode}")


@pytest.mark.parametrize("opts", ["-i bicgstab -p jacobi", "-i bicgstab -p ssor", "-i bicgstab -p none -maxiter 60"])
def test_bicgstab_fused_updates_same_bits(b200, oracle, monkeypatch, opts):
    """BiCGSTAB with its vector updates fused (p update; s = r - alpha v with ||s||; x, r updates with
    ||r||; <t,s> with <t,t>), with one launch per reference call for the updates, and with every
    fusion off: same iteration count, residual history and solution, bit for bit"""
    for name, (ptr, idx, val) in (("p7", H.poisson3d_7pt(12, 11, 10)), ("unsym", H.random_csr(1501, 6, 3)), ("odd", H.poisson1d(333))):
        b = oracle.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
        runs = {}
        for mode, env in (("fused", {}), ("calls", {"LIS_B200_BICGSTAB": "calls"}), ("off", {"LIS_B200_FUSE": "0"})):
            for key in ("LIS_B200_BICGSTAB", "LIS_B200_FUSE"):
                monkeypatch.delenv(key, raising=False)
            for key, v in env.items():
                monkeypatch.setenv(key, v)
            runs[mode] = b200.solve(ptr, idx, val, b, opts)
        for mode in ("calls", "off"):
            assert runs[mode]["iter"] == runs["fused"]["iter"] and runs[mode]["status"] == runs["fused"]["status"], (name, opts, mode)
            H.assert_bits_equal(runs[mode]["rhistory"], runs["fused"]["rhistory"], f"{name} {opts} rhistory fused vs {mode}")
            H.assert_bits_equal(runs[mode]["x"], runs["fused"]["x"], f"{name} {opts} x fused vs {mode}")


@pytest.mark.parametrize("fmt", ["msr", "coo", "bsc", "vbr", "dns"])
def test_other_formats_spmv_bits(b200, ref_serial, fmt):
    """MSR / COO / BSC / VBR / DNS through their row-ordered device mirrors (host/lis_formats_ext.c): the CSR
    kernels add the products in the order of the reference's serial lis_matvec_<fmt> -- same bits; a solve
    with -storage <fmt> ends on the reference's iteration count"""
    mats = [("p7", H.poisson3d_7pt(9, 8, 7)), ("unsym", H.random_csr(600, 6, 41, band=30)), ("p1d", H.poisson1d(333))]
    if fmt != "dns":
        mats.append(("p7_big", H.poisson3d_7pt(40, 30, 20)))               # row-block (TMA) kernel territory
    for name, (ptr, idx, val) in mats:
        n = len(ptr) - 1
        for seed, kind in ((3, "wide"), (5, "uniform")):
            x = H.rand_vec(n, seed, kind)
            y, _ = b200.spmv(fmt, ptr, idx, val, x, bnr=2, bnc=2)
            yr, _ = ref_serial.spmv(fmt, ptr, idx, val, x, bnr=2, bnc=2)
            H.assert_bits_equal(y, yr, f"{fmt}/{name}")
    ptr, idx, val = H.poisson3d_7pt(9, 8, 7)
    b, _ = ref_serial.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
    opts = f"-i cg -p jacobi -storage {fmt} -storage_block 2"
    g, r = b200.solve(ptr, idx, val, b, opts), ref_serial.solve(ptr, idx, val, b, opts)
    assert g["err"] == r["err"] == 0 and g["status"] == r["status"] == 0 and g["iter"] == r["iter"], (fmt, g["iter"], r["iter"])
    assert np.abs(g["x"] - 1.0).max() < 1e-8


@pytest.mark.parametrize("fmt", ["csc", "msr", "dia", "ell", "jad", "bsr", "bsc", "vbr", "coo", "dns"])
def test_matvech_and_bicg_in_every_format(b200, ref_serial, fmt):
    """lis_matvech through the transposed mirror of each storage format: the bits of the reference's serial scatter
    loop (JAD within rounding: equal-length rows sit in a different order in the reference's layout); BiCG -- the
    reference's default solver -- with -storage <fmt> ends on its iteration count"""
    import ctypes as C
    i32p, f64p = np.ctypeslib.ndpointer(np.int32), np.ctypeslib.ndpointer(np.float64)
    mats = [("p7", H.poisson3d_7pt(9, 8, 7))] + ([("unsym", H.random_csr(600, 6, 41, band=30))] if fmt != "dia" else [])
    if fmt not in ("dns", "dia"):
        mats.append(("p7_big", H.poisson3d_7pt(40, 30, 20)))
    for name, (ptr, idx, val) in mats:
        n = len(ptr) - 1
        x = H.rand_vec(n, 12, "wide")
        ys = []
        for shim in (b200, ref_serial):
            shim.lib.shim_matvech.argtypes = [C.c_int, C.c_int, i32p, i32p, f64p, C.c_int, f64p, f64p]
            y = np.zeros(n)
            assert shim.lib.shim_matvech(lis_b200.FMT[fmt], n, ptr, idx, val, 0, x, y) == 0, (fmt, name)
            ys.append(y)
        if fmt == "jad":
            assert np.abs(ys[0] - ys[1]).max() <= 1e-13 * np.abs(ys[1]).max()
        else:
            H.assert_bits_equal(ys[0], ys[1], f"matvech {fmt}/{name}")
    ptr, idx, val = H.poisson3d_7pt(9, 8, 7)
    b, _ = ref_serial.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
    opts = f"-i bicg -p jacobi -storage {fmt} -storage_block 2"
    g, r = b200.solve(ptr, idx, val, b, opts), ref_serial.solve(ptr, idx, val, b, opts)
    assert g["err"] == r["err"] == 0 and g["status"] == r["status"] == 0 and abs(g["iter"] - r["iter"]) <= 1, (fmt, g["iter"], r["iter"])
    assert np.abs(g["x"] - 1.0).max() < 1e-8


@pytest.mark.parametrize("opts", ["-i bicgstab -p hybrid -hybrid_i gs -hybrid_maxiter 3", "-i fgmres -p hybrid -hybrid_i bicgstab -hybrid_maxiter 4 -hybrid_p jacobi",
                                  "-i gmres -p hybrid -hybrid_i sor -hybrid_maxiter 3 -hybrid_tol 1e-30"])
def test_hybrid_preconditioner(b200, ref_serial, opts):
    """-p hybrid on the device kernels: converges like the serial reference (iteration count within its rounding spread).
    (GMRES needs a FIXED preconditioner: a fixed number of stationary inner steps.  With an inner Krylov solver stopped
    by -hybrid_tol it returns a wrong x in the reference as well -- the mock-device test reproduces even that bit for bit.)"""
    ptr, idx, val = H.poisson3d_7pt(10, 9, 8)
    b, _ = ref_serial.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
    g, r = b200.solve(ptr, idx, val, b, opts), ref_serial.solve(ptr, idx, val, b, opts)
    assert g["err"] == r["err"] == 0 and g["status"] == r["status"] == 0 and abs(g["iter"] - r["iter"]) <= max(1, r["iter"] // 10), (opts, g["iter"], r["iter"])
    assert np.abs(g["x"] - 1.0).max() < 1e-8


@pytest.mark.parametrize("opts", ["-i bicgstab -p ilut", "-i bicg -p ilut -iluc_drop 0.001", "-i gmres -p ilut -iluc_rate 0.5"])
def test_ilut_preconditioner(b200, ref_serial, opts):
    """-p ilut: host factorization (bit-equal to the reference's, tests/test_hostcheck.py), the two triangular solves on the
    one-launch sweep kernel -- converges like the serial reference"""
    for ptr, idx, val in (H.poisson3d_7pt(10, 9, 8), H.random_csr(1500, 7, 404, band=50)):
        b, _ = ref_serial.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
        g, r = b200.solve(ptr, idx, val, b, opts), ref_serial.solve(ptr, idx, val, b, opts)
        assert g["err"] == r["err"] == 0 and g["status"] == r["status"] == 0 and abs(g["iter"] - r["iter"]) <= max(1, r["iter"] // 10), (opts, g["iter"], r["iter"])
        assert np.abs(g["x"] - 1.0).max() < 1e-8


@pytest.mark.parametrize("opts", ["-i bicgstab -p is", "-i gmres -p is -is_alpha 0.5", "-i bicg -p is -is_m 1"])
def test_is_preconditioner(b200, ref_serial, opts):
    """-p is on the device kernels (one CSR product with the truncated upper part + axpyz per apply): converges like the
    serial reference"""
    for ptr, idx, val in (H.poisson3d_7pt(10, 9, 8), H.random_csr(1500, 7, 404, band=50)):
        b, _ = ref_serial.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
        g, r = b200.solve(ptr, idx, val, b, opts), ref_serial.solve(ptr, idx, val, b, opts)
        assert g["err"] == r["err"] == 0 and g["status"] == r["status"] == 0 and abs(g["iter"] - r["iter"]) <= max(1, r["iter"] // 10), (opts, g["iter"], r["iter"])
        assert np.abs(g["x"] - 1.0).max() < 1e-8


@pytest.mark.parametrize("fmt", ["ell", "dia", "msr", "jad"])
def test_ssor_and_stationary_sweeps_in_scalar_formats(b200, ref_serial, fmt):
    """SSOR / Gauss-Seidel / SOR with -storage <fmt>: sweeps on a private CSR copy, products in the format"""
    ptr, idx, val = H.poisson3d_7pt(9, 8, 7)
    b, _ = ref_serial.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
    for o in ("-i cg -p ssor", "-i bicg -p ssor", "-i bicgstab -p ssor -adds true", "-i sor -omega 1.3"):
        opts = f"{o} -storage {fmt} -maxiter 3000"
        g, r = b200.solve(ptr, idx, val, b, opts), ref_serial.solve(ptr, idx, val, b, opts)
        assert g["err"] == r["err"] == 0 and g["status"] == r["status"] == 0 and abs(g["iter"] - r["iter"]) <= 1, (opts, g["err"], g["iter"], r["iter"])
        assert np.abs(g["x"] - 1.0).max() < 1e-8


@pytest.mark.parametrize("opts", ["-i cg -p jacobi", "-i cg -p jacobi -maxiter 7", "-i cg -p jacobi -initx_zeros false"])
def test_cg_carried_jacobi_step_same_bits(b200, oracle, monkeypatch, opts):
    """CG + Jacobi with the update that ends an iteration also forming z = M^-1 r and <r,z> of the next one
    (one launch, one host wait less per iteration) and with the two as separate launches (LIS_B200_CG=split):
    same iteration count, residual history and solution, bit for bit.  (LIS_B200_FUSE=0 is not in this list:
    it also replaces the fused SpMV+dot, whose <p,q> is summed per row block -- envelope parity, not bits.)"""
    for name, (ptr, idx, val) in (("p7", H.poisson3d_7pt(12, 11, 10)), ("odd", H.poisson1d(333)), ("big", H.poisson3d_7pt(40, 30, 20))):
        b = oracle.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
        runs = {}
        for mode, env in (("carried", {}), ("split", {"LIS_B200_CG": "split"})):
            for key in ("LIS_B200_CG", "LIS_B200_FUSE"):
                monkeypatch.delenv(key, raising=False)
            for key, v in env.items():
                monkeypatch.setenv(key, v)
            runs[mode] = b200.solve(ptr, idx, val, b, opts)
        assert runs["carried"]["iter"] > 5
        for mode in ("split",):
            assert runs[mode]["iter"] == runs["carried"]["iter"] and runs[mode]["status"] == runs["carried"]["status"], (name, opts, mode)
            H.assert_bits_equal(runs[mode]["rhistory"], runs["carried"]["rhistory"], f"{name} {opts} rhistory carried vs {mode}")
            H.assert_bits_equal(runs[mode]["x"], runs["carried"]["x"], f"{name} {opts} x carried vs {mode}")


@pytest.mark.parametrize("opts", ["-i cg -scale jacobi", "-i bicgstab -scale symm_diag -p jacobi", "-i sor -p jacobi -omega 1.5 -maxiter 400",
                                  "-i cg -p ssor -adds true"])
def test_scaling_and_additive_schwarz_follow_the_reference(b200, ref_serial, opts):
    """-scale (A and b scaled in place before the loop), the stationary solvers with a preconditioner
    and the additive Schwarz wrapper against the compiled serial reference: same status, iteration
    count within one step (the reductions differ), same solution"""
    for name, (ptr, idx, val) in (("p7", H.poisson3d_7pt(10, 9, 8)), ("p1d", H.poisson1d(150))):
        n = len(ptr) - 1
        b = ref_serial.spmv("csr", ptr, idx, val, np.ones(n))[0]
        g = b200.solve(ptr, idx, val, b, opts)
        r = ref_serial.solve(ptr, idx, val, b, opts)
        assert g["err"] == r["err"] == 0 and g["status"] == r["status"], (name, opts, g["err"], g["status"], r["status"])
        assert abs(g["iter"] - r["iter"]) <= max(1, r["iter"] // 100), (name, opts, g["iter"], r["iter"])
        assert np.abs(g["x"] - r["x"]).max() < 1e-8 * max(1.0, np.abs(r["x"]).max()), (name, opts)


@pytest.mark.parametrize("fmt", ["ell", "dia", "jad", "bsr", "csc"])
def test_scaling_in_other_formats_follows_the_reference(b200, ref_serial, fmt):
    """lis_matrix_scale on a matrix that already is in another storage format (the reference's per-format loops), and the
    private split copy of -p ssor -storage <fmt> taking the same factors: status, iteration count within a step and the
    solution of the compiled serial reference (33 iterations where the unscaled solve takes 17 -- WD stays the one made
    from the unscaled diagonal, as there)"""
    ptr, idx, val = H.poisson3d_7pt(7, 6, 5)
    n = len(ptr) - 1
    b = ref_serial.spmv("csr", ptr, idx, val, np.ones(n))[0]
    cases = [("-i bicgstab -p none -scale jacobi", {"fmt": fmt}), ("-i cg -p none -scale symm_diag", {"fmt": fmt}),
             ("-i gmres -p jacobi -scale jacobi", {"fmt": fmt})]
    if fmt != "bsr":
        cases += [(f"-i cg -p ssor -scale jacobi -storage {fmt}", {}), (f"-i gs -p ssor -storage {fmt}", {})]
    for opts, kw in cases:
        g = b200.solve(ptr, idx, val, b, opts, **kw)
        r = ref_serial.solve(ptr, idx, val, b, opts, **kw)
        assert g["err"] == r["err"] == 0 and g["status"] == r["status"], (fmt, opts, g["err"], g["status"], r["status"])
        assert abs(g["iter"] - r["iter"]) <= max(1, r["iter"] // 100), (fmt, opts, g["iter"], r["iter"])
        assert np.abs(g["x"] - r["x"]).max() < 1e-8 * max(1.0, np.abs(r["x"]).max()), (fmt, opts)
