"""The product's CUDA kernel SOURCES executed on the host emulator (tests/cudaemu: every CUDA
thread a fiber, inline PTX restated, device buffers ending at guard pages) underneath the
product's unchanged host code, checked against the oracle with the very assertions of the GPU
parity tests.

This is not a parity claim for the GPU path and not a CPU fallback (the product library has
none): it is what lets kernel logic -- indexing, tiling, TMA/mbarrier pipeline phases, alignment
of 128-bit loads and bulk copies, out-of-bounds reads, the last-CTA fold, the dependency polling
of the one-launch sweeps -- be checked in the `-m "not gpu"` suite."""
import os
import subprocess

import pytest

import harness as H
import lis_b200
import test_gpu_parity as G
import test_z1_gpu_parity2 as G2

EMU_DIR = os.path.join(H.ROOT, "tests", "cudaemu")


@pytest.fixture(scope="module")
def b200(built):
    r = subprocess.run(["make", "-C", EMU_DIR, "-j8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return lis_b200.Shim(os.path.join(EMU_DIR, "_build", "liblis_emu_shim.so"))


test_spmv_bit_exact = G.test_spmv_bit_exact
test_spmv_csr_both_kernels = G.test_spmv_csr_both_kernels
test_spmv_bsr_block_shapes = G.test_spmv_bsr_block_shapes
test_spmv_csr_split_order = G.test_spmv_csr_split_order
test_spmv_long_rows = G.test_spmv_long_rows
test_scaling_and_additive_schwarz_follow_the_reference = G2.test_scaling_and_additive_schwarz_follow_the_reference
test_scaling_in_other_formats_follows_the_reference = G2.test_scaling_in_other_formats_follows_the_reference
test_bicgstab_fused_updates_same_bits = G2.test_bicgstab_fused_updates_same_bits
test_cg_carried_jacobi_step_same_bits = G2.test_cg_carried_jacobi_step_same_bits
test_other_formats_spmv_bits = G2.test_other_formats_spmv_bits
test_matvech_and_bicg_in_every_format = G2.test_matvech_and_bicg_in_every_format
test_ssor_and_stationary_sweeps_in_scalar_formats = G2.test_ssor_and_stationary_sweeps_in_scalar_formats
test_hybrid_preconditioner = G2.test_hybrid_preconditioner
test_ilut_preconditioner = G2.test_ilut_preconditioner
test_is_preconditioner = G2.test_is_preconditioner
test_gram_schmidt_fused_chain_same_bits = G2.test_gram_schmidt_fused_chain_same_bits
test_queued_products_same_bits_as_synchronous_calls = G2.test_queued_products_same_bits_as_synchronous_calls
test_device_conversion_same_arrays_as_host = G2.test_device_conversion_same_arrays_as_host
test_device_conversion_falls_back_to_host_builder = G2.test_device_conversion_falls_back_to_host_builder
test_blas1_elementwise_bit_exact = G.test_blas1_elementwise_bit_exact
test_blas1_length_mismatch_is_ill_arg = G.test_blas1_length_mismatch_is_ill_arg
test_reductions_bounded = G.test_reductions_bounded
test_reductions_deterministic = G.test_reductions_deterministic
test_empty_vector = G.test_empty_vector
test_get_diagonal = G.test_get_diagonal
test_psolve_ssor_bit_exact = G.test_psolve_ssor_bit_exact
test_psolve_ssor_level_launch_path = G.test_psolve_ssor_level_launch_path
test_psolve_sweep_look_ahead_same_bits = G.test_psolve_sweep_look_ahead_same_bits
test_psolve_ssor_repeated_sweeps = G.test_psolve_ssor_repeated_sweeps
test_psolve_jacobi_bit_exact = G.test_psolve_jacobi_bit_exact
test_bicg_default_solver = G.test_bicg_default_solver
test_solvers_match_oracle_poisson = G.test_solvers_match_oracle_poisson
test_solvers_match_oracle_unsymmetric = G.test_solvers_match_oracle_unsymmetric
test_ssor_block_count_changes_iterations_like_openmp = G.test_ssor_block_count_changes_iterations_like_openmp
test_solver_formats_give_same_iterations = G.test_solver_formats_give_same_iterations
test_solver_status_codes = G.test_solver_status_codes
test_golden_vectors = G.test_golden_vectors

import test_solvers_ext_gpu as E  # noqa: E402

test_further_solver_within_reference_envelope = E.test_further_solver_within_reference_envelope
test_ilu_and_transposed_sweeps_match_reference_bits = E.test_ilu_and_transposed_sweeps_match_reference_bits


# ---- the same SpMV checks with 2 emulated SMs: persistent grids (TMA row-block CSR kernel, BLAS-1
# grid-stride loops, reductions) then give every CTA several row blocks / many elements
@pytest.fixture()
def two_sms(monkeypatch):
    monkeypatch.setenv("LISB_EMU_SMS", "2")


def test_persistent_grids_two_sms(b200, oracle, two_sms, monkeypatch):
    G.test_spmv_csr_both_kernels(b200, oracle, monkeypatch)
    G.test_spmv_bit_exact(b200, oracle, "csr")
    G.test_reductions_bounded(b200, oracle, 100003)
    G.test_blas1_elementwise_bit_exact(b200, oracle, "axpy", 100003)
    for opts in ("-i cg -p jacobi", "-i bicgstab -p ssor"):
        import numpy as np
        ptr, idx, val = H.poisson3d_7pt(9, 8, 7)
        b = oracle.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
        r = b200.solve(ptr, idx, val, b, opts)
        assert r["status"] == 0 and abs(r["x"] - 1.0).max() < 1e-8


def test_cg_launch_pattern(b200, oracle):
    """CG + Jacobi launches per iteration: xpay, fused SpMV+dot, update carrying the next Jacobi step --
    jacobi_dot_kernel only once, before the first iteration"""
    import ctypes as C
    import numpy as np
    lib = C.CDLL(os.path.join(EMU_DIR, "_build", "liblis_emu.so"))
    lib.emu_launch_count.restype = C.c_long; lib.emu_launch_count.argtypes = [C.c_char_p]
    names = (b"cg_update_jacobi_kernel", b"jacobi_dot_kernel", b"cg_update_kernel", b"csr_tma_kernel", b"")
    ptr, idx, val = H.poisson3d_7pt(12, 11, 10)
    b = oracle.spmv("csr", ptr, idx, val, np.ones(len(ptr) - 1))
    before = {k: lib.emu_launch_count(k) for k in names}
    r = b200.solve(ptr, idx, val, b, "-i cg -p jacobi")
    d = {k: lib.emu_launch_count(k) - before[k] for k in names}
    assert r["status"] == 0
    assert d[b"cg_update_jacobi_kernel"] == r["iter"] and d[b"jacobi_dot_kernel"] == 1 and d[b"cg_update_kernel"] == 0, (d, r["iter"])


# ---- the emulator must notice what it exists to notice
_NEG = r"""
import ctypes as C, numpy as np, sys
lib = C.CDLL(sys.argv[1])
what = sys.argv[2]
vp = C.c_void_p
def dev(a):
    p = vp()
    assert lib.cudaMalloc(C.byref(p), C.c_size_t(a.nbytes)) == 0
    assert lib.cudaMemcpy(p, vp(a.ctypes.data), C.c_size_t(a.nbytes), 1) == 0
    return p
n = 300
ptr = np.arange(0, 3 * n + 1, 3, dtype=np.int32)
idx = (np.arange(3 * n) % n).astype(np.int32)
val = np.ones(3 * n)
x = np.ones(n); y = np.zeros(n)
d_ptr = dev(np.concatenate([ptr, np.zeros(4, np.int32)]))
d_idx = dev(np.concatenate([idx, np.zeros(8, np.int32)])); d_val = dev(np.concatenate([val, np.zeros(8)]))
d_x = dev(x); d_y = dev(y[:n - 8] if what == "short_y" else y)
lib.lisb200_spmv_csr_tma.argtypes = [C.c_int] * 4 + [vp] * 6
if what == "misaligned_ptr":
    d_ptr = vp(d_ptr.value + 4)
rc = lib.lisb200_spmv_csr_tma(n - (1 if what == "misaligned_ptr" else 0), 256, 1024, 2, d_ptr, d_idx, d_val, d_x, d_y, None)
print("rc", rc)
if what == "host_deref":
    print((C.c_double * n).from_address(d_y.value)[0])      # plain device memory read from the host
if what == "ok":
    back = np.zeros(n); assert lib.cudaMemcpy(vp(back.ctypes.data), d_y, C.c_size_t(back.nbytes), 2) == 0
    assert (back == 3.0).all()
"""


@pytest.mark.parametrize("what,expect", [("ok", 0), ("misaligned_ptr", "abort"), ("short_y", "segv"), ("host_deref", "segv")])
def test_emulator_catches_misalignment_and_overrun(b200, what, expect):
    """a TMA bulk copy from a 4-byte-aligned row-pointer slice aborts; an output vector
    that is 8 entries short makes the kernel write into the guard page behind it; the host
    reading a cudaMalloc block directly faults (its pages are open only while a kernel or a
    runtime copy runs)"""
    import signal
    import sys
    r = subprocess.run([sys.executable, "-c", _NEG, os.path.join(EMU_DIR, "_build", "liblis_emu.so"), what],
                       capture_output=True, text=True)
    if expect == 0:
        assert r.returncode == 0 and "rc 0" in r.stdout, r.stderr
    elif expect == "abort":
        assert r.returncode == -signal.SIGABRT and "not 16-byte aligned" in r.stderr, (r.returncode, r.stderr)
    else:
        assert r.returncode == -signal.SIGSEGV, (r.returncode, r.stderr)


# ---- the overlapped host-buffer product (lis_b200_matvec_host): three streams chained by events.
# The emulator's secondary streams are lazy (copies happen as late as the events allow), so a row
# chunk that starts before its x entries have landed reads the previous contents of x.
test_matvec_host_pipelined = G2.test_matvec_host_pipelined
test_matvec_host_pipelined_default_chunks = G2.test_matvec_host_pipelined_default_chunks


# ---- the reference's own drivers (test/*.c, unchanged) linked against the emulator build: their
# transcripts against the same sources linked with the compiled reference
import test_reference_drivers as D  # noqa: E402
import test_z2_reference_drivers2 as D2  # noqa: E402


@pytest.fixture()
def emu_drivers(b200, monkeypatch):
    monkeypatch.setattr(D, "OURS", os.path.join(EMU_DIR, "_build", "drivers"))


@pytest.mark.parametrize("driver,args,analytic", [("spmvtest1", (20000, 3), "1.414214e+00"), ("spmvtest3", (24, 24, 24, 2), None),
                                                  ("spmvtest3b", (9, 8, 7, 2), None)])
def test_spmvtest_drivers_emulated(emu_drivers, driver, args, analytic):
    D.test_spmvtest_drivers(driver, args, analytic)


@pytest.mark.parametrize("driver,args", [("spmvtest1", (3000, 2)), ("spmvtest3", (9, 8, 7, 2))])
def test_spmvtest_drivers_walk_all_formats_emulated(emu_drivers, driver, args):
    D.test_spmvtest_drivers_walk_all_formats(driver, args)


@pytest.mark.parametrize("driver", ["spmvtest4", "spmvtest5"])
def test_spmvtest_file_drivers_emulated(emu_drivers, tmp_path, driver):
    D2.test_spmvtest_file_drivers(tmp_path, driver)


@pytest.mark.parametrize("opts", ["-i cg -p jacobi", "-i bicgstab -p ssor", "-i gmres -restart 30 -p jacobi", "-i cg -p jacobi -storage ell"])
def test_test3_driver_emulated(emu_drivers, tmp_path, opts):
    D.test_test3_driver(tmp_path, opts)


def test_test1_driver_emulated(emu_drivers, tmp_path):
    D.test_test1_driver_matrix_market(tmp_path)


@pytest.mark.parametrize("opts", ["", "-e ii -i cg -p jacobi", "-e rqi"])
def test_etest1_driver_emulated(emu_drivers, tmp_path, opts):
    D2.test_etest1_driver(tmp_path, opts)


def test_etest5_driver_emulated(emu_drivers, tmp_path):
    D2.test_etest5_driver_lanczos(tmp_path)


def test_test3b_driver_hpcg_kernel_emulated(emu_drivers, tmp_path):
    D.test_test3b_driver_hpcg_kernel(tmp_path)


# ---- halo exchange inside the SpMV kernel (csr_tma_kernel<.., kHalo>): the kernel logic on the emulator
class _P2PTable(__import__("ctypes").Structure):
    import ctypes as _C
    _MAX = 16
    _fields_ = [("n_nbr", _C.c_int), ("n_export", _C.c_int), ("export_index", _C.c_void_p), ("exp_start", _C.c_int * (_MAX + 1)),
                ("nbr_rank", _C.c_int * _MAX), ("peer_inbox", _C.c_void_p * _MAX), ("peer_stride", _C.c_longlong * _MAX),
                ("peer_flag", _C.c_void_p * _MAX), ("inbox", _C.c_void_p), ("inbox_stride", _C.c_longlong), ("my_flag", _C.c_void_p),
                ("push_count", _C.c_void_p), ("error", _C.c_void_p)]


@pytest.mark.parametrize("with_dot", [0, 1])
@pytest.mark.parametrize("sms", ["148", "2"])
def test_in_kernel_halo_exchange_two_slabs(oracle, with_dot, sms):
    """Two row slabs of a 7-point cube, each multiplied by the kHalo kernel with the other slab's inbox as its push target
    (one address space stands in for CUDA IPC).  The emulator runs one kernel at a time, so the neighbour's arrival flag is
    raised by hand before a launch; what is checked is everything else the kernel does: the values and positions it
    pushes and the flag it raises, interior blocks first, halo columns read from the inbox of the epoch's parity, y with
    the bits of the one-process product, and the fused <x,y>.  Two epochs: both buffers."""
    import ctypes as C
    import numpy as np
    env_old = os.environ.get("LISB_EMU_SMS")
    os.environ["LISB_EMU_SMS"] = sms
    try:
        r = subprocess.run(["make", "-C", EMU_DIR, "-j8"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        lib = C.CDLL(os.path.join(EMU_DIR, "_build", "liblis_emu.so"))
        vp = C.c_void_p
        lib.lisb200_spmv_csr_tma_p2p.argtypes = [C.c_int] * 4 + [vp] * 5 + [C.c_int] + [vp] * 4 + [C.c_ulonglong, C.c_int, C.c_int, vp]
        lib.lisb200_reduce_slots.restype = C.c_int
        L, M, N = 24, 16, 16                       # 2 slabs of 12 planes, 256 rows per plane: row blocks align with planes
        ptr, idx, val = H.poisson3d_7pt(L, M, N, sort=True)
        nloc = L // 2 * M * N; plane = M * N
        keep = []                                  # buffers must outlive the launches
        slabs = []
        for rk in (0, 1):
            r0 = rk * nloc
            p = ptr[r0:r0 + nloc + 1].astype(np.int64)
            cols = idx[p[0]:p[-1]].astype(np.int64); vals = val[p[0]:p[-1]].copy()
            halo = np.unique(cols[(cols < r0) | (cols >= r0 + nloc)])
            loc = np.where((cols >= r0) & (cols < r0 + nloc), cols - r0, nloc + np.searchsorted(halo, cols))
            slabs.append(dict(ptr=(p - p[0]).astype(np.int32), idx=loc.astype(np.int32), val=vals, halo=halo, r0=r0))
        rng = np.random.default_rng(5)
        for epoch in (1, 2):
            xg = rng.uniform(-1, 1, L * M * N)
            yref = oracle.spmv("csr", ptr, idx, val, xg)
            stride = 256
            inbox = [np.zeros(2 * stride + 2 * 16 + 8) for _ in (0, 1)]         # doubles; flags live behind the two buffers
            flags = [ib[2 * stride:2 * stride + 32].view(np.uint64) for ib in inbox]
            for rk in (0, 1):
                s = slabs[rk]; other = 1 - rk
                x = xg[s["r0"]:s["r0"] + nloc].copy(); y = np.zeros(nloc)
                export = (slabs[other]["halo"] - s["r0"]).astype(np.int32)      # my rows the neighbour reads, its halo order
                par = epoch & 1
                # the neighbour's push, by hand: my inbox + its flag
                inbox[rk][par * stride:par * stride + len(s["halo"])] = xg[s["halo"]]
                flags[rk][par * 16 + other] = epoch
                tb = _P2PTable()
                tb.n_nbr = 1; tb.n_export = len(export); tb.export_index = export.ctypes.data
                tb.exp_start[0] = 0; tb.exp_start[1] = len(export); tb.nbr_rank[0] = other
                tb.peer_inbox[0] = inbox[other].ctypes.data; tb.peer_stride[0] = stride
                tb.peer_flag[0] = flags[other].ctypes.data + 8 * rk
                tb.inbox = inbox[rk].ctypes.data; tb.inbox_stride = stride; tb.my_flag = flags[rk].ctypes.data
                cnt = np.zeros(16, np.uint32); err = np.zeros(4, np.int32)
                tb.push_count = cnt.ctypes.data; tb.error = err.ctypes.data
                pp = np.concatenate([s["ptr"], np.zeros(4, np.int32)]); ii = np.concatenate([s["idx"], np.zeros(8, np.int32)])
                vv = np.concatenate([s["val"], np.zeros(8)])
                part = np.zeros(lib.lisb200_reduce_slots() + 8); counter = np.zeros(16, np.uint32); res = np.zeros(4)
                lo, hi = plane, nloc - plane                                    # all but the first and last plane
                keep += [x, y, export, pp, ii, vv, part, counter, res, cnt, err]
                rc = lib.lisb200_spmv_csr_tma_p2p(nloc, 256, 2048, 2, pp.ctypes.data, ii.ctypes.data, vv.ctypes.data, x.ctypes.data, y.ctypes.data,
                                                  with_dot, part.ctypes.data, counter.ctypes.data, res.ctypes.data, C.addressof(tb), epoch, lo, hi, None)
                assert rc == 0 and err[0] == 0
                H.assert_bits_equal(y, yref[s["r0"]:s["r0"] + nloc], f"slab {rk} epoch {epoch}")
                # what it pushed: the neighbour's halo values at the neighbour's halo positions, flag = epoch, counter reset
                np.testing.assert_array_equal(inbox[other][par * stride:par * stride + len(export)], xg[slabs[other]["halo"]])
                assert flags[other][par * 16 + rk] == epoch and cnt[0] == 0
                if with_dot:
                    exact = float(np.dot(x, y))
                    assert abs(res[0] - exact) <= 1e-12 * np.abs(x * y).sum()
    finally:
        if env_old is None:
            os.environ.pop("LISB_EMU_SMS", None)
        else:
            os.environ["LISB_EMU_SMS"] = env_old


def test_sweep_warp_per_row_for_long_rows(b200, oracle, monkeypatch):
    """LIS_B200_SWEEP_KERNEL=rows: the warp-per-row sweep kernel (sweep_rowwarp_kernel: all neighbours of a
    row polled at once, products added lane by lane in storage order): same bits as the oracle's sequential sweep, for the
    serial sweep and for block-SSOR; rows longer than one 64-entry chunk included (the long row of random_csr)"""
    import ctypes as C
    lib = C.CDLL(os.path.join(EMU_DIR, "_build", "liblis_emu.so"))
    lib.emu_launch_count.restype = C.c_long; lib.emu_launch_count.argtypes = [C.c_char_p]
    monkeypatch.setenv("LIS_B200_SWEEP_KERNEL", "rows")
    ptr, idx, val = H.random_csr(260, 30, 77)
    b = H.rand_vec(260, 5, "wide")
    for T in (1, 3):
        b200.set_threads(T)
        try:
            before = lib.emu_launch_count(b"sweep_rowwarp_kernel")
            x = b200.psolve(ptr, idx, val, b, "-p ssor -ssor_omega 1.2")
            assert lib.emu_launch_count(b"sweep_rowwarp_kernel") - before == 2, "expected the warp-per-row kernel for both sweeps"
            H.assert_bits_equal(x, oracle.psolve(ptr, idx, val, b, "ssor", omega=1.2, nthreads=T), f"warp-per-row SSOR, {T} block(s)")
        finally:
            b200.set_threads(1)


def test_in_kernel_halo_exchange_middle_slab_two_neighbours(oracle):
    """the middle one of three slabs: two neighbours, two export segments, two flags to wait for and two to raise"""
    import ctypes as C
    import numpy as np
    r = subprocess.run(["make", "-C", EMU_DIR, "-j8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lib = C.CDLL(os.path.join(EMU_DIR, "_build", "liblis_emu.so"))
    vp = C.c_void_p
    lib.lisb200_spmv_csr_tma_p2p.argtypes = [C.c_int] * 4 + [vp] * 5 + [C.c_int] + [vp] * 4 + [C.c_ulonglong, C.c_int, C.c_int, vp]
    L, M, N = 18, 16, 16                            # 3 slabs of 6 planes of 256 rows
    ptr, idx, val = H.poisson3d_7pt(L, M, N, sort=True)
    nloc = L // 3 * M * N; plane = M * N
    r0 = nloc                                       # the middle slab
    p = ptr[r0:r0 + nloc + 1].astype(np.int64)
    cols = idx[p[0]:p[-1]].astype(np.int64); vals = val[p[0]:p[-1]].copy()
    halo = np.unique(cols[(cols < r0) | (cols >= r0 + nloc)])          # lower neighbour's plane, then the upper one's
    loc = np.where((cols >= r0) & (cols < r0 + nloc), cols - r0, nloc + np.searchsorted(halo, cols)).astype(np.int32)
    lp = (p - p[0]).astype(np.int32)
    assert len(halo) == 2 * plane
    # what the neighbours read from me: rank 0 my first plane, rank 2 my last plane (their halo order = ascending)
    export = np.concatenate([np.arange(plane), np.arange(nloc - plane, nloc)]).astype(np.int32)
    rng = np.random.default_rng(9)
    stride = 2 * plane
    for epoch in (1, 2, 3):
        xg = rng.uniform(-1, 1, L * M * N)
        yref = oracle.spmv("csr", ptr, idx, val, xg)
        par = epoch & 1
        mine = np.zeros(2 * stride + 40); mflags = mine[2 * stride:2 * stride + 32].view(np.uint64)
        nb = {0: np.zeros(2 * 256 + 40), 2: np.zeros(2 * 256 + 40)}    # the neighbours' inboxes (stride 256)
        nflags = {k: v[512:512 + 32].view(np.uint64) for k, v in nb.items()}
        mine[par * stride:par * stride + 2 * plane] = xg[halo]          # both neighbours' pushes, by hand
        mflags[par * 16 + 0] = epoch; mflags[par * 16 + 2] = epoch
        tb = _P2PTable()
        tb.n_nbr = 2; tb.n_export = len(export); tb.export_index = export.ctypes.data
        tb.exp_start[0] = 0; tb.exp_start[1] = plane; tb.exp_start[2] = 2 * plane
        tb.nbr_rank[0] = 0; tb.nbr_rank[1] = 2
        for s, k in enumerate((0, 2)):
            tb.peer_inbox[s] = nb[k].ctypes.data; tb.peer_stride[s] = 256
            tb.peer_flag[s] = nflags[k].ctypes.data + 8 * 1            # my rank is 1
        tb.inbox = mine.ctypes.data; tb.inbox_stride = stride; tb.my_flag = mflags.ctypes.data
        cnt = np.zeros(16, np.uint32); err = np.zeros(4, np.int32)
        tb.push_count = cnt.ctypes.data; tb.error = err.ctypes.data
        x = xg[r0:r0 + nloc].copy(); y = np.zeros(nloc)
        pp = np.concatenate([lp, np.zeros(4, np.int32)]); ii = np.concatenate([loc, np.zeros(8, np.int32)]); vv = np.concatenate([vals, np.zeros(8)])
        rc = lib.lisb200_spmv_csr_tma_p2p(nloc, 256, 2048, 2, pp.ctypes.data, ii.ctypes.data, vv.ctypes.data, x.ctypes.data, y.ctypes.data,
                                          0, None, None, None, C.addressof(tb), epoch, plane, nloc - plane, None)
        assert rc == 0 and err[0] == 0
        H.assert_bits_equal(y, yref[r0:r0 + nloc], f"middle slab, epoch {epoch}")
        np.testing.assert_array_equal(nb[0][par * 256:par * 256 + plane], xg[r0:r0 + plane])
        np.testing.assert_array_equal(nb[2][par * 256:par * 256 + plane], xg[r0 + nloc - plane:r0 + nloc])
        assert nflags[0][par * 16 + 1] == epoch and nflags[2][par * 16 + 1] == epoch and cnt[0] == 0
