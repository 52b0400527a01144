"""One rank of the multi-process tests (launched by test_multi_rank.py with RANK / WORLD_SIZE /
MASTER_ADDR / MASTER_PORT set).  gloo carries the harness traffic (token broadcast, result
gathering); the library's own process group does the work being tested.

mode cpu : host-side logic only -- row partition, global->local numbering, halo lists,
           rank-ordered host allreduce.  No GPU needed.
mode gpu : one GPU per rank -- row-partitioned SpMV (halo exchange over NCCL) and solvers,
           compared by rank 0 with the single-process oracle.
mode hostcheck : the same flow on the mock-device build (tests/hostcheck), halo staged through
           host memory: the whole multi-rank host logic on a CPU-only machine."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402
import lis_b200  # noqa: E402


def overlap_case(rank, world, shim, lib):
    """rows that read no halo entry run on a second stream while the halo exchange is in flight
    (LIS_B200_OVERLAP=force: even when they are a minority), the others behind it: same bits as the
    one-process product; also through the overlapped host-buffer product"""
    L = shim.lib
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C"); f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    l, m, n = 3 * world + 1, 48, 40
    ptr, idx, val = H.poisson3d_7pt(l, m, n)
    gn = l * m * n
    q, r = divmod(gn, world)
    sizes = [q + 1 if k < r else q for k in range(world)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    is_, ie = int(starts[rank]), int(starts[rank + 1])
    lp = (ptr[is_:ie + 1] - ptr[is_]).astype(np.int32)
    li = np.ascontiguousarray(idx[ptr[is_]:ptr[ie]]); lv = np.ascontiguousarray(val[ptr[is_]:ptr[ie]])
    nl = ie - is_
    L.shim_mv_open_dist.argtypes = [C.c_int, C.c_int, i32p, i32p, f64p, C.c_int]
    L.shim_mv_set_x_local.argtypes = [C.c_int, f64p]; L.shim_mv_get_y_local.argtypes = [C.c_int, f64p]
    L.shim_mv_step_e2e_pipelined.argtypes = [C.c_int, f64p, f64p]
    o = H.Oracle()
    os.environ["LIS_B200_PIPE_CHUNKS"] = "5"
    for kernel in ("tma", "tile"):
        os.environ["LIS_B200_CSR_KERNEL"] = kernel
        h = L.shim_mv_open_dist(1, nl, lp, li, lv, 0)
        assert h >= 0
        for seed in (3, 4):
            x = H.rand_vec(gn, seed, "wide")
            want = o.spmv("csr", ptr, idx, val, x)[is_:ie]
            assert L.shim_mv_set_x_local(h, np.ascontiguousarray(x[is_:ie])) == 0
            assert L.shim_mv_matvec(h) == 0
            yl = np.zeros(nl)
            assert L.shim_mv_get_y_local(h, yl) == 0
            H.assert_bits_equal(yl, want, f"rank {rank} overlapped exchange ({kernel})")
            y2 = np.full(nl, np.nan)
            assert L.shim_mv_step_e2e_pipelined(h, np.ascontiguousarray(x[is_:ie]), y2) == 0
            H.assert_bits_equal(y2, want, f"rank {rank} host-buffer product ({kernel})")
        if hasattr(lib, "lis_b200_set_p2p"):
            # the halo exchange inside the kernel (real GPUs that can map each other; elsewhere the switch changes
            # nothing) against the NCCL / staged exchange: same bits, over several epochs (both inbox buffers)
            lib.lis_b200_p2p_products.restype = C.c_ulonglong
            for on in (0, 1, 0, 1):
                lib.lis_b200_set_p2p(on)
                before = lib.lis_b200_p2p_products()
                for seed in (11, 12, 13):
                    x = H.rand_vec(gn, seed, "wide")
                    want = o.spmv("csr", ptr, idx, val, x)[is_:ie]
                    assert L.shim_mv_set_x_local(h, np.ascontiguousarray(x[is_:ie])) == 0
                    assert L.shim_mv_matvec(h) == 0
                    yl = np.zeros(nl)
                    assert L.shim_mv_get_y_local(h, yl) == 0
                    H.assert_bits_equal(yl, want, f"rank {rank} in-kernel exchange={on} ({kernel})")
                used = lib.lis_b200_p2p_products() - before
                assert on or used == 0
                if kernel == "tile":
                    assert used == 0                     # only the TMA row-block kernel carries the exchange
            lib.lis_b200_set_p2p(1)
        L.shim_mv_close(h)
    try:
        lib.emu_launch_count.restype = C.c_long; lib.emu_launch_count.argtypes = [C.c_char_p]
        count = lambda: int(lib.emu_launch_count(b""))
        have_count = True
    except AttributeError:
        count = lambda: -1
        have_count = False                       # the product library keeps no launch log (emulator only)
    # CG on the same partition: q = A p with <p,q> fused, interior rows on the second stream during the
    # exchange, the dot assembled from the range shares.  Same iteration count as the oracle's
    # one-process CG and as the run without overlap; the overlapped run launches more kernels.
    os.environ["LIS_B200_CSR_KERNEL"] = "tma"
    L.shim_mv_solve_b.argtypes = [C.c_int, C.c_char_p, f64p, f64p, i32p, f64p, f64p, C.c_int]
    bvec = o.spmv("csr", ptr, idx, val, np.ones(gn))
    ref = o.solve("cg", ptr, idx, val, bvec, precon="jacobi")
    h = L.shim_mv_open_dist(1, nl, lp, li, lv, 0)
    assert h >= 0
    runs = {}
    for on in (1, 0):
        lib.lis_b200_set_overlap(on)
        xl = np.zeros(nl); oi = np.zeros(4, np.int32); od = np.zeros(4); rh = np.zeros(5000)
        c0 = count()
        rc = L.shim_mv_solve_b(h, b"-i cg -p jacobi", np.ascontiguousarray(bvec[is_:ie]), xl, oi, od, rh, 5000)
        assert rc == 0 and oi[1] == 0, (on, rc, oi)
        assert np.abs(xl - 1.0).max() < 1e-8, on
        assert int(oi[0]) == ref["iter"], (on, int(oi[0]), ref["iter"])
        n_it = int(oi[0])
        assert np.allclose(rh[:n_it + 1], ref["rhistory"][:n_it + 1], rtol=1e-6, atol=0), on
        runs[on] = (count() - c0, n_it)
    lib.lis_b200_set_overlap(1)
    if have_count:
        assert runs[1][0] >= runs[0][0] + runs[1][1], ("overlapped CG did not split its products", runs)
    L.shim_mv_close(h)
    launches = count()
    gathered = [None] * world
    dist.all_gather_object(gathered, {"rank": rank, "rows": nl, "launches": launches})
    dist.barrier()
    lib.lis_finalize()
    if rank == 0:
        print("MR_OK", gathered)


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tok = torch.zeros(1, dtype=torch.int64)
    if rank == 0:
        tok[0] = int.from_bytes(os.urandom(7), "little")
    dist.broadcast(tok, 0)
    # CPU-only builds of the host code: mock device (tests/hostcheck) or the kernel emulator (tests/cudaemu)
    hostcheck = os.environ.get("LIS_B200_HOSTCHECK_DIR")
    if hostcheck:
        name = os.environ.get("LIS_B200_HOSTCHECK_NAME", "hostcheck")
        shim = lis_b200.Shim(os.path.join(hostcheck, f"liblis_{name}_shim.so"))
        lib = C.CDLL(os.path.join(hostcheck, f"liblis_{name}.so"))
    else:
        lib = lis_b200.load_library()
        shim = lis_b200.load_shim()
    lib.lis_b200_comm_attach.argtypes = [C.c_int, C.c_int, C.c_ulonglong]
    assert lib.lis_b200_comm_attach(rank, world, int(tok[0])) == 0
    L = shim.lib

    if mode == "overlap":
        return overlap_case(rank, world, shim, lib)
    l, m, n = 3 * world + 1, 5, 4                      # planes do not divide evenly among the ranks
    ptr, idx, val = H.poisson3d_7pt(l, m, n)
    gn = l * m * n
    # reference partition LIS_GET_ISIE(rank, world, gn)
    q, r = divmod(gn, world)
    sizes = [q + 1 if k < r else q for k in range(world)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    is_, ie = int(starts[rank]), int(starts[rank + 1])
    lp = (ptr[is_:ie + 1] - ptr[is_]).astype(np.int32)
    li = np.ascontiguousarray(idx[ptr[is_]:ptr[ie]]); lv = np.ascontiguousarray(val[ptr[is_]:ptr[ie]])
    nl = ie - is_

    i32p = np.ctypeslib.ndpointer(np.int32, flags="C"); f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    # ---- host-side logic through the public API (works without a device)
    A = C.c_void_p()
    assert lib.lis_matrix_create(1, C.byref(A)) == 0
    assert lib.lis_matrix_set_size(A, 0, gn) == 0       # global size given: LIS_GET_ISIE split
    a, b = C.c_int(), C.c_int()
    assert lib.lis_matrix_get_range(A, C.byref(a), C.byref(b)) == 0
    assert (a.value, b.value) == (is_, ie), ((a.value, b.value), (is_, ie))
    libc = C.CDLL("libc.so.6"); libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]

    def to_malloc(arr):
        p = libc.malloc(max(arr.nbytes, 8))
        C.memmove(p, arr.ctypes.data, arr.nbytes)
        return p
    lib.lis_matrix_set_csr.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib.lis_matrix_set_csr(int(lp[-1]), to_malloc(lp), to_malloc(li), to_malloc(lv), A) == 0
    assert lib.lis_matrix_assemble(A) == 0
    lib.lis_b200_commtable_info.argtypes = [C.c_void_p, i32p, i32p, i32p, i32p, i32p, C.c_int]
    out = np.zeros(3, np.int32); imp = np.zeros(world + 1, np.int32); exp = np.zeros(world + 1, np.int32)
    cap = 4096
    ex = np.zeros(cap, np.int32); l2g = np.zeros(cap, np.int32)
    assert lib.lis_b200_commtable_info(A, out, imp, exp, ex, l2g, cap) == 0
    # expectation from the global matrix
    cols = idx[ptr[is_]:ptr[ie]]
    halo = np.unique(cols[(cols < is_) | (cols >= ie)])
    assert out[0] == len(halo) and np.array_equal(l2g[:len(halo)], halo), "halo list (l2g_map)"
    owner = np.searchsorted(starts, halo, side="right") - 1
    assert np.array_equal(np.diff(imp), np.bincount(owner, minlength=world)), "import_ptr"
    want_ex = []
    for k in range(world):
        if k == rank:
            continue
        ck = idx[ptr[starts[k]]:ptr[starts[k + 1]]]
        hk = np.unique(ck[(ck < starts[k]) | (ck >= starts[k + 1])])
        want_ex.append(hk[(hk >= is_) & (hk < ie)] - is_)
    want_ex = np.concatenate(want_ex) if want_ex else np.zeros(0, np.int64)
    assert out[1] == len(want_ex) and np.array_equal(ex[:len(want_ex)], want_ex), "export_index"
    # rank-ordered host allreduce: identical bits on every rank
    vals = np.array([0.1 * (rank + 1) + 1e-17 * rank, float(rank)], np.float64)
    lib.lis_b200_allreduce_sum.argtypes = [f64p, C.c_int]
    assert lib.lis_b200_allreduce_sum(vals, 2) == 0
    expect = 0.0
    for k in range(world):
        expect = expect + (0.1 * (k + 1) + 1e-17 * k) if k else 0.1 * (k + 1) + 1e-17 * k
    assert vals[0] == expect and vals[1] == sum(range(world)), (vals, expect)
    lib.lis_matrix_destroy(A)

    result = {"rank": rank, "mode": mode}
    if mode in ("gpu", "hostcheck"):
        os.environ["LIS_B200_PIPE_CHUNKS"] = "3"
        L.shim_mv_open_dist.argtypes = [C.c_int, C.c_int, i32p, i32p, f64p, C.c_int]
        L.shim_mv_set_x_local.argtypes = [C.c_int, f64p]; L.shim_mv_get_y_local.argtypes = [C.c_int, f64p]
        L.shim_mv_dot_xy.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.shim_mv_solve_b.argtypes = [C.c_int, C.c_char_p, f64p, f64p, i32p, f64p, f64p, C.c_int]
        x = H.rand_vec(gn, 5, "wide")
        o = H.Oracle()
        y_full = o.spmv("csr", ptr, idx, val, x)
        bvec = o.spmv("csr", ptr, idx, val, np.ones(gn))
        for fmt in ("csr", "ell", "dia", "jad", "bsr", "csc"):
            h = L.shim_mv_open_dist(lis_b200.FMT[fmt], nl, lp, li, lv, 0)
            assert h >= 0, (fmt, h)
            assert L.shim_mv_set_x_local(h, np.ascontiguousarray(x[is_:ie])) == 0
            for rep in range(2):                        # second product: the halo buffer is reused
                assert L.shim_mv_matvec(h) == 0
            yl = np.zeros(nl)
            assert L.shim_mv_get_y_local(h, yl) == 0
            if fmt == "csc":
                # CSC sums a row in ascending LOCAL column order (halo columns come after the owned ones, as in the
                # reference's MPI build): same products, another order than the one-process matrix
                assert np.allclose(yl, y_full[is_:ie], rtol=1e-13, atol=1e-13 * np.abs(y_full).max())
            elif fmt == "bsr":
                # BSR adds a row's products block column by block column in first-seen block order; with local + halo
                # numbering the blocks differ from the one-process matrix's: same products, another order
                assert np.allclose(yl, y_full[is_:ie], rtol=1e-13, atol=1e-13 * np.abs(y_full).max())
            elif fmt == "dia":
                # DIA sums by ascending LOCAL offset, and halo columns are numbered after the owned
                # ones (as in the reference's MPI build), so the order differs from the one-process
                # run for rows that touch the lower halo: same products, different rounding
                assert np.allclose(yl, y_full[is_:ie], rtol=1e-13, atol=1e-13 * np.abs(y_full).max())
            else:
                H.assert_bits_equal(yl, y_full[is_:ie], f"rank {rank} spmv {fmt}")
            if fmt == "csr":
                # the overlapped host-buffer product on local slices: chunks that read halo entries
                # run behind the exchange, the others as their inputs land (LIS_B200_PIPE_CHUNKS=3 here)
                L.shim_mv_step_e2e_pipelined.argtypes = [C.c_int, f64p, f64p]
                for seed in (8, 9):
                    x2 = H.rand_vec(gn, seed, "wide")
                    y2 = np.full(nl, np.nan)
                    assert L.shim_mv_step_e2e_pipelined(h, np.ascontiguousarray(x2[is_:ie]), y2) == 0
                    H.assert_bits_equal(y2, o.spmv("csr", ptr, idx, val, x2)[is_:ie], f"rank {rank} overlapped spmv")
                assert L.shim_mv_set_x_local(h, np.ascontiguousarray(x[is_:ie])) == 0
                assert L.shim_mv_matvec(h) == 0
            d = C.c_double()
            assert L.shim_mv_dot_xy(h, C.byref(d)) == 0
            assert abs(d.value - float(np.dot(x, y_full))) <= 1e-9 * abs(float(np.dot(np.abs(x), np.abs(y_full))))
            if fmt == "csr":
                # y = A^T x: local transposed product with the halo columns as extra rows, their sums sent back to the
                # owners and added in rank order (lis_reduce) -- against the one-process product of the transposed matrix
                import scipy.sparse as sp
                AT = sp.csr_matrix((val, idx, ptr), shape=(gn, gn)).T.tocsr()
                yt_full = AT @ x
                scale_t = np.abs(AT) @ np.abs(x)
                for rep in range(2):
                    assert L.shim_mv_matvech(h) == 0
                ytl = np.zeros(nl)
                assert L.shim_mv_get_y_local(h, ytl) == 0
                assert np.all(np.abs(ytl - yt_full[is_:ie]) <= 1e-14 * scale_t[is_:ie] + 1e-300), f"rank {rank} matvech"
                assert L.shim_mv_matvec(h) == 0                      # y back to A x for the dot check below
                xl = np.zeros(nl); oi = np.zeros(4, np.int32); od = np.zeros(4); rh = np.zeros(5000)
                rc = L.shim_mv_solve_b(h, b"-i bicg -p jacobi", np.ascontiguousarray(bvec[is_:ie]), xl, oi, od, rh, 5000)
                assert rc == 0 and oi[1] == 0 and np.abs(xl - 1.0).max() < 1e-8, ("bicg", rc, oi)
                result["bicg"] = int(oi[0])
                for opts, solver, pre in (("-i cg -p jacobi", "cg", "jacobi"), ("-i bicgstab -p jacobi", "bicgstab", "jacobi"),
                                          ("-i gmres -restart 20 -p none", "gmres", "none"), ("-i cg -p ssor", "cg", "ssor")):
                    xl = np.zeros(nl); oi = np.zeros(4, np.int32); od = np.zeros(4); rh = np.zeros(5000)
                    rc = L.shim_mv_solve_b(h, opts.encode(), np.ascontiguousarray(bvec[is_:ie]), xl, oi, od, rh, 5000)
                    assert rc == 0 and oi[1] == 0, (opts, rc, oi)
                    assert np.abs(xl - 1.0).max() < 1e-8, opts
                    kw = {"restart": 20} if solver == "gmres" else {}
                    if pre == "ssor":
                        # block-SSOR: one block per rank == the OpenMP reference with `world` threads
                        ref = o.solve(solver, ptr, idx, val, bvec, precon=pre, nthreads=1, ssor_blocks=world, **kw)
                    else:
                        ref = o.solve(solver, ptr, idx, val, bvec, precon=pre, **kw)
                    tol_it = 1 if solver == "bicgstab" else 0
                    assert abs(int(oi[0]) - ref["iter"]) <= tol_it, (opts, int(oi[0]), ref["iter"])
                    result[opts] = int(oi[0])
            L.shim_mv_close(h)
        # -scale on the row-partitioned matrix (symm_diag needs the diagonal entries of the halo columns: one halo
        # exchange of the diagonal).  A stays scaled after the solve, as in the reference, so each case takes a fresh
        # matrix.  Same iteration count (to the usual +-1 of the reduction order) and solution as the serial reference
        # on the whole matrix.
        refsh = H.ref_shim("serial")
        for opts in ("-i cg -p jacobi -scale symm_diag", "-i bicgstab -p none -scale jacobi", "-i bicgstab -p ssor -scale symm_diag",
                     "-i gmres -restart 20 -p jacobi -scale symm_diag"):
            h = L.shim_mv_open_dist(lis_b200.FMT["csr"], nl, lp, li, lv, 0)
            assert h >= 0
            xl = np.zeros(nl); oi = np.zeros(4, np.int32); od = np.zeros(4); rh = np.zeros(5000)
            rc = L.shim_mv_solve_b(h, opts.encode(), np.ascontiguousarray(bvec[is_:ie]), xl, oi, od, rh, 5000)
            assert rc == 0 and oi[1] == 0, (opts, rc, oi)
            assert np.abs(xl - 1.0).max() < 1e-8, (opts, np.abs(xl - 1.0).max())
            if refsh is not None and "ssor" not in opts:          # block SSOR per rank is another preconditioner than serial SSOR
                r = refsh.solve(ptr, idx, val, bvec, opts)
                assert abs(int(oi[0]) - r["iter"]) <= 1, (opts, int(oi[0]), r["iter"])
                k = max(2, len(r["rhistory"]) // 2)
                assert np.allclose(rh[:k], r["rhistory"][:k], rtol=1e-6), opts
            result[opts] = int(oi[0])
            L.shim_mv_close(h)
    gathered = [None] * world
    dist.all_gather_object(gathered, result)
    dist.barrier()
    lib.lis_finalize()
    if rank == 0:
        print("MR_OK", gathered)


if __name__ == "__main__":
    main()
