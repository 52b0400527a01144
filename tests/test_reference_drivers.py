"""Drop-in check: the reference's OWN drivers (test/spmvtest1.c, spmvtest3.c, spmvtest3b.c,
test1.c, test3.c, test3b.c), compiled unchanged against include/ + liblis_b200.so by the
top-level Makefile, run on the GPU and must print what the same sources print when linked
against the reference (oracle/_ref/drivers, built from the reference tree by oracle/Makefile)."""
import os
import re
import subprocess

import numpy as np
import pytest

import harness as H

OURS = os.path.join(H.ROOT, "lis_b200", "_lib", "drivers")
REFS = os.path.join(H.ROOT, "oracle", "_ref", "drivers")


def run(path, *args, cwd=None):
    r = subprocess.run([path, *map(str, args)], capture_output=True, text=True, timeout=600, cwd=cwd)
    assert r.returncode == 0, f"{path} {args} exited {r.returncode}\n{r.stdout[-1500:]}\n{r.stderr[-1500:]}"
    return r.stdout


def need(*names):
    for n in names:
        if not os.path.exists(os.path.join(OURS, n)):
            pytest.skip(f"driver {n} not built (needs the reference tree at build time)")


def test_drivers_were_built_from_unmodified_reference_sources():
    """all 28 C drivers of the reference compile and link against lis_b200 unchanged (checked where the tree exists);
    the three generalized-eigenproblem drivers getest1/5/5b stop at lis_gesolve(A, B, ...) with LIS_ERR_NOT_IMPLEMENTED"""
    if not os.path.isdir("/root/reference/test"):
        pytest.skip("reference tree not present")
    H.ensure_built()
    for n in ("spmvtest1", "spmvtest2", "spmvtest2b", "spmvtest3", "spmvtest3b", "spmvtest4", "spmvtest5", "test1", "test2", "test2b",
              "test3", "test3b", "test3c", "test4", "test5", "etest1", "etest2", "etest3", "etest4", "etest5", "etest5b", "etest6", "etest7", "test6", "test7", "getest1", "getest5", "getest5b"):
        assert os.path.exists(os.path.join(OURS, n)), n


def test_dense_helper_drivers_print_what_the_reference_prints():
    """test6 (dense Gaussian elimination / matvec through lis_array_*) and etest7 (QR iteration on a
    dense matrix): host-only drivers, so they run here; transcripts equal to the reference-linked binaries"""
    need("test6", "etest7")
    for d in ("test6", "etest7"):
        if not os.path.exists(os.path.join(REFS, d)):
            pytest.skip("oracle/_ref drivers not built (reference tree absent)")
        for args in ((3, 2), (4, 4), (6, 5)):
            keep = lambda s: [ln for ln in s.splitlines() if "sec" not in ln and "time" not in ln]
            assert keep(run(os.path.join(OURS, d), *args)) == keep(run(os.path.join(REFS, d), *args)), (d, args)
    if os.path.exists(os.path.join(OURS, "test7")) and os.path.exists(os.path.join(REFS, "test7")):
        assert run(os.path.join(OURS, "test7")) == run(os.path.join(REFS, "test7"))         # the complex-number smoke driver (real build)


def norms(out):
    return {int(m.group(1)): float(m.group(2)) for m in re.finditer(r"matrix_type\s*=\s*(\d+).*2-norm = (\S+)", out)}


@pytest.mark.gpu
@pytest.mark.parametrize("driver,args,analytic", [
    ("spmvtest1", (100000, 20), "1.414214e+00"),
    ("spmvtest3", (24, 24, 24, 5), None),
    ("spmvtest3b", (12, 12, 12, 5), None),
])
def test_spmvtest_drivers(driver, args, analytic):
    need(driver)
    for fmt in (1, 2, 4, 5, 6, 7):
        out = run(os.path.join(OURS, driver), *args, fmt)
        got = norms(out)
        assert fmt in got, out
        if analytic:
            assert f"{got[fmt]:e}" == analytic
        if os.path.exists(os.path.join(REFS, driver)):
            ref = norms(run(os.path.join(REFS, driver), *args, fmt))
            assert f"{got[fmt]:e}" == f"{ref[fmt]:e}", (driver, fmt)
    if driver == "spmvtest3":
        N = 24
        assert abs(got[7] - np.sqrt(6 * (N - 2) ** 2 + 48 * (N - 2) + 72)) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("driver,args", [("spmvtest1", (30000, 5)), ("spmvtest2", (40, 30, 5)), ("spmvtest3", (12, 11, 10, 5))])
def test_spmvtest_drivers_walk_all_formats(driver, args):
    """without a format argument the drivers convert to and multiply in formats 1..10 (CSR, CSC, MSR, DIA,
    ELL, JAD, BSR, BSC, VBR, COO; test/spmvtest1.c:188-204): every one answers, with the reference's norm"""
    need(driver)
    got = norms(run(os.path.join(OURS, driver), *args))
    assert sorted(got) == list(range(1, 11)), got
    if driver == "spmvtest1":
        assert all(f"{v:e}" == "1.414214e+00" for v in got.values()), got
    if os.path.exists(os.path.join(REFS, driver)):
        ref = norms(run(os.path.join(REFS, driver), *args))
        assert {k: f"{v:e}" for k, v in got.items()} == {k: f"{v:e}" for k, v in ref.items()}


def solver_lines(out):
    it = re.search(r"number of iterations\s*=\s*(\d+)", out)
    rs = re.search(r"relative residual\s*=\s*(\S+)", out)
    return int(it.group(1)), float(rs.group(1))


@pytest.mark.gpu
@pytest.mark.parametrize("opts", ["-i cg -p jacobi", "-i cg -p ssor", "-i bicgstab -p ssor", "-i gmres -restart 30 -p jacobi",
                                  "-i cg -p jacobi -storage ell", "-i cg -storage dia"])
def test_test3_driver(tmp_path, opts):
    need("test3")
    args = (20, 20, 20, 1, tmp_path / "sol.txt", tmp_path / "rh.txt", *opts.split())
    it, res = solver_lines(run(os.path.join(OURS, "test3"), *args))
    sol = np.loadtxt(tmp_path / "sol.txt", skiprows=2)[:, 1]
    assert np.abs(sol - 1.0).max() < 1e-8
    rh = np.loadtxt(tmp_path / "rh.txt")
    assert len(rh) == it + 1 and rh[0] == 1.0
    if os.path.exists(os.path.join(REFS, "test3")):
        args_r = (20, 20, 20, 1, tmp_path / "sol_r.txt", tmp_path / "rh_r.txt", *opts.split())
        it_r, res_r = solver_lines(run(os.path.join(REFS, "test3"), *args_r))
        assert abs(it - it_r) <= (1 if "bicgstab" in opts else 0), (opts, it, it_r)
        if it == it_r:
            rh_r = np.loadtxt(tmp_path / "rh_r.txt")
            k = (3 * len(rh)) // 4
            assert np.allclose(rh[:k], rh_r[:k], rtol=1e-5)


@pytest.mark.gpu
def test_test1_driver_matrix_market(tmp_path):
    """test1 reads a Matrix Market file (here: a 2-D 5-point Laplacian with its right-hand side
    appended in Lis' extended format, like test/testmat.mtx)"""
    need("test1")
    m = 12
    n = m * m
    lines = []
    for i in range(n):
        r, c = divmod(i, m)
        for j, v in ((i - m, -1.0), (i - 1, -1.0), (i, 4.0), (i + 1, -1.0), (i + m, -1.0)):
            if j < 0 or j >= n or (j == i - 1 and c == 0) or (j == i + 1 and c == m - 1):
                continue
            lines.append(f"{i + 1} {j + 1} {v:.20e}")
    A = np.zeros((n, n))
    for ln in lines:
        a, b, v = ln.split(); A[int(a) - 1, int(b) - 1] = float(v)
    rhs = A @ np.ones(n)
    text = ["%%MatrixMarket matrix coordinate real general", f"{n} {n} {len(lines)} 1 0"] + lines + [f"{i + 1} {rhs[i]:.20e}" for i in range(n)]
    (tmp_path / "lap.mtx").write_text("\n".join(text) + "\n")
    args = (tmp_path / "lap.mtx", 0, tmp_path / "sol.txt", tmp_path / "rh.txt", "-i", "cg", "-p", "jacobi")
    it, res = solver_lines(run(os.path.join(OURS, "test1"), *args))
    sol = np.loadtxt(tmp_path / "sol.txt", skiprows=2)[:, 1]
    assert np.abs(sol - 1.0).max() < 1e-9 and res < 1e-12
    if os.path.exists(os.path.join(REFS, "test1")):
        args_r = (tmp_path / "lap.mtx", 0, tmp_path / "sol_r.txt", tmp_path / "rh_r.txt", "-i", "cg", "-p", "jacobi")
        it_r, _ = solver_lines(run(os.path.join(REFS, "test1"), *args_r))
        assert it == it_r


@pytest.mark.gpu
def test_test3b_driver_hpcg_kernel(tmp_path):
    """test3b (installed as hpcg_kernel) hard-wires -i cg -p ssor -adds true: CG with the additive Schwarz
    wrapper around SSOR, on its own 27-point matrix; same iteration count as the reference-linked binary"""
    need("test3b")
    args = (12, 12, 12, 1, tmp_path / "sol.txt", tmp_path / "rh.txt")
    it, res = solver_lines(run(os.path.join(OURS, "test3b"), *args))
    assert res < 1e-12 and it > 0
    sol = np.loadtxt(tmp_path / "sol.txt", skiprows=2)[:, 1]
    assert np.abs(sol - 1.0).max() < 1e-8
    if os.path.exists(os.path.join(REFS, "test3b")):
        args_r = (12, 12, 12, 1, tmp_path / "sol_r.txt", tmp_path / "rh_r.txt")
        it_r, _ = solver_lines(run(os.path.join(REFS, "test3b"), *args_r))
        assert it == it_r, (it, it_r)
