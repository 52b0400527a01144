"""Drop-in check, part 2 (-m gpu): the reference drivers added after round 1's GPU minutes were spent --
spmvtest4 / spmvtest5 (matrices from Matrix Market and Harwell-Boeing files), etest1 / etest5
(eigensolvers).  Same idea as test_reference_drivers.py: unchanged reference sources linked against
lis_b200 must print what they print when linked against the reference."""
import os
import re
import subprocess

import numpy as np
import pytest

import harness as H
import test_reference_drivers as D1
from test_reference_drivers import REFS, need, run


def _ours(name):
    return os.path.join(D1.OURS, name)          # looked up at call time: the emulator suite points OURS elsewhere


def _write_mtx(path, ptr, idx, val):
    n = len(ptr) - 1
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate real general\n{n} {n} {int(ptr[-1])}\n")
        for i in range(n):
            for j in range(ptr[i], ptr[i + 1]):
                f.write(f"{i + 1} {idx[j] + 1} {val[j]:.20e}\n")


@pytest.mark.gpu
@pytest.mark.parametrize("driver", ["spmvtest4", "spmvtest5"])
def test_spmvtest_file_drivers(tmp_path, driver):
    """spmvtest4 / spmvtest5 read their matrices from files (a list of file names / one file): one
    Matrix Market file and one Harwell-Boeing file, every storage format the drivers cycle through
    that lis_b200 has; the printed 2-norms must equal the reference-linked driver's"""
    need(driver)
    from test_host_logic import _write_hb
    ptr, idx, val = H.poisson3d_7pt(9, 8, 7)
    _write_mtx(tmp_path / "a.mtx", ptr, idx, val)
    ptr2, idx2, val2 = H.random_csr(300, 5, 4)
    _write_hb(str(tmp_path / "b.rua"), ptr2, idx2, val2)
    (tmp_path / "list.txt").write_text(f"{tmp_path / 'a.mtx'}\n{tmp_path / 'b.rua'}\n")
    if driver == "spmvtest4":
        runs = [((tmp_path / "list.txt", 3), None)]
    else:
        runs = [((tmp_path / "a.mtx", fmt, 3), fmt) for fmt in (1, 2, 4, 5, 6, 7)] + [((tmp_path / "b.rua", 1, 3), 1)]
    for args, fmt in runs:
        r = subprocess.run([_ours(driver), *map(str, args)], capture_output=True, text=True, timeout=600)
        got = re.findall(r"matrix_type\s*=\s*(\d+).*2-norm = (\S+)", r.stdout)
        assert got, (r.returncode, r.stdout[-1500:], r.stderr[-800:])
        if os.path.exists(os.path.join(REFS, driver)):
            q = subprocess.run([os.path.join(REFS, driver), *map(str, args)], capture_output=True, text=True, timeout=600)
            ref = re.findall(r"matrix_type\s*=\s*(\d+).*2-norm = (\S+)", q.stdout)
            assert got == ref[:len(got)], (driver, args, got, ref[:len(got)])      # same lines, as far as ours goes
        # formats lis_b200 does not carry end the driver's cycle with LIS_ERR_NOT_IMPLEMENTED (exit code 5)
        assert r.returncode in (0, 5), (r.returncode, r.stderr[-800:])
        assert {int(t) for t, _ in got} >= ({1, 2} if fmt is None else {fmt}), got


@pytest.mark.gpu
@pytest.mark.parametrize("opts", ["", "-e ii -i cg -p jacobi", "-e rqi", "-e cg -i cg"])
def test_etest1_driver(tmp_path, opts):
    """etest1 (the reference's `make check` eigen case): eigenvalue of a Matrix Market matrix with the
    default CR eigensolver and three others; printed eigenvalue equal to the reference-linked driver's
    to the 7 digits it prints, iteration count within a few steps"""
    need("etest1")
    ptr, idx, val = H.poisson3d_7pt(6, 5, 4)
    _write_mtx(tmp_path / "a.mtx", ptr, idx, val)

    def go(binary, tag):
        out = run(binary, tmp_path / "a.mtx", tmp_path / f"evec_{tag}.txt", tmp_path / f"rh_{tag}.txt", *opts.split())
        ev = float(re.search(r"eigenvalue\s*=\s*(\S+)", out).group(1))
        it = int(re.search(r"number of iterations\s*=\s*(\d+)", out).group(1))
        return ev, it
    ev, it = go(_ours("etest1"), "ours")
    vec = np.loadtxt(tmp_path / "evec_ours.txt", skiprows=2)[:, 1]
    assert abs(np.linalg.norm(vec) - 1.0) < 1e-12
    import scipy.sparse as sp
    A = sp.csr_matrix((val, idx, ptr), shape=(len(ptr) - 1,) * 2)
    assert np.linalg.norm(A @ vec - ev * vec) < 1e-6 * abs(ev)
    if os.path.exists(os.path.join(REFS, "etest1")):
        ev_r, it_r = go(os.path.join(REFS, "etest1"), "ref")
        assert f"{ev:e}" == f"{ev_r:e}" and abs(it - it_r) <= max(2, it_r // 10), (opts, ev, ev_r, it, it_r)


@pytest.mark.gpu
def test_etest5_driver_lanczos(tmp_path):
    """etest5: several eigenpairs (Lanczos, refined by inverse iteration), written with
    lis_esolver_get_evalues / get_evectors / get_residualnorms / get_iters"""
    need("etest5")
    ptr, idx, val = H.poisson1d(40)
    _write_mtx(tmp_path / "a.mtx", ptr, idx, val)

    def go(binary, tag):
        files = [tmp_path / f"{k}_{tag}.txt" for k in ("evalues", "evectors", "resid", "iters")]
        run(binary, tmp_path / "a.mtx", *files, "-e", "li", "-ss", "3")
        return np.loadtxt(files[0], skiprows=2)[:, 1], np.loadtxt(files[1], skiprows=2)
    ev, vecs = go(_ours("etest5"), "ours")
    exact = 2.0 - 2.0 * np.cos(np.arange(1, 41) * np.pi / 41)
    for e in ev:
        assert np.abs(exact - e).min() < 1e-9, e
    assert vecs.shape == (40 * 3, 3)
    if os.path.exists(os.path.join(REFS, "etest5")):
        ev_r, vecs_r = go(os.path.join(REFS, "etest5"), "ref")
        assert np.allclose(ev, ev_r, rtol=1e-9)
        # same (row, mode) entries; the reference's COO -> CSR pass leaves them in its quicksort's order
        o, o_r = np.lexsort((vecs[:, 1], vecs[:, 0])), np.lexsort((vecs_r[:, 1], vecs_r[:, 0]))
        assert np.array_equal(vecs[o, :2], vecs_r[o_r, :2])
        for m in (1, 2, 3):
            a, b = vecs[o][vecs[o, 1] == m, 2], vecs_r[o_r][vecs_r[o_r, 1] == m, 2]
            assert min(np.abs(a - b).max(), np.abs(a + b).max()) < 1e-6, m
