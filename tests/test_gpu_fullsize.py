"""Full-size checks (-m gpu): the BASELINE configuration itself, 3-D 7-point Poisson 512^3
(n = 134 217 728, nnz = 937 951 232), where the CPU oracle would take minutes.  Parity at this size
rests on properties that do not depend on it:
  * the reference drivers' known answer  ||A*1||_2 = sqrt(6(N-2)^2 + 48(N-2) + 72)  (exact: every
    entry of A*1 is a small integer, so the sum of squares is exact in fp64);
  * the four SpMV kernels (CSR product-tile, CSR TMA, ELL, DIA) give bit-identical y for the same
    random x (same products, same order, different memory layouts);
  * symmetry of the operator: <Ax, y> == <x, Ay> to rounding;
  * linearity: A(ax + by) == a*Ax + b*Ay to rounding.
The kernels are driven through the C-ABI on device arrays built with torch (as bench.py does)."""
import ctypes as C
import math
import os
import sys

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu
sys.path.insert(0, H.ROOT)


@pytest.fixture(scope="module")
def big():
    import torch
    import bench
    import lis_b200
    N = int(os.environ.get("LIS_B200_FULLSIZE_GRID", "512"))
    dev = torch.device("cuda", 0)
    K = lis_b200.load_kernels()
    ptr, idx, val = bench.poisson7_device(torch, N, N, N, 0, N, dev)
    n, nnz = ptr.numel() - 1, idx.numel()
    idx_p = torch.cat([idx, torch.zeros(8, device=dev, dtype=torch.int32)])
    val_p = torch.cat([val, torch.zeros(8, device=dev, dtype=torch.float64)])
    ptr_p = torch.cat([ptr, torch.zeros(4, device=dev, dtype=torch.int32)])
    del idx, val
    stream = torch.cuda.Stream(device=dev)
    partial = torch.zeros(K.lisb200_reduce_slots(), device=dev, dtype=torch.float64)
    counter = torch.zeros(16, device=dev, dtype=torch.int32)
    result = torch.zeros(4, device=dev, dtype=torch.float64)
    return dict(torch=torch, K=K, N=N, n=n, nnz=nnz, ptr=ptr_p, idx=idx_p, val=val_p, dev=dev, stream=stream,
                sp=C.c_void_p(stream.cuda_stream), partial=partial, counter=counter, result=result)


def spmv_tma(b, x, y):
    b["torch"].cuda.synchronize()            # torch filled x on ITS stream; the kernel runs on ours
    rc = b["K"].lisb200_spmv_csr_tma(b["n"], 256, 2048, 4, b["ptr"].data_ptr(), b["idx"].data_ptr(), b["val"].data_ptr(),
                                     x.data_ptr(), y.data_ptr(), b["sp"])
    assert rc == 0
    b["stream"].synchronize()


def dot(b, x, y):
    b["torch"].cuda.synchronize()
    rc = b["K"].lisb200_reduce(0, b["n"], x.data_ptr(), y.data_ptr(), b["partial"].data_ptr(), b["counter"].data_ptr(),
                               b["result"].data_ptr(), b["sp"])
    assert rc == 0
    b["stream"].synchronize()
    return float(b["result"][0].item())


def test_known_answer_norm_of_A_times_ones(big):
    t = big["torch"]
    x = t.ones(big["n"], device=big["dev"], dtype=t.float64); y = t.zeros_like(x)
    spmv_tma(big, x, y)
    N = big["N"]
    exact_sq = 6 * (N - 2) ** 2 + 48 * (N - 2) + 72
    assert dot(big, y, y) == float(exact_sq)                 # integers: exact whatever the summation tree
    assert math.sqrt(dot(big, y, y)) == math.sqrt(exact_sq)
    assert float(y.abs().max().item()) == 3.0 and float(y.min().item()) == 0.0


def test_four_kernels_same_bits(big):
    t, K, n, nnz, dev, sp = big["torch"], big["K"], big["n"], big["nnz"], big["dev"], big["sp"]
    t.manual_seed(7)
    x = t.rand(n, device=dev, dtype=t.float64) * 2 - 1
    y0 = t.zeros_like(x); y = t.zeros_like(x)
    spmv_tma(big, x, y0)
    t.cuda.synchronize()
    assert K.lisb200_spmv_csr(n, big["ptr"].data_ptr(), big["idx"].data_ptr(), big["val"].data_ptr(), x.data_ptr(), y.data_ptr(), sp) == 0
    big["stream"].synchronize()
    assert t.equal(y.view(t.int64), y0.view(t.int64)), "product-tile CSR != TMA CSR"
    ptr = big["ptr"][:n + 1]; idx = big["idx"][:nnz]; val = big["val"][:nnz]
    rows = t.repeat_interleave(t.arange(n, device=dev, dtype=t.int64), (ptr[1:] - ptr[:-1]).to(t.int64))
    slot = t.arange(nnz, device=dev, dtype=t.int64) - ptr[:-1].to(t.int64)[rows]
    ell_i = t.arange(n, device=dev, dtype=t.int32).repeat(7); ell_v = t.zeros(7 * n, device=dev, dtype=t.float64)
    ell_i[slot * n + rows] = idx; ell_v[slot * n + rows] = val
    t.cuda.synchronize()
    assert K.lisb200_spmv_ell(n, 7, n, ell_i.data_ptr(), ell_v.data_ptr(), x.data_ptr(), y.data_ptr(), sp) == 0
    big["stream"].synchronize()
    assert t.equal(y.view(t.int64), y0.view(t.int64)), "ELL != CSR"
    del ell_i, ell_v, slot
    N = big["N"]
    offs = t.tensor([-N * N, -N, -1, 0, 1, N, N * N], device=dev, dtype=t.int32)
    dia_v = t.zeros(7 * n, device=dev, dtype=t.float64)
    dia_v[t.searchsorted(offs.to(t.int64), idx.to(t.int64) - rows) * n + rows] = val
    del rows
    t.cuda.synchronize()
    assert K.lisb200_spmv_dia(n, n, 7, n, offs.data_ptr(), dia_v.data_ptr(), x.data_ptr(), y.data_ptr(), sp) == 0
    big["stream"].synchronize()
    assert t.equal(y.view(t.int64), y0.view(t.int64)), "DIA != CSR"


def test_symmetry_and_linearity(big):
    t, n, dev = big["torch"], big["n"], big["dev"]
    t.manual_seed(11)
    x = t.rand(n, device=dev, dtype=t.float64) * 2 - 1
    y = t.rand(n, device=dev, dtype=t.float64) * 2 - 1
    ax = t.zeros_like(x); ay = t.zeros_like(x)
    spmv_tma(big, x, ax); spmv_tma(big, y, ay)
    lhs, rhs = dot(big, ax, y), dot(big, x, ay)
    scale = dot(big, ax.abs(), y.abs())
    assert abs(lhs - rhs) <= 1e-13 * scale, (lhs, rhs)
    a, b_ = 0.37, -1.9
    z = a * x + b_ * y
    az = t.zeros_like(x)
    spmv_tma(big, z, az)
    err = (az - (a * ax + b_ * ay)).abs().max().item()
    assert err <= 256 * np.finfo(float).eps * 12.0, err       # |row sum| <= 12 * max|z|, a few roundings each side


@pytest.mark.parametrize("grid", [256, 512])
def test_cg_jacobi_to_1e12_iteration_count_equals_reference(b200, grid):
    """BASELINE.json config 3 at its stated size: test/test3.c's system (7-pt Poisson, rows in test3.c order, b = A*1,
    x0 = 0), `-i cg -p jacobi -tol 1e-12` through lis_solve on the GPU against the compiled reference's run of the same
    system (tests/golden/cg_poisson_<grid>.npz, made by tests/golden/make_cg_fullsize.py with the OpenMP build):
    identical iteration count (764 at 256^3, 1504 at 512^3); residual history within the reference's own
    thread-count envelope (~1e-12..1e-9 relative over the first three quarters, growing
    towards convergence where conditioning takes over); solution error of the same size as the reference's."""
    import bench
    out = bench.cg_to_convergence(b200.lib, grid)
    assert out["reference_iters"] is not None, out["reference_source"]
    assert out["cg_iters_to_1e-12"] == out["reference_iters"], out
    assert out["cg_final_relres"] < 1e-12
    assert out["history_gap_first_three_quarters"] < 1e-8, out
    assert out["history_gap"] < 5e-2, out
    assert out["cg_max_abs_x_minus_1"] < 1e-9
