"""Multi-process tests of the row-partitioned path: world_size 2 (and 3) processes on one host.
CPU part (always): partition, global->local numbering, halo lists and the rank-ordered host
allreduce, with gloo carrying the harness traffic.  GPU part (-m gpu, needs >= 2 GPUs):
SpMV with NCCL halo exchange and the solvers against the single-process oracle."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def launch(world, mode, timeout=300, extra_env=None):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), GLOO_SOCKET_IFNAME="lo")
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "mr_worker.py"), mode], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
    assert "MR_OK" in outs[0], outs[0][-2000:]
    return outs[0]


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_halo_lists_cpu(built, world):
    launch(world, "cpu")


@pytest.mark.parametrize("world", [2, 3])
def test_row_partitioned_spmv_and_solvers_hostcheck(built, world):
    """the complete multi-rank flow -- assemble, halo exchange (host-staged transport), SpMV in four
    formats, CG / BiCGSTAB / GMRES / block-SSOR -- on the mock-device build, CPU only"""
    import harness
    d = harness.ensure_hostcheck()
    launch(world, "hostcheck", timeout=600, extra_env={"LIS_B200_HOSTCHECK_DIR": d, "LIS_B200_TRANSPORT": "host"})


@pytest.mark.parametrize("world", [2, 3])
def test_row_partitioned_spmv_and_solvers_emulated_kernels(built, world):
    """the same flow with the product's kernel sources on the host emulator (tests/cudaemu) instead of
    the mock: gather/pack, the SpMV kernels on local+halo numbering, reductions, block-SSOR sweeps"""
    d = os.path.join(HERE, "cudaemu")
    r = subprocess.run(["make", "-C", d, "-j8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    launch(world, "hostcheck", timeout=900, extra_env={"LIS_B200_HOSTCHECK_DIR": os.path.join(d, "_build"),
                                                       "LIS_B200_HOSTCHECK_NAME": "emu", "LIS_B200_TRANSPORT": "host"})


@pytest.mark.parametrize("world", [2, 3])
def test_interior_rows_overlap_the_halo_exchange_emulated(built, world):
    """row-partitioned CSR product with the interior rows on a second stream during the exchange, on
    the kernel emulator (19 k rows per rank at world 3, both CSR kernels), plus the host-buffer product"""
    d = os.path.join(HERE, "cudaemu")
    r = subprocess.run(["make", "-C", d, "-j8"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    launch(world, "overlap", timeout=900, extra_env={"LIS_B200_HOSTCHECK_DIR": os.path.join(d, "_build"), "LIS_B200_HOSTCHECK_NAME": "emu",
                                                     "LIS_B200_TRANSPORT": "host", "LIS_B200_OVERLAP": "force"})


@pytest.mark.gpu
def test_row_partitioned_spmv_and_solvers_2gpu(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    launch(2, "gpu", timeout=600)


@pytest.mark.gpu
def test_interior_rows_overlap_and_in_kernel_exchange_2gpu(built):
    """19 k rows per rank, an interior run of row blocks: products with the NCCL exchange (interior rows on a second stream)
    and with the halo exchange inside the kernel over peer memory, both CSR kernels, against the one-process oracle bit for
    bit; CG with the split fused step; the overlapped host-buffer product"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    launch(2, "overlap", timeout=600, extra_env={"LIS_B200_OVERLAP": "force"})
