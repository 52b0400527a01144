#!/bin/bash
# Round 2, GPU session 7 (2 GPUs): is the un-narrowed multi-rank run slower, or was it the box?  N=1 on this box, then
# N=2 with each rank seeing only its GPU, then N=2 with every GPU visible (in-kernel exchange available).  Sweep v4 timing.
O=gpurun_out; mkdir -p $O
D=lis_b200/_lib/drivers
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
export LIS_B200_VERBOSE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== sweep v4: CG+SSOR 256^3"; $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 2000 2>&1 | grep -E "number of iterations|CG:   linear solver" | head -3
timeout 300 python -m pytest tests -m gpu -x -q -k "ssor or sweep or ilu or psolve" 2>&1 | tail -2
echo "== N=1"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-format-extras --no-cg-converge 2>&1 >/dev/null | grep -E "^CSR|CG\+Jacobi:" | cut -c1-200
echo "== N=2 narrowed"; LIS_B200_NARROW=1 timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 >$O/r02_bench_2gpu_narrow.json | grep -E "rank 0.*(ms/product|in-kernel|CG)|^2 GPUs|lis_b200:" | cut -c1-250
echo "== N=2 all visible"; timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 >$O/r02_bench_2gpu_visible.json | grep -E "rank 0.*(ms/product|in-kernel|CG)|^2 GPUs|lis_b200:" | cut -c1-250
echo "== N=2 all visible, vectors in device memory"; LIS_B200_VECTORS=device timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 >$O/r02_bench_2gpu_visible_dev.json | grep -E "rank 0.*(ms/product|in-kernel|CG)|^2 GPUs|lis_b200:" | cut -c1-250
