#!/bin/bash
# Round 2, GPU session 13 (2 GPUs): the in-kernel halo exchange with the inbox mapped through the virtual-memory API.
O=gpurun_out; mkdir -p $O
export LIS_B200_VERBOSE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q 2>&1 | tail -8
timeout 600 $TR --master-port 29901 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r02_bench_2gpu_vmm.json 2> $O/r02_bench_2gpu_vmm.log; echo "bench rc=$?"
grep -E "rank 0.*(ms/product|in-kernel|CG)|^2 GPUs|lis_b200:|Error|error" $O/r02_bench_2gpu_vmm.log | cut -c1-260
