#!/bin/bash
# Round 2, GPU session 15 (4 GPUs): the in-kernel halo exchange with middle ranks (two neighbours each); bench at N=4.
O=gpurun_out; mkdir -p $O
export LIS_B200_VERBOSE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29951 bench.py --gpus 4 --steps 20 --warmup 3 > $O/r02_bench_4gpu_b.json 2> $O/r02_bench_4gpu_b.log; echo "bench rc=$?"
grep -E "rank 0.*(ms/product|in-kernel|CG)|^4 GPUs|lis_b200:|Error|error" $O/r02_bench_4gpu_b.log | cut -c1-300 | sort -u
cut -c1-300 $O/r02_bench_4gpu_b.json
