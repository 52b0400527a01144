#!/bin/bash
# Round 2, GPU session 16 (8 GPUs): the final tree at N=8 (what the driver's scaling run executes).
O=gpurun_out; mkdir -p $O
export LIS_B200_VERBOSE=1
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 500 $TR8 --master-port 29961 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r02_bench_8gpu_c.json 2> $O/r02_bench_8gpu_c.log; echo "bench8 rc=$?"
grep -E "rank 0.*(ms/product|in-kernel|CG)|^8 GPUs|lis_b200:|Error|error" $O/r02_bench_8gpu_c.log | cut -c1-300 | sort -u
cut -c1-200 $O/r02_bench_8gpu_c.json
timeout 300 $TR8 --master-port 29962 bench.py --impl reference --gpus 8 --steps 20 --warmup 3 2>/dev/null | cut -c1-700
