#!/bin/bash
# Round 2, GPU session 12 (8 GPUs): the multi-GPU default after session 11 (every GPU visible, NCCL exchange, no peer
# access enabled by the library): bench at 8; config 5 at 8; does cudaMallocManaged fail because of peer access?
O=gpurun_out; mkdir -p $O
export LIS_B200_VERBOSE=1
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR8 --master-port 29801 profiles/managed_probe.py 2>&1 | grep -E "rank 0\]|max_map" | head -6
PROBE_PEER=1 timeout 200 $TR8 --master-port 29802 profiles/managed_probe.py 2>&1 | grep -E "rank 0\]" | head -6
timeout 600 $TR8 --master-port 29803 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r02_bench_8gpu_b.json 2> $O/r02_bench_8gpu_b.log; echo "bench8 rc=$?"
grep -E "rank 0.*(ms/product|in-kernel|CG|e2e)|^8 GPUs|lis_b200:|Error|error" $O/r02_bench_8gpu_b.log | cut -c1-260 | sort -u | head -20
timeout 600 $TR8 --master-port 29804 profiles/run_configs.py gm27 --size 768 --slab 96 --maxiter 600 --out $O/r02_configs_n8b.jsonl 2>&1 | grep -E '^\{|Error|error|lis_b200:' | sort -u | cut -c1-700
