#!/bin/bash
# Round 2, GPU session 9 (8 GPUs, every GPU visible to every rank): managed-memory probe, bench at 8 and 4,
# configs 3/4/5 at 8 GPUs.
O=gpurun_out; mkdir -p $O
export LIS_B200_VERBOSE=1
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
nvidia-smi -L | head -8; nproc; free -g | head -2
timeout 200 $TR8 --master-port 29601 profiles/managed_probe.py 2>&1 | grep -E "rank [07]\]|max_map" | head -10
timeout 600 $TR8 --master-port 29602 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r02_bench_8gpu.json 2> $O/r02_bench_8gpu.log; rc=$?; echo "bench8 rc=$rc"
grep -E "rank 0.*(ms/product|in-kernel|CG|e2e)|^8 GPUs|lis_b200:|Error|error" $O/r02_bench_8gpu.log | cut -c1-260 | head -20
if [ $rc -ne 0 ]; then
  tail -20 $O/r02_bench_8gpu.log
  LIS_B200_NARROW=1 timeout 600 $TR8 --master-port 29603 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r02_bench_8gpu_narrow.json 2> $O/r02_bench_8gpu_narrow.log; echo "bench8 narrowed rc=$?"
  grep -E "rank 0.*(ms/product|in-kernel|CG|e2e)|^8 GPUs" $O/r02_bench_8gpu_narrow.log | cut -c1-260 | head
fi
nvcc -arch=sm_100a -O3 -o /tmp/hop profiles/hop_latency.cu && /tmp/hop | tee $O/r02_hop_latency.txt
rm -f $O/r02_configs_n8.jsonl
timeout 600 $TR8 --master-port 29604 profiles/run_configs.py gm27 --size 768 --slab 96 --maxiter 600 --out $O/r02_configs_n8.jsonl 2>&1 | grep -E '^\{|Error|error' | cut -c1-700
timeout 400 $TR8 --master-port 29605 profiles/run_configs.py su --size 10000000 --threads 2 --out $O/r02_configs_n8.jsonl 2>&1 | grep -E '^\{|Error|error' | cut -c1-700
timeout 400 $TR8 --master-port 29606 profiles/run_configs.py cg7 --size 512 --maxiter 600 --out $O/r02_configs_n8.jsonl 2>&1 | grep -E '^\{|Error|error' | cut -c1-700
timeout 600 $TR4 --master-port 29607 bench.py --gpus 4 --steps 20 --warmup 3 > $O/r02_bench_4gpu.json 2> $O/r02_bench_4gpu.log; echo "bench4 rc=$?"
grep -E "rank 0.*(ms/product|in-kernel|CG)|^4 GPUs" $O/r02_bench_4gpu.log | cut -c1-260 | head
