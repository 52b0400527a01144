#!/bin/bash
# Round 2, GPU session 6 (2 GPUs, every GPU visible to every rank): managed-memory probe, in-kernel halo exchange
# (multi-rank parity test, bench A/B), configs at N=2.
O=gpurun_out; mkdir -p $O
export LIS_B200_VERBOSE=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29501 profiles/managed_probe.py 2>&1 | grep -E "rank|max_map" | head -12
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q 2>&1 | tail -6
timeout 900 $TR --master-port 29502 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r02_bench_2gpu_p2p.json 2> $O/r02_bench_2gpu_p2p.log; echo "bench rc=$?"
grep -E "ms/product|in-kernel|CG|e2e|halo overlap|lis_b200:" $O/r02_bench_2gpu_p2p.log | grep -v "rank 1" | cut -c1-250
cut -c1-400 $O/r02_bench_2gpu_p2p.json
rm -f $O/r02_configs_n2.jsonl
for cfg in "cg7 --size 256" "gm27 --size 256 --opts -maxiter_300" "su --size 10000000 --threads 8"; do
  timeout 900 $TR --master-port 29503 profiles/run_configs.py ${cfg//_/ } --out $O/r02_configs_n2.jsonl 2>&1 | grep '^{' | cut -c1-600
done
