#!/bin/bash
# Round 2, GPU session 20 (2 GPUs): row-partitioned GMRES(30) with fused Gram-Schmidt links vs the separate dot/axpy calls.
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mode in chain fused; do
  echo "== LIS_B200_MGS=$mode"
  LIS_B200_MGS=$mode timeout 300 $TR --master-port 2998$((RANDOM % 10)) profiles/run_configs.py gm27 --size 256 --maxiter 300 --out $O/r02_configs_n2b.jsonl 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('what', 'n_ranks', 'iters', 'ms_per_iter', 'gflops', 'relres')})"
done
