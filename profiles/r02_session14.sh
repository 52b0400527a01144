#!/bin/bash
# Round 2, GPU session 14 (1 GPU): final state -- full GPU suite, config 5's slab at N=1, config 4 with the warp-per-row
# sweep kernel, final bench line + reference arm, smoke.
O=gpurun_out; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu_final.txt 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu_final.txt ); tail -4 $O/r02_pytest_gpu_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
rm -f $O/r02_configs_n1c.jsonl
for cfg in "gm27 --size 768 --slab 96 --maxiter 600" "su --size 10000000 --threads 16" "su --size 10000000" "su --size 1000000 --threads 16"; do
  timeout 900 python profiles/run_configs.py $cfg --out $O/r02_configs_n1c.jsonl 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print({k: d[k] for k in ('config', 'what', 'n_ranks', 'threads_or_blocks_per_rank', 'iters', 'status', 'ms_per_iter', 'gflops', 'relres')})"
done
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02_bench_1gpu_final.json 2> $O/r02_bench_1gpu_final.log; echo "bench rc=$?"
grep -E "convert|e2e|CG|^CSR|^ELL|^DIA" $O/r02_bench_1gpu_final.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02_bench_ref_final.json 2>/dev/null; cut -c1-300 $O/r02_bench_ref_final.json
