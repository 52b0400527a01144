#!/bin/bash
# Round 2, GPU session 17 (1 GPU): BSR 2x2 with 256-bit loads.
O=gpurun_out; mkdir -p $O
D=lis_b200/_lib/drivers
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
timeout 600 python -m pytest tests -m gpu -x -q -k "bsr or format or conver or smoke" 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bsr_tile -c 2 -o $O/r02_bsr_v5 -f \
    $D/spmvtest3 256 256 256 3 7 > $O/r02_ncu_bsr_v5.log 2>&1; echo "ncu bsr rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cg-converge > $O/r02_bench_1gpu_e.json 2> $O/r02_bench_1gpu_e.log; echo "bench rc=$?"
grep -E "bsr|jad" $O/r02_bench_1gpu_e.log | cut -c1-200
