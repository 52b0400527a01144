#!/bin/bash
# Round 2, GPU session 5 (1 GPU): full GPU suite with the new kernels + device conversion default; sweep v3, BSR, bench.
O=gpurun_out; mkdir -p $O
D=lis_b200/_lib/drivers
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu_s5.txt 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu_s5.txt ); tail -4 $O/r02_pytest_gpu_s5.txt
for c in 8 4; do
  echo "== LIS_B200_SWEEP_CTAS=$c"
  LIS_B200_SWEEP_CTAS=$c $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 2000 2>&1 | grep -E "number of iterations|CG:   linear solver" | head -3
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_sell -c 2 -o $O/r02_sweep_v3 -f \
    $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 3 > $O/r02_ncu_sweep_v3.log 2>&1; echo "ncu sweep rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bsr_tile -c 2 -o $O/r02_bsr_v4 -f \
    $D/spmvtest3 256 256 256 3 7 > $O/r02_ncu_bsr_v4.log 2>&1; echo "ncu bsr rc=$?"
timeout 900 python profiles/run_configs.py su --size 10000000 --threads 16 2>&1 | tail -1 | cut -c1-700
timeout 900 python profiles/run_configs.py su --size 10000000 2>&1 | tail -1 | cut -c1-700
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02_bench_1gpu_c.json 2> $O/r02_bench_1gpu_c.log; echo "bench rc=$?"
grep -E "convert|e2e|CG" $O/r02_bench_1gpu_c.log | cut -c1-300
