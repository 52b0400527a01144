#!/bin/bash
# Round 2, GPU session 10 (1 GPU): full GPU suite on the current tree; sweep v5 timing + ncu; csr_tma ncu --set full at 512^3
# (roofline.traffic); bench (conversion times incl. DIA); launch list.
O=gpurun_out; mkdir -p $O
D=lis_b200/_lib/drivers
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu_s10.txt 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu_s10.txt ); tail -4 $O/r02_pytest_gpu_s10.txt
for c in 0 6; do
  echo "== CG+SSOR 256^3, LIS_B200_SWEEP_CTAS=$c"
  LIS_B200_SWEEP_CTAS=$c $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 2000 2>&1 | grep -E "number of iterations|CG:   linear solver" | head -3
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_sell -c 2 -o $O/r02_sweep_v5 -f \
    $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 3 > $O/r02_ncu_sweep_v5.log 2>&1; echo "ncu sweep rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:csr_tma -c 1 -s 3 -o $O/r02_csr_tma_512 -f \
    $D/spmvtest3 512 512 512 6 1 > $O/r02_ncu_csr_tma_512.log 2>&1; echo "ncu csr_tma rc=$?"
timeout 900 python profiles/run_configs.py su --size 10000000 --threads 16 2>&1 | grep '^{' | cut -c1-500
timeout 900 python profiles/run_configs.py su --size 10000000 2>&1 | grep '^{' | cut -c1-500
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02_bench_1gpu_d.json 2> $O/r02_bench_1gpu_d.log; echo "bench rc=$?"
grep -E "convert|e2e|CG|^CSR|^ELL|^DIA" $O/r02_bench_1gpu_d.log | cut -c1-260
