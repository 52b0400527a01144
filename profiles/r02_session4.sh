#!/bin/bash
# Round 2, GPU session 4 (1 GPU): sweep v2.2 + BSR v3 timing/ncu; configs 3/4/5 at N=1.
O=gpurun_out; mkdir -p $O
D=lis_b200/_lib/drivers
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
timeout 900 python -m pytest tests -m gpu -x -q -k "ssor or sweep or ilu or psolve or bsr or smoke" 2>&1 | tail -3
for c in 6 3; do
  echo "== LIS_B200_SWEEP_CTAS=$c"
  LIS_B200_SWEEP_CTAS=$c $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 2000 2>&1 | grep -E "number of iterations|CG:   linear solver|CG:   precond" | head -3
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_sell -c 2 -o $O/r02_sweep_v22 -f \
    $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 3 > $O/r02_ncu_sweep_v22.log 2>&1; echo "ncu sweep rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bsr_tile -c 2 -o $O/r02_bsr_v3 -f \
    $D/spmvtest3 256 256 256 3 7 > $O/r02_ncu_bsr_v3.log 2>&1; echo "ncu bsr rc=$?"
rm -f $O/r02_configs_n1.jsonl
for cfg in "cg7 --size 256" "su --size 1000000" "su --size 1000000 --threads 16" "su --size 10000000" "su --size 10000000 --threads 16" "gm27 --size 128" "gm27 --size 256"; do
  timeout 900 python profiles/run_configs.py $cfg --out $O/r02_configs_n1.jsonl 2>&1 | tail -1 | cut -c1-900
done
for cfg in "su --size 1000000" "gm27 --size 128"; do
  timeout 900 python profiles/run_configs.py $cfg --impl reference --out $O/r02_configs_n1.jsonl 2>&1 | tail -1 | cut -c1-900
done
