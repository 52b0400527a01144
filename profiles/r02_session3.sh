#!/bin/bash
# Round 2, GPU session 3: sweep v2.1 timing + ncu, JAD/BSR kernels at 256^3/512^3 + ncu.
O=gpurun_out; mkdir -p $O
D=lis_b200/_lib/drivers
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
timeout 900 python -m pytest tests -m gpu -x -q -k "ssor or sweep or ilu or psolve or bsr or jad or spmv or smoke" 2>&1 | tail -4
for c in 6 3; do
  echo "== LIS_B200_SWEEP_CTAS=$c"
  LIS_B200_SWEEP_CTAS=$c $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 2000 2>&1 | grep -E "iterations|linear solver" | head -3
done
echo "== spmvtest3 512^3: CSR ELL JAD BSR (driver's own MFLOPS)"
for f in 1 5 6 7; do $D/spmvtest3 512 512 512 20 $f 2>&1 | grep -E "MFLOPS|matrix|storage" | head -3; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_sell -c 2 -o $O/r02_sweep_v21 -f \
    $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 3 > $O/r02_ncu_sweep_v21.log 2>&1; echo "ncu sweep rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jad_ -c 2 -o $O/r02_jad_v2 -f \
    $D/spmvtest3 256 256 256 3 6 > $O/r02_ncu_jad_v2.log 2>&1; echo "ncu jad rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bsr_tile -c 2 -o $O/r02_bsr_v2 -f \
    $D/spmvtest3 256 256 256 3 7 > $O/r02_ncu_bsr_v2.log 2>&1; echo "ncu bsr rc=$?"
