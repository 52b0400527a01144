#!/bin/bash
# Round 2, GPU session 2: sweep kernel v2 (parity tests, timing at 1..8 CTAs/SM, ncu), new bench line.
O=gpurun_out; mkdir -p $O
D=lis_b200/_lib/drivers
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
timeout 900 python -m pytest tests -m gpu -x -q -k "ssor or sweep or ilu or psolve or solver or smoke" 2>&1 | tail -4
for c in 8 4 2 1; do
  echo "== LIS_B200_SWEEP_CTAS=$c"
  LIS_B200_SWEEP_CTAS=$c $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 2000 2>&1 | grep -E "iteration|elapsed|itr|precon|matvec|CG:" | head -12
done
echo "== banded 1M x 70, BiCGSTAB+SSOR"
timeout 600 python profiles/run_solvers.py 128 1000000 64 --noref 2>&1 | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_sell -c 2 -o $O/r02_sweep_v2 -f \
    $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 3 > $O/r02_ncu_sweep_v2.log 2>&1; echo "ncu sweep rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r02_bench_1gpu_b.json 2> $O/r02_bench_1gpu_b.log; echo "bench rc=$?"
tail -25 $O/r02_bench_1gpu_b.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.log; echo "ref rc=$?"; cat $O/r02_bench_ref.json
