#!/bin/bash
# Round 2, GPU session 1 (one B200): full GPU test suite, bench line, launch list, solver table,
# ncu --set full captures of the kernels VERDICT r01 names (SSOR sweep, JAD, BSR).
# usage (from the repo root on the GPU box):  bash profiles/r02_session1.sh
O=gpurun_out
mkdir -p $O
D=lis_b200/_lib/drivers
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu.txt )
tail -5 $O/r02_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.log; echo "bench rc=$?"
tail -c 1500 $O/r02_bench_1gpu.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file $O/r02_launches.csv python bench.py --steps 3 --warmup 3 --cg-iters 4 --no-cpu-baseline > /dev/null 2>&1; echo "launch list rc=$?"
timeout 900 python profiles/run_solvers.py > $O/r02_solvers.txt 2>&1; echo "solvers rc=$?"; tail -30 $O/r02_solvers.txt
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ssor_syncfree -c 4 -o $O/r02_ssor -f \
    $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 3 > $O/r02_ncu_ssor.log 2>&1; echo "ncu ssor rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jad_ -c 2 -o $O/r02_jad -f \
    $D/spmvtest3 256 256 256 3 6 > $O/r02_ncu_jad.log 2>&1; echo "ncu jad rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bsr -c 2 -o $O/r02_bsr -f \
    $D/spmvtest3 256 256 256 3 7 > $O/r02_ncu_bsr.log 2>&1; echo "ncu bsr rc=$?"
ls -la $O
