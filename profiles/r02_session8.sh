#!/bin/bash
# Round 2, GPU session 8 (1 GPU): sweep v4 ncu + timing, N=1 points of the weak-scaling tables (configs 3/4/5), launch list.
O=gpurun_out; mkdir -p $O
D=lis_b200/_lib/drivers
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_sell -c 2 -o $O/r02_sweep_v4 -f \
    $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 3 > $O/r02_ncu_sweep_v4.log 2>&1; echo "ncu sweep rc=$?"
rm -f $O/r02_configs_n1b.jsonl
for cfg in "gm27 --size 768 --slab 96 --opts=-maxiter_600" "cg7 --size 512" "su --size 1250000 --threads 2" "su --size 10000000 --threads 16" "su --size 10000000"; do
  timeout 900 python profiles/run_configs.py ${cfg//_/ } --out $O/r02_configs_n1b.jsonl 2>&1 | grep '^{' | cut -c1-700
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
    --log-file $O/r02_launches_b.csv python bench.py --steps 3 --warmup 3 --cg-iters 4 --no-cpu-baseline --no-cg-converge > /dev/null 2>&1; echo "launch list rc=$?"
