#!/bin/bash
# Round 2, GPU session 19 (1 GPU): launch lists (time + DRAM bytes per launch) of a GMRES(30) and a BiCGSTAB+SSOR solve:
# evidence for mgs_step_kernel, bicgstab_update_kernel, dot2, the product-tile CSR kernel and the sweep on long rows.
O=gpurun_out; mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file $O/r02_launches_gmres.csv python profiles/run_configs.py gm27 --size 256 --maxiter 40 > /dev/null 2>&1; echo "gmres rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file $O/r02_launches_bicgstab.csv python profiles/run_configs.py su --size 4000000 --threads 16 --maxiter 6 > /dev/null 2>&1; echo "bicgstab rc=$?"
