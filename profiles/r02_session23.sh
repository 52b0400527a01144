#!/bin/bash
# Round 2, GPU session 23 (1 GPU, the last seconds of the budget): set-up phases of CG + SSOR at 256^3 on the box's host cores.
O=gpurun_out; mkdir -p $O
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
( nproc; LIS_B200_TRACE_SETUP=1 timeout 12 lis_b200/_lib/drivers/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 2000 2>&1 | grep -E "setup:|number of iterations|linear solver  |elapsed" ) | tee $O/r02_setup_trace_256.txt
