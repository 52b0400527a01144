#!/bin/bash
# Round 2, GPU session 22 (1 GPU, last seconds of the budget): the sweep's wait address moved a = 0/1/2 levels ahead
# (LIS_B200_SWEEP_AHEAD, host/lis_precon.c): CG + SSOR at 256^3, "linear solver" seconds for 363 iterations; then the
# sweep parity tests with a = 1.
O=gpurun_out; mkdir -p $O
D=lis_b200/_lib/drivers
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
for a in 0 1 2; do
  echo "== LIS_B200_SWEEP_AHEAD=$a" | tee -a $O/r02_sweep_ahead.txt
  LIS_B200_SWEEP_AHEAD=$a timeout 40 $D/test3 256 256 256 1 /dev/null /dev/null -i cg -p ssor -maxiter 2000 2>&1 | grep -E "number of iterations|CG:   linear solver|relative residual" | tee -a $O/r02_sweep_ahead.txt
done
LIS_B200_SWEEP_AHEAD=1 timeout 45 python -m pytest tests/test_gpu_parity.py tests/test_z1_gpu_parity2.py -m gpu -x -q -k "ssor or psolve or ilu" 2>&1 | tail -3 | tee -a $O/r02_sweep_ahead.txt
