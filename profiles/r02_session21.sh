#!/bin/bash
# Round 2, GPU session 21 (1 GPU): the GPU parity suite on the final tree (after per-format -scale, the reference's
# order among equal keys in lis_sort_id, the strict sortedness check of the DIA conversion, host views of device-only vectors).
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/r02_pytest_gpu_final4.txt
echo "pytest rc=${PIPESTATUS[0]}" | tee -a $O/r02_pytest_gpu_final4.txt
