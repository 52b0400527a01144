#!/bin/bash
# Round 2, GPU session 18 (1 GPU): 27-point CSR SpMV roofline (spmvtest3b 256^3), final full GPU suite.
O=gpurun_out; mkdir -p $O
D=lis_b200/_lib/drivers
export LD_LIBRARY_PATH=$PWD/lis_b200/_lib:$LD_LIBRARY_PATH
timeout 300 ncu --set full --clock-control none -k regex:csr_ -c 1 -s 3 -o $O/r02_csr_27pt_256 -f \
    $D/spmvtest3b 256 256 256 6 1 > $O/r02_ncu_csr_27pt.log 2>&1; echo "ncu 27pt rc=$?"; grep -E "MFLOPS|nonzero" $O/r02_ncu_csr_27pt.log | head -3
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu_final2.txt 2>&1; echo "pytest rc=$?" >> $O/r02_pytest_gpu_final2.txt ); tail -4 $O/r02_pytest_gpu_final2.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
