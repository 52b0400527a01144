#!/bin/bash
# Round 2, GPU session 11 (2 GPUs): why is the product 2.49 ms per GPU in multi-rank runs when it is 2.09 ms at N=1?
# A: every GPU visible, peer access never enabled by the library (LIS_B200_P2P=0), NCCL default
# B: as A, and NCCL without its P2P transport (NCCL_P2P_DISABLE=1: shared-memory transport, no peer mappings at all)
# C: every GPU visible, library default (probe enables peer access, in-kernel exchange available)
# D: each rank sees only its GPU (round 1's setting)
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='rank 0.*(ms/product|in-kernel)|^2 GPUs|lis_b200:'
echo "== A: P2P off in the library"; LIS_B200_VERBOSE=1 LIS_B200_P2P=0 timeout 600 $TR --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 >/dev/null | grep -E "$F" | cut -c1-230
echo "== B: and NCCL_P2P_DISABLE=1"; LIS_B200_VERBOSE=1 LIS_B200_P2P=0 NCCL_P2P_DISABLE=1 timeout 600 $TR --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 >/dev/null | grep -E "$F" | cut -c1-230
echo "== C: default"; LIS_B200_VERBOSE=1 timeout 600 $TR --master-port 29703 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 >/dev/null | grep -E "$F" | cut -c1-230
echo "== D: narrowed"; LIS_B200_VERBOSE=1 LIS_B200_NARROW=1 timeout 600 $TR --master-port 29704 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 >/dev/null | grep -E "$F" | cut -c1-230
