#!/bin/bash
# Round 2, GPU call 19 (2 GPUs): sharded-vs-one-GPU check with the multicast fall-back case (failed set-up -> NCCL loop
# on VMM-backed factors).
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r02s_multigpu_check_n2.txt 2>&1; echo "rc=$?"; grep -c "bit-identical=True" gpurun_out/r02s_multigpu_check_n2.txt; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r02s_multigpu_check_n2.txt | grep "mcfail\|OK\|Error\|error\|assert" | tail -8
