#!/bin/bash
# Round 2, first multi-GPU call: the in-process multi-GPU path behind the reference entry point (RCPPML_NUM_GPUS,
# abi_reference.cu fit_in_process_multi_gpu) was written at the end of round 1 with no GPU minutes left.
#   gpurun --gpus 2 --timeout 600 -- bash tools/gpu_jobs/round2_inprocess_multigpu.sh
set -u
mkdir -p gpurun_out
echo "== parity (bit-identical to one GPU, small cases)"
RCPPML_B200_TEST_INPROCESS_MULTIGPU=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k in_process_multi_gpu 2>&1 | tail -5
echo "== C4 through the reference ABI, RCPPML_NUM_GPUS = 1, 2"
timeout 400 python tools/inprocess_multigpu_probe.py --gpus 1,2 --out gpurun_out/inprocess_multigpu.json 2>&1 | tail -6
