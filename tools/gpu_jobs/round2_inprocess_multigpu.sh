#!/bin/bash
# Round 2, first multi-GPU call: the in-process multi-GPU path behind the reference entry point (RCPPML_NUM_GPUS,
# abi_reference.cu fit_in_process_multi_gpu) and the block-wise factor I/O (rcppml_b200_set/get_factor_blocks_f32) were
# written at the end of round 1 with no GPU minutes left. RCPPML_B200_TEST_ROUND2=1 opens their tests; once green, drop the gates.
#   gpurun --gpus 2 --timeout 600 -- bash tools/gpu_jobs/round2_inprocess_multigpu.sh
set -u
mkdir -p gpurun_out
echo "== parity (bit-identical to one GPU, small cases)"
RCPPML_B200_TEST_ROUND2=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k 'in_process_multi_gpu or factor_blocks or multi_gpu_matches' 2>&1 | tail -5
RCPPML_B200_TEST_ROUND2=1 timeout 300 python -m pytest tests/test_nmf_api.py -q -x -m gpu -k multiple_initialisations 2>&1 | tail -3
echo "== C4 through the reference ABI, RCPPML_NUM_GPUS = 1, 2"
timeout 400 python tools/inprocess_multigpu_probe.py --gpus 1,2 --out gpurun_out/inprocess_multigpu.json 2>&1 | tail -6
