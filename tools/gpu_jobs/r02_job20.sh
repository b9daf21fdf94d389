#!/bin/bash
# Round 2, GPU call 20 (2 GPUs): RCPPML_NUM_GPUS behind the cross-validation entry and the masked extension entry
# (in-process multi-GPU test, all replication modes).
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "in_process_multi_gpu" 2>&1 | tail -12
