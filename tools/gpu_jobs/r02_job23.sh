#!/bin/bash
# Round 2, GPU call 23 (1 GPU): the pre-stored-transpose entry + the whole GPU suite once more.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02u_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02u_pytest_gpu.log
