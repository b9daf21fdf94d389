#!/bin/bash
# GPU call 14: right-looking per-warp LLT, lower-triangle downdates, 12 warps/CTA in the CV / explicit-mask kernels.
set -u
mkdir -p gpurun_out
echo "== pytest (cv, masked)"; timeout 900 python -m pytest tests -m gpu -q -x -k "cv or masked" > gpurun_out/pytest_cv.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_cv.log
echo "== timings"; timeout 900 python tools/next_rows_bench.py --out gpurun_out/next_rows.json 2>&1 | head -2 | cut -c1-300
