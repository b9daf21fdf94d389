#!/bin/bash
# GPU call 13: timings of the §8f rows at C4 (CV fit, predict, evaluate).
set -u
mkdir -p gpurun_out
timeout 900 python tools/next_rows_bench.py --out gpurun_out/next_rows.json 2>&1 | tail -12 | cut -c1-400
