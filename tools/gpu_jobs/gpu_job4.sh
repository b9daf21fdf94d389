#!/bin/bash
# GPU call 4: full GPU suite (graph test, real datasets), small configs with/without graphs, CD default geometry, bench.
set -u
mkdir -p gpurun_out
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu.log
echo "== small configs"; timeout 600 python tools/small_configs.py --out gpurun_out/small_configs.jsonl > gpurun_out/small_configs.log 2>&1; echo "rc=$?"; cut -c1-420 gpurun_out/small_configs.log | tail -8
echo "== cd default geometry"; for k in 16 32 64; do timeout 300 python tools/cd_explore.py --k $k --variants default --out gpurun_out/cd_default_k$k.jsonl 2>&1 | cut -c1-300 | head -1; done
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
