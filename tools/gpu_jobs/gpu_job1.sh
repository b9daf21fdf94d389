#!/bin/bash
# GPU call 1 of this session: CD kernel exploration + parity, full GPU test-suite, default bench, ncu captures.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
echo "== cd_explore (CD)"; timeout 300 python tools/cd_explore.py --out gpurun_out/cd_explore.jsonl > gpurun_out/cd_explore.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/cd_explore.log | cut -c1-400
echo "== cd_explore (chol variants)"; timeout 200 python tools/cd_explore.py --solver 1 --steps 10 --warmup 3 --out gpurun_out/chol_explore.jsonl > gpurun_out/chol_explore.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/chol_explore.log | cut -c1-300
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
echo "== ncu launch list (cd)"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cd.csv python bench.py --solver cd --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_list_cd.log 2>&1; echo "rc=$?"
echo "== ncu full (cd kernel)"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:cd_half_step -s 2 -c 2 -o gpurun_out/prof_cd -f python bench.py --solver cd --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_cd.log 2>&1; echo "rc=$?"
ls -la gpurun_out
