#!/bin/bash
# GPU call 2: CD kernel v3 (rolled blocks, x in shared memory, 3 CTAs/SM) — parity + geometry sweep + ncu.
set -u
mkdir -p gpurun_out
echo "== pytest (cd geometries + full fits)"; timeout 600 python -m pytest tests -m gpu -q -x -k "cd_kernel or full_fit or half_steps or panel" > gpurun_out/pytest_cd.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_cd.log
echo "== cd_explore v3 (3 CTAs/SM)"; timeout 300 python tools/cd_explore.py --out gpurun_out/cd_explore_v3.jsonl > gpurun_out/cd_explore_v3.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/cd_explore_v3.log | tail -8
echo "== cd_explore v3 (2 CTAs/SM build)"; RCPPML_B200_LIB=$PWD/rcppml_b200/lib/RcppML_gpu_cd2.so timeout 300 python tools/cd_explore.py --variants v2_geom702,v2_geom304 --out gpurun_out/cd_explore_v3_2cta.jsonl > gpurun_out/cd_explore_v3_2cta.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/cd_explore_v3_2cta.log | tail -4
echo "== cd_explore k=32 / k=128 / k=16"; for k in 16 32 128; do timeout 300 python tools/cd_explore.py --k $k --out gpurun_out/cd_explore_v3_k$k.jsonl > gpurun_out/cd_explore_v3_k$k.log 2>&1; echo "k=$k rc=$?"; cut -c1-200 gpurun_out/cd_explore_v3_k$k.log | tail -6; done
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== ncu full (cd kernel)"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:cd_half_step -s 2 -c 2 -o gpurun_out/prof_cd_v3 -f python bench.py --solver cd --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_cd_v3.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -40
