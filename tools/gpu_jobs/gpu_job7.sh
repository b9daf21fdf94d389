#!/bin/bash
# GPU call 7: final policy — full GPU suite, small configs, C4 both solvers, bench, ncu launch list + full captures.
set -u
mkdir -p gpurun_out
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== small configs"; timeout 600 python tools/small_configs.py --iters 100 --out gpurun_out/small_configs.jsonl > gpurun_out/small_configs.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/small_configs.jsonl'):
    d=json.loads(l); print("  %-38s solver %d  graph %.3f  plain %.3f"%(d['config'],d['solver_mode'],d['graph']['device_ms_per_iter'],d['plain']['device_ms_per_iter']))
PY
echo "== C4 chol"; timeout 300 python tools/cd_explore.py --solver 1 --steps 10 --warmup 3 --variants chol_default,chol_untiled --out gpurun_out/chol_c4.jsonl > /dev/null 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/chol_c4.jsonl'):
    d=json.loads(l)
    if 'variant' in d:
        s=d['sections_ms_per_iter']; print("  %-16s %8.3f ms/iter  H %.3f  W %.3f"%(d['variant'],d['ms_per_iter'],s['fused_rhs_nnls_H'],s['fused_rhs_nnls_W']))
PY
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 500 ncu --set full --clock-control none --import-source on -k regex:'half_step|normalize_gram' -s 6 -c 8 -o gpurun_out/prof_halfstep -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
