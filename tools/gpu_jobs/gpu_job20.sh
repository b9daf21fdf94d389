#!/bin/bash
# GPU call 20: last validation — full GPU suite, smoke, default bench, k = 32 default policy.
set -u
mkdir -p gpurun_out
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/cd_explore.py --solver 1 --k 32 --steps 8 --warmup 2 --variants chol_default --out gpurun_out/chol_default_k32.jsonl 2>&1 | cut -c1-130 | head -1
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['ms_per_step'], d['value'], d['clocks'], d['gpu_launches'])
print('roofline', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3))
print('cd', d['solver_mode_0']['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['seconds_total'])
PY
