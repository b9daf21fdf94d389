#!/bin/bash
# GPU call 9: one-barrier LLT, overlapped upload in the reference entry points — tests + bench.
set -u
mkdir -p gpurun_out
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['ms_per_step'], d['value'], d['clocks'])
print({k:round(v,4) for k,v in d['roofline']['sections_ms_per_step'].items()})
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['mean_launch_ms'], d['roofline']['traffic'])
print('cd', d['solver_mode_0']['ms_per_step'])
print('e2e', d['e2e']['value'], d['e2e']['seconds_total'], d['e2e']['phases'])
print('cpu', d['cpu_baseline'])
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-600
