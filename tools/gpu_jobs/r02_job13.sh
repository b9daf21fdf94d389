#!/bin/bash
# Round 2, GPU call 13 (1 GPU): GPU suite + bench N = 1 with the unconditional local store (multicast as a trailing store).
set -u
mkdir -p gpurun_out
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02m_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r02m_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02m_bench_n1.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'],4), d['value'], {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
print('e2e', d['e2e']['seconds_total'], d['e2e']['phases'], 'cd', d['solver_mode_0']['ms_per_step'], d['solver_mode_0'].get('vs_cpu_reference'))
print('parity', d['parity']['ok'], d['parity']['rel_err'], d['parity']['matrix'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
