#!/bin/bash
# Round 2, GPU call 21 (1 GPU): final validation of the shipped tree — GPU suite (with the C4-shaped parity test), smoke,
# default bench line.
set -u
mkdir -p gpurun_out
echo "== pytest full"; timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r02t_pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/r02t_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (default flags)"; timeout 900 python bench.py > gpurun_out/r02t_bench_n1.json 2> gpurun_out/r02t_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r02t_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02t_bench_n1.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'],4), d['value'], d['steps'], d['warmup'], {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
print('e2e', d['e2e']['seconds_total'], d['e2e']['value'], 'cd', d['solver_mode_0']['ms_per_step'], d['solver_mode_0'].get('vs_cpu_reference'))
print('parity', d['parity']['ok'], d['parity']['rel_err'], d['parity']['matrix'])
print('roofline', round(d['roofline']['frac'],3), round(d['roofline']['frac_dram'],3), d['roofline']['frac_l2'], d['roofline']['bound'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
