#!/bin/bash
# Round 2, GPU call 1 (1 GPU): full GPU suite incl. the single-GPU tests gated in round 1, the new bench line (parity
# object, frac_dram, CD ratio), and the per-rank kernel-selection sweep (tools/rank_shape_sweep.py).
set -u
mkdir -p gpurun_out
echo "== pytest full"; RCPPML_B200_TEST_ROUND2=1 timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02a_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 900 python bench.py > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r02a_bench_n1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02a_bench_n1.json'))
    print(d['ms_per_step'], d['value'], d['clocks'], d['gpu_launches'])
    print('roofline', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), d['roofline']['frac_dram'])
    print('sections', {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
    print('cd', d['solver_mode_0']['ms_per_step'], d['solver_mode_0'].get('vs_cpu_reference'))
    print('e2e', d['e2e']['value'], d['e2e']['seconds_total'], d['e2e']['phases'])
    print('parity', d['parity'])
    print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'))
except Exception as ex:
    print('bench parse failed', ex)
PY
echo "== rank shapes"; timeout 900 python tools/rank_shape_sweep.py --ns 2,4,8 --out gpurun_out/r02a_rank_shapes.jsonl 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    if 'variant' in d: print(d['N'], d['half_step'], d['variant'], round(d['half_step_ms'],4), round(d['iteration_ms'],4), d['checksum'])
    else: print(d)
"
