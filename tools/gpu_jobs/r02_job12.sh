#!/bin/bash
# Round 2, GPU call 12 (1 GPU): bench N = 1 after the multimem asm lost its "memory" clobber and the Gram kernel's
# replicating variant became a separate instantiation (the common kernels must be back at their round-2 times).
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-cd > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r02l_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02l_bench_n1.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'],4), d['value'], {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
print('e2e', d['e2e']['seconds_total'], d['e2e']['phases'])
print('parity', d['parity']['ok'], d['parity']['rel_err'], d['parity']['matrix'])
print('clocks', d['clocks'])
PY
