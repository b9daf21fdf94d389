#!/bin/bash
# GPU call 19: refined Cholesky kernel selection — full suite, C5 default, C4 defaults at k = 16, 32, 64, 128.
set -u
mkdir -p gpurun_out
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_gpu.log
echo "== C5 default"; timeout 400 python bench.py --m 5000000 --n 500000 --density 5e-4 --k 128 --L1 0.01 --L2 0.01 --steps 8 --warmup 3 --no-e2e --no-cpu --no-cd > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_c5_n1.json')); print('C5', round(d['ms_per_step'],2), d['value'], {k:round(v,2) for k,v in d['roofline']['sections_ms_per_step'].items()})"
for k in 16 32 64 128; do timeout 200 python tools/cd_explore.py --solver 1 --k $k --steps 8 --warmup 2 --variants chol_default --out gpurun_out/chol_default_k$k.jsonl 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'variant' in d:
        s=d['sections_ms_per_iter']; print(d['variant'], 'k', d['k'], round(d['ms_per_iter'],3), 'H', round(s['fused_rhs_nnls_H'],3), 'W', round(s['fused_rhs_nnls_W'],3))
"; done
