#!/bin/bash
# GPU call 17: C5 (5M x 500K, k = 128, L1 = L2 = 0.01) on one GPU — regression check of the default policy, and the
# tiled kernel forced for comparison.
set -u
mkdir -p gpurun_out
echo "== C5 default"; timeout 400 python bench.py --m 5000000 --n 500000 --density 5e-4 --k 128 --L1 0.01 --L2 0.01 --steps 8 --warmup 3 --no-e2e --no-cpu --no-cd > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; echo "rc=$?"; tail -1 gpurun_out/bench_c5_n1.err
echo "== C5 tiled forced"; RCPPML_B200_TILED=2 timeout 400 python bench.py --m 5000000 --n 500000 --density 5e-4 --k 128 --L1 0.01 --L2 0.01 --steps 8 --warmup 3 --no-e2e --no-cpu --no-cd > gpurun_out/bench_c5_n1_tiled.json 2> gpurun_out/bench_c5_n1_tiled.err; echo "rc=$?"; tail -1 gpurun_out/bench_c5_n1_tiled.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_c5_n1.json','gpurun_out/bench_c5_n1_tiled.json'):
    try:
        d=json.load(open(f)); print(f, round(d['ms_per_step'],2), d['value'], {k:round(v,2) for k,v in d['roofline']['sections_ms_per_step'].items()})
    except Exception as e: print(f, 'failed', e)
PY
