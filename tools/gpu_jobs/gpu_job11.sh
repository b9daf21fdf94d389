#!/bin/bash
# GPU call 11: software-pipelined gather rounds in the tiled kernel — parity + timings.
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x -k "tiled or full_fit or panel or pbmc3k or movielens" > gpurun_out/pytest_tiled.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_tiled.log
timeout 300 python tools/cd_explore.py --solver 1 --steps 10 --warmup 3 --variants chol_default,chol_untiled --out gpurun_out/chol_c4.jsonl > /dev/null 2>&1
timeout 300 python tools/cd_explore.py --k 64 --variants default,untiled --out gpurun_out/cd_c4.jsonl > /dev/null 2>&1
timeout 300 python tools/cd_explore.py --k 32 --variants default,untiled --out gpurun_out/cd_c4_k32.jsonl > /dev/null 2>&1
python - <<'PY'
import json
for f in ('gpurun_out/chol_c4.jsonl','gpurun_out/cd_c4.jsonl','gpurun_out/cd_c4_k32.jsonl'):
    for l in open(f):
        d=json.loads(l)
        if 'variant' in d:
            s=d['sections_ms_per_iter']; print("  %-16s k=%d %8.3f ms/iter  H %.3f  W %.3f"%(d['variant'],d['k'],d['ms_per_iter'],s['fused_rhs_nnls_H'],s['fused_rhs_nnls_W']))
        else: print(d)
PY
