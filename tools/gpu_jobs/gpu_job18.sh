#!/bin/bash
# GPU call 18: Cholesky kernel selection at other ranks (C4 shape, k = 16, 32, 128): one-geometry vs tiled W vs tiled both.
set -u
mkdir -p gpurun_out
for k in 16 32 128; do timeout 200 python tools/cd_explore.py --solver 1 --k $k --steps 8 --warmup 2 --variants chol_untiled,chol_tiled_both --out gpurun_out/chol_k$k.jsonl > /dev/null 2>&1; done
python - <<'PY'
import json
for k in (16,32,128):
    for l in open(f'gpurun_out/chol_k{k}.jsonl'):
        d=json.loads(l)
        if 'variant' in d:
            s=d['sections_ms_per_iter']; print("  %-18s k=%-3d %8.3f ms/iter  H %.3f  W %.3f"%(d['variant'],d['k'],d['ms_per_iter'],s['fused_rhs_nnls_H'],s['fused_rhs_nnls_W']))
        else: print(d)
PY
