#!/bin/bash
# GPU call 3: CD kernel with tolerance early-out + one-thread-per-column geometries (k <= 32).
set -u
mkdir -p gpurun_out
echo "== pytest (cd)"; timeout 600 python -m pytest tests -m gpu -q -x -k "cd_kernel or full_fit or half_steps or panel or pbmc3k or movielens" > gpurun_out/pytest_cd.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_cd.log
for k in 64 32 16 128; do timeout 300 python tools/cd_explore.py --k $k --variants v2_geom301,v2_geom102,v2_geom701,v2_geom302,v2_geom702,v2_geom304,v2_geom704,v2_geom308,v1_nv1 --out gpurun_out/cd_explore_v4_k$k.jsonl > gpurun_out/cd_explore_v4_k$k.log 2>&1; echo "k=$k rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/cd_explore_v4*.jsonl')):
    for l in open(f):
        d=json.loads(l)
        if 'variant' in d:
            s=d['sections_ms_per_iter']
            print("  %-12s k=%-3d %8.2f ms/iter  H %.2f  W %.2f  sweeps %d %s"%(d['variant'],d['k'],d['ms_per_iter'],s['fused_rhs_nnls_H'],s['fused_rhs_nnls_W'],d['cd_sweeps_total'],d['digest']))
        else: print(f, d)
PY
