#!/bin/bash
# GPU call 5: tiled kernel (wide gather -> smem tile -> narrow solve): parity, then timings.
set -u
mkdir -p gpurun_out
echo "== pytest tiled"; timeout 900 python -m pytest tests -m gpu -q -x -k "tiled or cd_kernel or full_fit or half_steps or panel or pbmc3k or movielens or aml or graph" > gpurun_out/pytest_tiled.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_tiled.log
echo "== chol variants"; timeout 300 python tools/cd_explore.py --solver 1 --steps 10 --warmup 3 --out gpurun_out/chol_tiled.jsonl > gpurun_out/chol_tiled.log 2>&1; echo "rc=$?"
echo "== cd variants"; for k in 64 32 16; do timeout 300 python tools/cd_explore.py --k $k --variants default,untiled --out gpurun_out/cd_tiled_k$k.jsonl > gpurun_out/cd_tiled_k$k.log 2>&1; echo "k=$k rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/chol_tiled.jsonl'))+sorted(glob.glob('gpurun_out/cd_tiled*.jsonl')):
    for l in open(f):
        d=json.loads(l)
        if 'variant' in d:
            s=d['sections_ms_per_iter']
            print("  %-24s k=%-3d %8.3f ms/iter  H %.3f  W %.3f  sweeps %d %s"%(d['variant'],d['k'],d['ms_per_iter'],s['fused_rhs_nnls_H'],s['fused_rhs_nnls_W'],d['cd_sweeps_total'],d['digest']))
        else: print(f, d)
PY
echo "== small configs"; timeout 600 python tools/small_configs.py --out gpurun_out/small_configs.jsonl > gpurun_out/small_configs.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/small_configs.jsonl'):
    d=json.loads(l); print("  %-40s solver %d  graph %.3f ms  plain %.3f ms"%(d['config'],d['solver_mode'],d['graph']['device_ms_per_iter'],d['plain']['device_ms_per_iter']))
PY
