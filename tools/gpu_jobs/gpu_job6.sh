#!/bin/bash
# GPU call 6: tiled policy check (small configs x kernel variants), C4 CD/Cholesky defaults, ncu of the tiled W half-step.
set -u
mkdir -p gpurun_out
echo "== pytest tiled"; timeout 900 python -m pytest tests -m gpu -q -x -k "tiled or full_fit or pbmc3k or movielens or aml" > gpurun_out/pytest_tiled.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_tiled.log
echo "== small configs x variants"; timeout 600 python tools/small_configs.py --kernel-variants --iters 100 --out gpurun_out/small_variants.jsonl > gpurun_out/small_variants.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/small_variants.jsonl'):
    d=json.loads(l); print("  %-38s solver %d  default %.3f  tiled_always %.3f  untiled %.3f  wide_cd %s"%(d['config'],d['solver_mode'],d['graph']['device_ms_per_iter'],d['tiled_always'],d['untiled'],d.get('untiled_wide_cd')))
PY
echo "== C4 defaults"; timeout 300 python tools/cd_explore.py --k 64 --variants default,untiled --out gpurun_out/cd_c4.jsonl 2>&1 | cut -c1-120 | head -3
timeout 300 python tools/cd_explore.py --solver 1 --steps 10 --warmup 3 --variants chol_default,chol_untiled --out gpurun_out/chol_c4.jsonl 2>&1 | cut -c1-120 | head -3
python - <<'PY'
import json
for f in ('gpurun_out/cd_c4.jsonl','gpurun_out/chol_c4.jsonl'):
    for l in open(f):
        d=json.loads(l)
        if 'variant' in d:
            s=d['sections_ms_per_iter']; print("  %-16s %8.3f ms/iter  H %.3f  W %.3f"%(d['variant'],d['ms_per_iter'],s['fused_rhs_nnls_H'],s['fused_rhs_nnls_W']))
PY
echo "== ncu full (tiled chol)"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:tiled_half_step -s 2 -c 1 -o gpurun_out/prof_tiled_chol -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-cd > gpurun_out/ncu_tiled_chol.log 2>&1; echo "rc=$?"
