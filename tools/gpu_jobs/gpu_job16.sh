#!/bin/bash
# GPU call 16: final validation of the round — full GPU suite, smoke, default bench, ncu of the tiled CD kernel.
set -u
mkdir -p gpurun_out
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['ms_per_step'], d['value'], d['clocks'], d['gpu_launches'])
print('roofline', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],3), round(d['roofline']['mean_launch_ms'],4))
print('cd', d['solver_mode_0']['ms_per_step'])
print('e2e', d['e2e']['value'], d['e2e']['seconds_total'])
print('cpu', {k:v for k,v in d['cpu_baseline'].items() if k!='sample'})
PY
echo "== ncu full (tiled CD)"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:tiled_half_step -s 4 -c 2 -o gpurun_out/prof_tiled_cd -f python bench.py --solver cd --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_tiled_cd.log 2>&1; echo "rc=$?"
