#!/bin/bash
# Round 2, GPU call 9 (1 GPU): ncu evidence for the headline configuration (C4, Cholesky) with the round-2 kernels:
# launch list of the default bench command, `--set full` captures of the H half-step (half_step_kernel), the W half-step
# (tiled_half_step_kernel, 8-column batches), the DMMA Gram and the LLT kernel. Numbers printed under ncu are never
# bench values.
set -u
mkdir -p gpurun_out
export RCPPML_B200_GRAPH=0
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-cd --no-parity"
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02i_halfstep_launches.csv $B > gpurun_out/r02i_ncu_list.log 2>&1; echo "rc=$?"
echo "== full: solve kernels (iteration 1: H then W)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:half_step_kernel -s 2 -c 2 -o gpurun_out/r02i_prof_halfstep -f $B > gpurun_out/r02i_ncu_halfstep.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02i_ncu_halfstep.log
echo "== full: gram + LLT"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"normalize_gram_mma_kernel|prepare_solver_kernel" -s 4 -c 4 -o gpurun_out/r02i_prof_dense -f $B > gpurun_out/r02i_ncu_dense.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r02i_ncu_dense.log
ls -la gpurun_out/*.ncu-rep
unset RCPPML_B200_GRAPH
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02i_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench n1"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r02i_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02i_bench_n1.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'],4), d['value'], {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
print('e2e', d['e2e']['seconds_total'], d['e2e']['phases'], 'cd', d['solver_mode_0']['ms_per_step'], d['solver_mode_0'].get('vs_cpu_reference'))
print('parity', d['parity']['ok'], d['parity']['rel_err'], d['parity']['matrix'])
print('roofline', d['roofline']['frac'], d['roofline']['frac_dram'], d['roofline']['frac_l2'])
PY
