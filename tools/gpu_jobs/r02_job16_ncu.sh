#!/bin/bash
# Round 2, GPU call 16 (1 GPU): ncu evidence of the FINAL build for the headline configuration (C4, Cholesky): launch list
# of the default bench command + `--set full` captures of the H half-step, the W half-step, the Gram and the LLT kernel.
set -u
mkdir -p gpurun_out
export RCPPML_B200_GRAPH=0
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-cd --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02p_halfstep_launches.csv $B > gpurun_out/r02p_ncu_list.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:half_step_kernel -s 2 -c 2 -o gpurun_out/r02p_prof_halfstep -f $B > gpurun_out/r02p_ncu_halfstep.log 2>&1; echo "halfstep rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"normalize_gram_mma_kernel|prepare_solver_kernel" -s 4 -c 4 -o gpurun_out/r02p_prof_dense -f $B > gpurun_out/r02p_ncu_dense.log 2>&1; echo "dense rc=$?"
ls -la gpurun_out/r02p*.ncu-rep
