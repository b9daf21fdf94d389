#!/bin/bash
# Round 2, GPU call 17 (1 GPU): the W half-step experiment VERDICT item 5 asks for, built: (b) one 768-thread CTA per SM
# (L / Lt once per SM) and (a)+(b) that layout with the hybrid register + cp.async shared-memory ring gather — parity
# test, then the C4 W half-step and its per-rank shapes, variant by variant.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "cta768 or tiled_kernel_is_bit" 2>&1 | tail -3
timeout 600 python tools/rank_shape_sweep.py --ns 1,8 --half-steps W --variants tiled8,tiled8_cta768,tiled8_cta768_hybrid,untiled --steps 8 --out gpurun_out/r02q_wstep_variants.jsonl 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    if 'variant' in d: print(d['N'], d['half_step'], d['variant'], round(d['half_step_ms'],4), round(d['iteration_ms'],4), d['checksum'])
    else: print(d)
"
