#!/bin/bash
# Round 2, GPU call 18 (1 GPU): 4-column batches (solve geometry = gather geometry, 8 lanes x 2 words) for the tiled
# Cholesky kernel, against the 8-column default, on the C4 W half-step and its N = 8 rank shape.
set -u
mkdir -p gpurun_out
timeout 600 python tools/rank_shape_sweep.py --ns 1,8 --half-steps W --variants tiled8,tiled4 --steps 8 --out gpurun_out/r02r_wstep_tiled4.jsonl 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    if 'variant' in d: print(d['N'], d['half_step'], d['variant'], round(d['half_step_ms'],4), round(d['iteration_ms'],4), d['checksum'])
    else: print(d)
"
