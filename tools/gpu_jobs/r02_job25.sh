#!/bin/bash
# Round 2, GPU call 25 (1 GPU): the shapes of the reference's own GPU vignette, end to end through the reference ABI.
set -u
mkdir -p gpurun_out
timeout 200 python tools/vignette_shapes.py --out gpurun_out/r02w_vignette_shapes.json 2>&1 | cut -c1-330
