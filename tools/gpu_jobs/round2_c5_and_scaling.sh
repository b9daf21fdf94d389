#!/bin/bash
# Round 2: the measurements round 1 left open (DESIGN.md §8 "still open" (3), (0b)) — needs 8 GPUs:
#   gpurun --gpus 8 --timeout 900 -- bash tools/gpu_jobs/round2_c5_and_scaling.sh
# 1. C4 at N = 8 with the current kernels (last measured before the tiled kernel and the side-stream overlap), default
#    e2e and the block-wise factor I/O variant;
# 2. BASELINE.json configs[4]: C5 (5M x 500K, 0.05 %, k = 128, L1 = L2 = 0.01) on 8 GPUs, Cholesky and CD.
set -u
mkdir -p gpurun_out
run() {  # name, bench args...
    local name=$1; shift
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
        bench.py --gpus 8 "$@" > gpurun_out/${name}.json 2> gpurun_out/${name}.err
    echo "== ${name}: rc=$?"; tail -c 600 gpurun_out/${name}.json; echo
}
run bench_c4_n8 --steps 20 --warmup 3 --no-cpu
run bench_c4_n8_blocks --steps 20 --warmup 3 --no-cpu --no-cd --e2e-blocks
run bench_c5_n8 --steps 5 --warmup 3 --m 5000000 --n 500000 --density 5e-4 --k 128 --L1 0.01 --L2 0.01 --no-cpu
