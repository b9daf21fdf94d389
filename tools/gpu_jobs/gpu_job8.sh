#!/bin/bash
# GPU call 8 (2 GPUs): sharded fits with the new kernels — bit-identity vs one GPU, then the N=2 bench.
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== multigpu check"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/multigpu_check.log 2>&1; echo "rc=$?"; grep -E "bit-identical|MULTIGPU|Error|error" gpurun_out/multigpu_check.log | cut -c1-200 | tail -16
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; grep '"metric"' gpurun_out/bench_n2.json | cut -c1-400; tail -2 gpurun_out/bench_n2.err
echo "== bench N=2 CD"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --solver cd --no-e2e > gpurun_out/bench_n2_cd.json 2> gpurun_out/bench_n2_cd.err; echo "rc=$?"; grep '"metric"' gpurun_out/bench_n2_cd.json | cut -c1-300
echo "== bench N=2 untiled"; RCPPML_B200_TILED=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_n2_untiled.json 2> gpurun_out/bench_n2_untiled.err; echo "rc=$?"; grep '"metric"' gpurun_out/bench_n2_untiled.json | cut -c1-300
