#!/bin/bash
# GPU call 12 (4 GPUs): sharded bit-identity at world=4 and the N=4 bench.
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
echo "== multigpu check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/multigpu_check_n$N.log 2>&1; echo "rc=$?"; grep -E "bit-identical|MULTIGPU|Error|error" gpurun_out/multigpu_check_n$N.log | cut -c1-160 | tail -14
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
python - <<PY
import json
for l in open('gpurun_out/bench_n$N.json'):
    if l.startswith('{"metric"'):
        d=json.loads(l); print(d['n_gpus'], round(d['ms_per_step'],3), d['value'], {k:round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()}, 'e2e', d['e2e'] and (d['e2e']['value'], d['e2e']['seconds_total']))
PY
tail -2 gpurun_out/bench_n$N.err
