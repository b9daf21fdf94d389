#!/bin/bash
# Round 2, GPU call 24 (2 GPUs): last sanity of the committed tree in one-process-per-GPU mode — bench N = 2.
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02v_bench_n2.json 2> gpurun_out/r02v_bench_n2.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02v_bench_n2.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'],4), d['value'], d['gpu_launches'], d['config']['parallelism'][:100])
e=d['e2e']; print(' e2e', e['value'], e['seconds_total'], e.get('factors_bit_identical_to_sharded_engine'))
print(' parity', d['parity']['ok'], d['parity']['bit_identical_to_n1'])
PY
