#!/bin/bash
# Round 2, GPU call 14 (2 GPUs): final multi-GPU validation — in-process ABI path (overlapped factor-block upload) in every
# replication mode, sharded-vs-one-GPU check, bench N = 2 with the e2e trace.
set -u
mkdir -p gpurun_out
echo "== pytest (multi-GPU tests)"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "in_process_multi_gpu or factor_blocks or multi_gpu" 2>&1 | tail -3
echo "== bench n2"
RCPPML_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02n_bench_n2.json 2> gpurun_out/r02n_bench_n2.err; echo "rc=$?"; grep "RcppML_gpu" gpurun_out/r02n_bench_n2.err | tail -1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02n_bench_n2.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'],4), d['value'], d['gpu_launches'], d['config']['parallelism'][:80])
print(' sections', {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
e=d['e2e']; print(' e2e', e['value'], e['seconds_total'], e.get('phases'), e.get('factors_bit_identical_to_sharded_engine'))
print(' parity', d['parity'])
PY
