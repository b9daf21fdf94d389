#!/bin/bash
# Round 2, GPU call 15 (8 GPUs): the final code on 8 GPUs — C4 bench line (multicast stores from the solve kernel,
# overlapped factor-block upload in the e2e call).
set -u
mkdir -p gpurun_out
RCPPML_B200_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02o_bench_c4_n8.json 2> gpurun_out/r02o_bench_c4_n8.err
echo "rc=$?"; grep "RcppML_gpu" gpurun_out/r02o_bench_c4_n8.err | tail -1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02o_bench_c4_n8.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'],4), d['value'], d['gpu_launches'], d['config']['parallelism'][:80])
print(' sections', {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
o=d['roofline']['over_ranks']; print(' over_ranks', o['loop_ms_per_step'], o['profiled_loop_ms_per_step'])
e=d['e2e']; print(' e2e', e['value'], e['seconds_total'], e.get('phases'), e.get('factors_bit_identical_to_sharded_engine'))
print(' parity', d['parity'])
PY
