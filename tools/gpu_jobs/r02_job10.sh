#!/bin/bash
# Round 2, GPU call 10 (2 GPUs): multicast from the solve kernel (mode 2) vs unicast at N = 2 with the stream priorities;
# in-process ABI path in every replication mode (pytest).
set -u
mkdir -p gpurun_out
echo "== pytest (multi-GPU tests)"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "in_process_multi_gpu or factor_blocks or multi_gpu" 2>&1 | tail -3
for mc in 2 0; do
  RCPPML_B200_MC=$mc timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$mc bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r02j_bench_n2_mc$mc.json 2> gpurun_out/r02j_bench_n2_mc$mc.err; echo "== bench n2 mc=$mc rc=$?"
  RCPPML_B200_MC=$mc timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$mc bench.py --gpus 2 --steps 40 --warmup 5 --rows 250000 --cols 25000 --density 0.004 --no-e2e --no-parity > gpurun_out/r02j_bench_quarter_n2_mc$mc.json 2> gpurun_out/r02j_bench_quarter_n2_mc$mc.err; echo "== quarter n2 mc=$mc rc=$?"
done
python - <<'PY'
import json
for f in ('r02j_bench_n2_mc2', 'r02j_bench_n2_mc0', 'r02j_bench_quarter_n2_mc2', 'r02j_bench_quarter_n2_mc0'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, round(d['ms_per_step'],4), d['value'], d['gpu_launches'], d['config']['parallelism'][:70])
        print(' sections', {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
        print(' parity', d['parity'])
    except Exception as ex:
        print(f, 'parse failed', ex)
PY
