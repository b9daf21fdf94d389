#!/bin/bash
# Round 2, GPU call 8 (4 GPUs): replication mode at N = 4 — unicast peer stores (RCPPML_B200_MC=0), multicast from the
# Gram kernel (=1), multicast from the solve kernel (=2): sharded-vs-one-GPU check in every mode, then the bench.
set -u
mkdir -p gpurun_out
echo "== multigpu_check n4"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r02h_multigpu_check_n4.txt 2>&1; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r02h_multigpu_check_n4.txt | grep -c "bit-identical=True"; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r02h_multigpu_check_n4.txt | grep -v "bit-identical=True" | tail -5
for mc in 2 1 0; do
  extra="--no-e2e --no-parity"; [ $mc = 2 ] && extra=""
  RCPPML_B200_MC=$mc RCPPML_B200_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2952$mc bench.py --gpus 4 --steps 20 --warmup 5 $extra > gpurun_out/r02h_bench_n4_mc$mc.json 2> gpurun_out/r02h_bench_n4_mc$mc.err; echo "== bench n4 mc=$mc rc=$?"
done
python - <<'PY'
import json
for f in ('r02h_bench_n4_mc2', 'r02h_bench_n4_mc1', 'r02h_bench_n4_mc0'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, round(d['ms_per_step'],4), d['value'], d['gpu_launches'], d['config']['parallelism'][:70])
        print(' sections', {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
        e=d['e2e']
        if e: print(' e2e', e['value'], e['seconds_total'], e.get('phases'), e.get('factors_bit_identical_to_sharded_engine'))
        print(' parity', d['parity'])
    except Exception as ex:
        print(f, 'parse failed', ex)
PY
grep "RcppML_gpu" gpurun_out/r02h_bench_n4_mc2.err | tail -2
