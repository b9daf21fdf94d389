#!/bin/bash
# Round 2, GPU call 11 (8 GPUs): the round-2 loop on 8 GPUs — multicast stores from the solve kernel (default),
# unicast peer stores (RCPPML_B200_MC=0) and the un-throttled re-normalisation kernel (RCPPML_B200_SIDE_CTAS=8) for
# comparison; C5; the sharded-vs-one-GPU check in every replication mode.
#   gpurun --gpus 8 --timeout 780 -- bash tools/gpu_jobs/r02_job11_n8.sh
set -u
mkdir -p gpurun_out
tr() { # name timeout port args...
    local name=$1 to=$2 port=$3; shift 3
    RCPPML_B200_TRACE=1 timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus 8 "$@" > gpurun_out/${name}.json 2> gpurun_out/${name}.err
    echo "== ${name}: rc=$?"; grep "RcppML_gpu" gpurun_out/${name}.err | tail -1
}
tr r02k_bench_c4_n8 200 29541 --steps 20 --warmup 5
RCPPML_B200_MC=0 tr r02k_bench_c4_n8_unicast 150 29542 --steps 20 --warmup 5 --no-e2e --no-parity
RCPPML_B200_SIDE_CTAS=8 tr r02k_bench_c4_n8_side8 150 29543 --steps 20 --warmup 5 --no-e2e --no-parity
tr r02k_bench_c5_n8 330 29544 --steps 5 --warmup 3 --rows 5000000 --cols 500000 --density 5e-4 --rank 128 --L1 0.01 --L2 0.01
echo "== multigpu_check n8"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 tests/multigpu_check.py > gpurun_out/r02k_multigpu_check_n8.txt 2>&1; echo "rc=$?"; grep -c "bit-identical=True" gpurun_out/r02k_multigpu_check_n8.txt; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r02k_multigpu_check_n8.txt | grep -v "bit-identical=True" | tail -4
python - <<'PY'
import json
for f in ('r02k_bench_c4_n8', 'r02k_bench_c4_n8_unicast', 'r02k_bench_c4_n8_side8', 'r02k_bench_c5_n8'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, round(d['ms_per_step'],4), d['value'], d['gpu_launches'], d['config']['parallelism'][:70])
        print(' sections', {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
        o=d['roofline']['over_ranks']; print(' over_ranks', o['loop_ms_per_step'], o['profiled_loop_ms_per_step'])
        print(' sections min/max', {k: (round(v['min'],3), round(v['max'],3)) for k,v in o['sections_ms_per_step'].items()})
        e=d['e2e']
        if e: print(' e2e', e['value'], e['seconds_total'], e.get('phases'), e.get('factors_bit_identical_to_sharded_engine'))
        print(' parity', d['parity'])
    except Exception as ex:
        print(f, 'parse failed', ex)
PY
