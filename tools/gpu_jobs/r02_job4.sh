#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): sharded masked / CV fits vs one GPU, the ungated GPU suite, where the e2e wall time goes
# (library-side wall clock vs the caller's), graph on / off on a quarter-size problem (per-rank sizes of N = 8).
set -u
mkdir -p gpurun_out
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r02d_pytest_gpu.log
echo "== multigpu_check n2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r02d_multigpu_check_n2.txt 2>&1; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r02d_multigpu_check_n2.txt | tail -14
echo "== bench n2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err; echo "rc=$?"
echo "== bench n2 (trace)"
RCPPML_B200_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-parity > gpurun_out/r02d_bench_n2_trace.json 2> gpurun_out/r02d_bench_n2_trace.err; echo "rc=$?"; grep "RcppML_gpu" gpurun_out/r02d_bench_n2_trace.err | tail -3
echo "== quarter-size problem on 2 GPUs (per-rank sizes of N = 8), graph on / off"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 40 --warmup 5 --rows 250000 --cols 25000 --density 0.004 --no-e2e > gpurun_out/r02d_bench_quarter_n2.json 2> gpurun_out/r02d_bench_quarter_n2.err; echo "rc=$?"
RCPPML_B200_GRAPH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 40 --warmup 5 --rows 250000 --cols 25000 --density 0.004 --no-e2e --no-parity > gpurun_out/r02d_bench_quarter_n2_nograph.json 2> gpurun_out/r02d_bench_quarter_n2_nograph.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('r02d_bench_n2', 'r02d_bench_n2_trace', 'r02d_bench_quarter_n2', 'r02d_bench_quarter_n2_nograph'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, round(d['ms_per_step'],4), d['value'], d['gpu_launches'])
        print(' sections', {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
        print(' over_ranks', d['roofline']['over_ranks']['loop_ms_per_step'], d['roofline']['over_ranks']['profiled_loop_ms_per_step'])
        e=d['e2e']
        if e: print(' e2e', e['value'], e['seconds_total'], e.get('phases'), e.get('factors_bit_identical_to_sharded_engine'), e.get('warmup_call_seconds'))
        print(' parity', d['parity'])
    except Exception as ex:
        print(f, 'parse failed', ex)
PY
