#!/bin/bash
# Round 2, GPU call 3 (2 GPUs): where the in-process ABI call's wall time goes (RCPPML_B200_TRACE), CUDA-graph replay of
# the sharded iteration on/off at N = 2, and a quarter-size problem on 2 GPUs as a proxy of the per-rank sizes of N = 8.
set -u
mkdir -p gpurun_out
run() { # name, env..., -- args
    local name=$1; shift
    timeout 600 env "$@"
}
echo "== bench n2 (trace)"
RCPPML_B200_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err; echo "rc=$?"; grep "RcppML_gpu" gpurun_out/r02c_bench_n2.err | tail -4
echo "== bench n2, no graph"
RCPPML_B200_GRAPH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-parity > gpurun_out/r02c_bench_n2_nograph.json 2> gpurun_out/r02c_bench_n2_nograph.err; echo "rc=$?"
echo "== quarter-size problem on 2 GPUs (per-rank sizes of N = 8), graph on / off"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 40 --warmup 5 --m 250000 --n 25000 --density 0.004 --no-e2e > gpurun_out/r02c_bench_quarter_n2.json 2> gpurun_out/r02c_bench_quarter_n2.err; echo "rc=$?"
RCPPML_B200_GRAPH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 40 --warmup 5 --m 250000 --n 25000 --density 0.004 --no-e2e --no-parity > gpurun_out/r02c_bench_quarter_n2_nograph.json 2> gpurun_out/r02c_bench_quarter_n2_nograph.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('r02c_bench_n2', 'r02c_bench_n2_nograph', 'r02c_bench_quarter_n2', 'r02c_bench_quarter_n2_nograph'):
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
echo "== bench n1"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r02c_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02c_bench_n1.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'],4), d['value'], {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()}, d['e2e']['seconds_total'], d['solver_mode_0']['ms_per_step'])
PY
