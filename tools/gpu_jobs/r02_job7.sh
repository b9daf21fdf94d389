#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): NVSwitch multicast replication of the factors (vmm.hpp, multimem.st from the Gram kernel):
# GPU suite (in-process ABI path with multicast and unicast), sharded-vs-one-GPU check in every mode, bench N = 2.
set -u
mkdir -p gpurun_out
echo "== pytest (multi-GPU tests first)"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "in_process_multi_gpu or factor_blocks or multi_gpu" 2>&1 | tail -15
echo "== multigpu_check n2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r02g_multigpu_check_n2.txt 2>&1; echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r02g_multigpu_check_n2.txt | tail -36
echo "== bench n2"
RCPPML_B200_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err; echo "rc=$?"; grep "RcppML_gpu" gpurun_out/r02g_bench_n2.err | tail -2
echo "== bench n2 unicast"
RCPPML_B200_MC=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > gpurun_out/r02g_bench_n2_unicast.json 2> gpurun_out/r02g_bench_n2_unicast.err; echo "rc=$?"
echo "== quarter-size problem on 2 GPUs, multicast / unicast"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 40 --warmup 5 --rows 250000 --cols 25000 --density 0.004 --no-e2e > gpurun_out/r02g_bench_quarter_n2.json 2> gpurun_out/r02g_bench_quarter_n2.err; echo "rc=$?"
RCPPML_B200_MC=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 40 --warmup 5 --rows 250000 --cols 25000 --density 0.004 --no-e2e --no-parity > gpurun_out/r02g_bench_quarter_n2_unicast.json 2> gpurun_out/r02g_bench_quarter_n2_unicast.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('r02g_bench_n2', 'r02g_bench_n2_unicast', 'r02g_bench_quarter_n2', 'r02g_bench_quarter_n2_unicast'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json' % f) if l.startswith('{')][-1])
        print(f, round(d['ms_per_step'],4), d['value'], d['gpu_launches'], d['config']['parallelism'][:90])
        print(' sections', {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
        e=d['e2e']
        if e: print(' e2e', e['value'], e['seconds_total'], e.get('phases'), e.get('factors_bit_identical_to_sharded_engine'))
        print(' parity', d['parity'])
    except Exception as ex:
        print(f, 'parse failed', ex)
PY
tail -5 gpurun_out/r02g_bench_n2.err | cut -c1-300
