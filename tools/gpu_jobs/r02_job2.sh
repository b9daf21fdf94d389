#!/bin/bash
# Round 2, GPU call 2 (2 GPUs): everything multi-GPU that round 1 left unvalidated, on the new code —
# in-process multi-GPU behind the reference ABI (sharded ingest), block-wise factor I/O, sharded fits vs one GPU,
# CUDA-graph replay of the sharded iteration, the bench line at N = 2 (parity + e2e through the ABI) and N = 1.
#   gpurun --gpus 2 --timeout 1500 -- bash tools/gpu_jobs/r02_job2.sh
set -u
mkdir -p gpurun_out
export RCPPML_B200_TEST_ROUND2=1
echo "== pytest full"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r02b_pytest_gpu.log
echo "== multigpu_check n2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r02b_multigpu_check_n2.txt 2>&1; echo "rc=$?"; grep -v "^W\|^\*\*\*" gpurun_out/r02b_multigpu_check_n2.txt | tail -16
echo "== bench n2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "rc=$?"; tail -5 gpurun_out/r02b_bench_n2.err
echo "== bench n1"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "rc=$?"; tail -3 gpurun_out/r02b_bench_n1.err
python - <<'PY'
import json
for f in ('gpurun_out/r02b_bench_n2.json', 'gpurun_out/r02b_bench_n1.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['ms_per_step'], d['value'], d['gpu_launches'])
        print(' sections', {k: round(v,3) for k,v in d['roofline']['sections_ms_per_step'].items()})
        print(' over_ranks', d['roofline']['over_ranks']['loop_ms_per_step'])
        e=d['e2e']; print(' e2e', e['value'], e['seconds_total'], e.get('phases'), e.get('factors_bit_identical_to_sharded_engine'), e.get('warmup_call_seconds'))
        print(' parity', d['parity'])
        if 'solver_mode_0' in d: print(' cd', d['solver_mode_0']['ms_per_step'], d['solver_mode_0'].get('vs_cpu_reference'))
        if 'cpu_baseline' in d: print(' cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:60])
    except Exception as ex:
        print(f, 'parse failed', ex)
PY
echo "== reference arm under torchrun (threads)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | grep '^{' | cut -c1-700
echo "== rank shapes N=1"; timeout 600 python tools/rank_shape_sweep.py --ns 1 --out gpurun_out/r02b_rank_shapes_n1.jsonl 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: d=json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    if 'variant' in d: print(d['N'], d['half_step'], d['variant'], round(d['half_step_ms'],4), round(d['iteration_ms'],4), d['checksum'])
    else: print(d)
"
