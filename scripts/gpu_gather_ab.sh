#!/bin/bash
# K1a on the GPU box: parity tests of the gather kernels, then the per-stage times of C2 / C1 with the streaming kernel.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "gather or impute_golden_fp32" > gpurun_out/gather_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gather_tests.log
tail -3 gpurun_out/gather_tests.log
B="--no-cpu-baseline --no-e2e --no-train --no-c3"
HSG_GATHER_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 $B 2>&1 | grep "overflow tokens" | sort | uniq -c | head -5
timeout 300 python bench.py --steps 6 --warmup 3 $B > gpurun_out/ab_c2_stream.json 2> gpurun_out/ab_c2_stream.err
timeout 300 python bench.py --workload C1 --cells-per-step 620 --steps 6 --warmup 3 $B > gpurun_out/ab_c1_stream.json 2> gpurun_out/ab_c1_stream.err
HSG_GATHER_STREAM=0 timeout 300 python bench.py --workload C1 --cells-per-step 620 --steps 6 --warmup 3 $B > gpurun_out/ab_c1_old.json 2> gpurun_out/ab_c1_old.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['roofline']['stage_ms_per_step'], d['roofline'].get('gather_K1a', {}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e)
P
