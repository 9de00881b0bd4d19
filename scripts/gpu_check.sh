#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training.py tests/test_gpu_wrapper.py -x -q -k "gather or impute_golden or training or forward or backward or module or pipeline or stage1 or full_size" > gpurun_out/check_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/check_tests.log
tail -4 gpurun_out/check_tests.log
B="--no-cpu-baseline --no-e2e --no-train --no-c3"
HSG_GATHER_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 1 $B 2>&1 | grep "overflow tokens" | sort | uniq -c | head -5
timeout 300 python bench.py --steps 6 --warmup 3 $B > gpurun_out/chk_c2.json 2> gpurun_out/chk_c2.err
HSG_GATHER_HASH=0 timeout 300 python bench.py --steps 6 --warmup 3 $B > gpurun_out/chk_c2_nohash.json 2> gpurun_out/chk_c2_nohash.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/chk_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['roofline']['stage_ms_per_step'], d['roofline'].get('gather_K1a', {}).get('frac'))
    except Exception as e:
        print(f, 'ERR', e)
P
