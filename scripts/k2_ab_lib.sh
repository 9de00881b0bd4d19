#!/bin/bash
# A/B of two builds of the library on one B200: parity of the experiment build on the scorer tests, then alternating bench lines.
# usage: scripts/k2_ab_lib.sh higashi_b200/libhsg_exp.so [more experiment libraries ...]
EXP=$(readlink -f "$1")
mkdir -p gpurun_out
HSG_LIB_PATH=$EXP timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -k "impute_golden_tensor_core or synthetic_vs_oracle or small_genomes or full_size or range_guard" > gpurun_out/ab_exp_tests.log 2>&1; echo "exp tests rc=$?"; tail -2 gpurun_out/ab_exp_tests.log
B="--steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-c3"
for rep in 1 2; do
  for which in base "$@"; do
    if [ $which = base ]; then unset HSG_LIB_PATH; else export HSG_LIB_PATH=$(readlink -f $which); fi
    timeout 200 python bench.py $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$which %.4g triplets/s  %.2f ms/step ' % (d['value'], d['ms_per_step']), d['roofline']['stage_ms_per_step'], 'frac %.3f' % d['roofline']['frac'])" || echo "$which FAILED/timeout"
  done
done
