#!/bin/bash
# A/B of the K2 run-time variants (one bench line each): usage scripts/k2_ab.sh "ENV1=.. ENV2=.." "..."
# every run sits under its own `timeout` so that a hung kernel costs 2 minutes, not the whole GPU call
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 150 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-train --no-c3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g triplets/s  %.2f ms/step ' % (d['value'], d['ms_per_step']), d['roofline']['stage_ms_per_step'], 'frac %.3f' % d['roofline']['frac'])" || echo "FAILED/timeout"
done
