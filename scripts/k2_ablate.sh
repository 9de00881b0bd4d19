#!/bin/bash
# K2 ablations on a shrunk C2 (1200 cells, 600 per step): usage scripts/k2_ablate.sh "ENV=.." ...   (results are WRONG with HSG_K2_EXP != 0)
for cfg in "$@"; do
  echo -n "== $cfg : "
  env $cfg timeout 120 python bench.py --steps 4 --warmup 2 --n-cells 1200 --cells-per-step 600 --no-cpu-baseline --no-e2e --no-train --no-c3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('K2 %.2f ms  step %.2f ms' % (d['roofline']['stage_ms_per_step']['scorer_K2'], d['ms_per_step']))" || echo "FAILED/timeout"
done
