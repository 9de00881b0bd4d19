#!/bin/bash
# Round-2 evidence on one B200: the whole GPU test suite, the ncu launch list + one full capture of the generation-4 scorer, and the
# bench lines of C1 / C3 / C5 at their full cell counts and of the default invocation.  Output under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -14 gpurun_out/gpu_tests.log
S="--steps 2 --warmup 1 --n-cells 240 --cells-per-step 120 --no-cpu-baseline --no-e2e --no-train --no-c3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py $S > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:score_umma4 -s 2 -c 1 -o gpurun_out/k2gen4 -f python bench.py $S > gpurun_out/ncu_k2.log 2>&1
B="--no-cpu-baseline --no-train --no-c3"
timeout 300 python bench.py --workload C1 --cells-per-step 620 $B > gpurun_out/r02_bench_c1.json 2> gpurun_out/c1.err
timeout 500 python bench.py --workload C3 --cells-per-step 250 --steps 6 --warmup 3 $B > gpurun_out/r02_bench_c3.json 2> gpurun_out/c3.err
timeout 600 python bench.py --workload C5 --cells-per-step 200 --steps 6 --warmup 3 $B > gpurun_out/r02_bench_c5.json 2> gpurun_out/c5.err
timeout 600 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/default.err
ls -la gpurun_out
