#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_upcat.py tests/test_gpu_options.py -q -s -p no:cacheprovider ) > gpurun_out/j_up_tests.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/j_tests.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/j_bench_cfg2.json 2> gpurun_out/j_bench_cfg2.err
( timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/j_bench_cfg5.json 2> gpurun_out/j_bench_cfg5.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j_launches_cfg2.csv python tools/one_step.py cfg2 2 > gpurun_out/j_ncu2.log 2>&1
grep -E "^upzs|passed|failed" gpurun_out/j_up_tests.log | tail -8; tail -3 gpurun_out/j_tests.log
for f in gpurun_out/j_bench_*.json; do echo $f; grep '^{' $f | head -c 260; echo; done
