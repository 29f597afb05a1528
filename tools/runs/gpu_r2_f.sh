#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/f_tests.log 2>&1
timeout 600 python tools/zs_debug.py > gpurun_out/f_zs_debug.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/f_bench_cfg2.json 2> gpurun_out/f_bench_cfg2.err
( timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/f_bench_cfg3.json 2> gpurun_out/f_bench_cfg3.err
( timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/f_bench_cfg5.json 2> gpurun_out/f_bench_cfg5.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_cfg3.csv python tools/one_step.py cfg3 2 > gpurun_out/f_ncu3.log 2>&1
tail -4 gpurun_out/f_tests.log; grep -E " zs  |halo" gpurun_out/f_zs_debug.log
for f in gpurun_out/f_bench_*.json; do echo $f; head -c 260 $f; echo; done
