#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/zs_debug.py > gpurun_out/c_zs_debug.log 2>&1
( time timeout 600 python -m pytest tests/test_gpu_kernels.py -q -s -k "hstacked or wgrad" -p no:cacheprovider ) > gpurun_out/c_hs_tests.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/c_bench_cfg2_hs1.json 2> gpurun_out/c_bench_cfg2_hs1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_launches_cfg3.csv python tools/one_step.py cfg3 2 > gpurun_out/c_ncu3.log 2>&1
cat gpurun_out/c_zs_debug.log; tail -5 gpurun_out/c_hs_tests.log; head -c 300 gpurun_out/c_bench_cfg2_hs1.json
