#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/ac_tests.log 2>&1
tail -4 gpurun_out/ac_tests.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/ac_bench_cfg2.json 2> gpurun_out/ac_bench_cfg2.err
( timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/ac_bench_cfg3.json 2> gpurun_out/ac_bench_cfg3.err
( timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/ac_bench_cfg4.json 2> gpurun_out/ac_bench_cfg4.err
for f in gpurun_out/ac_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
