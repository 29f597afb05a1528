#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "stem" -p no:cacheprovider ) > gpurun_out/q_stem.log 2>&1
tail -4 gpurun_out/q_stem.log
( time timeout 600 python -m pytest tests/test_gpu_upcat.py -q -s -k "up_wgrad_stacked or up_phase" -p no:cacheprovider ) > gpurun_out/q_up.log 2>&1
grep -E "^up wgrad|passed|failed|Error|error" gpurun_out/q_up.log | cut -c1-220 | tail -30
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/q_tests.log 2>&1
tail -5 gpurun_out/q_tests.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/q_bench_cfg2.json 2> gpurun_out/q_bench_cfg2.err
( B200UNET_UP_WGRAD_HS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/q_bench_cfg2_noup.json 2> gpurun_out/q_bench_cfg2_noup.err
for f in gpurun_out/q_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
