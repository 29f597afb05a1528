#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "zstacked or tcgen05_conv" -p no:cacheprovider ) > gpurun_out/d_zs_tests.log 2>&1
timeout 600 python tools/zs_debug.py > gpurun_out/d_zs_debug.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/d_bench_cfg2.json 2> gpurun_out/d_bench_cfg2.err
tail -3 gpurun_out/d_zs_tests.log; cat gpurun_out/d_zs_debug.log; head -c 300 gpurun_out/d_bench_cfg2.json
