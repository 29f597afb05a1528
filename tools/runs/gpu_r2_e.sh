#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/e_tests.log 2>&1
timeout 600 python tools/zs_debug.py > gpurun_out/e_zs_debug.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -s -p no:cacheprovider ) > gpurun_out/e_parity.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/e_bench_cfg2.json 2> gpurun_out/e_bench_cfg2.err
( timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/e_bench_cfg3.json 2> gpurun_out/e_bench_cfg3.err
( B200UNET_EXPLICIT_GN=0 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/e_bench_cfg3_fold.json 2> gpurun_out/e_bench_cfg3_fold.err
tail -4 gpurun_out/e_tests.log; grep -E "^\(2, 1|^\(2, 6" gpurun_out/e_zs_debug.log | grep -E " zs  |halo" ; tail -3 gpurun_out/e_parity.log
for f in gpurun_out/e_bench_*.json; do echo $f; head -c 260 $f; echo; done
