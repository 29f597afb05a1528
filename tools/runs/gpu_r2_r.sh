#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/r_tests.log 2>&1
tail -5 gpurun_out/r_tests.log
( time timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -s -p no:cacheprovider -k "cfg2" ) > gpurun_out/r_parity.log 2>&1
grep -E "^(UNet3D|Residual)|passed|failed" gpurun_out/r_parity.log | cut -c1-330
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/r_bench_cfg2.json 2> gpurun_out/r_bench_cfg2.err
( B200UNET_SIDE_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/r_bench_cfg2_noside.json 2> gpurun_out/r_bench_cfg2_noside.err
( timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/r_bench_cfg3.json 2> gpurun_out/r_bench_cfg3.err
( B200UNET_SIDE_STREAM=0 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/r_bench_cfg3_noside.json 2> gpurun_out/r_bench_cfg3_noside.err
for f in gpurun_out/r_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
