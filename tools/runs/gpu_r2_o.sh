#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "stem" -p no:cacheprovider ) > gpurun_out/o_stem.log 2>&1
tail -15 gpurun_out/o_stem.log
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/o_tests.log 2>&1
tail -5 gpurun_out/o_tests.log
( time timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -s -p no:cacheprovider -k "cfg2" ) > gpurun_out/o_parity.log 2>&1
grep -E "^(UNet3D|Residual)|passed|failed" gpurun_out/o_parity.log | cut -c1-330
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/o_bench_cfg2.json 2> gpurun_out/o_bench_cfg2.err
( B200UNET_STEM_FMA=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/o_bench_cfg2_fma.json 2> gpurun_out/o_bench_cfg2_fma.err
( timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/o_bench_cfg5.json 2> gpurun_out/o_bench_cfg5.err
for f in gpurun_out/o_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
