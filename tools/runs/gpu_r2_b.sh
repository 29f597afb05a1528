#!/bin/bash
# round 2, GPU call B: the z-stacked conv kernel -- parity, then step time with it on / off, launch list
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_kernels.py -q -s -k "zstacked or tcgen05_conv" -p no:cacheprovider ) > gpurun_out/b_zs_tests.log 2>&1
echo "rc=$?" >> gpurun_out/b_zs_tests.log
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider ) > gpurun_out/b_tests.log 2>&1
echo "rc=$?" >> gpurun_out/b_tests.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/b_bench_cfg2_zs1.json 2> gpurun_out/b_bench_cfg2_zs1.err
( B200UNET_ZS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/b_bench_cfg2_zs0.json 2> gpurun_out/b_bench_cfg2_zs0.err
( timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/b_bench_cfg5_zs1.json 2> gpurun_out/b_bench_cfg5_zs1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches_cfg2.csv python tools/one_step.py cfg2 2 > gpurun_out/b_ncu.log 2>&1
tail -4 gpurun_out/b_zs_tests.log; tail -4 gpurun_out/b_tests.log
for f in gpurun_out/b_bench_*.json; do echo $f; head -c 300 $f; echo; done
