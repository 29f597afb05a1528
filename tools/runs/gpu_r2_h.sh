#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_model.py -q -s -k "fp16" -p no:cacheprovider ) > gpurun_out/h_fp16_tests.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -s -k "cfg4" -p no:cacheprovider ) > gpurun_out/h_parity4.log 2>&1
( timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/h_bench_cfg4_fp16.json 2> gpurun_out/h_bench_cfg4_fp16.err
( timeout 900 python bench.py --workload cfg4 --operand-dtype bf16 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/h_bench_cfg4_bf16.json 2> gpurun_out/h_bench_cfg4_bf16.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/h_launches_cfg4.csv python tools/one_step.py cfg4 2 > gpurun_out/h_ncu4.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/h_launches_cfg2.csv python tools/one_step.py cfg2 2 > gpurun_out/h_ncu2.log 2>&1
tail -4 gpurun_out/h_fp16_tests.log; grep -E "^(Residual|  flip)|passed|failed" gpurun_out/h_parity4.log
for f in gpurun_out/h_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
