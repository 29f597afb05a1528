#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/y_tests.log 2>&1
tail -5 gpurun_out/y_tests.log
( time timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -s -p no:cacheprovider -k cfg2 ) > gpurun_out/y_parity.log 2>&1
grep -E "^(UNet3D|Residual)|passed|failed" gpurun_out/y_parity.log | cut -c1-330
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/y_bench_cfg2.json 2> gpurun_out/y_bench_cfg2.err
( timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/y_bench_cfg5.json 2> gpurun_out/y_bench_cfg5.err
for f in gpurun_out/y_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_cfg2.csv python tools/one_step.py cfg2 2 > gpurun_out/y_one_step.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r02_cfg2.csv
