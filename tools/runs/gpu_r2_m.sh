#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/m_tests.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_upcat.py -q -s -k "zstacked" -p no:cacheprovider ) > gpurun_out/m_up.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -s -p no:cacheprovider -k "cfg2" ) > gpurun_out/m_parity.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/m_bench_cfg2.json 2> gpurun_out/m_bench_cfg2.err
( timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/m_bench_cfg3.json 2> gpurun_out/m_bench_cfg3.err
tail -3 gpurun_out/m_tests.log; grep -E "^up|passed|failed" gpurun_out/m_up.log | cut -c1-200; grep -E "^(UNet3D|Residual)|passed|failed" gpurun_out/m_parity.log | cut -c1-330
for f in gpurun_out/m_bench_*.json; do echo $f; grep '^{' $f | head -c 260; echo; done
