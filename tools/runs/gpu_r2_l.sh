#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/l_tests.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -s -p no:cacheprovider -k "cfg2 or cfg3" ) > gpurun_out/l_parity.log 2>&1
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/l_bench_cfg2.json 2> gpurun_out/l_bench_cfg2.err
( B200UNET_FUSED_GN_BWD=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/l_bench_cfg2_nofuse.json 2> gpurun_out/l_bench_cfg2_nofuse.err
( timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/l_bench_cfg3.json 2> gpurun_out/l_bench_cfg3.err
( timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/l_bench_cfg4.json 2> gpurun_out/l_bench_cfg4.err
tail -3 gpurun_out/l_tests.log; grep -E "^(UNet3D|Residual)|passed|failed" gpurun_out/l_parity.log | cut -c1-300
for f in gpurun_out/l_bench_*.json; do echo $f; grep '^{' $f | head -c 260; echo; done
bash tools/gpu_ncu_full_r02.sh
