#!/bin/bash
# round 2, GPU call A: full GPU test suite, smoke, and the four benchmark workloads at N=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/a_gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 --deselect tests/test_gpu_parity_configs.py -p no:cacheprovider ) > gpurun_out/a_tests.log 2>&1
echo "rc=$?" >> gpurun_out/a_tests.log
( time timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -s -p no:cacheprovider ) > gpurun_out/a_parity.log 2>&1
echo "rc=$?" >> gpurun_out/a_parity.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/a_smoke.log 2>&1
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/a_bench_cfg2.json 2> gpurun_out/a_bench_cfg2.err
( time timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/a_bench_cfg3.json 2> gpurun_out/a_bench_cfg3.err
( time timeout 900 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/a_bench_cfg4.json 2> gpurun_out/a_bench_cfg4.err
( time timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/a_bench_cfg5.json 2> gpurun_out/a_bench_cfg5.err
tail -3 gpurun_out/a_tests.log; tail -3 gpurun_out/a_parity.log; tail -2 gpurun_out/a_smoke.log
for f in gpurun_out/a_bench_*.json; do echo $f; head -c 600 $f; echo; done
