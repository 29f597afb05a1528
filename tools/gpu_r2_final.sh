#!/bin/bash
# round-2 final records: smoke, the default bench line (with cpu / torch-gpu baselines), the other workloads, launch list of cfg2
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/final_smoke.log 2>&1
tail -3 gpurun_out/final_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/final_bench_cfg2.json 2> gpurun_out/final_bench_cfg2.err
( timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/final_bench_cfg3.json 2> gpurun_out/final_bench_cfg3.err
( timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/final_bench_cfg4.json 2> gpurun_out/final_bench_cfg4.err
( timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/final_bench_cfg5.json 2> gpurun_out/final_bench_cfg5.err
for f in gpurun_out/final_bench_*.json; do echo $f; grep '^{' $f | head -c 400; echo; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_cfg2.csv python tools/one_step.py cfg2 2 > gpurun_out/final_one_step.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r02_cfg2.csv
