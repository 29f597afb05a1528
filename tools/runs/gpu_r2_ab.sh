#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/ab_tests.log 2>&1
tail -5 gpurun_out/ab_tests.log
( time timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -s -p no:cacheprovider -k cfg2 ) > gpurun_out/ab_parity.log 2>&1
grep -E "^\.?(UNet3D|Residual)|passed|failed" gpurun_out/ab_parity.log | cut -c1-330
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/ab_bench_cfg2.json 2> gpurun_out/ab_bench_cfg2.err
( timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/ab_bench_cfg3.json 2> gpurun_out/ab_bench_cfg3.err
( timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/ab_bench_cfg5.json 2> gpurun_out/ab_bench_cfg5.err
for f in gpurun_out/ab_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
python - <<'PY'
import json
d=json.loads(open("gpurun_out/ab_bench_cfg2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d["roofline"]["per_kernel"].items()}, d["roofline"]["frac"])
PY
