#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "wgrad" -p no:cacheprovider ) > gpurun_out/v_wg.log 2>&1
tail -4 gpurun_out/v_wg.log
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/v_tests.log 2>&1
tail -5 gpurun_out/v_tests.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/v_bench_cfg2.json 2> gpurun_out/v_bench_cfg2.err
( B200UNET_WGRAD_HS_PG1=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/v_bench_cfg2_nopg1.json 2> gpurun_out/v_bench_cfg2_nopg1.err
for f in gpurun_out/v_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
python - <<'PY'
import json
for c in ("cfg2","cfg2_nopg1"):
    d=json.loads(open(f"gpurun_out/v_bench_{c}.json").read().strip().splitlines()[-1])
    pb=d["roofline"]["per_block"]
    print(c, d["ms_per_step"], " ".join(f"{k.split('.')[0]}.{k.split('.')[1]}.{k.split('.')[-1]}:{v['wgrad']['ms']:.3f}" for k,v in pb.items() if 'wgrad' in v))
PY
