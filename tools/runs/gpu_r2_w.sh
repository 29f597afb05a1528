#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "zs or zstack or z_stack" -p no:cacheprovider ) > gpurun_out/w_zs.log 2>&1
tail -4 gpurun_out/w_zs.log
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider --deselect tests/test_gpu_parity_configs.py ) > gpurun_out/w_tests.log 2>&1
tail -5 gpurun_out/w_tests.log
( time timeout 900 python -m pytest tests/test_gpu_parity_configs.py -q -s -p no:cacheprovider -k "cfg2" ) > gpurun_out/w_parity.log 2>&1
grep -E "^(UNet3D|Residual)|passed|failed" gpurun_out/w_parity.log | cut -c1-330
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/w_bench_cfg2.json 2> gpurun_out/w_bench_cfg2.err
( B200UNET_ZS_EPI=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/w_bench_cfg2_epi1.json 2> gpurun_out/w_bench_cfg2_epi1.err
( timeout 600 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/w_bench_cfg5.json 2> gpurun_out/w_bench_cfg5.err
for f in gpurun_out/w_bench_*.json; do echo $f; grep '^{' $f | head -c 300; echo; done
python - <<'PY'
import json
for c in ("cfg2","cfg2_epi1"):
    d=json.loads(open(f"gpurun_out/w_bench_{c}.json").read().strip().splitlines()[-1])
    pk=d["roofline"]["per_kernel"]
    print(c, d["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in pk.items()})
PY
