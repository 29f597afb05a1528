#!/bin/bash
# Full GPU test pass + smoke + a short bench.  Output -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/tests.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s >> gpurun_out/tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench.log
grep -E "passed|failed|rc=|FAILED|Error|error|^block_|^unet|probe|smoke" gpurun_out/tests.log | tail -80
tail -5 gpurun_out/smoke.log
tail -5 gpurun_out/bench.log
