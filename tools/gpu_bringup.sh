#!/bin/bash
# Bring-up on a fresh B200 box: every tcgen05 test in its own process (a device-side trap poisons the CUDA
# context of the process that hit it), everything bounded by `timeout`.  Output -> gpurun_out/bringup.log
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
run() { echo "=== $*" >> $LOG; timeout 240 "$@" >> $LOG 2>&1; echo "--- rc=$?" >> $LOG; }
run python -m pytest tests/test_gpu_kernels.py -q -x -k "direct and not tcgen05" -p no:cacheprovider
for id in $(python -m pytest tests/test_gpu_kernels.py --collect-only -q -k "tcgen05 or probe" 2>/dev/null | grep "::"); do
  run python -m pytest "$id" -q -p no:cacheprovider
done
B200UNET_ONLY=direct run python -m pytest tests/test_gpu_model.py -q -k "direct" -p no:cacheprovider
run python -m pytest tests/test_gpu_model.py -q -k "not direct" -p no:cacheprovider
grep -E "^(===|--- rc|FAILED|ERROR|[0-9]+ (passed|failed))|passed|failed|Error|error|rel=" $LOG | tail -150
