#!/bin/bash
# launch list (ncu device times) of one cfg4 training step (ResidualUNetSE3D f64, 1x160^3, fp16 operands)
mkdir -p gpurun_out
B200UNET_OPERAND_DTYPE=fp16 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_cfg4.csv python tools/one_step.py cfg4 2 > gpurun_out/cfg4_one_step.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r02_cfg4.csv; tail -2 gpurun_out/cfg4_one_step.log
