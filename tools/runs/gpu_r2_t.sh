#!/bin/bash
# launch list of one cfg2 training step (ncu device times, serialised) + host-side enqueue profile
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_cfg2.csv python tools/one_step.py cfg2 2 > gpurun_out/t_one_step.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_r02_cfg2.csv
timeout 600 python tools/host_profile.py UNet3D 32 2 128 > gpurun_out/t_host.log 2>&1
cat gpurun_out/t_host.log
