#!/bin/bash
# ncu --set full of the three virtual-concat launches of decoder 0 (first training step)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3_wgrad_igemm_kernel -s 0 -c 1 -f -o gpurun_out/prof_up_wgrad \
    python tools/one_step.py 1 > gpurun_out/ncu_up_wgrad.log 2>&1
echo "wgrad rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3_igemm_kernel -s 10 -c 2 -f -o gpurun_out/prof_up_fwd_dgrad \
    python tools/one_step.py 1 > gpurun_out/ncu_up_fwd_dgrad.log 2>&1
echo "fwd/dgrad rc=$?"
ls -la gpurun_out/*.ncu-rep
