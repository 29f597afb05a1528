#!/bin/bash
# ncu --set full captures of the dominant tensor-core kernels (one launch each, second training step).
mkdir -p gpurun_out
for k in wgrad_halo_kernel conv3_halo_kernel conv3_igemm_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 14 -c 1 -f -o gpurun_out/prof_r01_$k \
      python tools/one_step.py 2 > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
