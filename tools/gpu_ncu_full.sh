#!/bin/bash
# ncu --set full captures of the dominant tensor-core kernels (one launch each, second training step of tools/one_step.py):
#   wgrad_halo_kernel  -s 11 : dec0.conv2 weight gradient, 32->32 @ 2x128^3
#   conv3_halo_kernel  -s 14 : dec0.conv2 fprop, 32->32 @ 2x128^3
#   conv3_igemm_kernel -s 14 : a 64^3-level decoder layer
mkdir -p gpurun_out
for ks in wgrad_halo_kernel:11 conv3_halo_kernel:14 conv3_igemm_kernel:14; do
  k=${ks%%:*}; sk=${ks##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $sk -c 1 -f -o gpurun_out/prof_r01_$k \
      python tools/one_step.py 2 > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
