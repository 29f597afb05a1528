#!/bin/bash
# ncu --set full of every tensor-core kernel family in ONE training step of cfg 2 (all their launches), plus two HBM-bound kernels
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:"conv3_zs_kernel|wgrad_hs_kernel|conv3_upzs_kernel|conv3_igemm_kernel|conv3_wgrad_igemm_kernel|wgrad_halo_kernel|stem_conv_fwd_tiled|gn_bwd_apply_kernel|maxpool_bwd_kernel" \
  -c 80 -f -o gpurun_out/prof_r02_cfg2 python tools/one_step.py cfg2 1 > gpurun_out/ncu_r02_cfg2.log 2>&1
echo "rc=$?"; ls -la gpurun_out/*.ncu-rep
ncu -i gpurun_out/prof_r02_cfg2.ncu-rep --page raw --csv > gpurun_out/prof_r02_cfg2_raw.csv 2>/dev/null
wc -l gpurun_out/prof_r02_cfg2_raw.csv
