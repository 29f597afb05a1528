#!/bin/bash
# ncu --set full of the tensor-core kernel families in ONE training step of cfg 2 (all their launches).  The .ncu-rep stays on the box
# (> 64 MiB); the raw metric page and the source page of the dominant kernel come back as CSV.
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:"conv3_zs_kernel|wgrad_hs_kernel|wgrad_up_kernel|conv3_upzs_kernel|conv3_updzs_kernel|stem_mma" \
  -c 34 -f -o /tmp/prof_r02_cfg2 python tools/one_step.py cfg2 1 > gpurun_out/ncu_r02_cfg2.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_r02_cfg2.ncu-rep --page raw --csv > gpurun_out/prof_r02_cfg2_raw.csv 2>/dev/null
wc -l gpurun_out/prof_r02_cfg2_raw.csv
