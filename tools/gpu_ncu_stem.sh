#!/bin/bash
# ncu --set full of the two first-conv kernels (one launch each) in one cfg2 training step; raw + details pages come back as CSV/text
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stem_" -c 2 -f -o /tmp/prof_stem python tools/one_step.py cfg2 1 > gpurun_out/ncu_stem.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_stem.ncu-rep --page raw --csv > gpurun_out/prof_stem_raw.csv 2>/dev/null
ncu -i /tmp/prof_stem.ncu-rep --page details > gpurun_out/prof_stem_details.txt 2>/dev/null
ncu -i /tmp/prof_stem.ncu-rep --page source --csv > gpurun_out/prof_stem_source.csv 2>/dev/null
ls -la gpurun_out/prof_stem*
