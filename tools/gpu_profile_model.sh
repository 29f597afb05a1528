#!/bin/bash
# ncu launch list of ONE training step of a named model: tools/gpu_profile_model.sh <name> <f_maps> <batch> <size> <tag>
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$5.csv \
    python tools/one_step_model.py $1 $2 $3 $4 2 > gpurun_out/prof_$5.log 2>&1
echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches_$5.csv "$1 f$2 batch $3x$4^3" > gpurun_out/launches_$5.md 2>&1
head -60 gpurun_out/launches_$5.md
