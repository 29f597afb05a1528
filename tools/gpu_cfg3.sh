#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_cfg3.py > gpurun_out/cfg3.log 2>&1
echo "rc=$?" >> gpurun_out/cfg3.log
tail -8 gpurun_out/cfg3.log
